"""TEST INFRASTRUCTURE — generates tests/golden/ref_*.npz by driving the reference's OWN code
(oracle/_ref/libvsom_ref.so = /root/reference/src/*.cpp compiled unmodified, see oracle/Makefile).

Needs /root/reference (this container); the fixtures it writes are committed and travel to the GPU box.
Usage: python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, W, H, Din, transform, decay, eta, rows, sigmas (one train_rows segment each)
CASES = [
    ("std_exp", 12, 9, 17, po.STANDARD, po.EXPONENTIAL, 0.2, 240, (3.0, 1.4, 1.0)),
    ("std_inv", 9, 12, 10, po.STANDARD, po.INVERSE, 0.2, 240, (2.2, 1.3, 1.0)),
    ("med_exp", 8, 8, 33, po.MEDIAN, po.EXPONENTIAL, 0.05, 240, (2.5, 1.2, 1.0)),
    ("med_inv", 5, 7, 4, po.MEDIAN, po.INVERSE, 0.05, 120, (1.9, 1.0)),
    ("clr_exp", 6, 5, 6, po.CLR, po.EXPONENTIAL, 0.01, 200, (2.0, 1.5, 1.0)),
    ("clr_inv", 4, 6, 5, po.CLR, po.INVERSE, 0.01, 160, (1.8, 1.0)),
]


def main():
    if not po.Reference.available():
        po.build(quiet=False)
    assert po.Reference.available(), "oracle/_ref/libvsom_ref.so missing (needs /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    for ci, (name, W, H, Din, tr, dec, eta, rows, sigmas) in enumerate(CASES):
        rng = np.random.default_rng(1000 + ci)
        r = po.Reference(W, H, Din, tr)
        r.random_initialize(42, 1.0)
        init = r.get_state()
        if tr == po.CLR:
            z = rng.standard_normal((rows * len(sigmas), 1)).astype(np.float32)
            a = rng.uniform(0.5, 1.5, (1, Din)).astype(np.float32)
            x = (a * z + 0.1 * rng.standard_normal((rows * len(sigmas), Din))).astype(np.float32)
        else:
            x = rng.standard_normal((rows * len(sigmas), Din)).astype(np.float32)
        out = dict(W=W, H=H, Din=Din, transform=tr, decay=dec, eta=eta, sigmas=np.array(sigmas), rows=rows, x=x, init_mean=init["mean"])
        for si, sg in enumerate(sigmas):
            seg = x[si * rows:(si + 1) * rows]
            bmu, dist, resid2, last = r.train_rows(seg, eta, sg, dec)
            st = r.get_state()
            out[f"bmu{si}"], out[f"dist{si}"], out[f"resid2{si}"], out[f"last{si}"] = bmu, dist, resid2, last
            for k, v in st.items():
                out[f"{k}{si}"] = v
        out["umatrix"] = r.update_umatrix()
        q = x[:64]
        out["score_bmu"], out["score_dist"] = r.find_bmu(q)
        out["restricted_bmu"] = r.find_restricted_bmu(q, 2)
        out["all_dists_row0"] = r.all_dists(q[0])
        if tr != po.CLR:  # Som::evaluate mixes Dm- and Din-sized vectors for CLR (src/Som.cpp:509): undefined in the reference
            out["evaluate"] = np.float64(r.evaluate(q))
        out["eigen_kind"] = np.array(po.Reference.eigen_kind())
        np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
        print(name, "ok", x.shape)
    # a whole Som::train run (epoch schedule + chunked DataSet protocol) on a small map
    rng = np.random.default_rng(77)
    x = rng.standard_normal((150, 9)).astype(np.float32)
    r = po.Reference(7, 6, 9, po.STANDARD)
    r.random_initialize(5, 0.5)
    init = r.get_state()
    mse = r.train(x, 64, 5, 0.3, 0.2, 3.0, 0.15, po.EXPONENTIAL, True)
    st = r.get_state()
    np.savez_compressed(os.path.join(OUT, "ref_train_std_exp.npz"), x=x, init_mean=init["mean"], mse=mse, chunk=64, epochs=5, eta0=0.3, eta_decay=0.2,
                        sigma0=3.0, sigma_decay=0.15, **{f"final_{k}": v for k, v in st.items()})
    print("train ok", mse)


if __name__ == "__main__":
    main()
