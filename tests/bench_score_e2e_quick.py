"""Quick timing of the host-buffer scoring call (vsom_find_bmu on pinned rows: H2D / search / D2H pipelined over slabs) on the
trained 128x128x256 map, for several slab sizes (not a test)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
rng = np.random.default_rng(0)
W, H, D = 128, 128, 256
cent = (rng.standard_normal((64, D)) * 3).astype(np.float32)
data = lambda k: (rng.standard_normal((k, D), dtype=np.float32) + cent[rng.integers(0, 64, k)])
q = torch.empty((n, D), dtype=torch.float32).pin_memory()
for i in range(0, n, 1 << 18):
    q[i:i + (1 << 18)] = torch.from_numpy(data(min(1 << 18, n - i)))
b = torch.empty(n, dtype=torch.int32).pin_memory(); d = torch.empty(n, dtype=torch.float32).pin_memory()
ctx = v.VsomContext(W, H, D, v.STANDARD, v.ORDER_EIGEN_SSE)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * H, D)) / 1000).astype(np.float32))
for sg, eta in ((32, .5), (16, .3), (8, .2), (4, .1), (2, .05)):
    ctx.train_chunk(data(4000), eta, float(sg), v.EXPONENTIAL)
for tier in ("", "2"):
    for lg in (18, 19, 20, 21):
        os.environ["VSOM_TC_HOST_SLAB_LOG2"] = str(lg)
        if tier:
            os.environ["VSOM_TC_TIER"] = tier
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.find_bmu_host_ptr(q.data_ptr(), n, b.data_ptr(), d.data_ptr())
            best = min(best, time.perf_counter() - t0)
        print(f"tier {'auto' if not tier else tier} slab 2^{lg}: {n / best / 1e6:7.2f} M rows/s ({best * 1e3:.1f} ms), ran tier {ctx.last_score_tc}", flush=True)
