"""Generates tests/golden/api_driver_big.json: SHA-256 of the reference drivers' outputs (oracle/_ref/api_driver_ref and
api_driver_ref_sse = the reference's own translation units, sequential / Eigen-SSE2 dot order) on the large LIGHT case of
tests/test_gpu_host_api.py.  Needs /root/reference (oracle/Makefile builds the drivers); several CPU-minutes per driver."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_gpu_host_api as t  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    case = os.path.join(d, "case.bin")
    t.write_case(case, t.BIG_CASE, t.make_big_x(), t.BIG_TRAIN_ROWS)
    out = {"case_sha256": hashlib.sha256(open(case, "rb").read()).hexdigest()}
    procs = {}
    for order, drv in (("reference", t.REF_DRIVER), ("eigen_sse", t.REF_DRIVER_SSE)):
        procs[order] = subprocess.Popen([drv, case, os.path.join(d, order + ".bin")])
    for order, p in procs.items():
        assert p.wait() == 0
        b = open(os.path.join(d, order + ".bin"), "rb").read()
        out[order] = hashlib.sha256(b).hexdigest()
        out["bytes"] = len(b)
    json.dump(out, open(os.path.join(HERE, "api_driver_big.json"), "w"), indent=1)
    print(out)
