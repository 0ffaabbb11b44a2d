"""Quick timing of the online step in the findLocalBmu regime (sigma <= 1) at config 2's shape — not a test, not the bench."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(0)
W, H, D = 64, 64, 128
ctx = v.VsomContext(W, H, D, v.MEDIAN)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * H, D)) / 1000).astype(np.float32))
x = (rng.standard_normal((n, D)) + 3 * rng.standard_normal((64, D))[rng.integers(0, 64, n)]).astype(np.float32)
xd = torch.from_numpy(x).cuda(); ob = torch.empty(n, dtype=torch.int32, device="cuda"); od = torch.empty(n, dtype=torch.float32, device="cuda")
# a few global-regime samples first so that the map is organised like after the early epochs
ctx.train_chunk_device(xd, min(n, 50000), 0.05, 8.0, v.EXPONENTIAL, ob, od)
for sigma in (1.0, 1.5):
    ctx.train_chunk_device(xd, n, 0.05, sigma, v.EXPONENTIAL, ob, od)
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.train_chunk_device(xd, n, 0.05, sigma, v.EXPONENTIAL, ob, od)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    ctx.debug_profile(True)
    ctx.train_chunk_device(xd, n, 0.05, sigma, v.EXPONENTIAL, ob, od)
    ctx.synchronize()
    print(f"sigma={sigma}: {n / dt:10.0f} samples/s  fast={ctx.last_train_fast}", {k: round(val) for k, val in ctx.debug_phase_cycles_raw().items()})
    ctx.debug_profile(False)
