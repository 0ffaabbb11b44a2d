"""Quick timing of K6 (one chunk-epoch of the batch-map trainer) on bench.py's shape (not a test).  Prints the wall time of the
host-buffer call (H2D inside) for the first (global search) and a later (local walks) epoch."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
bw, bh, bd, brows = 20, 20, 784, 60000
rng = np.random.default_rng(5)
ctx = v.VsomContext(bw, bh, bd, v.STANDARD, v.ORDER_EIGEN_SSE)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (bw * bh, bd)) / 1000).astype(np.float32))
bx = np.floor(256 * rng.random((brows, bd), dtype=np.float32) ** 2).astype(np.float32)
import torch
keep = torch.from_numpy(bx).pin_memory()
bx = keep.numpy()
ctx.batch_epoch(bx[:4096], 5.0, True)
for first in (True, False, False):
    t0 = time.perf_counter()
    mse, _ = ctx.batch_epoch(bx, 5.0, first)
    dt = time.perf_counter() - t0
    print(f"K6 chunk-epoch ({'global search' if first else 'local walks'}): {dt * 1e3:.2f} ms, mse {mse}", flush=True)
# phase A of the first epoch on its own: dispatch (tensor cores + probes) vs the exact scan
import torch
xd = torch.from_numpy(bx).cuda()
ob = torch.empty(brows, dtype=torch.int32, device="cuda"); od = torch.empty(brows, dtype=torch.float32, device="cuda")
for name, fn in (("dispatch", lambda: ctx.find_bmu_batch_device(xd, brows, ob, od)), ("exact scan", lambda: ctx.find_bmu_exact_device(xd, brows, ob, od))):
    for _ in range(2):
        ctx.synchronize(); t0 = time.perf_counter(); fb = fn(); ctx.synchronize(); dt = time.perf_counter() - t0
    print(f"phase A {name}: {dt * 1e3:.2f} ms, tier {ctx.last_score_tc}, fallback {fb}", flush=True)
