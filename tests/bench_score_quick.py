"""Quick timing of K2 batch scoring at the config-4 shape (not a test, not the bench)."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n = 1 << 22
rng = np.random.default_rng(0)
W, H, D = 128, 128, 256
ctx = v.VsomContext(W, H, D, v.STANDARD)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * H, D)) / 1000).astype(np.float32))
q = torch.empty((n, D), dtype=torch.float32, device="cuda")
blk = 1 << 18
cent = (rng.standard_normal((64, D)) * 3).astype(np.float32)
for i in range(0, n, blk):
    q[i:i + blk] = torch.from_numpy((rng.standard_normal((blk, D), dtype=np.float32) + cent[rng.integers(0, 64, blk)]))
ob = torch.empty(n, dtype=torch.int32, device="cuda"); od = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    ctx.find_bmu_batch_device(q, n, ob, od)
ctx.synchronize()
best = 1e9
for _ in range(4):
    t0 = time.perf_counter()
    fb = ctx.find_bmu_batch_device(q, n, ob, od)
    ctx.synchronize()
    best = min(best, time.perf_counter() - t0)
print(f"{n / best / 1e6:8.2f} M rows/s  {best * 1e3:7.2f} ms  {2 * W * H * D * n / best / 1e12 / 1401.7 * 100:5.1f}% of sustained bf16 peak, fallback rows {fb}")
