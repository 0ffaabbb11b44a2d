"""Quick timing of K2 batch scoring at the config-4 shape on maps of different character (not a test, not the bench):
random init; trained with a decaying sigma (what bench.py scores against); over-smoothed (few samples, wide sigma)."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
rng = np.random.default_rng(0)
W, H, D = 128, 128, 256
cent = (rng.standard_normal((64, D)) * 3).astype(np.float32)
data = lambda k: (rng.standard_normal((k, D), dtype=np.float32) + cent[rng.integers(0, 64, k)])
q = torch.empty((n, D), dtype=torch.float32, device="cuda")
blk = 1 << 18
for i in range(0, n, blk):
    q[i:i + blk] = torch.from_numpy(data(min(blk, n - i)))
ob = torch.empty(n, dtype=torch.int32, device="cuda"); od = torch.empty(n, dtype=torch.float32, device="cuda")
init = (rng.integers(-1000, 1000, (W * H, D)) / 1000).astype(np.float32)
for name, schedule in (("random init", ()), ("trained, sigma 32 -> 2", ((32, .5), (16, .3), (8, .2), (4, .1), (2, .05))), ("over-smoothed", ((24, .3), (8, .15)))):
    ctx = v.VsomContext(W, H, D, v.STANDARD, v.ORDER_EIGEN_SSE)
    ctx.upload_state(mean=init)
    t0 = time.perf_counter()
    for sg, eta in schedule:
        ctx.train_chunk(data(4000), eta, float(sg), v.EXPONENTIAL)
    ttrain = time.perf_counter() - t0
    ctx.find_bmu_batch_device(q, min(n, 1 << 20), ob, od)
    ctx.synchronize()
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        fb = ctx.find_bmu_batch_device(q, n, ob, od)
        ctx.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{name:24s} {n / best / 1e6:8.2f} M rows/s  {best * 1e3:8.2f} ms  {2 * W * H * D * n / best / 1e12 / 1401.7 * 100:5.1f}% of sustained peak, "
          f"fallback {fb} ({100.0 * fb / n:.3f}%), tier {ctx.last_score_tc}, causes {ctx.tc_stats()}, distinct BMUs {len(torch.unique(ob))}, training {ttrain:.2f}s", flush=True)
    ctx.close()
