"""Quick timing of the online step on an HBM-resident map (config-5 shape on one GPU; not a test, not the bench)."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 16.0
D = 784
rng = np.random.default_rng(0)
order = {'seq': v.ORDER_REFERENCE, 'eigen': v.ORDER_EIGEN_SSE}[sys.argv[4] if len(sys.argv) > 4 else 'eigen']
ctx = v.VsomContext(W, W, D, v.STANDARD, order)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * W, D)) / 1000).astype(np.float32))
x = rng.standard_normal((n, D)).astype(np.float32)
xd = torch.from_numpy(x).cuda(); ob = torch.empty(n, dtype=torch.int32, device="cuda"); od = torch.empty(n, dtype=torch.float32, device="cuda")
ctx.train_chunk_device(xd, n, 0.1, sigma, v.EXPONENTIAL, ob, od)
ctx.synchronize()
t0 = time.perf_counter()
ctx.train_chunk_device(xd, n, 0.1, sigma, v.EXPONENTIAL, ob, od)
ctx.synchronize()
dt = time.perf_counter() - t0
ctx.debug_profile(True)
ctx.train_chunk_device(xd, n, 0.1, sigma, v.EXPONENTIAL, ob, od)
ctx.synchronize()
ph = ctx.debug_phase_cycles_raw()
bytes_per_sample = 4 * W * W * D
print(f"{W}x{W}x{D} sigma={sigma}: {dt / n * 1e6:8.1f} us/sample  scan-only HBM bound {bytes_per_sample / 6388e9 * 1e6:6.1f} us  resident={ctx.planes_resident} fast={ctx.last_train_fast}",
      {k: round(val) for k, val in ph.items()})
