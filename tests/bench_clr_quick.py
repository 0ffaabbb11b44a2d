"""Quick timing of the online step at config 3 (50x50, CLR J=32) — not a test, not the bench."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
rng = np.random.default_rng(0)
W, H, J = 50, 50, 32
ctx = v.VsomContext(W, H, J, v.CLR)
dm = v.model_length(J, v.CLR)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * H, dm)) / 1000).astype(np.float32))
z = rng.standard_normal((n, 1)).astype(np.float32)
x = (rng.uniform(0.5, 1.5, (1, J)).astype(np.float32) * z + 0.1 * rng.standard_normal((n, J))).astype(np.float32)
xd = torch.from_numpy(x).cuda(); ob = torch.empty(n, dtype=torch.int32, device="cuda"); od = torch.empty(n, dtype=torch.float32, device="cuda")
ctx.train_chunk_device(xd, n, 0.001, 25.0, v.EXPONENTIAL, ob, od)
ctx.synchronize()
t0 = time.perf_counter()
ctx.train_chunk_device(xd, n, 0.001, 25.0, v.EXPONENTIAL, ob, od)
ctx.synchronize()
dt = time.perf_counter() - t0
ctx.debug_profile(True)
ctx.train_chunk_device(xd, n, 0.001, 25.0, v.EXPONENTIAL, ob, od)
ctx.synchronize()
print(f"{n / dt:10.0f} samples/s  fast={ctx.last_train_fast} die={ctx.die_aware} ", {k: round(val) for k, val in ctx.debug_phase_cycles_raw().items()})
