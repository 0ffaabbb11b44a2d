"""Quick A/B timing of the online step at config 2 (not a test, not the bench): samples/s and phase cycles of a 200k chunk."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
order = {"seq": v.ORDER_REFERENCE, "eigen": v.ORDER_EIGEN_SSE}[sys.argv[2] if len(sys.argv) > 2 else "eigen"]
rng = np.random.default_rng(0)
W, H, D = (int(a) for a in sys.argv[3:6]) if len(sys.argv) > 5 else (64, 64, 128)
tr = int(sys.argv[6]) if len(sys.argv) > 6 else v.MEDIAN
ETA, SIG = (0.001, W / 2.0) if tr == 2 else (0.05, W / 2.0)
ctx = v.VsomContext(W, H, D, tr, order)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * H, ctx.Dm)) / 1000).astype(np.float32))
x = (rng.standard_normal((n, D)) + 3 * rng.standard_normal((64, D))[rng.integers(0, 64, n)]).astype(np.float32)
xd = torch.from_numpy(x).cuda(); ob = torch.empty(n, dtype=torch.int32, device="cuda"); od = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    ctx.train_chunk_device(xd, n, ETA, SIG, v.EXPONENTIAL, ob, od)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ctx.train_chunk_device(xd, n, ETA, SIG, v.EXPONENTIAL, ob, od)
ctx.synchronize()
dt = time.perf_counter() - t0
ctx.debug_profile(True)
ctx.train_chunk_device(xd, n, ETA, SIG, v.EXPONENTIAL, ob, od)
ctx.synchronize()
ph = ctx.debug_phase_cycles_raw()
print(f"{W}x{H}x{D} tr={tr} order={order}: {3 * n / dt:10.0f} samples/s  fast={ctx.last_train_fast} die={ctx.die_aware} ", {k: round(val) for k, val in ph.items()}, "peaks", ctx.measure_peaks() if len(sys.argv) > 7 else "")
