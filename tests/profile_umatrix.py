"""Small driver for ncu captures of the U-matrix kernel (not a test)."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D = 784
rng = np.random.default_rng(0)
order = v.ORDER_EIGEN_SSE if (len(sys.argv) > 2 and sys.argv[2] == "eigen") else v.ORDER_REFERENCE
ctx = v.VsomContext(W, W, D, v.STANDARD, order)
m = (rng.integers(-1000, 1000, (W * W, D)) / 1000).astype(np.float32)
ctx.upload_state(mean=m, sigma=np.abs(m) + 0.5)
import torch
for _ in range(3):
    t0 = time.perf_counter()
    u = ctx.update_umatrix()
    dt = time.perf_counter() - t0
st = torch.cuda.ExternalStream(ctx.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record(st)
    v.lib().vsom_update_umatrix(ctx._h, None)
    e1.record(st)
ctx.synchronize()
print(f"umatrix {W}x{W}x{D} order {order}: kernel {e0.elapsed_time(e1):.3f} ms ({(8 * W * W * D + 8 * W * W) / e0.elapsed_time(e1) / 1e6:.0f} GB/s algorithmic); {dt * 1e3:.2f} ms incl. download, mean {float(u.mean()):.4f}")
