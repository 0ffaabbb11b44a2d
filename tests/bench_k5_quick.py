"""Quick timing of K5 (SomIndex build on the device: histogram + scan + stable radix sort) at bench.py's size (not a test)."""
import importlib, sys
sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
import torch
n, W, H = 1 << 24, 128, 128
ctx = v.VsomContext(W, H, 4)
bmu = torch.randint(0, W * H, (n,), dtype=torch.int32, device="cuda")
counts = torch.empty(W * H, dtype=torch.int64, device="cuda")
offsets = torch.empty(W * H + 1, dtype=torch.int64, device="cuda")
rows = torch.empty(n, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ctx.stream)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        ctx.build_index_device(bmu, n, counts, offsets, rows)
        e1.record(st)
    ctx.synchronize()
    assert int(counts.sum()) == n and int(offsets[-1]) == n
    print(f"K5 index build, {n} rows, {W * H} nodes: {e0.elapsed_time(e1):.3f} ms", flush=True)
