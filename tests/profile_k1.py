"""Small driver for ncu captures of the online step (not a test): trains config-2-shaped chunks so that the persistent
kernel is launched a few times.  Usage: python tests/profile_k1.py [samples_per_chunk] [launches]"""
import importlib
import sys

import numpy as np

sys.path.insert(0, ".")
v = importlib.import_module("variational-self-organizing-maps_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
W, H, D = (int(a) for a in sys.argv[3:6]) if len(sys.argv) > 5 else (64, 64, 128)
tr = int(sys.argv[6]) if len(sys.argv) > 6 else v.MEDIAN
ctx = v.VsomContext(W, H, D, tr, v.ORDER_EIGEN_SSE)
ctx.upload_state(mean=(rng.integers(-1000, 1000, (W * H, D)) / 1000).astype(np.float32))
x = (rng.standard_normal((n, D)) + 3 * rng.standard_normal((64, D))[rng.integers(0, 64, n)]).astype(np.float32)
for _ in range(launches):
    bmu, d, r2, last = ctx.train_chunk(x, 0.05, W / 2.0, v.EXPONENTIAL)
print("profile_k1 ok: fast kernel" if ctx.last_train_fast else "profile_k1 ok: generic kernel", "die-aware rows" if ctx.die_aware else "rows in pool order", float(d.mean()))
