import importlib, sys, numpy as np
sys.path.insert(0, '.')
v = importlib.import_module('variational-self-organizing-maps_b200')
rng = np.random.default_rng(0)
ctx = v.VsomContext(32, 32, 64, v.STANDARD)
ctx.upload_state(mean=rng.standard_normal((1024, 64)).astype(np.float32))
x = rng.standard_normal((1500, 64)).astype(np.float32)
a = ctx.find_bmu(x); b = ctx.find_bmu_batch(x)
assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), "K2 != K3"
c = ctx.train_chunk(x[:300], 0.1, 1.0, 0)   # local-walk regime
u = ctx.update_umatrix(); idx = ctx.build_index(c[0])
print("sanitizer script ok", b[2])
