"""Small driver for compute-sanitizer runs (not a test): touches every kernel family on small shapes."""
import importlib, sys, numpy as np
sys.path.insert(0, '.')
v = importlib.import_module('variational-self-organizing-maps_b200')
rng = np.random.default_rng(0)
ctx = v.VsomContext(32, 32, 64, v.STANDARD)
ctx.upload_state(mean=rng.standard_normal((1024, 64)).astype(np.float32))
x = rng.standard_normal((1500, 64)).astype(np.float32)
a = ctx.find_bmu(x); b = ctx.find_bmu_batch(x)
assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), "K2 != K3"
f = ctx.train_chunk(x[:200], 0.1, 4.0, 0)   # K1F (register-resident rows)
assert ctx.last_train_fast
c = ctx.train_chunk(x[:300], 0.1, 1.0, 0)   # generic K1, local-walk regime
u = ctx.update_umatrix(); idx = ctx.build_index(c[0])
ctx.close()
# K1F with rows in shared memory (more than 64 items per CTA) and the 4-byte sample path
k = v.VsomContext(150, 90, 7, v.MEDIAN)
k.upload_state(mean=rng.standard_normal((150 * 90, 7)).astype(np.float32))
k.train_chunk(rng.standard_normal((60, 7)).astype(np.float32), 0.1, 5.0, 1)
assert k.last_train_fast
k.close()
# generic K1 on an HBM-resident map: streamed scan + window-cell enumeration
g = v.VsomContext(112, 112, 784, v.STANDARD)
g.upload_state(mean=rng.standard_normal((112 * 112, 784)).astype(np.float32))
g.train_chunk(rng.standard_normal((6, 784)).astype(np.float32), 0.1, 4.0, 0)
assert not g.planes_resident and not g.last_train_fast
g.close()
# CLR on the generic kernel, batch-map epoch
r = v.VsomContext(12, 12, 8, v.CLR)
r.upload_state(mean=(rng.standard_normal((144, 56)) * 0.1).astype(np.float32))
r.train_chunk(rng.standard_normal((40, 8)).astype(np.float32), 0.001, 3.0, 0)
r.batch_epoch(rng.standard_normal((50, 8)).astype(np.float32), 2.0, True)
r.close()
print("sanitizer script ok", b[2])
