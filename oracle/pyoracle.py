"""TEST INFRASTRUCTURE — ctypes front-ends for the two CPU checkers.

  * ``Oracle``    -> oracle/liboracle.so        (plain-C restatement, oracle/vsom_oracle.c)
  * ``Reference`` -> oracle/_ref/libvsom_ref.so (the reference's own translation units compiled
                                                 unmodified by oracle/Makefile + oracle/ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker.  The product (libvsom_b200.so and the package that binds it)
never imports it.  Both classes expose the same method names so a test can drive either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvsom_ref.so")
REF_SSE_SO = os.path.join(HERE, "_ref", "libvsom_ref_sse.so")  # same TUs, stand-in Eigen reducing in Eigen's SSE2 packet order

STANDARD, MEDIAN, CLR = 0, 1, 2
EXPONENTIAL, INVERSE = 0, 1
ORDER_SEQUENTIAL, ORDER_EIGEN_SSE = 0, 2  # summation order of dot() / squaredNorm(); values match vsom_reduction_order

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(quiet: bool = True) -> None:
    """Compile the checkers (make -C oracle).  Builds _ref only where /root/reference exists."""
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _p(a, t):
    if a is None:
        return C.cast(None, t)
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def model_length(d_in: int, transform: int) -> int:
    """Transformation::Length — src/Transformation.cpp:31-35, :69-73, :162-165."""
    return d_in * (d_in - 1) if transform == CLR else d_in


class _Base:
    """Shared numpy-facing surface; subclasses bind the symbols."""

    W: int
    H: int
    N: int
    Din: int
    Dm: int

    # -- state ---------------------------------------------------------------------------
    def get_state(self):
        mean = np.empty((self.N, self.Dm), np.float32)
        S = np.empty_like(mean)
        sigma = np.empty_like(mean)
        weight = np.empty(self.N, np.float32)
        hits = np.empty(self.N, np.uint64)
        self._get_state(self._h, _p(mean, _f32p), _p(S, _f32p), _p(sigma, _f32p), _p(weight, _f32p), _p(hits, _u64p))
        return dict(mean=mean, S=S, sigma=sigma, weight=weight, hits=hits)

    def set_state(self, mean=None, S=None, sigma=None, weight=None, hits=None):
        mean = None if mean is None else _f32(mean)
        S = None if S is None else _f32(S)
        sigma = None if sigma is None else _f32(sigma)
        weight = None if weight is None else _f32(weight)
        hits = None if hits is None else np.ascontiguousarray(hits, dtype=np.uint64)
        self._set_state(self._h, _p(mean, _f32p), _p(S, _f32p), _p(sigma, _f32p), _p(weight, _f32p), _p(hits, _u64p))

    def random_initialize(self, seed: int, sigma: float):
        self._random_initialize(self._h, int(seed), C.c_float(sigma))

    # -- training ------------------------------------------------------------------------
    def train_rows(self, x, eta, sigma, decay, last_bmu=None):
        x = _f32(x).reshape(-1, self.Din)
        n = x.shape[0]
        bmu = np.empty(n, np.uint32)
        dist = np.empty(n, np.float32)
        resid2 = np.empty(n, np.float32)
        last = np.zeros(n, np.uint64) if last_bmu is None else np.ascontiguousarray(last_bmu, dtype=np.uint64)
        self._train_rows(self._h, _p(x, _f32p), n, C.c_double(eta), C.c_double(sigma), int(decay), _p(last, _u64p),
                         _p(bmu, _u32p), _p(dist, _f32p), _p(resid2, _f32p))
        return bmu, dist, resid2, last

    def train(self, x, chunk_rows, epochs, eta0, eta_decay, sigma0, sigma_decay, decay, umatrix_after_epoch=False):
        x = _f32(x).reshape(-1, self.Din)
        mse = np.zeros(epochs, np.float32)
        self._train(self._h, _p(x, _f32p), x.shape[0], int(chunk_rows), int(epochs), C.c_double(eta0), C.c_double(eta_decay),
                    C.c_double(sigma0), C.c_double(sigma_decay), int(decay), int(bool(umatrix_after_epoch)), _p(mse, _f32p))
        return mse

    def train_batch(self, x, chunk_rows, epochs, sigma0, sigma_decay, umatrix_after_epoch=False):
        """Som::train(..., BatchMap) — the batch-map trainer (src/Som.cpp:716-879)."""
        x = _f32(x).reshape(-1, self.Din)
        mse = np.zeros(epochs, np.float32)
        self._train_batch(self._h, _p(x, _f32p), x.shape[0], int(chunk_rows), int(epochs), C.c_double(sigma0), C.c_double(sigma_decay),
                          int(bool(umatrix_after_epoch)), _p(mse, _f32p))
        return mse

    # -- scoring -------------------------------------------------------------------------
    def find_bmu(self, x):
        x = _f32(x).reshape(-1, self.Din)
        n = x.shape[0]
        bmu = np.empty(n, np.uint32)
        dist = np.empty(n, np.float32)
        self._find_bmu(self._h, _p(x, _f32p), n, _p(bmu, _u32p), _p(dist, _f32p))
        return bmu, dist

    def all_dists(self, v):
        v = _f32(v).reshape(self.Din)
        out = np.empty(self.N, np.float64)
        self._all_dists(self._h, _p(v, _f32p), _p(out, _f64p))
        return out

    def find_local_bmu(self, v, start):
        v = _f32(v).reshape(self.Din)
        return int(self._find_local_bmu(self._h, _p(v, _f32p), int(start)))

    def find_restricted_bmu(self, x, min_hits):
        x = _f32(x).reshape(-1, self.Din)
        out = np.empty(x.shape[0], np.uint32)
        self._find_restricted_bmu(self._h, _p(x, _f32p), x.shape[0], int(min_hits), _p(out, _u32p))
        return out

    def find_restricted_bmd(self, v, min_hits):
        v = _f32(v).reshape(self.Din)
        out = np.empty(self.N, np.float64)
        self._find_restricted_bmd(self._h, _p(v, _f32p), int(min_hits), _p(out, _f64p))
        return out

    def evaluate(self, x):
        x = _f32(x).reshape(-1, self.Din)
        return float(self._evaluate(self._h, _p(x, _f32p), x.shape[0]))

    def measure_similarity(self, x, num_sigmas, min_hits):
        x = _f32(x).reshape(-1, self.Din)
        return int(self._measure_similarity(self._h, _p(x, _f32p), x.shape[0], int(num_sigmas), int(min_hits)))

    def update_umatrix(self):
        out = np.empty(self.N, np.float64)
        self._update_umatrix(self._h, _p(out, _f64p))
        return out

    def dist_raw(self, pos, u):
        u = _f32(u).reshape(self.Dm)
        return float(self._dist_raw(self._h, int(pos), _p(u, _f32p)))


def _sig(fn, res, args):
    fn.restype = res
    fn.argtypes = args
    return fn


class Oracle(_Base):
    """oracle/liboracle.so — the plain-C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build()
            L = C.CDLL(ORACLE_SO)
            vp = C.c_void_p
            _sig(L.oracle_create, vp, [C.c_int] * 4)
            _sig(L.oracle_destroy, None, [vp])
            _sig(L.oracle_depth, C.c_int, [vp])
            _sig(L.oracle_set_order, None, [vp, C.c_int])
            _sig(L.oracle_random_initialize, None, [vp, C.c_int, C.c_float])
            _sig(L.oracle_get_state, None, [vp, _f32p, _f32p, _f32p, _f32p, _u64p])
            _sig(L.oracle_set_state, None, [vp, _f32p, _f32p, _f32p, _f32p, _u64p])
            _sig(L.oracle_neighbourhood_weight, C.c_double, [C.c_uint64] * 4 + [C.c_double])
            _sig(L.oracle_dist, C.c_double, [vp, C.c_size_t, _f32p])
            _sig(L.oracle_dist_f64, C.c_double, [vp, C.c_size_t, _f32p])
            _sig(L.oracle_all_dists, None, [vp, _f32p, _f64p])
            _sig(L.oracle_dist_raw, C.c_double, [vp, C.c_size_t, _f32p])
            _sig(L.oracle_find_local_bmu, C.c_uint32, [vp, _f32p, C.c_uint64])
            _sig(L.oracle_find_bmu, None, [vp, _f32p, C.c_size_t, _u32p, _f32p])
            _sig(L.oracle_find_restricted_bmu, None, [vp, _f32p, C.c_size_t, C.c_uint64, _u32p])
            _sig(L.oracle_find_restricted_bmd, None, [vp, _f32p, C.c_uint64, _f64p])
            _sig(L.oracle_train_rows, None, [vp, _f32p, C.c_size_t, C.c_double, C.c_double, C.c_int, _u64p, _u32p, _f32p, _f32p])
            _sig(L.oracle_train, None, [vp, _f32p, C.c_size_t, C.c_size_t, C.c_size_t] + [C.c_double] * 4 + [C.c_int, C.c_int, _f32p])
            _sig(L.oracle_train_batch, None, [vp, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_int, _f32p])
            _sig(L.oracle_batch_epoch, C.c_float, [vp, _f32p, C.c_size_t, C.c_double, C.c_int, _u64p])
            _sig(L.oracle_evaluate, C.c_double, [vp, _f32p, C.c_size_t])
            _sig(L.oracle_measure_similarity, C.c_int, [vp, _f32p, C.c_size_t, C.c_int, C.c_uint64])
            _sig(L.oracle_update_umatrix, None, [vp, _f64p])
            _sig(L.oracle_build_index, None, [_u32p, C.c_size_t, C.c_int, _u64p, _u64p, _u32p])
            cls._lib = L
        return cls._lib

    def __init__(self, W, H, d_in, transform=STANDARD, order=ORDER_SEQUENTIAL):
        L = self.lib()
        self.W, self.H, self.N, self.Din, self.transform = W, H, W * H, d_in, transform
        self._h = L.oracle_create(W, H, d_in, transform)
        L.oracle_set_order(self._h, int(order))
        self.order = order
        self.Dm = L.oracle_depth(self._h)
        self._get_state, self._set_state = L.oracle_get_state, L.oracle_set_state
        self._random_initialize = L.oracle_random_initialize
        self._train_rows, self._train = L.oracle_train_rows, L.oracle_train
        self._train_batch = L.oracle_train_batch
        self._find_bmu, self._all_dists = L.oracle_find_bmu, L.oracle_all_dists
        self._find_local_bmu = L.oracle_find_local_bmu
        self._find_restricted_bmu, self._find_restricted_bmd = L.oracle_find_restricted_bmu, L.oracle_find_restricted_bmd
        self._evaluate, self._measure_similarity = L.oracle_evaluate, L.oracle_measure_similarity
        self._update_umatrix, self._dist_raw = L.oracle_update_umatrix, L.oracle_dist_raw

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().oracle_destroy(self._h)
            self._h = None

    def batch_epoch(self, x, sigma, is_first, last_bmu=None):
        x = _f32(x).reshape(-1, self.Din)
        last = np.zeros(x.shape[0], np.uint64) if last_bmu is None else np.ascontiguousarray(last_bmu, dtype=np.uint64).copy()
        mse = self.lib().oracle_batch_epoch(self._h, _p(x, _f32p), x.shape[0], C.c_double(sigma), int(bool(is_first)), _p(last, _u64p))
        return float(np.float32(mse)), last

    def all_dists_f64(self, v):
        """Same f32 residuals, f64 accumulation — the near-tie classifier."""
        v = _f32(v).reshape(self.Din)
        L = self.lib()
        return np.array([L.oracle_dist_f64(self._h, p, _p(v, _f32p)) for p in range(self.N)])

    @classmethod
    def neighbourhood_weight(cls, cx, cy, bx, by, sigma):
        return float(cls.lib().oracle_neighbourhood_weight(cx, cy, bx, by, C.c_double(sigma)))

    @classmethod
    def build_index(cls, bmu, N):
        bmu = np.ascontiguousarray(bmu, dtype=np.uint32)
        counts = np.empty(N, np.uint64)
        offsets = np.empty(N + 1, np.uint64)
        rows = np.empty(bmu.shape[0], np.uint32)
        cls.lib().oracle_build_index(_p(bmu, _u32p), bmu.shape[0], int(N), _p(counts, _u64p), _p(offsets, _u64p), _p(rows, _u32p))
        return counts, offsets, rows


class Reference(_Base):
    """oracle/_ref/libvsom_ref.so — the reference's own code behind oracle/ref_shim.cpp."""

    _lib = None
    SO = REF_SO

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(cls.SO)
            vp = C.c_void_p
            _sig(L.ref_create, vp, [C.c_int] * 4)
            _sig(L.ref_destroy, None, [vp])
            _sig(L.ref_depth, C.c_int, [vp])
            _sig(L.ref_random_initialize, None, [vp, C.c_int, C.c_float])
            _sig(L.ref_get_state, None, [vp, _f32p, _f32p, _f32p, _f32p, _u64p])
            _sig(L.ref_set_state, None, [vp, _f32p, _f32p, _f32p, _f32p, _u64p])
            _sig(L.ref_train_rows, None, [vp, _f32p, C.c_size_t, C.c_double, C.c_double, C.c_int, _u64p, _u32p, _f32p, _f32p])
            _sig(L.ref_train, None, [vp, _f32p, C.c_size_t, C.c_size_t, C.c_size_t] + [C.c_double] * 4 + [C.c_int, C.c_int, _f32p])
            _sig(L.ref_find_bmu, None, [vp, _f32p, C.c_size_t, _u32p, _f32p])
            _sig(L.ref_all_dists, None, [vp, _f32p, _f64p])
            _sig(L.ref_find_local_bmu, C.c_uint32, [vp, _f32p, C.c_uint64])
            _sig(L.ref_find_restricted_bmu, None, [vp, _f32p, C.c_size_t, C.c_uint64, _u32p])
            _sig(L.ref_find_restricted_bmd, None, [vp, _f32p, C.c_uint64, _f64p])
            _sig(L.ref_evaluate, C.c_double, [vp, _f32p, C.c_size_t])
            _sig(L.ref_measure_similarity, C.c_int, [vp, _f32p, C.c_size_t, C.c_int, C.c_uint64])
            _sig(L.ref_update_umatrix, None, [vp, _f64p])
            _sig(L.ref_dist_raw, C.c_double, [vp, C.c_uint64, _f32p])
            _sig(L.ref_neighbourhood_weight, C.c_double, [C.c_uint64] * 4 + [C.c_double])
            _sig(L.ref_train_batch, None, [vp, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_int, _f32p])
            _sig(L.ref_eigen_kind, C.c_char_p, [])
            cls._lib = L
        return cls._lib

    def __init__(self, W, H, d_in, transform=STANDARD):
        L = self.lib()
        self.W, self.H, self.N, self.Din, self.transform = W, H, W * H, d_in, transform
        self._h = L.ref_create(W, H, d_in, transform)
        self.Dm = L.ref_depth(self._h)
        self._get_state, self._set_state = L.ref_get_state, L.ref_set_state
        self._random_initialize = L.ref_random_initialize
        self._train_rows, self._train = L.ref_train_rows, L.ref_train
        self._train_batch = L.ref_train_batch
        self._find_bmu, self._all_dists = L.ref_find_bmu, L.ref_all_dists
        self._find_local_bmu = L.ref_find_local_bmu
        self._find_restricted_bmu, self._find_restricted_bmd = L.ref_find_restricted_bmu, L.ref_find_restricted_bmd
        self._evaluate, self._measure_similarity = L.ref_evaluate, L.ref_measure_similarity
        self._update_umatrix, self._dist_raw = L.ref_update_umatrix, L.ref_dist_raw

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().ref_destroy(self._h)
            self._h = None

    @classmethod
    def neighbourhood_weight(cls, cx, cy, bx, by, sigma):
        return float(cls.lib().ref_neighbourhood_weight(cx, cy, bx, by, C.c_double(sigma)))

    @classmethod
    def eigen_kind(cls) -> str:
        return cls.lib().ref_eigen_kind().decode()


class ReferenceSse(Reference):
    """oracle/_ref/libvsom_ref_sse.so — the same translation units compiled with -DVSOM_COMPAT_EIGEN_SSE_REDUX: dot() /
    squaredNorm() reduce in the order real Eigen has in an -msse2 build (what users who installed libeigen3-dev run)."""

    _lib = None
    SO = REF_SSE_SO


def best_cpu_checker(order=ORDER_SEQUENTIAL):
    """Reference (in the requested summation order) when the compiled library is present, else the C port.
    Returns (factory(W, H, d_in, transform), kind)."""
    ref = ReferenceSse if order == ORDER_EIGEN_SSE else Reference
    if ref.available():
        return ref, "reference"
    return (lambda W, H, d, t=STANDARD: Oracle(W, H, d, t, order)), "port"
