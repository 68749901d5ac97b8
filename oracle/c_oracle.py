"""
ctypes loader for oracle/libvegas_oracle.so (test infrastructure, NOT product code).
See oracle/vegas_oracle.c for the restated reference lines.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvegas_oracle.so")

MODE_PLAIN, MODE_VEGAS = 0, 1
INTEGRAND_IDS = {"symgauss": 0, "product": 1}


def _stale():
    src = os.path.join(_HERE, "vegas_oracle.c")
    return not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)


def build(force=False):
    if force or _stale():
        import fcntl

        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:  # one builder at a time
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or _stale():
                    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libvegas_oracle.so"])
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().vfo_philox4x32_10(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def set_rng_bits(bits):
    """Select the engine stream the oracle reproduces: 52 (default) or 32 bits per uniform."""
    lib().vfo_set_rng_bits(C.c_int(int(bits)))


def uniforms(seed, iteration, ev_begin, n, n_dim):
    out = np.empty((n, n_dim), dtype=np.float64)
    lib().vfo_uniforms(C.c_uint64(seed), C.c_uint32(iteration), C.c_uint64(ev_begin), C.c_int64(n),
                       C.c_int(n_dim), _p(out, C.c_double))
    return out


def integrand(name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, d = x.shape
    f = np.empty(n, dtype=np.float64)
    lib().vfo_integrand(C.c_int(INTEGRAND_IDS[name]), C.c_int(d), C.c_int64(n), _p(x, C.c_double),
                        _p(f, C.c_double))
    return f


def digest_from_uniforms(mode, name, rnds, divisions, xjac, xmin=None, xdelta=None):
    rnds = np.ascontiguousarray(rnds, dtype=np.float64)
    n, d = rnds.shape
    divisions = np.ascontiguousarray(divisions, dtype=np.float64)
    xmin = None if xmin is None else np.ascontiguousarray(xmin, dtype=np.float64)
    xdelta = None if xdelta is None else np.ascontiguousarray(xdelta, dtype=np.float64)
    x = np.empty((n, d)); w = np.empty(n); ind = np.empty((n, d), dtype=np.int32); wf = np.empty(n)
    lib().vfo_digest_from_uniforms(
        C.c_int(mode), C.c_int(INTEGRAND_IDS[name]), C.c_int(d), C.c_int64(n), _p(rnds, C.c_double),
        _p(divisions, C.c_double), C.c_double(xjac), _p(xmin, C.c_double), _p(xdelta, C.c_double),
        _p(x, C.c_double), _p(w, C.c_double), _p(ind, C.c_int32), _p(wf, C.c_double))
    return x, w, ind, wf


def run_event(mode, name, n_dim, ev_begin, n_events, xjac, seed, iteration, train, divisions,
              xmin=None, xdelta=None, nthreads=0):
    divisions = np.ascontiguousarray(divisions, dtype=np.float64)
    xmin = None if xmin is None else np.ascontiguousarray(xmin, dtype=np.float64)
    xdelta = None if xdelta is None else np.ascontiguousarray(xdelta, dtype=np.float64)
    sums = np.zeros(2)
    hist = np.zeros((n_dim, 50))
    lib().vfo_run_event(
        C.c_int(mode), C.c_int(INTEGRAND_IDS[name]), C.c_int(n_dim), C.c_uint64(ev_begin),
        C.c_int64(n_events), C.c_double(xjac), C.c_uint64(seed), C.c_uint32(iteration),
        C.c_int(int(train)), _p(divisions, C.c_double), _p(xmin, C.c_double),
        _p(xdelta, C.c_double), _p(sums, C.c_double), _p(hist, C.c_double), C.c_int(nthreads))
    return sums[0], sums[1], hist


def refine_grid(hist, divisions):
    hist = np.ascontiguousarray(hist, dtype=np.float64)
    out = np.array(divisions, dtype=np.float64, order="C", copy=True)
    lib().vfo_refine_grid(C.c_int(out.shape[0]), _p(hist, C.c_double), _p(out, C.c_double))
    return out


def plus_run_event(name, n_dim, n_strat, n_ev, xjac, seed, iteration, train, divisions, xmin=None,
                   xdelta=None, rnds=None, detail=False):
    n_ev = np.ascontiguousarray(n_ev, dtype=np.int32)
    n_cubes = n_ev.shape[0]
    n = int(n_ev.sum())
    divisions = np.ascontiguousarray(divisions, dtype=np.float64)
    xmin = None if xmin is None else np.ascontiguousarray(xmin, dtype=np.float64)
    xdelta = None if xdelta is None else np.ascontiguousarray(xdelta, dtype=np.float64)
    rnds = None if rnds is None else np.ascontiguousarray(rnds, dtype=np.float64)
    ress = np.zeros(n_cubes); var = np.zeros(n_cubes); hist = np.zeros((n_dim, 50))
    x = w = ind = wf = None
    if detail:
        x = np.empty((n, n_dim)); w = np.empty(n); ind = np.empty((n, n_dim), dtype=np.int32)
        wf = np.empty(n)
    lib().vfo_plus_run_event(
        C.c_int(INTEGRAND_IDS[name]), C.c_int(n_dim), C.c_int(n_strat), C.c_int64(n_cubes),
        _p(n_ev, C.c_int32), C.c_double(xjac), C.c_uint64(seed), C.c_uint32(iteration),
        C.c_int(int(train)), _p(divisions, C.c_double), _p(xmin, C.c_double),
        _p(xdelta, C.c_double), _p(rnds, C.c_double), _p(ress, C.c_double), _p(var, C.c_double),
        _p(hist, C.c_double), _p(x, C.c_double), _p(w, C.c_double), _p(ind, C.c_int32),
        _p(wf, C.c_double))
    return ress, var, hist, dict(x=x, w=w, ind=ind, wf=wf)
