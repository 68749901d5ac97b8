"""
ctypes binding of libvegasflow_b200.so (the C ABI in include/vegasflow_b200.h).

The product path has NO CPU fallback: if the shared library is missing or no
CUDA device is visible, every compute entry raises RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libvegasflow_b200.so")

MODE_PLAIN = 0
MODE_VEGAS = 1
MODE_RNG32 = 0x100  # OR'ed into a mode word: 4 x 32-bit uniforms per Philox block

_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "vf_version": (C.c_int, []),
    "vf_last_error": (C.c_char_p, []),
    "vf_integrand_id": (C.c_int, [C.c_char_p]),
    "vf_register_user_integrand": (C.c_int, [C.c_char_p]),
    "vf_supported": (C.c_int, [C.c_int, C.c_int]),
    "vf_flops_per_event": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vf_workspace_bytes": (C.c_size_t, [C.c_int]),
    "vf_run_event": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_double,
                               C.c_uint64, C.c_uint32, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P,
                               C.c_size_t, _P]),
    "vf_run_iterations": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_uint32,
                                    C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vf_exchange_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int64]),
    "vf_run_iterations_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64,
                                            C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                            _P, _P, _P, _P, _P, _P, _P, C.c_size_t, C.c_int,
                                            C.c_int, C.POINTER(C.c_uint64), C.c_uint64, _P]),
    "vf_refine_grid": (C.c_int, [C.c_int, _P, _P, _P]),
    "vf_iteration_epilogue": (C.c_int, [C.c_int, C.c_int64, C.c_int, _P, _P, _P, _P, _P]),
    "vf_digest_from_uniforms": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, C.c_double,
                                          _P, _P, _P, _P, _P, _P, _P]),
    "vf_uniforms": (C.c_int, [C.c_int, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, _P,
                              _P]),
    "vf_sample": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_double, C.c_uint64,
                            C.c_uint32, _P, _P, _P, _P, _P, _P, _P]),
    "vf_accumulate": (C.c_int, [C.c_int, C.c_int64, _P, _P, _P, C.c_int, _P, _P, C.c_int, _P,
                                C.c_size_t, _P]),
    "vfp_run_event": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P,
                                C.c_double, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _P, _P, _P,
                                _P, _P,
                                _P, C.c_int, _P, C.c_size_t, _P, _P, _P, _P, _P, _P]),
    "vfp_iteration_epilogue": (C.c_int, [C.c_int64, _P, _P, C.c_int, C.c_int, C.c_int64, _P, _P,
                                         _P, _P, _P, _P]),
    "vfp_run_iterations": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_uint32,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P,
                                     _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, C.c_int,
                                     C.c_int, C.POINTER(C.c_uint64), C.c_uint64, _P]),
    "vf_fp64_peak_probe": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "vf_kernel_timing": (C.c_int, [C.c_int]),
    "vf_kernel_time_ms": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "vf_epilogue_time_ms": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "vf_sm_count": (C.c_int, []),
    "vf_launch_count": (C.c_int64, [C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class VegasFlowB200Error(RuntimeError):
    """Raised when a C-ABI call returns a non-zero code."""


def load():
    """Load the shared library (once).  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -m vegasflow_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().vf_last_error().decode()


def check(rc):
    """Map ABI return codes onto the reference's exception types."""
    if rc == 0:
        return
    msg = last_error()
    if rc in (-1, -2):
        raise ValueError(msg)
    raise VegasFlowB200Error(f"[{rc}] {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "vegasflow_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return load()


def ptr(t):
    """Raw pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def host_doubles(seq):
    """Host double array (or None) kept alive by the caller."""
    if seq is None:
        return None
    arr = (C.c_double * len(seq))(*[float(v) for v in seq])
    return arr


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
