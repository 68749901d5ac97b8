"""
Handles of the integrands that are evaluated inline by the fused CUDA kernel.

`compile()` accepts one of these objects (or its name) and then runs the fused
path (one kernel launch per iteration).  Any other Python callable goes through
the unfused sample -> callable -> accumulate path.

Reference bodies (file:line relative to /root/reference):
  symgauss      examples/simgauss_tf.py:22-32
  product       README.md:63-68
  drellyan_lo   examples/drellyan_lo_tf.py:27-249   (n_dim = 4)
  singletop_lo  examples/singletop_lo_tf.py:45-270  (n_dim = 3)
"""
import fcntl
import hashlib
import math
import os
import subprocess

import torch

# largest n_dim whose fused-kernel shared memory (vf_event.cuh::choose_smem: 4 table copies and
# 16 histogram copies, n_dim * 9600 B, at the top end) fits the 227 KB opt-in limit
MAX_FUSED_DIM = 24


class BuiltinIntegrand:
    """A named integrand compiled into libvegasflow_b200.so."""

    def __init__(self, name, fixed_dim=None, torch_impl=None):
        self.__name__ = name
        self.name = name
        self.fixed_dim = fixed_dim
        self._torch_impl = torch_impl
        self._id = None

    def integrand_id(self):
        if self._id is None:
            from vegasflow_b200 import _lib

            iid = _lib.load().vf_integrand_id(self.name.encode())
            _lib.check(min(iid, 0))
            self._id = iid
        return self._id

    def supported(self, n_dim):
        from vegasflow_b200 import _lib

        return bool(_lib.load().vf_supported(self.integrand_id(), int(n_dim)))

    def __call__(self, xarr, **kwargs):
        """Tensor form used by the unfused path (CUDA tensors in, CUDA tensors out)."""
        if self._torch_impl is None:
            raise NotImplementedError(f"{self.name} is only available on the fused path")
        return self._torch_impl(xarr)

    def __repr__(self):
        return f"<builtin integrand {self.name}>"


class CudaIntegrand(BuiltinIntegrand):
    """A user integrand written as a CUDA device function and compiled into the fused kernels.

    The native counterpart of the reference's "integrand in C / CUDA" examples
    (examples/simgauss_cffi.py:25-62, examples/cuda/integrand.cpp:41-89): instead of a C function
    called through cffi on host arrays, or a TensorFlow custom op, the source is instantiated
    INSIDE the event kernel, so it runs inline like the built-in integrands (no per-event data in
    HBM).  Use `cuda_integrand(source, n_dim)` to build one.
    """

    def __init__(self, name, n_dim, module_path):
        super().__init__(name, fixed_dim=n_dim)
        self.module_path = module_path

    def integrand_id(self):
        if self._id is None:
            from vegasflow_b200 import _lib

            iid = _lib.load().vf_register_user_integrand(self.module_path.encode())
            _lib.check(min(iid, 0))
            self._id = iid
        return self._id


def cuda_integrand(source, n_dim, name="user_integrand", heavy=False, verbose=False):
    """Compile `source` (CUDA C++ defining
    ``__device__ double integrand(const double* x, int n_dim)``) into a module that
    instantiates the fused kernels for it, and return a handle for `compile()`.

    `heavy=True` gives the kernel the 128-register budget (one block per SM), for integrands
    the size of a matrix element.  nvcc cross-compiles, so this works without a GPU; modules
    are cached under vegasflow_b200/build/user/ by source hash.
    """
    from vegasflow_b200 import build as vf_build

    n_dim = int(n_dim)
    # the fused kernels keep n_dim*9600 B of tables + histograms in shared memory: 24 dimensions
    # fill the 227 KB a block can opt in to on sm_100
    if not 1 <= n_dim <= MAX_FUSED_DIM:
        raise ValueError(f"cuda_integrand supports 1 <= n_dim <= {MAX_FUSED_DIM}")
    vf_build.build()
    with open(os.path.join(vf_build.CSRC, "vf_user_integrand.cu.in")) as fh:
        template = fh.read()
    unit = (template.replace("@USER_SOURCE@", source).replace("@N_DIM@", str(n_dim))
            .replace("@HEAVY@", "true" if heavy else "false"))
    # the module is tied to this build of the library (struct layouts, ABI version)
    stamp = vf_build.source_hash()
    digest = hashlib.sha256((unit + stamp).encode()).hexdigest()[:20]
    out_dir = os.path.join(vf_build.OBJDIR, "user")
    os.makedirs(out_dir, exist_ok=True)
    cu_path = os.path.join(out_dir, f"{name}_{digest}.cu")
    so_path = os.path.join(out_dir, f"{name}_{digest}.so")
    if not os.path.exists(so_path):
        # one builder per module at a time (torchrun: every rank sees the empty cache at once)
        with open(os.path.join(out_dir, f".{name}_{digest}.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if not os.path.exists(so_path):  # another rank built it while we waited
                    _compile_user_module(vf_build, unit, cu_path, so_path, verbose)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return CudaIntegrand(name, n_dim, so_path)


def _compile_user_module(vf_build, unit, cu_path, so_path, verbose):
    tmp_cu = f"{cu_path}.{os.getpid()}.tmp.cu"
    tmp_so = f"{so_path}.{os.getpid()}.tmp"
    with open(tmp_cu, "w") as fh:
        fh.write(unit)
    os.replace(tmp_cu, cu_path)
    flags = [f for f in vf_build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    cmd = [vf_build._nvcc(), *flags, "-shared", "-I", vf_build.INCLUDE, "-I", vf_build.CSRC,
           cu_path, "-o", tmp_so, "-L", vf_build.LIBDIR, "-lvegasflow_b200",
           "-Xlinker", "-rpath", "-Xlinker", vf_build.LIBDIR]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp_so):
            os.remove(tmp_so)
        raise ValueError(f"nvcc failed on the user integrand:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    os.replace(tmp_so, so_path)  # atomic: a concurrent loader never sees a half-written module


def _symgauss_torch(xarr):
    n_dim = xarr.shape[-1]
    a = 0.1
    n100 = float(100 * n_dim)
    pref = math.pow(1.0 / a / math.sqrt(math.pi), n_dim)
    c = (n100 + 1) * n100 / 2.0
    s = torch.zeros(xarr.shape[0], dtype=xarr.dtype, device=xarr.device)
    for j in range(n_dim):
        t = (xarr[:, j] - 0.5) / a
        s = t * t if j == 0 else s + t * t
    coef = (c + s) - c
    return pref * torch.exp(-coef)


def _product_torch(xarr):
    p = xarr[:, 0].clone()
    for j in range(1, xarr.shape[-1]):
        p = p * xarr[:, j]
    return p


symgauss = BuiltinIntegrand("symgauss", torch_impl=_symgauss_torch)
product = BuiltinIntegrand("product", torch_impl=_product_torch)
drellyan_lo = BuiltinIntegrand("drellyan_lo", fixed_dim=4)
singletop_lo = BuiltinIntegrand("singletop_lo", fixed_dim=3)

BUILTINS = {b.name: b for b in (symgauss, product, drellyan_lo, singletop_lo)}


def resolve(integrand):
    """Return the BuiltinIntegrand for a handle or a name, else None."""
    if isinstance(integrand, BuiltinIntegrand):
        return integrand
    if isinstance(integrand, str):
        if integrand not in BUILTINS:
            raise ValueError(f"unknown built-in integrand '{integrand}'; have {sorted(BUILTINS)}")
        return BUILTINS[integrand]
    return None
