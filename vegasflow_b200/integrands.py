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
import math

import torch


class BuiltinIntegrand:
    """A named integrand compiled into libvegasflow_b200.so."""

    def __init__(self, name, fixed_dim=None, torch_impl=None):
        self.__name__ = name
        self.name = name
        self.fixed_dim = fixed_dim
        self._torch_impl = torch_impl

    def integrand_id(self):
        from vegasflow_b200 import _lib

        iid = _lib.load().vf_integrand_id(self.name.encode())
        _lib.check(min(iid, 0))
        return iid

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
