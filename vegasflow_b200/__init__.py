"""
vegasflow_b200 -- B200-native VEGAS Monte Carlo integration.

Public surface (same names as N3PDF/vegasflow v1.4.0 exports):
    integrators   PlainFlow, VegasFlow, VegasFlowPlus
    wrappers      plain_wrapper, vegas_wrapper, vegasflowplus_wrapper
    samplers      plain_sampler, vegas_sampler, vegasflowplus_sampler
    dtype helpers DTYPE, DTYPEINT, float_me, int_me, run_eager
plus `integrands` (built-in fused integrands and `cuda_integrand` for user CUDA source).
"""
from importlib import import_module as _import_module

__version__ = "0.1.0"

_PUBLIC = {
    "configflow": ("DTYPE", "DTYPEINT", "float_me", "int_me", "run_eager"),
    "plain": ("PlainFlow", "plain_sampler", "plain_wrapper"),
    "vflow": ("VegasFlow", "vegas_sampler", "vegas_wrapper"),
    "vflowplus": ("VegasFlowPlus", "vegasflowplus_sampler", "vegasflowplus_wrapper"),
}
__all__ = ["integrands", "__version__"]
for _module, _names in _PUBLIC.items():
    _loaded = _import_module(f"{__name__}.{_module}")
    for _name in _names:
        globals()[_name] = getattr(_loaded, _name)
        __all__.append(_name)
integrands = _import_module(f"{__name__}.integrands")
del _import_module, _module, _names, _loaded, _name
