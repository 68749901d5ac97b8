"""README minimal example of the reference (README.md:55-74): x1 * ... * x8, three ways.

    python examples/basic_example.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import torch

import vegasflow_b200 as vf

dimensions, iterations, events = 8, 5, int(1e6)


def integrand(x, **kwargs):
    """Any callable on CUDA torch tensors: x[events, n] -> [events] (unfused path)."""
    return torch.prod(x, dim=1)


USER_CUDA = """
__device__ double integrand(const double* x, int n_dim) {
    double p = x[0];
    for (int i = 1; i < n_dim; ++i) p *= x[i];
    return p;
}
"""

if __name__ == "__main__":
    print("python callable (sample -> torch -> accumulate):")
    vf.vegas_wrapper(integrand, dimensions, iterations, events)
    print("built-in integrand (one fused kernel per iteration):")
    vf.vegas_wrapper(vf.integrands.product, dimensions, iterations, events)
    print("user CUDA source compiled into the fused kernel:")
    mine = vf.integrands.cuda_integrand(USER_CUDA, dimensions, name="product_user")
    vf.vegas_wrapper(mine, dimensions, iterations, events)
