"""symgauss with the fused kernel (reference: examples/simgauss_tf.py).

    python examples/simgauss.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import time

import vegasflow_b200 as vf

dim, ncalls, n_iter = 4, int(1e6), 5

if __name__ == "__main__":
    print(f"VEGAS MC, ncalls={ncalls}:")
    start = time.time()
    res, err = vf.vegas_wrapper(vf.integrands.symgauss, dim, n_iter, ncalls)
    print(f"result {res:.6f} +/- {err:.6f}; took {time.time() - start:.3f} s")
