"""LO matrix elements of the reference's examples, evaluated inline in the fused kernel
(examples/drellyan_lo_tf.py, examples/singletop_lo_tf.py).

    python examples/matrix_elements.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import vegasflow_b200 as vf

if __name__ == "__main__":
    print("single-top LO (t-channel), 3 dimensions -- expected 423.9 +- 0.2 pb:")
    vf.vegas_wrapper(vf.integrands.singletop_lo, 3, 5, int(1e7))
    print("Drell-Yan LO, 4 dimensions (ln^2-divergent at kappa -> 0: the value depends on the cut):")
    vf.vegas_wrapper(vf.integrands.drellyan_lo, 4, 5, int(1e7))
