"""VEGAS+ with adaptive stratification, grid checkpoint and sampling.

    python examples/vegasplus_and_grid.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import tempfile

import vegasflow_b200 as vf

if __name__ == "__main__":
    plus = vf.VegasFlowPlus(8, int(1e7), adaptive=True)
    plus.compile(vf.integrands.symgauss)
    print("VegasFlowPlus:", plus.run_integration(5))

    vegas = vf.VegasFlow(8, int(1e7))
    vegas.compile(vf.integrands.symgauss)
    vegas.run_integration(5)
    path = tempfile.mktemp(suffix=".json")
    vegas.save_grid(path)  # same JSON schema as the reference (vflow.py:271-292)
    fresh = vf.VegasFlow(8, int(1e7), train=False)
    fresh.compile(vf.integrands.symgauss)
    fresh.load_grid(file_name=path)
    print("frozen grid loaded from", path, "->", fresh.run_integration(3))
    x, px = fresh.generate_random_array(10)
    print("samples", tuple(x.shape), "p(x)", px[:3].tolist())
