"""
Generates tests/golden/*.npz from the oracle (oracle/vegas_ref.py, numpy) --
the reference itself cannot be imported here (TensorFlow is absent), so these
are ORACLE-generated vectors (parity unpinned, see oracle/vegas_ref.py header).

    python tests/golden/make_golden.py

Inputs are seeded; outputs are what the CUDA path must reproduce
(bit-exact ind/x/w, <=1e-12 relative w*f).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vegas_ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 2048


def trained_grid(integrand, d, seed, iters=3, n=20000):
    rng = np.random.default_rng(seed)
    grid = R.initial_divisions(d)
    for _ in range(iters):
        r = R.TECH_CUT + rng.random((n, d)) * (1 - 2 * R.TECH_CUT)
        _, _, h, _ = R.vegas_run_event(r, grid, integrand, n)
        grid = R.refine_grid(h, grid)
    return grid


def main():
    out = {}
    cases = [("symgauss", 2), ("symgauss", 4), ("symgauss", 8), ("symgauss", 20), ("product", 1),
             ("product", 3), ("product", 8), ("drellyan_lo", 4), ("singletop_lo", 3)]
    for k, (name, d) in enumerate(cases):
        f = R.INTEGRANDS[name]
        grid = trained_grid(f, d, seed=100 + k)
        rng = np.random.default_rng(200 + k)
        r = R.TECH_CUT + rng.random((N, d)) * (1 - 2 * R.TECH_CUT)
        # edge rows: the extremes of the allowed interval
        r[0, :] = R.TECH_CUT
        r[1, :] = np.nextafter(1 - R.TECH_CUT, 0)
        res, res2, hist, det = R.vegas_run_event(r, grid, f, N)
        key = f"{name}_d{d}"
        out[key + "_rnds"] = r
        out[key + "_grid"] = grid
        out[key + "_x"] = det["x"]
        out[key + "_w"] = det["w"]
        out[key + "_ind"] = det["ind"].astype(np.int32)
        out[key + "_wf"] = det["wf"]
        out[key + "_hist"] = hist
        out[key + "_newgrid"] = R.refine_grid(hist, grid)
    # integration limits
    d = 3
    xmin = np.array([-1.0, 0.25, 2.0]); xmax = np.array([3.0, 0.75, 2.5])
    grid = trained_grid(R.product, d, seed=300)
    r = R.TECH_CUT + np.random.default_rng(301).random((N, d)) * (1 - 2 * R.TECH_CUT)
    _, _, hist, det = R.vegas_run_event(r, grid, R.product, N, xmin, xmax - xmin)
    out.update(limits_rnds=r, limits_grid=grid, limits_xmin=xmin, limits_xmax=xmax,
               limits_x=det["x"], limits_w=det["w"], limits_wf=det["wf"],
               limits_ind=det["ind"].astype(np.int32))
    # plain
    r = R.TECH_CUT + np.random.default_rng(302).random((N, 4)) * (1 - 2 * R.TECH_CUT)
    _, _, det = R.plain_run_event(r, R.symgauss, N)
    out.update(plain_rnds=r, plain_wf=det["wf"])
    # VEGAS+ digest, d=3, n_strat from plus_setup(3, 4000)
    st = R.plus_setup(3, 4000, adaptive=True)
    cubes = R.hypercube_coords(st["n_strat"], 3)
    rng = np.random.default_rng(303)
    n_ev = (st["n_ev"] + rng.integers(0, 4, size=st["n_cubes"])).astype(np.int32)
    n = int(n_ev.sum())
    r = R.TECH_CUT + rng.random((n, 3)) * (1 - 2 * R.TECH_CUT)
    grid = trained_grid(R.symgauss, 3, seed=304)
    ress, var, hist, det = R.plus_run_event(r, st["n_strat"], n_ev, cubes, grid, R.symgauss,
                                            st["xjac"])
    res, sigma = R.plus_result(ress, var, n_ev)
    new_n_ev, new_total = R.plus_redistribute(var, st["min_neval_hcube"], st["init_calls"])
    out.update(plus_rnds=r, plus_grid=grid, plus_n_ev=n_ev, plus_n_strat=st["n_strat"],
               plus_x=det["x"], plus_w=det["w"], plus_ind=det["ind"].astype(np.int32),
               plus_wf=det["wf"], plus_ress=ress, plus_var=var, plus_hist=hist,
               plus_res=res, plus_sigma=sigma, plus_new_n_ev=new_n_ev,
               plus_min_neval=st["min_neval_hcube"], plus_init_calls=st["init_calls"])
    np.savez_compressed(os.path.join(HERE, "vegas_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "vegas_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
