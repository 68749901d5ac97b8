"""
CPU tests that PIN the oracle to the reference's own code.

tests/golden/reference_golden.npz holds what the UNMODIFIED reference (/root/reference,
N3PDF/vegasflow v1.4.0) computes when executed on the numpy-backed tensorflow stand-in of
tests/ref_shim (generator: tests/golden/make_golden_from_reference.py).  Here:

  * the numpy oracle (oracle/vegas_ref.py) and the C oracle (oracle/vegas_oracle.c) must
    reproduce those vectors -- bit-exact ind / x / w / per-event w*f for the numpy oracle;
  * where /root/reference is present (this container, not the GPU box) the reference's own
    test-suite runs on the shim, and the committed golden file is regenerated and compared
    bit for bit, so the fixture cannot drift from the reference + shim that produced it.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import vegas_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("VEGASFLOW_REFERENCE", "/root/reference")
HAVE_REFERENCE = os.path.isdir(os.path.join(REFERENCE, "src", "vegasflow"))
needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present")

PER_EVENT = ["symgauss_d2", "symgauss_d4", "symgauss_d8", "symgauss_d20", "product_d1",
             "product_d3", "product_d8", "drellyan_lo_d4", "singletop_lo_d3",
             "flat_symgauss_d8", "flat_product_d8"]


def _name(key):
    return key.replace("flat_", "").rsplit("_d", 1)[0]


def _shim_env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [os.path.join(ROOT, "tests", "ref_shim"), os.path.join(REFERENCE, "src"),
         os.path.join(REFERENCE, "examples"), ROOT])
    env["VEGASFLOW_LOG_LEVEL"] = "0"
    return env


# ------------------------------------------------------------------ the reference itself
@needs_reference
def test_reference_own_tests_pass_on_the_shim():
    """src/vegasflow/tests/{test_utils,test_algs,test_misc}.py of the reference, unmodified, run
    in a subprocess against the shim (a subprocess, so that the stand-in `tensorflow` never
    enters this interpreter).  test_utils.py:11-30 is the reference's one value-pinned test of
    this path; the others are its statistical integration tests."""
    tests = [os.path.join(REFERENCE, "src", "vegasflow", "tests", f)
             for f in ("test_utils.py", "test_algs.py", "test_misc.py")]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                        *tests], cwd="/tmp", env=_shim_env(), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@needs_reference
def test_committed_golden_is_what_the_reference_computes(tmp_path):
    """Regenerate the golden file from the reference and compare with the committed one."""
    env = _shim_env()
    script = os.path.join(ROOT, "tests", "golden", "make_golden_from_reference.py")
    env["VEGASFLOW_GOLDEN_OUT"] = str(tmp_path / "regen.npz")
    r = subprocess.run([sys.executable, script], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    new = dict(np.load(tmp_path / "regen.npz"))
    old = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz")))
    assert sorted(new) == sorted(old)
    for k in old:
        np.testing.assert_array_equal(new[k], old[k], err_msg=k)


# ------------------------------------------------------------------ numpy oracle, per event
@pytest.mark.parametrize("key", PER_EVENT)
def test_numpy_oracle_reproduces_reference_per_event(golden, key):
    name = _name(key)
    r, grid = golden[key + "_rnds"], golden[key + "_grid"]
    n = r.shape[0]
    res, res2, hist, det = R.vegas_run_event(r, grid, R.INTEGRANDS[name], n)
    np.testing.assert_array_equal(det["ind"], golden[key + "_ind"])
    np.testing.assert_array_equal(det["x"], golden[key + "_x"])
    np.testing.assert_array_equal(det["w"], golden[key + "_w"])
    np.testing.assert_array_equal(det["f"], golden[key + "_f"])
    np.testing.assert_array_equal(det["wf"], golden[key + "_wf"])
    np.testing.assert_array_equal(hist, golden[key + "_hist"])
    # scalar totals: the oracle sums pairwise, the shim left to right (order is a convention)
    assert abs(res - golden[key + "_res"]) <= 1e-14 * abs(golden[key + "_res"])
    assert abs(res2 - golden[key + "_res2"]) <= 1e-14 * abs(golden[key + "_res2"])
    np.testing.assert_array_equal(R.refine_grid(hist, grid), golden[key + "_newgrid"])


@pytest.mark.parametrize("key", [k for k in PER_EVENT if _name(k) in co.INTEGRAND_IDS])
def test_c_oracle_reproduces_reference_per_event(golden, key):
    name = _name(key)
    r, grid = golden[key + "_rnds"], golden[key + "_grid"]
    n = r.shape[0]
    x, w, ind, wf = co.digest_from_uniforms(co.MODE_VEGAS, name, r, grid, 1.0 / n)
    np.testing.assert_array_equal(ind, golden[key + "_ind"])
    np.testing.assert_array_equal(x, golden[key + "_x"])
    np.testing.assert_array_equal(w, golden[key + "_w"])
    rel = np.abs(wf - golden[key + "_wf"]) / np.abs(golden[key + "_wf"])
    assert rel.max() <= 1e-12
    np.testing.assert_allclose(co.refine_grid(golden[key + "_hist"], grid),
                               golden[key + "_newgrid"], rtol=0, atol=1e-14)


def test_oracle_limits_plain_scatter(golden):
    g = golden
    xmin, xmax = g["limits_xmin"], g["limits_xmax"]
    n = g["limits_rnds"].shape[0]
    _, _, hist, det = R.vegas_run_event(g["limits_rnds"], g["limits_grid"], R.product, n, xmin,
                                        xmax - xmin)
    for k in ("x", "w", "ind", "wf"):
        np.testing.assert_array_equal(det[k], g["limits_" + k])
    np.testing.assert_array_equal(hist, g["limits_hist"])
    n = g["plain_rnds"].shape[0]
    res, res2, det = R.plain_run_event(g["plain_rnds"], R.symgauss, n)
    np.testing.assert_array_equal(det["wf"], g["plain_wf"])
    assert abs(res - g["plain_res"]) <= 1e-14 * abs(g["plain_res"])
    assert abs(R.plain_sigma(res, res2, n) - g["plain_it"][1]) <= 1e-12 * g["plain_it"][1]
    out = R.consume_array_into_indices(g["scatter_in"], g["scatter_idx"].reshape(-1, 1), 50)
    np.testing.assert_array_equal(out, g["scatter_out"])
    onehot = R.consume_array_into_indices_onehot(g["scatter_in"], g["scatter_idx"].reshape(-1, 1),
                                                 50)
    np.testing.assert_allclose(onehot, g["scatter_out"], rtol=1e-15)


def test_oracle_refine_edge_cases(golden):
    """Empty bins, a single spike, 20 decades of dynamic range, all-zero and flat histograms, on
    a flat and on a trained grid (vflow.py:135-211 executed by the reference)."""
    for h, sub, want in zip(golden["refine_hist"], golden["refine_sub"], golden["refine_new"]):
        assert np.isfinite(want).all() and (np.diff(want) >= 0).all()
        np.testing.assert_array_equal(R.refine_grid_per_dimension(h, sub), want)
        got_c = co.refine_grid(h.reshape(1, 50), sub.reshape(1, 51))[0]
        np.testing.assert_allclose(got_c, want, rtol=0, atol=1e-14)


def test_oracle_vegas_plus_per_event(golden):
    g = golden
    n_ev, n_strat = g["plus_n_ev"], int(g["plus_n_strat"])
    cubes = R.hypercube_coords(n_strat, 3)
    ress, var, hist, det = R.plus_run_event(g["plus_rnds"], n_strat, n_ev, cubes, g["plus_grid"],
                                            R.symgauss, 1.0 / len(n_ev))
    for k in ("x", "w", "ind", "wf"):
        np.testing.assert_array_equal(det[k], g["plus_" + k])
    np.testing.assert_array_equal(ress, g["plus_ress"])
    np.testing.assert_array_equal(var, g["plus_var"])
    np.testing.assert_array_equal(hist, g["plus_hist"])
    res, sigma = R.plus_result(ress, var, n_ev)
    assert abs(res - g["plus_res"]) <= 1e-14 * abs(g["plus_res"])
    assert abs(sigma - g["plus_sigma"]) <= 1e-13 * g["plus_sigma"]
    new_n_ev, _ = R.plus_redistribute(var, int(g["plus_min_neval"]), int(g["plus_init_calls"]))
    np.testing.assert_array_equal(new_n_ev, g["plus_new_n_ev"])
    np.testing.assert_array_equal(R.refine_grid(hist, g["plus_grid"]), g["plus_newgrid"])


# ------------------------------------------------------------------ stream family
def _philox_draw(seed):
    def draw(n, d, iteration=0, offset=0):
        return co.uniforms(seed, iteration, offset, n, d)

    return draw


STREAM_VEGAS = [("c1", "symgauss"), ("c2", "product"), ("c4dy", "drellyan_lo"),
                ("c4st", "singletop_lo"), ("c5", "symgauss"), ("lim", "product")]


@pytest.mark.parametrize("tag,name", STREAM_VEGAS)
def test_oracle_run_integration_on_engine_stream(golden, tag, name):
    """run_integration of the reference fed with the engine's Philox stream vs the oracle's
    iteration loop on the same stream: (res, sigma) per iteration and the final grid."""
    d, n, seed, n_iter = (int(v) for v in golden[f"stream_{tag}_meta"])
    kw = {}
    if tag == "lim":
        kw = dict(xmin=golden["limits_xmin"], xmax=golden["limits_xmax"])
    _, _, results, grid = R.vegas_integrate(R.INTEGRANDS[name], d, n, n_iter, _philox_draw(seed),
                                            **kw)
    want = golden[f"stream_{tag}_results"]
    np.testing.assert_allclose(np.array(results), want, rtol=1e-11)
    np.testing.assert_allclose(grid, golden[f"stream_{tag}_grid"], rtol=0, atol=1e-12)


def test_oracle_plain_and_plus_streams(golden):
    d, n, seed, n_iter = (int(v) for v in golden["stream_plain_meta"])
    _, _, results = R.plain_integrate(R.symgauss, d, n, n_iter, _philox_draw(seed))
    np.testing.assert_allclose(np.array(results), golden["stream_plain_results"], rtol=1e-12)
    for tag, name, adaptive in (("c3", "symgauss", True), ("plus4", "symgauss", False),
                                ("plus3a", "product", True)):
        d, n, seed, n_iter = (int(v) for v in golden[f"stream_{tag}_meta"])
        _, _, results, grid, n_ev = R.plus_integrate(R.INTEGRANDS[name], d, n, n_iter,
                                                     _philox_draw(seed), adaptive=adaptive)
        np.testing.assert_allclose(np.array(results), golden[f"stream_{tag}_results"], rtol=1e-10)
        np.testing.assert_allclose(grid, golden[f"stream_{tag}_grid"], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(n_ev, golden[f"stream_{tag}_n_ev"][-1])
        assert int(golden[f"stream_{tag}_n_strat"]) == R.plus_setup(d, n, adaptive)["n_strat"]


def test_c_oracle_streams(golden):
    """The fused C restatement on its own Philox stream against the reference fed with it."""
    for tag, name in (("c1", "symgauss"), ("c2", "product"), ("c5", "symgauss")):
        d, n, seed, n_iter = (int(v) for v in golden[f"stream_{tag}_meta"])
        grid = R.initial_divisions(d)
        for it in range(n_iter):
            s1, s2, hist = co.run_event(co.MODE_VEGAS, name, d, 0, n, 1.0 / n, seed, it, True, grid)
            want = golden[f"stream_{tag}_results"][it]
            assert abs(s1 - want[0]) <= 1e-10 * abs(want[0])
            assert abs(R.vegas_sigma(s1, s2, n) - want[1]) <= 1e-9 * want[1]
            grid = co.refine_grid(hist, grid)
        np.testing.assert_allclose(grid, golden[f"stream_{tag}_grid"], rtol=0, atol=1e-11)
