"""
CPU tests of the DEVICE source: vegasflow_b200/csrc/vf_common.cuh and vf_integrands.cuh are
compiled unchanged with g++ through tests/host_shim/cuda_shim.h (every CUDA intrinsic restated
with its IEEE meaning, -ffp-contract=off) and checked against the oracle.  The GPU tests check
the same functions as compiled by nvcc; these catch regressions without a GPU and reach inputs
the GPU tests do not (exp tail, wide division ranges).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import vegas_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "host_shim")
OUT = os.path.join(ROOT, "tests", "host_shim", "libdevice_source_host.so")


@pytest.fixture(scope="module")
def hs():
    srcs = [os.path.join(SHIM, "integrands_host.cpp"), os.path.join(SHIM, "cuda_shim.h"),
            os.path.join(ROOT, "vegasflow_b200", "csrc", "vf_common.cuh"),
            os.path.join(ROOT, "vegasflow_b200", "csrc", "vf_integrands.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-frounding-math", "-fPIC",
             "-shared", "-I", SHIM, "-I", os.path.join(ROOT, "vegasflow_b200", "csrc"), "-I",
             os.path.join(ROOT, "include"), srcs[0], "-o", OUT])
    return C.CDLL(OUT)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _integrand(hs, iid, x, pref=0.0, c=0.0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.shape[0])
    rc = hs.hs_integrand(C.c_int(iid), C.c_int(x.shape[1]), C.c_long(x.shape[0]), _p(x),
                         C.c_double(pref), C.c_double(c), _p(out))
    assert rc == 0
    return out


def test_philox_and_uniform_conversions(hs):
    out = np.zeros(4, dtype=np.uint32)
    for ctr, seed in (([0, 0, 0, 0], 0), ([5, 1, 2, 9], 0xDEADBEEF12345678)):
        c = np.array(ctr, dtype=np.uint32)
        hs.hs_philox(_p(c, C.c_uint32), C.c_uint64(seed), _p(out, C.c_uint32))
        want = co.philox4x32_10(c, [seed & 0xFFFFFFFF, seed >> 32])
        assert list(out) == list(want)
    # the two stream definitions against the oracle's, via the raw Philox words
    seed, it, n, d = 77, 3, 4096, 8
    for bits in (52, 32):
        co.set_rng_bits(bits)
        try:
            want = co.uniforms(seed, it, 0, n, d)
        finally:
            co.set_rng_bits(52)
        per = 2 if bits == 52 else 4
        got = np.empty((n, d))
        for ev in range(0, n, 512):  # a subset of events is enough
            for p in range(d // per):
                c = np.array([ev, 0, p, it], dtype=np.uint32)
                w = co.philox4x32_10(c, [seed, 0])
                words = (np.array([w[0], w[1], w[2], w[3]], dtype=np.uint32) if bits == 52 else
                         np.array([w[0], 0, w[1], 0, w[2], 0, w[3], 0], dtype=np.uint32))
                u = np.empty(per)
                hs.hs_uniforms(C.c_long(per), _p(words, C.c_uint32), C.c_int(bits), _p(u))
                got[ev, per * p: per * p + per] = u
            np.testing.assert_array_equal(got[ev], want[ev])


def test_exact_division_sequences(hs):
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.random(200000) - 0.5, (rng.random(200000) - 0.5) * 8,
                        np.ldexp(1 + rng.random(100000), rng.integers(-300, 300, 100000)),
                        [0.0, -0.0, 0.5, -0.5, 1e-300, 5e-324]])
    out = np.empty_like(y)
    hs.hs_div_by_tenth(C.c_long(y.size), _p(y), _p(out))
    np.testing.assert_array_equal(out, y / 0.1)
    for b in (2.0, 3.0, 7.0, 10.0, 70.0, 7620.0, 15241.0, 99996201.0):
        yy = np.concatenate([rng.random(100000) * 50 * min(b, 100), np.ldexp(1 + rng.random(50000),
                                                                          rng.integers(-40, 40, 50000))])
        out = np.empty_like(yy)
        hs.hs_div_rn_by(C.c_long(yy.size), _p(yy), C.c_double(b), _p(out))
        np.testing.assert_array_equal(out, yy / b)


def test_exp_nonpositive_accuracy_and_tail(hs):
    rng = np.random.default_rng(1)
    x = -np.concatenate([rng.random(200000) * 700, rng.random(50000) * 50, rng.random(20000) * 1e-3,
                         [0.0, 1e-320, 699.9999, 700.0, 700.0001, 708.3, 720.0, 744.0, 745.2, 799.9,
                          800.1, 1e4, np.inf]])
    out = np.empty_like(x)
    hs.hs_exp_nonpositive(C.c_long(x.size), _p(x), _p(out))
    want = np.exp(x)
    normal = want > 1e-300
    rel = np.abs(out[normal] - want[normal]) / want[normal]
    assert rel.max() < 4.5e-16  # <= 2 ulp against libm
    # subnormal tail: absolute error within one subnormal step, exact zero far out
    assert np.abs(out[~normal] - want[~normal]).max() <= 1e-300 * 1e-15 + 5e-324
    assert out[-1] == 0.0 and out[-2] == 0.0 and (out >= 0).all()
    assert out[np.argmax(x == 0.0)] == 1.0


def test_vegas_map_dim_against_oracle(hs, golden):
    key = "symgauss_d4"
    r, grid = golden[key + "_rnds"], golden[key + "_grid"]
    n = r.shape[0]
    for j in range(4):
        xn = 50.0 * (1.0 - r[:, j])
        x = np.empty(n); wf = np.empty(n); b = np.empty(n, dtype=np.int32)
        row = np.ascontiguousarray(grid[j])
        hs.hs_vegas_map(C.c_long(n), _p(np.ascontiguousarray(xn)), _p(row), _p(x), _p(wf),
                        _p(b, C.c_int))
        np.testing.assert_array_equal(b, golden[key + "_ind"][:, j])  # RD-add floor == trunc
        np.testing.assert_array_equal(x, golden[key + "_x"][:, j])
        np.testing.assert_array_equal(wf, (grid[j][b + 1] - grid[j][b]) * 50.0)
    # floor/truncation on exact integers and just below them
    xn = np.array([0.0, 1.0, np.nextafter(1.0, 0), 49.0, np.nextafter(50.0, 0), 17.5, 2.0**-60])
    x = np.empty(xn.size); wf = np.empty(xn.size); b = np.empty(xn.size, dtype=np.int32)
    hs.hs_vegas_map(C.c_long(xn.size), _p(xn), _p(np.ascontiguousarray(R.initial_divisions(1)[0])),
                    _p(x), _p(wf), _p(b, C.c_int))
    np.testing.assert_array_equal(b, xn.astype(np.int32))


@pytest.mark.parametrize("d", [1, 2, 4, 8, 20])
def test_symgauss_device_source(hs, d):
    rng = np.random.default_rng(d)
    x = rng.random((50000, d))
    pref, c = R.symgauss_constants(d)
    got = _integrand(hs, 0, x, pref, c)
    want = R.symgauss(x)
    ok = want > 1e-300
    assert (np.abs(got[ok] - want[ok]) / want[ok]).max() <= 1e-12
    # identical up to the exp implementation: the argument of exp is bit-exact
    got_c = co.integrand("symgauss", x)
    assert (np.abs(got[ok] - got_c[ok]) / got_c[ok]).max() <= 5e-16


def test_product_device_source(hs):
    x = np.random.default_rng(3).random((20000, 8))
    np.testing.assert_array_equal(_integrand(hs, 1, x), R.product(x))


@pytest.mark.parametrize("iid,name,d", [(2, "drellyan_lo", 4), (3, "singletop_lo", 3)])
def test_matrix_elements_device_source(hs, golden, iid, name, d):
    """The spinor chains as written for the GPU (exact-zero terms dropped, sincos pairing)
    against the literal numpy restatement: 1e-12 for the bulk, small exceedance."""
    x = golden[f"{name}_d{d}_x"]
    got = _integrand(hs, iid, x)
    want = R.INTEGRANDS[name](x)
    rel = np.abs(got - want) / np.abs(want)
    assert np.isfinite(got).all()
    assert np.quantile(rel, 0.99) <= 1e-12 and (rel > 1e-12).mean() < 5e-3
    x2 = 1e-8 + np.random.default_rng(8).random((100000, d)) * (1 - 2e-8)
    got, want = _integrand(hs, iid, x2), R.INTEGRANDS[name](x2)
    rel = np.abs(got - want) / np.abs(want)
    assert np.median(rel) < 5e-15 and np.quantile(rel, 0.99) <= 1e-12
    assert (rel > 1e-12).mean() < 5e-3
    assert abs(got.sum() - want.sum()) <= 1e-12 * np.abs(want).sum()
    if name == "drellyan_lo":
        closed = R.drellyan_closed_form(x2)
        assert np.median(np.abs(got - closed) / closed) < 2e-15


def test_singletop_threshold_and_edge_regions(hs):
    """single-top where the reference's own rounding is amplified (vf_integrands.cuh::angles_acos):
    near threshold (x0 -> 0) the projected top momentum is anti-parallel to the beam and
    cos(theta/2), sin(theta) carry the rounding of theta = rn(acos(c)); deep in that region
    cos(phi) carries the rounding of phi = rn(acos(cx)) ~ pi/2.  The device source reproduces both
    roundings without acos; the plain half-angle forms would leave 14 % of the first sample and
    0.4 % of uniformly drawn points outside the bar."""
    rng = np.random.default_rng(3)
    n = 100000
    for lo in (-8, -4, -2):
        x = np.column_stack([10.0 ** rng.uniform(lo, 0, n), rng.random(n), rng.random(n)])
        x = np.clip(x, R.TECH_CUT, 1 - R.TECH_CUT)
        got, want = _integrand(hs, 3, x), R.INTEGRANDS["singletop_lo"](x)
        rel = np.abs(got - want) / np.abs(want)
        assert np.isfinite(got).all()
        assert np.quantile(rel, 0.999) <= 1e-12 and (rel > 1e-12).mean() < 1e-3, lo
    # both ends of x2 (cos of the scattering angle -> +-1) and of x1
    ends = np.where(rng.random(n) < 0.5, 10.0 ** rng.uniform(-8, -1, n),
                    1 - 10.0 ** rng.uniform(-8, -1, n))
    for col in (1, 2):
        x = rng.random((n, 3))
        x[:, col] = ends
        x = np.clip(x, R.TECH_CUT, 1 - R.TECH_CUT)
        got, want = _integrand(hs, 3, x), R.INTEGRANDS["singletop_lo"](x)
        rel = np.abs(got - want) / np.abs(want)
        assert np.isfinite(got).all()
        assert np.quantile(rel, 0.999) <= 1e-12 and (rel > 1e-12).mean() < 1e-3, col


def test_drellyan_edge_regions(hs):
    """Drell-Yan at the ends of every variable (x2 -> 1 makes Delta - 1 ill-conditioned: ptmax,
    pta and Delta keep the reference's operations bit for bit there)."""
    rng = np.random.default_rng(4)
    n = 100000
    ends = np.where(rng.random(n) < 0.5, 10.0 ** rng.uniform(-8, -1, n),
                    1 - 10.0 ** rng.uniform(-8, -1, n))
    for col in range(4):
        x = rng.random((n, 4))
        x[:, col] = ends
        x = np.clip(x, R.TECH_CUT, 1 - R.TECH_CUT)
        got, want = _integrand(hs, 2, x), R.INTEGRANDS["drellyan_lo"](x)
        rel = np.abs(got - want) / np.abs(want)
        assert np.isfinite(got).all()
        assert np.quantile(rel, 0.99) <= 1e-12 and (rel > 1e-12).mean() < 5e-3, col


def test_drellyan_tiny_kappa(hs):
    """Drell-Yan grows like |ln kappa|/kappa, so a trained grid zooms in on kappa -> 0 iteration
    after iteration and mV = sqrt(s kappa) runs through hundreds of binades.  The device source
    collects its scalar quotients into one division on an exactly rescaled mV
    (vf_integrands.cuh): finite and inside the bar wherever the reference is."""
    rng = np.random.default_rng(5)
    for ex in (-20, -60, -80, -120, -200, -290):
        n = 2000
        x = rng.random((n, 4))
        x[:, 0] = 10.0 ** ex * (0.5 + rng.random(n))
        got = _integrand(hs, 2, x)
        with np.errstate(all="ignore"):
            want = R.INTEGRANDS["drellyan_lo"](x)
        assert np.isfinite(want).all() and np.isfinite(got).all(), ex
        rel = np.abs(got - want) / np.abs(want)
        assert np.quantile(rel, 0.99) <= 1e-12, ex


def test_implemented_chain_op_counts():
    """F_alg of the matrix elements as SURVEY 8(d) prescribes: vf_integrands.cuh compiled with an
    op-counting scalar (tests/host_shim/count_flops_host.cpp; add/sub/mul/div/sqrt/
    transcendental = 1, fma = 2), averaged over uniformly drawn events, must be what
    vf_flops_per_event returns on top of the 12d+5 of the VEGAS step."""
    from vegasflow_b200 import _lib

    out = os.path.join(SHIM, "libcount_flops_host.so")
    src = os.path.join(SHIM, "count_flops_host.cpp")
    deps = [src, os.path.join(ROOT, "vegasflow_b200", "csrc", "vf_common.cuh"),
            os.path.join(ROOT, "vegasflow_b200", "csrc", "vf_integrands.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", SHIM, "-I",
                               os.path.join(ROOT, "vegasflow_b200", "csrc"), "-I",
                               os.path.join(ROOT, "include"), src, "-o", out])
    lib = C.CDLL(out)
    lib.hs_count_flops.restype = C.c_double
    abi = _lib.load()
    rng = np.random.default_rng(1)
    for iid, name, d in ((2, "drellyan_lo", 4), (3, "singletop_lo", 3)):
        x = R.TECH_CUT + rng.random((100000, d)) * (1 - 2 * R.TECH_CUT)
        ops = lib.hs_count_flops(C.c_int(iid), C.c_int(d), C.c_long(x.shape[0]), _p(x),
                                 C.c_double(0.0), C.c_double(0.0))
        table = abi.vf_flops_per_event(1, abi.vf_integrand_id(name.encode()), d, 0) - (12 * d + 5)
        assert abs(ops - table) <= 1.0, (name, ops, table)
    # the counter itself: product of 8 numbers is 7 multiplications
    x = rng.random((1000, 8))
    assert lib.hs_count_flops(C.c_int(1), C.c_int(8), C.c_long(1000), _p(x), C.c_double(0.0),
                              C.c_double(0.0)) == 7.0
