"""
CPU tests of the oracle (oracle/): the numpy restatement against the reference's
only value-pinned test, against analytic answers, against the committed golden
vectors, and against the independent C restatement.
"""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import vegas_ref as R


def test_histogram_scatter_like_reference_test_utils():
    """Restates src/vegasflow/tests/test_utils.py:11-30 (the one pinned test)."""
    rng = np.random.default_rng(0)
    size_in = int(rng.integers(5, 100))
    size_out = int(rng.integers(1, size_in - 3))
    input_array = rng.random(size_in)
    indices = rng.integers(0, size_out, size=size_in)
    result = R.consume_array_into_indices(input_array, indices.reshape(-1, 1), size_out)
    onehot = R.consume_array_into_indices_onehot(input_array, indices.reshape(-1, 1), size_out)
    np.testing.assert_allclose(result, onehot, rtol=1e-14)
    np.testing.assert_almost_equal(np.sum(input_array), np.sum(result))
    check = np.zeros(size_out)
    for val, i in zip(input_array, indices):
        check[i] += val
    np.testing.assert_allclose(check, result)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert list(co.philox4x32_10(ctr, key)) == want


def test_uniform_stream_range_and_independence_of_chunking():
    r = co.uniforms(3, 2, 0, 50000, 5)
    assert r.min() > R.TECH_CUT * 0.99 and r.max() <= 1 - R.TECH_CUT
    assert abs(r.mean() - 0.5) < 5e-3
    a = co.uniforms(3, 2, 1000, 100, 5)
    np.testing.assert_array_equal(a, r[1000:1100])  # counter-based: offset == slice
    assert not np.array_equal(co.uniforms(3, 3, 0, 10, 5), r[:10])  # iteration changes stream
    assert not np.array_equal(co.uniforms(4, 2, 0, 10, 5), r[:10])  # seed changes stream


def test_uniform_stream_definition_makes_the_first_reference_steps_exact():
    """r = 2 - v with v = fma(m, S, 3T) in [1+T, 2-T): r is a multiple of 2^-52, so 1-r (vflow.py:117)
    is exact and rn(50*(1-r)) == rn(50*v - 50) -- what the fused kernel evaluates as ONE fma on v
    (vf_common.cuh::u52_to_v).  Checked in exact rational arithmetic against the Philox words."""
    from fractions import Fraction
    seed, it, d = 11, 3, 6
    r = co.uniforms(seed, it, 40, 64, d)
    assert np.array_equal(r * 2.0**52, np.floor(r * 2.0**52))  # on the 2^-52 grid
    T, S = Fraction(R.TECH_CUT), Fraction(1.0 - 2.0 * R.TECH_CUT)
    for ev in (40, 41, 103):
        for p in range(d // 2):
            w = co.philox4x32_10(np.array([ev, 0, p, it], dtype=np.uint32), [seed, 0])
            for h in range(2):
                hi, lo = int(w[2 * h]), int(w[2 * h + 1])
                m = 1 + Fraction(((hi & 0xFFFFF) << 32) | lo, 2**52)
                exact_v = m * S + Fraction(3.0 * R.TECH_CUT)
                got = Fraction(float(r[ev - 40, 2 * p + h]))
                v = 2 - got                                   # exact by construction
                assert abs(v - exact_v) <= Fraction(1, 2**53)  # one rounding to the grid of [1,2)
                # the reference's first two operations on r, and the kernel's single fma on v
                one_minus_r = 1.0 - float(got)
                assert Fraction(one_minus_r) == 1 - got        # exact
                xn_ref = 50.0 * one_minus_r                   # rn(50 * (1 - r))
                exact_xn = 50 * v - 50
                ulp = Fraction(np.spacing(xn_ref))
                assert abs(Fraction(xn_ref) - exact_xn) <= ulp / 2  # == rn(50 v - 50), the fma


@pytest.mark.parametrize("name,d", [("symgauss", 2), ("symgauss", 4), ("symgauss", 8),
                                     ("symgauss", 20), ("product", 1), ("product", 3),
                                     ("product", 8)])
def test_numpy_vs_c_vs_golden_digest(golden, name, d):
    key = f"{name}_d{d}"
    r, grid = golden[key + "_rnds"], golden[key + "_grid"]
    n = r.shape[0]
    _, _, hist, det = R.vegas_run_event(r, grid, R.INTEGRANDS[name], n)
    np.testing.assert_array_equal(det["ind"], golden[key + "_ind"])
    np.testing.assert_array_equal(det["x"], golden[key + "_x"])
    np.testing.assert_array_equal(det["w"], golden[key + "_w"])
    np.testing.assert_allclose(det["wf"], golden[key + "_wf"], rtol=1e-13)
    x, w, ind, wf = co.digest_from_uniforms(co.MODE_VEGAS, name, r, grid, 1.0 / n)
    np.testing.assert_array_equal(ind, golden[key + "_ind"])
    np.testing.assert_array_equal(x, golden[key + "_x"])
    np.testing.assert_array_equal(w, golden[key + "_w"])
    np.testing.assert_allclose(wf, golden[key + "_wf"], rtol=1e-12, atol=0)
    # bins are in range and the extremes of the uniform interval land in bins 49 / 0
    assert ind.min() >= 0 and ind.max() <= 49
    assert (ind[0] == 49).all() and (ind[1] == 0).all()
    np.testing.assert_allclose(co.refine_grid(hist, grid), golden[key + "_newgrid"], rtol=0,
                               atol=1e-13)


def test_limits_and_plain_against_golden(golden):
    r, grid = golden["limits_rnds"], golden["limits_grid"]
    xmin, xmax = golden["limits_xmin"], golden["limits_xmax"]
    n = r.shape[0]
    x, w, ind, wf = co.digest_from_uniforms(co.MODE_VEGAS, "product", r, grid, 1.0 / n, xmin,
                                            xmax - xmin)
    np.testing.assert_array_equal(x, golden["limits_x"])
    np.testing.assert_array_equal(w, golden["limits_w"])
    np.testing.assert_array_equal(wf, golden["limits_wf"])
    assert (x >= xmin).all() and (x <= xmax).all()
    r = golden["plain_rnds"]
    _, _, _, wf = co.digest_from_uniforms(co.MODE_PLAIN, "symgauss", r, None, 1.0 / n)
    np.testing.assert_allclose(wf, golden["plain_wf"], rtol=1e-12)


def test_refine_grid_properties():
    rng = np.random.default_rng(9)
    grid = R.initial_divisions(3)
    hist = rng.random((3, 50)) ** 4
    hist[1, 10:20] = 0.0  # empty bins hit the 1e-30 floor
    new = R.refine_grid(hist, grid)
    assert (np.diff(new, axis=1) > 0).all()
    assert (new[:, 0] == 0).all() and (new[:, -1] == 1).all()
    np.testing.assert_allclose(co.refine_grid(hist, grid), new, rtol=0, atol=1e-13)
    # a flat histogram leaves a flat grid unchanged (to rounding)
    flat = R.refine_grid(np.ones((1, 50)), R.initial_divisions(1))
    np.testing.assert_allclose(flat, R.initial_divisions(1), atol=1e-12)


def test_known_integrals_numpy_oracle():
    draw = R.uniform_source_numpy(1)
    res, err, _, _ = R.vegas_integrate(R.symgauss, 4, 100000, 5, draw)
    assert abs(res - 1.0) < 3 * err
    res, err, _, _ = R.vegas_integrate(R.product, 8, 100000, 5, draw)
    assert abs(res - 2.0**-8) < 3 * err
    xmin, xmax = [-0.5, 0.2], [3.1, 3.9]
    res, err, _, _ = R.vegas_integrate(R.product, 2, 20000, 5, draw, xmin=xmin, xmax=xmax)
    want = np.prod(np.array(xmax) ** 2 / 2 - np.array(xmin) ** 2 / 2)
    assert abs(res - want) < 3 * err
    res, err, _ = R.plain_integrate(R.product, 3, 100000, 3, draw)
    assert abs(res - 0.125) < 3 * err


def test_c_oracle_integration_on_philox_stream():
    """Whole VEGAS loop of the C restatement on the engine's own stream."""
    d, n = 4, 200000
    grid = R.initial_divisions(d)
    results = []
    for it in range(5):
        s1, s2, hist = co.run_event(co.MODE_VEGAS, "symgauss", d, 0, n, 1.0 / n, 42, it, True, grid)
        results.append((s1, R.vegas_sigma(s1, s2, n)))
        grid = co.refine_grid(hist, grid)
    res, err = R.combine_iterations(results)
    assert abs(res - 1.0) < 3 * err
    assert err < 2e-3
    # chunking / sharding invariance: two half ranges sum to the full range
    a = co.run_event(co.MODE_VEGAS, "symgauss", d, 0, n // 2, 1.0 / n, 42, 5, True, grid)
    b = co.run_event(co.MODE_VEGAS, "symgauss", d, n // 2, n - n // 2, 1.0 / n, 42, 5, True, grid)
    full = co.run_event(co.MODE_VEGAS, "symgauss", d, 0, n, 1.0 / n, 42, 5, True, grid)
    np.testing.assert_allclose(a[0] + b[0], full[0], rtol=1e-12)
    np.testing.assert_allclose(a[2] + b[2], full[2], rtol=1e-10)


def test_singletop_and_drellyan_known_answers(golden):
    draw = R.uniform_source_numpy(2)
    res, err, _, _ = R.vegas_integrate(R.singletop_lo, 3, 100000, 5, draw)
    assert abs(res - 423.9) < 4 * err + 0.5  # 423.9 +- 0.2 pb
    x = 1e-8 + np.random.default_rng(0).random((20000, 4)) * (1 - 2e-8)
    f, g = R.drellyan_lo(x), R.drellyan_closed_form(x)
    rel = np.abs(f - g) / np.abs(g)
    assert np.median(rel) < 2e-15 and np.quantile(rel, 0.99) < 1e-13
    for name, d in (("drellyan_lo", 4), ("singletop_lo", 3)):
        key = f"{name}_d{d}"
        n = golden[key + "_rnds"].shape[0]
        _, _, _, det = R.vegas_run_event(golden[key + "_rnds"], golden[key + "_grid"],
                                         R.INTEGRANDS[name], n)
        np.testing.assert_allclose(det["wf"], golden[key + "_wf"], rtol=1e-13)
        assert np.isfinite(det["wf"]).all() and (det["wf"] > 0).all()


def test_plus_setup_matches_survey_table():
    """SURVEY 9.1 sizes (vflowplus.py:113-139 evaluated in float32)."""
    want = {(2, 10**4, False): (70, 4900, 2, 9800), (2, 10**4, True): (50, 2500, 2, 5000),
            (4, 10**6, False): (10, 10000, 100, 10**6),
            (8, 10**8, False): (3, 6561, 15241, 99996201),
            (8, 10**8, True): (3, 6561, 7620, 49994820)}
    for (d, n, ad), (ns, nc, mn, ne) in want.items():
        st = R.plus_setup(d, n, ad)
        assert (st["n_strat"], st["n_cubes"], st["min_neval_hcube"], st["n_events"]) == (ns, nc, mn, ne)


def test_plus_numpy_vs_c_vs_golden(golden):
    r, grid, n_ev = golden["plus_rnds"], golden["plus_grid"], golden["plus_n_ev"]
    n_strat = int(golden["plus_n_strat"])
    ress, var, hist, det = co.plus_run_event("symgauss", 3, n_strat, n_ev, 1.0 / len(n_ev), 0, 0,
                                             True, grid, rnds=r, detail=True)
    np.testing.assert_array_equal(det["ind"], golden["plus_ind"])
    np.testing.assert_array_equal(det["x"], golden["plus_x"])
    np.testing.assert_array_equal(det["w"], golden["plus_w"])
    np.testing.assert_allclose(det["wf"], golden["plus_wf"], rtol=1e-12)
    np.testing.assert_allclose(ress, golden["plus_ress"], rtol=1e-12)
    np.testing.assert_allclose(hist, golden["plus_hist"], rtol=1e-11)
    res, sigma = R.plus_result(golden["plus_ress"], golden["plus_var"], n_ev)
    # np.sum is pairwise, the reference (on the shim) sums left to right
    assert abs(res - golden["plus_res"]) <= 1e-14 * abs(res)
    assert abs(sigma - golden["plus_sigma"]) <= 1e-13 * sigma
    new_n_ev, total = R.plus_redistribute(golden["plus_var"], int(golden["plus_min_neval"]),
                                          int(golden["plus_init_calls"]))
    np.testing.assert_array_equal(new_n_ev, golden["plus_new_n_ev"])
    assert new_n_ev.min() >= int(golden["plus_min_neval"]) and total == new_n_ev.sum()


def test_plus_integrates_to_one():
    draw = R.uniform_source_numpy(4)
    res, err, *_ = R.plus_integrate(R.symgauss, 2, 10000, 4, draw)
    assert abs(res - 1.0) < 3 * err
    res, err, *_ = R.plus_integrate(R.symgauss, 2, 10000, 4, draw, adaptive=True)
    assert abs(res - 1.0) < 3 * err


def test_rng32_stream_definition():
    """Optional 32-bit stream: r = (1-T) - (k + 1/2) * 2^-32 * S, inside (TECH_CUT, 1-TECH_CUT)."""
    co.set_rng_bits(32)
    try:
        r = co.uniforms(9, 0, 0, 100000, 7)
        ctr = np.array([5, 0, 1, 0], dtype=np.uint32)
        words = co.philox4x32_10(ctr, [9, 0])
        row = co.uniforms(9, 0, 5, 1, 7)[0]
    finally:
        co.set_rng_bits(52)
    assert r.min() > R.TECH_CUT and r.max() < 1 - R.TECH_CUT and abs(r.mean() - 0.5) < 3e-3
    want = (1 - R.TECH_CUT) - (words[:3].astype(np.float64) + 0.5) * 2.0**-32 * (1 - 2 * R.TECH_CUT)
    np.testing.assert_allclose(row[4:7], want, rtol=0, atol=2e-16)  # dims 4..6 <- block 1
    default = co.uniforms(9, 0, 0, 100, 7)
    assert not np.array_equal(default, r[:100])


def test_flop_counts_match_the_abi_table():
    """F_alg is counted, not guessed.  symgauss / product: the op-counting ndarray of
    oracle/count_flops.py on the numpy restatement agrees with vf_flops_per_event (SURVEY 8d
    formulas).  Matrix elements: the same count of the reference's LITERAL chain is what bench.py
    reports as `flops_per_event_reference_chain`; vf_flops_per_event itself carries the count of
    the implemented chain (tests/test_device_source_on_host.py::test_implemented_chain_op_counts)."""
    import bench
    from oracle.count_flops import count
    from vegasflow_b200 import _lib

    lib = _lib.load()
    for name, d in (("symgauss", 4), ("symgauss", 8), ("product", 8)):
        flops, _ = count(R.INTEGRANDS[name], d)
        abi = lib.vf_flops_per_event(1, lib.vf_integrand_id(name.encode()), d, 0) - (12 * d + 5)
        # symgauss: SURVEY counts d adds for the reduce_sum, the restatement starts from term 0
        assert abs(abi - flops) <= 1, (name, d, abi, flops)
    for name, d in (("drellyan_lo", 4), ("singletop_lo", 3)):
        flops, _ = count(R.INTEGRANDS[name], d)
        assert abs(bench.REFERENCE_CHAIN_OPS[name] - flops) <= 1, (name, flops)
        abi = lib.vf_flops_per_event(1, lib.vf_integrand_id(name.encode()), d, 0) - (12 * d + 5)
        assert abi < flops  # the implemented chain never forms the exact-zero terms


def test_reference_shaped_cpu_port_is_the_same_algorithm():
    """The torch-CPU port timed by `bench.py --impl reference` (oracle/ref_shaped_torch.py) and
    the numpy oracle are two restatements of the same reference lines: identical bins, x and
    histogram on the same uniforms (weights/sums up to the reduction order of torch.prod/sum)."""
    import torch

    from oracle import ref_shaped_torch as T

    rng = np.random.default_rng(12)
    d, n = 5, 4000
    r = R.TECH_CUT + rng.random((n, d)) * (1 - 2 * R.TECH_CUT)
    grid = np.sort(rng.random((d, 51)), axis=1)
    grid[:, 0], grid[:, -1] = 0.0, 1.0
    x, w, ind = T.generate_random_array(torch.from_numpy(r), torch.from_numpy(grid))
    xo, wo, io = R.vegas_digest(r, grid)
    np.testing.assert_array_equal(ind.numpy(), io)
    np.testing.assert_array_equal(x.numpy(), xo)
    np.testing.assert_allclose(w.numpy(), wo, rtol=1e-15)
    f = T.symgauss(x).numpy()
    np.testing.assert_allclose(f, R.symgauss(xo), rtol=1e-9)  # sum order inside (C + s) - C
    tmp2 = (wo / n * R.symgauss(xo)) ** 2
    h = T.consume_array_into_indices(torch.from_numpy(tmp2), ind[:, 2:3], 50).numpy()
    np.testing.assert_allclose(h, R.consume_array_into_indices(tmp2, io[:, 2:3], 50), rtol=1e-12)
    new = R.refine_grid(np.stack([h] * d), grid)
    assert (np.diff(new, axis=1) >= 0).all()


def test_unpinned_conventions_are_quantified():
    """What "parity unpinned" means in numbers: TensorFlow/Eigen may sum the d terms
    of symgauss in another order or contract a*b+c; the oracle fixes left-to-right, unfused.
    The (C + s) - C quantisation turns a 1-ulp change of s into ulp(C) for ~1e-3 of the events;
    everything that does not pass through that quantisation is insensitive at the 1e-15 level."""
    rng = np.random.default_rng(5)
    d = 8
    x = rng.random((200000, d))
    a = np.float64(0.1)
    sq = ((x - 0.5) / a) ** 2
    s_seq = sq[:, 0].copy()
    for j in range(1, d):
        s_seq = s_seq + sq[:, j]
    s_pair = ((sq[:, 0] + sq[:, 1]) + (sq[:, 2] + sq[:, 3])) + ((sq[:, 4] + sq[:, 5]) +
                                                                (sq[:, 6] + sq[:, 7]))
    _, C = R.symgauss_constants(d)
    f_seq = np.exp(-((C + s_seq) - C))
    f_pair = np.exp(-((C + s_pair) - C))
    rel = np.abs(f_seq - f_pair) / f_seq
    flipped = (rel > 1e-12).mean()
    assert 1e-5 < flipped < 5e-3          # ~ulp(s)/ulp(C) of the events move ...
    assert rel.max() < 2 * 5.9e-11        # ... by one ulp(C) = 5.8e-11 at d = 8
    # without the quantisation the two orders agree to a few ulp of s
    plain = np.abs(np.exp(-s_seq) - np.exp(-s_pair)) / np.exp(-s_seq)
    assert plain.max() < 1e-13
    # the product integrand and the VEGAS weight are order-sensitive only at the ulp level
    p_seq = R.product(x)
    p_pair = ((x[:, 0] * x[:, 1]) * (x[:, 2] * x[:, 3])) * ((x[:, 4] * x[:, 5]) * (x[:, 6] * x[:, 7]))
    assert (np.abs(p_seq - p_pair) / p_seq).max() < 1e-15
