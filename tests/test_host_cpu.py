"""
CPU tests of the host layer: C-ABI surface, argument validation, grid
checkpoint format, VEGAS+ sizing, sharding + all-reduce over gloo (world 2).
No compute call is made (there is no GPU here and no CPU fallback).
"""
import ctypes
import json
import os
import re
import tempfile

import numpy as np
import pytest
import torch

import vegasflow_b200 as vf
from vegasflow_b200 import _lib, parallel
from vegasflow_b200.monte_carlo import print_iteration

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vegasflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vfp?_[a-z0-9_]+)\s*\(", text)))


def test_abi_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/vegasflow_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_abi_host_only_entry_points():
    lib = _lib.load()
    assert lib.vf_version() == 2
    assert lib.vf_integrand_id(b"symgauss") == 0
    assert lib.vf_integrand_id(b"product") == 1
    assert lib.vf_integrand_id(b"drellyan_lo") == 2
    assert lib.vf_integrand_id(b"singletop_lo") == 3
    assert lib.vf_integrand_id(b"nope") < 0
    assert "nope" in _lib.last_error()
    for d in range(1, 21):
        assert lib.vf_supported(0, d) == 1 and lib.vf_supported(1, d) == 1
    assert lib.vf_supported(0, 21) == 0 and lib.vf_supported(1, 0) == 0
    assert lib.vf_supported(2, 4) == 1 and lib.vf_supported(2, 3) == 0
    assert lib.vf_supported(3, 3) == 1 and lib.vf_supported(3, 4) == 0
    # F_alg(symgauss) = 16d+9, F_alg(product) = 13d+4, plus = 17d+10
    assert lib.vf_flops_per_event(1, 0, 4, 0) == 73
    assert lib.vf_flops_per_event(1, 0, 8, 0) == 137
    assert lib.vf_flops_per_event(1, 0, 20, 0) == 329
    assert lib.vf_flops_per_event(1, 1, 8, 0) == 108
    assert lib.vf_flops_per_event(1, 0, 8, 1) == 146
    assert lib.vf_workspace_bytes(8) >= (296 * 2 + 400) * 8
    # argument validation happens before any CUDA call
    assert lib.vf_run_event(1, 0, 0, 0, 10, 1.0, 0, 0, 1, None, None, None, None, None, 0, None, 0,
                            None) == -1
    assert lib.vf_run_event(7, 0, 4, 0, 10, 1.0, 0, 0, 1, None, None, None, None, None, 0, None, 0,
                            None) == -1


@pytest.mark.skipif(not NO_GPU, reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    inst = vf.VegasFlow(4, 1000, verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        inst.compile(vf.integrands.symgauss)
    with pytest.raises(RuntimeError):
        inst.compile(lambda x: x.sum(dim=1))
    with pytest.raises(RuntimeError):
        inst.run_integration(1)


def test_integration_limits_checks():
    """Restates src/vegasflow/tests/test_misc.py:85-97."""
    with pytest.raises(ValueError):
        vf.PlainFlow(1, 10, xmin=[10], xmax=[1])
    with pytest.raises(ValueError):
        vf.PlainFlow(1, 10, xmin=[10])
    with pytest.raises(ValueError):
        vf.PlainFlow(1, 10, xmax=[10])
    with pytest.raises(ValueError):
        vf.PlainFlow(2, 10, xmin=[0], xmax=[1])
    with pytest.raises(ValueError):
        vf.PlainFlow(2, 10, xmin=[0, 1], xmax=[1])
    ok = vf.VegasFlow(2, 10, xmin=[0, 1], xmax=[1, 3])
    assert ok._xdeltajac == 2.0


def test_grid_save_and_load_roundtrip_and_errors():
    """Restates src/vegasflow/tests/test_algs.py:102-148 (host side)."""
    inst = vf.VegasFlow(2, 100, verbose=False)
    assert tuple(inst.divisions.shape) == (2, 51)
    np.testing.assert_array_equal(inst.divisions.numpy()[0], np.linspace(0, 1, 51))
    tmp = tempfile.mktemp()
    grid = np.sort(np.random.default_rng(0).random((2, 51)), axis=1)
    inst.load_grid(numpy_grid=grid)
    np.testing.assert_array_equal(inst.divisions.numpy(), grid)
    inst.save_grid(tmp)
    with open(tmp) as f:
        jd = json.load(f)
    assert set(jd) == {"dimensions", "ALPHA", "BINS", "integrand", "grid"}
    assert jd["dimensions"] == 2 and jd["BINS"] == 51 and jd["ALPHA"] == 1.5
    np.testing.assert_array_equal(np.array(jd["grid"]), grid)
    other = vf.VegasFlow(2, 100, verbose=False)
    other.load_grid(file_name=tmp)
    np.testing.assert_array_equal(other.divisions.numpy(), grid)
    with pytest.raises(ValueError):
        other.load_grid(file_name=tmp, numpy_grid=grid)
    with pytest.raises(ValueError):
        other.load_grid()
    jd["BINS"] = 0
    with open(tmp, "w") as f:
        json.dump(jd, f)
    with pytest.raises(ValueError):
        other.load_grid(file_name=tmp)
    jd["BINS"] = 51
    jd["dimensions"] = -4
    with open(tmp, "w") as f:
        json.dump(jd, f)
    with pytest.raises(ValueError):
        other.load_grid(file_name=tmp)
    with pytest.raises(ValueError):
        other.load_grid(numpy_grid=np.zeros((3, 51)))


def test_events_per_run_and_xjac():
    inst = vf.VegasFlow(3, 2_500_000, verbose=False)
    assert inst.events_per_run == 1_000_000  # MAX_EVENTS_LIMIT, configflow.py:22
    assert inst.xjac == 1.0 / 2_500_000
    small = vf.PlainFlow(3, 1000, verbose=False)
    assert small.events_per_run == 1000
    with pytest.raises(RuntimeError):
        small._recompile()
    with pytest.raises(RuntimeError, match="Compile must be ran"):
        small.run_event()


def test_vegasflowplus_sizes():
    """SURVEY 9.1 table (vflowplus.py:113-139)."""
    p = vf.VegasFlowPlus(2, 10**4, verbose=False)
    assert (p._n_strat, p._n_cubes, p.min_neval_hcube, p.n_events) == (70, 4900, 2, 9800)
    p = vf.VegasFlowPlus(2, 10**4, adaptive=True, verbose=False)
    assert (p._n_strat, p._n_cubes, p.min_neval_hcube, p.n_events) == (50, 2500, 2, 5000)
    p = vf.VegasFlowPlus(8, 10**8, adaptive=True, verbose=False)
    assert (p._n_strat, p._n_cubes, p.min_neval_hcube, p.n_events) == (3, 6561, 7620, 49994820)
    assert p.xjac == 1.0 / 6561
    p = vf.VegasFlowPlus(14, 10**6, adaptive=True, verbose=False)
    assert p._adaptive is False  # vflowplus.py:106-110
    assert int(p.n_ev.sum()) == p.n_events


def test_builtin_dimension_check_and_names():
    assert vf.integrands.resolve("symgauss") is vf.integrands.symgauss
    assert vf.integrands.resolve(lambda x: x) is None
    with pytest.raises(ValueError):
        vf.integrands.resolve("unknown")
    inst = vf.VegasFlow(5, 100, verbose=False)
    with pytest.raises(ValueError):
        inst.compile(vf.integrands.drellyan_lo)


def test_print_iteration_format():
    """monte_carlo.py:61-69."""
    assert print_iteration(0, 0.5, 0.01) == "Result for iteration 0: 0.5000 +/- 0.0100"
    assert print_iteration(3, 0.0039, 1e-6, extra="!") == "Result for iteration 3: 3.900e-03 +/- 1.000e-06!"


def test_configflow_surface():
    assert vf.DTYPE is torch.float64 and vf.DTYPEINT is torch.int32
    assert vf.float_me(1).dtype is torch.float64 and vf.int_me(1.7).dtype is torch.int32
    from vegasflow_b200 import configflow as c

    assert (c.BINS_MAX, c.ALPHA, c.BETA, c.TECH_CUT) == (50, 1.5, 0.75, 1e-8)
    assert c.MAX_EVENTS_LIMIT == 10**6 and c.MAX_NEVAL_HCUBE == 10**4


def test_utils_cpu_tensors():
    """src/vegasflow/tests/test_utils.py restated on CPU torch tensors."""
    from vegasflow_b200.utils import consume_array_into_indices, generate_condition_function

    rng = np.random.default_rng(1)
    vals = rng.random(60)
    idx = rng.integers(0, 7, size=60)
    out = consume_array_into_indices(torch.from_numpy(vals), torch.from_numpy(idx).reshape(-1, 1), 7)
    check = np.zeros(7)
    for v, i in zip(vals, idx):
        check[i] += v
    np.testing.assert_allclose(out.numpy(), check)
    masks = rng.integers(0, 2, size=(4, 15)).astype(bool)
    tmasks = [torch.from_numpy(m) for m in masks]
    m_and, i_and = generate_condition_function(4, "and")(*tmasks)
    np.testing.assert_array_equal(m_and.numpy(), masks.all(axis=0))
    np.testing.assert_array_equal(i_and.numpy(), np.array(masks.all(axis=0).nonzero()).T)
    m_c, _ = generate_condition_function(3, ["and", "or"])(*tmasks[:3])
    np.testing.assert_array_equal(m_c.numpy(), masks[0] & masks[1] | masks[2])
    for bad in [(1, "and"), (5, "bad"), (5, ["or", "and"]), (3, ["or", "bad"])]:
        with pytest.raises(ValueError):
            generate_condition_function(*bad)


def test_shard_range_partitions_exactly():
    for n in (1, 7, 10**6, 10**9 + 7):
        for world in (1, 2, 3, 8):
            ranges = [parallel.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1


def _gloo_worker(rank, world, port, n, d, out_dir):
    import torch.distributed as dist

    from oracle import c_oracle as co
    from oracle import vegas_ref as R

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.world() == (rank, world)
        begin, end = parallel.shard_range(n)
        grid = R.initial_divisions(d)
        # stand-in for the per-rank CUDA launch: the oracle on this rank's event range
        s1, s2, hist = co.run_event(co.MODE_VEGAS, "symgauss", d, begin, end - begin, 1.0 / n, 5, 0,
                                    True, grid)
        packed = torch.from_numpy(np.concatenate([hist.reshape(-1), [s1, s2]]))
        parallel.allreduce_sum_(packed)
        new_grid = co.refine_grid(packed[: d * 50].numpy().reshape(d, 50), grid)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"),
                np.concatenate([packed.numpy(), new_grid.reshape(-1)]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank_gloo(tmp_path):
    """World-size-2 gloo run of the multi-GPU plumbing (shard_range + packed all-reduce):
    both ranks end with the same reduced buffer and the same refined grid, equal to the
    single-rank result over the same Philox event space."""
    import torch.multiprocessing as mp

    from oracle import c_oracle as co
    from oracle import vegas_ref as R

    n, d, world = 40000, 4, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, n, d, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0, r1)  # identical sums -> identical grids on every rank
    grid = R.initial_divisions(d)
    s1, s2, hist = co.run_event(co.MODE_VEGAS, "symgauss", d, 0, n, 1.0 / n, 5, 0, True, grid)
    single = np.concatenate([hist.reshape(-1), [s1, s2]])
    np.testing.assert_allclose(r0[: d * 50 + 2], single, rtol=1e-11)
    np.testing.assert_allclose(r0[d * 50 + 2 :], co.refine_grid(hist, grid).reshape(-1), atol=1e-12)


def test_cube_shard_partitions_cubes_exactly():
    rng = np.random.default_rng(7)
    for n_cubes in (1, 5, 81, 6561):
        n_ev = rng.integers(2, 5000, size=n_cubes)
        off = np.concatenate([[0], np.cumsum(n_ev)])
        for world in (1, 2, 3, 8):
            parts = [parallel.cube_shard(off, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n_cubes
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and a <= b
            # balanced on events to within one cube
            ev = [off[b] - off[a] for a, b in parts]
            assert max(ev) - min(ev) <= 2 * n_ev.max()


def _gloo_plus_worker(rank, world, port, d, n_req, out_dir):
    import torch.distributed as dist

    from oracle import c_oracle as co
    from oracle import vegas_ref as R

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = R.plus_setup(d, n_req, adaptive=True)
        n_ev, grid = st["n_ev"].copy(), R.initial_divisions(d)
        n_cubes = st["n_cubes"]
        out = []
        for it in range(2):
            off = np.concatenate([[0], np.cumsum(n_ev.astype(np.int64))])
            lo, hi = parallel.cube_shard(off)
            mine = np.zeros_like(n_ev)
            mine[lo:hi] = n_ev[lo:hi]
            # stand-in for this rank's CUDA launch: the oracle on the events of its cube range,
            # fed the slice of the GLOBAL Philox stream that belongs to those events
            rnds = co.uniforms(9, it, int(off[lo]), int(off[hi] - off[lo]), d)
            cubes = R.hypercube_coords(st["n_strat"], d)
            # weights divide by the true n_ev of the cube: evaluate with the masked allocation
            # for the event list and the true one for the weights
            ress, _, hist, det = R.plus_run_event(rnds, st["n_strat"], mine, cubes, grid,
                                                  R.symgauss, st["xjac"])
            fn = n_ev.astype(np.float64)
            ress2 = np.zeros(n_cubes)
            np.add.at(ress2, det["segm"], det["wf"] ** 2)
            var = np.where(mine > 0, ress2 * fn - ress * ress, 0.0)
            t_var = torch.from_numpy(var.copy())
            dist.all_reduce(t_var)  # disjoint supports: the sum IS the all-gather
            sig2_part = float(np.sum(np.maximum(var, 0.0)[lo:hi] / (fn[lo:hi] - 1.0)))
            packed = torch.from_numpy(np.concatenate([hist.reshape(-1),
                                                      [float(ress.sum()), sig2_part]]))
            parallel.allreduce_sum_(packed)
            n_ev, _ = R.plus_redistribute(t_var.numpy(), st["min_neval_hcube"], st["init_calls"])
            grid = R.refine_grid(packed[: d * 50].numpy().reshape(d, 50), grid)
            out.append(np.concatenate([packed.numpy()[-2:], n_ev.astype(np.float64),
                                       grid.reshape(-1)]))
        np.save(os.path.join(out_dir, f"plus_rank{rank}.npy"), np.concatenate(out))
    finally:
        dist.destroy_process_group()


def test_two_rank_vegasflowplus_cube_sharding_gloo(tmp_path):
    """World-size-2 gloo run of the VEGAS+ multi-GPU logic (cube_shard + all-reduce of the
    histogram and the partial (res, sigma^2) + all-gather of the per-cube variances + redundant
    redistribute): both ranks end with identical n_ev and grids, equal to the single-rank
    iteration on the same global Philox event space."""
    import torch.multiprocessing as mp

    from oracle import c_oracle as co
    from oracle import vegas_ref as R

    d, n_req, world = 3, 6000, 2
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_gloo_plus_worker, args=(world, port, d, n_req, str(tmp_path)), nprocs=world,
             join=True)
    r0 = np.load(tmp_path / "plus_rank0.npy")
    r1 = np.load(tmp_path / "plus_rank1.npy")
    np.testing.assert_array_equal(r0, r1)
    _, _, results, grid, n_ev = R.plus_integrate(
        R.symgauss, d, n_req, 2, lambda n, dd, iteration=0, offset=0: co.uniforms(9, iteration,
                                                                                  offset, n, dd),
        adaptive=True)
    n_cubes = len(n_ev)
    rec = 2 + n_cubes + d * 51
    last = r0[rec:]
    assert abs(last[0] - results[1][0]) <= 1e-12 * abs(results[1][0])
    assert abs(np.sqrt(last[1]) - results[1][1]) <= 1e-10 * results[1][1]
    np.testing.assert_array_equal(last[2 : 2 + n_cubes].astype(np.int32), n_ev)
    np.testing.assert_allclose(last[2 + n_cubes :], grid.reshape(-1), atol=1e-12)


USER_SYMGAUSS = r"""
__device__ double integrand(const double* x, int n_dim) {
    // examples/simgauss_cffi.py:25-47 restated as a CUDA device function
    const double a = 0.1;
    double pref = 1.0;
    for (int i = 0; i < n_dim; ++i) pref *= 1.0 / a / sqrt(M_PI);
    double coef = 0.0;
    for (int i = 0; i < n_dim; ++i) { const double t = (x[i] - 0.5) / a; coef += t * t; }
    return pref * exp(-coef);
}
"""


def test_user_cuda_integrand_compiles_and_registers():
    """nvcc cross-compiles the module without a GPU; the library loads and registers it."""
    h = vf.integrands.cuda_integrand(USER_SYMGAUSS, 3, name="test_user_symgauss")
    assert os.path.exists(h.module_path)
    iid = h.integrand_id()
    assert iid >= 16 and h.integrand_id() == iid  # registered once
    assert h.supported(3) and not h.supported(4)
    lib = _lib.load()
    assert lib.vf_register_user_integrand(b"/nonexistent/module.so") == -1
    assert "cannot load" in _lib.last_error()
    with pytest.raises(ValueError, match="nvcc failed"):
        vf.integrands.cuda_integrand("this is not CUDA", 2, name="test_bad")
    inst = vf.VegasFlow(4, 1000, verbose=False)
    with pytest.raises(ValueError):
        inst.compile(h)  # 3-dimensional integrand, 4-dimensional integrator


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on the host cores and prints ONE JSON line with the
    driver's keys (the GPU arm is exercised on the B200 box)."""
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "events/s" and d["value"] > 1e4
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"]
