"""
GPU tests of the Python API, mirroring src/vegasflow/tests/test_algs.py and
test_misc.py of the reference with torch integrands, plus the built-in (fused)
integrands on independent streams against the oracle.
"""
import json
import tempfile

import numpy as np
import pytest
import torch

import vegasflow_b200 as vf
from oracle import vegas_ref as R
from vegasflow_b200 import PlainFlow, VegasFlow, VegasFlowPlus, plain_sampler, vegas_sampler

pytestmark = pytest.mark.gpu

dim = 2
ncalls = int(1e4)
n_iter = 4


def example_integrand(xarr, weight=None):
    """symgauss written with torch ops (tests/test_algs.py:27-36)."""
    n_dim = xarr.shape[-1]
    a = 0.1
    n100 = 100 * n_dim
    pref = pow(1.0 / a / np.sqrt(np.pi), n_dim)
    coef = float(sum(range(n100 + 1)))
    coef = coef + torch.sum(torch.square((xarr - 1.0 / 2.0) / a), dim=1)
    coef = coef - (n100 + 1) * n100 / 2.0
    return pref * torch.exp(-coef)


def instance_and_compile(Integrator, mode=0, integrand_function=example_integrand):
    if mode == 0:
        integrand = integrand_function
    elif mode == 1:
        def integrand(xarr, n_dim=None):
            return integrand_function(xarr)
    elif mode == 2:
        def integrand(xarr):
            return integrand_function(xarr)
    elif mode == 3:
        def integrand(xarr, n_dim=None, weight=None):
            return integrand_function(xarr, weight=None)
    int_instance = Integrator(dim, ncalls, verbose=False)
    int_instance.set_seed(1000 + 17 * mode + len(Integrator.__name__))  # deterministic streams
    int_instance.compile(integrand)
    return int_instance


def check_is_one(result, sigmas=3, target_result=1.0):
    res = result[0]
    err = np.mean(result[1] * sigmas)
    np.testing.assert_allclose(res, target_result, atol=err)


@pytest.mark.parametrize("mode", range(4))
def test_VegasFlow(mode):
    vegas_instance = instance_and_compile(VegasFlow, mode)
    _ = vegas_instance.run_integration(n_iter)
    vegas_instance.freeze_grid()
    result = vegas_instance.run_integration(n_iter)
    check_is_one(result)


def test_VegasFlow_grid_management():
    vegas_instance = instance_and_compile(VegasFlow, 1)
    _ = vegas_instance.run_integration(n_iter)
    vegas_instance.freeze_grid()
    frozen = vegas_instance.divisions.clone()
    vegas_instance.n_events = 2 * ncalls
    check_is_one(vegas_instance.run_integration(n_iter))
    assert torch.equal(frozen, vegas_instance.divisions)
    vegas_instance.unfreeze_grid()
    check_is_one(vegas_instance.run_integration(n_iter))
    assert not torch.equal(frozen, vegas_instance.divisions)
    vegas_instance.n_events = 3 * ncalls
    check_is_one(vegas_instance.run_integration(n_iter))


def test_VegasFlow_save_and_load_grid():
    tmp_filename = tempfile.mktemp()
    vegas_instance = instance_and_compile(VegasFlow)
    _ = vegas_instance.run_integration(1)
    current_grid = vegas_instance.divisions.cpu().numpy()
    assert not np.array_equal(current_grid, R.initial_divisions(dim))
    vegas_instance.save_grid(tmp_filename)
    with open(tmp_filename, "r") as f:
        json_grid = np.array(json.load(f)["grid"])
    np.testing.assert_equal(current_grid, json_grid)
    tmp_grid = np.sort(np.random.rand(*current_grid.shape), axis=1)
    vegas_instance.load_grid(numpy_grid=tmp_grid)
    np.testing.assert_equal(vegas_instance.divisions.cpu().numpy(), tmp_grid)


@pytest.mark.parametrize("mode", range(4))
def test_PlainFlow(mode):
    plain_instance = instance_and_compile(PlainFlow, mode)
    check_is_one(plain_instance.run_integration(n_iter))


def test_PlainFlow_change_nevents():
    plain_instance = instance_and_compile(PlainFlow, 0)
    check_is_one(plain_instance.run_integration(n_iter))
    plain_instance.n_events = 2 * ncalls
    check_is_one(plain_instance.run_integration(n_iter))


def helper_rng_tester(sampling_function, n_events):
    rnds, px = sampling_function(n_events)
    np.testing.assert_equal(tuple(rnds.shape), (n_events, dim))
    return rnds, px


def test_rng_generation_plain(n_events=100):
    inst = instance_and_compile(PlainFlow)
    _, px = helper_rng_tester(inst.generate_random_array, n_events)
    np.testing.assert_equal(px.cpu().numpy(), 1.0 / n_events)


def test_rng_generation_vegasflow(n_events=100):
    inst = instance_and_compile(VegasFlow)
    inst.run_integration(2)
    a, px = helper_rng_tester(inst.generate_random_array, n_events)
    np.testing.assert_equal(tuple(px.shape), (n_events,))
    b, _ = inst.generate_random_array(n_events)
    assert not torch.equal(a, b)  # successive calls use fresh streams
    # p(x) integrates to one: mean of 1/(n p) ... sum of px over samples of the unit volume
    x, p = inst.generate_random_array(200000)
    assert abs(float(p.sum()) - 1.0) < 0.05


def test_rng_generation_vegasflowplus(n_events=100):
    inst = instance_and_compile(VegasFlowPlus)
    _, px = helper_rng_tester(inst.generate_random_array, n_events)
    np.testing.assert_equal(tuple(px.shape), (n_events,))
    _, px = helper_rng_tester(inst.generate_random_array, 3 * inst.n_events + 17)


def test_rng_generation_wrappers(n_events=100):
    p = plain_sampler(example_integrand, dim, n_events, training_steps=2, return_class=True)
    _ = helper_rng_tester(p.generate_random_array, n_events)
    v = vegas_sampler(example_integrand, dim, n_events, training_steps=2)
    _ = helper_rng_tester(v, n_events)


@pytest.mark.parametrize("mode", range(4))
def test_VegasFlowPlus_default(mode):
    inst = instance_and_compile(VegasFlowPlus, mode)
    check_is_one(inst.run_integration(n_iter))


def test_VegasFlowPlus_adaptive_python_integrand():
    inst = VegasFlowPlus(dim, ncalls, adaptive=True, verbose=False)
    inst.set_seed(79)
    inst.compile(example_integrand)
    check_is_one(inst.run_integration(n_iter))
    assert int(inst.n_ev.sum()) == inst.n_events


# ---- tests/test_misc.py ------------------------------------------------------
def _vector_integrand(xarr, weight=None):
    res = torch.square((xarr - 1.0) ** 2)
    return torch.exp(-res) / 0.845


def _wrong_integrand(xarr):
    return torch.sum(xarr)


def _simple_integrand(xarr):
    return torch.prod(xarr, dim=1)


def _simple_integral(xmin, xmax):
    xm = np.array(xmin) ** 2 / 2.0
    xp = np.array(xmax) ** 2 / 2.0
    return np.prod(xp - xm)


def _wrong_vector_integrand(xarr):
    return xarr.T


@pytest.mark.parametrize("mode", range(4))
@pytest.mark.parametrize("alg", [VegasFlow, PlainFlow])
def test_working_vectorial(alg, mode):
    inst = instance_and_compile(alg, mode=mode, integrand_function=_vector_integrand)
    result = inst.run_integration(2)
    assert result[0].shape == (dim,)
    check_is_one(result, sigmas=5)


def test_notworking_vectorial():
    with pytest.raises(NotImplementedError):
        _ = instance_and_compile(VegasFlowPlus, integrand_function=_vector_integrand)


def test_check_wrong_main_dimension():
    inst = VegasFlow(3, 100, main_dimension=5, verbose=False)
    with pytest.raises(ValueError):
        inst.compile(_vector_integrand)


@pytest.mark.parametrize("wrong_fun", [_wrong_vector_integrand, _wrong_integrand])
def test_wrong_shape(wrong_fun):
    with pytest.raises(ValueError):
        _ = instance_and_compile(PlainFlow, integrand_function=wrong_fun)


@pytest.mark.parametrize("alg", [PlainFlow, VegasFlow, VegasFlowPlus])
@pytest.mark.parametrize("builtin", [False, True])
def test_integration_limits(alg, builtin, ncalls=int(1e4)):
    rng = np.random.default_rng(len(alg.__name__) * 2 + int(builtin))
    dims = int(rng.integers(1, 5))
    xmin = -1.0 + rng.random(dims) * 2.0
    xmax = 3.0 + rng.random(dims)
    inst = alg(dims, ncalls, xmin=xmin, xmax=xmax, verbose=False)
    inst.set_seed(4242 + dims)
    inst.compile(vf.integrands.product if builtin else _simple_integrand)
    result = inst.run_integration(5)
    check_is_one(result, target_result=_simple_integral(xmin, xmax))


# ---- built-in (fused) integrands ---------------------------------------------
@pytest.mark.parametrize("alg", [PlainFlow, VegasFlow, VegasFlowPlus])
def test_fused_symgauss_is_one(alg):
    inst = alg(dim, ncalls, verbose=False)
    inst.set_seed(77)
    inst.compile(vf.integrands.symgauss)
    check_is_one(inst.run_integration(n_iter))
    assert len(inst.history) == n_iter and all(len(h) == 3 for h in inst.history)
    assert all(isinstance(h[0], float) for h in inst.history)


def test_fused_vegasflowplus_adaptive():
    inst = VegasFlowPlus(4, 200000, adaptive=True, verbose=False)
    inst.set_seed(78)
    inst.compile("symgauss")
    before = inst.n_ev.clone().cpu()
    check_is_one(inst.run_integration(5))
    assert not torch.equal(before, inst.n_ev.cpu())  # samples were redistributed
    assert int(inst.n_ev.sum()) == inst.n_events
    assert int(inst.n_ev.min()) >= inst.min_neval_hcube


def test_vegas_wrapper_configs_1_and_2():
    """BASELINE configs[0] and a reduced configs[1] through the convenience wrapper."""
    res, err = vf.vegas_wrapper(vf.integrands.symgauss, 4, 5, int(1e6))
    assert abs(res - 1.0) < 3 * err and err < 5e-4
    res, err = vf.vegas_wrapper(vf.integrands.product, 8, 5, int(1e6))
    assert abs(res - 2.0**-8) < 3 * err and err < 2e-6


def test_independent_stream_agreement_with_oracle():
    """Independent streams: integral within 3 combined sigma of the oracle's, refined grids
    within 3x the oracle-vs-oracle (two seeds) noise floor at the same N."""
    d, n, iters = 4, 200000, 5
    inst = VegasFlow(d, n, verbose=False)
    inst.set_seed(5)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(iters)
    o_res, o_err, _, grid_a = R.vegas_integrate(R.symgauss, d, n, iters, R.uniform_source_numpy(1))
    _, _, _, grid_b = R.vegas_integrate(R.symgauss, d, n, iters, R.uniform_source_numpy(2))
    assert abs(res - o_res) < 3 * np.hypot(err, o_err)
    assert 0.5 < err / o_err < 2.0
    noise = np.abs(grid_a - grid_b).max()
    assert np.abs(inst.divisions.cpu().numpy() - grid_a).max() < 3 * noise + 1e-4


def test_singletop_and_drellyan_fused():
    res, err = vf.vegas_wrapper(vf.integrands.singletop_lo, 3, 5, int(1e6))
    assert abs(res - 423.9) < 4 * err + 0.5
    # Drell-Yan: ln^2-divergent at kappa -> 0: only finiteness/positivity is checked
    inst = VegasFlow(4, int(2e5), verbose=False)
    inst.compile(vf.integrands.drellyan_lo)
    r, e = inst.run_integration(3)
    assert np.isfinite(r) and r > 0 and np.isfinite(e)


def test_seed_reproducibility_and_instance_independence():
    def run(seed):
        inst = VegasFlow(3, 50000, verbose=False)
        if seed is not None:
            inst.set_seed(seed)
        inst.compile(vf.integrands.product)
        return inst.run_integration(3)

    # same seed -> same events.  The scalar reductions run in a fixed order; the histogram bins
    # are summed by shared-memory atomics, so the refined grid (hence later iterations) can
    # differ in the last bits between runs.
    a, b = run(3), run(3)
    assert abs(a[0] - b[0]) <= 1e-10 * abs(a[0]) and abs(a[1] - b[1]) <= 1e-8 * a[1]
    assert run(3) != run(4)
    assert run(None) != run(None)  # two default instances use different Philox keys


def test_run_before_compile_raises():
    inst = VegasFlow(2, 100, verbose=False)
    with pytest.raises(RuntimeError, match="Compile must be ran"):
        inst.run_integration(1)
    with pytest.raises(NotImplementedError):
        inst.make_differentiable()


def test_utils_consume_array_into_indices_gpu():
    """src/vegasflow/tests/test_utils.py:11-30 on CUDA tensors."""
    from vegasflow_b200.utils import consume_array_into_indices, py_consume_array_into_indices

    size_in = np.random.randint(5, 100)
    size_out = np.random.randint(1, size_in - 3)
    input_array = np.random.rand(size_in)
    indices = np.random.randint(0, size_out, size=size_in)
    t_in = torch.from_numpy(input_array).cuda()
    t_idx = torch.from_numpy(indices.reshape(-1, 1)).cuda()
    result = consume_array_into_indices(t_in, t_idx, size_out).cpu().numpy()
    py_result = py_consume_array_into_indices(t_in, t_idx, size_out).cpu().numpy()
    np.testing.assert_allclose(result, py_result, rtol=1e-14)  # index_add order is free
    check_result = np.zeros(size_out)
    for val, i in zip(input_array, indices):
        check_result[i] += val
    np.testing.assert_allclose(check_result, result)


def test_batched_and_verbose_paths_agree():
    """run_integration with verbose=False enqueues all iterations through one
    vf_run_iterations call; verbose=True goes iteration by iteration.  Same seed, same events."""
    out = []
    for verbose in (False, True):
        inst = VegasFlow(4, 100000, verbose=verbose)
        inst.set_seed(99)
        inst.compile(vf.integrands.symgauss)
        out.append(inst.run_integration(4, log_time=False))
        assert len(inst.history) == 4
    assert abs(out[0][0] - out[1][0]) <= 1e-9 * abs(out[0][0])
    assert abs(out[0][1] - out[1][1]) <= 1e-7 * out[0][1]


def _torchrun_dist_check(tmp_path, tag, alg, exchange, port, **extra_env):
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "dist_check.py")
    out = tmp_path / f"dist_{tag}.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, str(out), alg]
    env = dict(os.environ, VEGASFLOW_B200_EXCHANGE=exchange, **extra_env)
    subprocess.run(cmd, check=True, timeout=600, cwd=root, env=env)
    return json.load(open(out))


def test_two_gpu_sharding_matches_single_gpu(tmp_path):
    """torchrun with 2 ranks over NCCL: same seed -> same integral as one rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    inst = VegasFlow(4, 400000, verbose=False)
    inst.set_seed(2718)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(4)
    grid = inst.divisions.cpu().numpy()
    # both collectives: the fused NVLink peer-memory kernel and the NCCL all-reduce path
    for k, exchange in enumerate(("p2p", "nccl")):
        data = _torchrun_dist_check(tmp_path, exchange, "vegas", exchange, 29611 + k)
        assert data["exchange"] == exchange
        assert abs(res - data["res"]) <= 1e-9 * abs(res)
        assert abs(err - data["err"]) <= 1e-7 * err
        np.testing.assert_allclose(grid, np.array(data["grid"]), atol=1e-11)
    # PlainFlow shards the same way (no grid, two sums)
    inst = PlainFlow(4, 400000, verbose=False)
    inst.set_seed(2718)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(4)
    data = _torchrun_dist_check(tmp_path, "plain", "plain", "p2p", 29615)
    assert abs(res - data["res"]) <= 1e-9 * abs(res) and abs(err - data["err"]) <= 1e-7 * err


def test_two_gpu_missing_peer_poisons_instead_of_hanging(tmp_path):
    """A rank that never joins an exchange (crash, mismatched iteration counts): the waiting rank
    gives up after VEGASFLOW_B200_EXCHANGE_TIMEOUT_S, poisons every rank's buffer, returns NaN and
    the host raises -- no hung GPU, no silently wrong sums (exchange_epilogue_kernel)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    data = _torchrun_dist_check(tmp_path, "timeout", "timeout", "p2p", 29619,
                                VEGASFLOW_B200_EXCHANGE_TIMEOUT_S="2")
    assert data["outcome"].startswith("raised"), data
    assert data["poisoned"] == 1 and 1.5 < data["seconds"] < 30.0


def test_two_gpu_vegasflowplus_cube_sharding_matches_single_gpu(tmp_path):
    """VEGAS+ over 2 ranks -- contiguous cube ranges balanced on the event
    prefix sum, all-reduce of the histogram and the partial (res, sigma^2), all-gather of the
    per-cube variances, redundant redistribute_samples.  Must equal the 1-rank run: the Philox
    counters are global event indices, so only the summation order differs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    inst = VegasFlowPlus(4, 400000, adaptive=True, verbose=False)
    inst.set_seed(2718)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(4)
    data = _torchrun_dist_check(tmp_path, "plus", "plus", "p2p", 29617)
    assert data["exchange"] == "p2p"
    assert abs(res - data["res"]) <= 1e-9 * abs(res)
    assert abs(err - data["err"]) <= 1e-7 * err
    np.testing.assert_allclose(inst.divisions.cpu().numpy(), np.array(data["grid"]), atol=1e-11)
    assert data["events_log"] == inst.events_log and data["n_events"] == inst.n_events
    np.testing.assert_array_equal(inst.n_ev.cpu().numpy(), np.array(data["n_ev"]))


@pytest.mark.parametrize("alg", [PlainFlow, VegasFlow, VegasFlowPlus])
def test_rng_bits_32_option(alg):
    """rng_bits=32: one Philox word per uniform; same integrals within errors."""
    inst = alg(4, 200000, verbose=False, rng_bits=32)
    inst.set_seed(5)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(4)
    assert abs(res - 1.0) < 3 * err
    with pytest.raises(ValueError):
        alg(4, 1000, rng_bits=24)


@pytest.mark.parametrize("alg", [PlainFlow, VegasFlow])
def test_user_histograms(alg):
    """examples/histogram_ex.py: the integrand accumulates weight*f into a 2-bin histogram of
    x[:, 2]; run_integration empties it every iteration and leaves the weighted average
    (monte_carlo.py:688-729).  The bins of a symmetric integrand each hold half the integral."""
    from vegasflow_b200.utils import consume_array_into_indices

    d, nbins = 3, 2
    cumul = torch.zeros(nbins, dtype=torch.float64, device="cuda")

    def integrand(xarr, weight=None):
        res = example_integrand(xarr)
        idx = torch.clamp((xarr[:, 2] * nbins).to(torch.int64), 0, nbins - 1)
        cumul.add_(consume_array_into_indices(res * weight, idx.reshape(-1, 1), nbins))
        return res

    inst = alg(d, 100000, verbose=False)
    inst.set_seed(321)
    inst.compile(integrand, check=False)
    res, err = inst.run_integration(4, histograms=(cumul,))
    assert abs(res - 1.0) < 4 * err
    h = cumul.cpu().numpy()
    assert abs(h.sum() - res) < 1e-9 * abs(res)  # histogram entries add up to the integral
    assert abs(h[0] - h[1]) < 8 * err
    assert all(len(entry[2]) == 1 for entry in inst.history)  # per-iteration copies kept


def test_user_cuda_integrand_runs_fused():
    """A user integrand written in CUDA C++ (the native analogue of examples/simgauss_cffi.py and
    examples/cuda/) is compiled into the fused kernels: same events as the built-in symgauss on
    the same seed, so the integrals agree to rounding; the parity entry reproduces the formula."""
    from tests.test_host_cpu import USER_SYMGAUSS
    from vegasflow_b200 import _lib

    d, n = 4, 200000
    user = vf.integrands.cuda_integrand(USER_SYMGAUSS, d, name="test_user_symgauss")
    out = []
    for integrand in (user, vf.integrands.symgauss):
        for alg in (VegasFlow, VegasFlowPlus, PlainFlow):
            inst = alg(d, n, verbose=False)
            inst.set_seed(11)
            inst.compile(integrand)
            out.append(inst.run_integration(3))
    for (ru, eu), (rb, eb) in zip(out[:3], out[3:]):
        assert abs(ru - rb) <= 1e-7 * abs(rb) and abs(eu - eb) <= 1e-5 * eb
        assert abs(ru - 1.0) < 4 * eu
    # per-event values through the parity entry
    lib = _lib.load()
    rng = np.random.default_rng(0)
    r = 1e-8 + rng.random((5000, d)) * (1 - 2e-8)
    grid = R.initial_divisions(d)
    dev = torch.device("cuda")
    t_r, t_g = torch.from_numpy(r).to(dev), torch.from_numpy(grid).to(dev)
    wf = torch.empty(5000, dtype=torch.float64, device=dev)
    _lib.check(lib.vf_digest_from_uniforms(1, user.integrand_id(), d, 5000, _lib.ptr(t_r),
                                           _lib.ptr(t_g), 1.0 / 5000, None, None, None, None, None,
                                           _lib.ptr(wf), _lib.stream_ptr()))
    x, w, _ = R.vegas_digest(r, grid)
    f = (1.0 / 0.1 / np.sqrt(np.pi)) ** d * np.exp(-(((x - 0.5) / 0.1) ** 2).sum(axis=1))
    want = w / 5000 * f
    assert (np.abs(wf.cpu().numpy() - want) / want).max() < 1e-12


@pytest.mark.gpu
def test_integration_md_binding_declarations_drive_the_abi():
    """INTEGRATION.md 1(b) sketches the reference-side ctypes binding.  The reference and a GPU
    never share a machine here, so the sketch cannot run as written -- but its ctypes declarations
    and its call sequences can: they are taken verbatim from the document, pointed at the built
    library, and one VEGAS iteration loop driven through them (vf_run_event -> sigma ->
    vf_refine_grid, exactly the calls of the sketch's `_run_event` / `refine_grid`) must give what
    the shipped host layer gives for the same seed."""
    import ctypes as C
    import os
    import re

    from vegasflow_b200 import _lib as L

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# src/vegasflow/b200.py.*?```", doc, re.S).group(0)
    decl = re.search(r"(_lib\.vf_run_event\.restype.*?)\ndef _check", block, re.S).group(1)
    lib = C.CDLL(L.SO_PATH)  # a fresh handle: only the document's declarations apply to it
    exec(decl, {"C": C, "_lib": lib})
    assert "_lib.vf_run_event(1, self._iid, d, 0, n, 1.0 / self.n_events, self._seed, self._it," in block
    assert "_lib.vf_refine_grid(self.n_dim, self._packed.data_ptr(), self._grid.data_ptr(), None)" in block

    d, n, n_iter, seed = 4, 200000, 3, 4242
    iid = lib.vf_integrand_id(b"symgauss")
    assert iid >= 0
    ws = torch.zeros(lib.vf_workspace_bytes(d) // 8, dtype=torch.float64, device="cuda")
    packed = torch.zeros(d * 50 + 2, dtype=torch.float64, device="cuda")
    grid = torch.as_tensor(R.initial_divisions(d), device="cuda").contiguous()
    got = []
    for it in range(n_iter):  # the sketch's _run_event + refine_grid, argument for argument
        rc = lib.vf_run_event(1, iid, d, 0, n, 1.0 / n, seed, it, 1, grid.data_ptr(), None, None,
                              packed[d * 50:].data_ptr(), packed.data_ptr(), 0, ws.data_ptr(),
                              ws.numel() * 8, None)
        assert rc == 0, lib.vf_last_error().decode()
        torch.cuda.synchronize()
        res, res2 = packed[d * 50].item(), packed[d * 50 + 1].item()
        got.append((res, float(R.vegas_sigma(res, res2, n))))
        assert lib.vf_refine_grid(d, packed.data_ptr(), grid.data_ptr(), None) == 0
    torch.cuda.synchronize()

    inst = VegasFlow(d, n, verbose=False)
    inst.set_seed(seed)
    inst.compile(vf.integrands.symgauss)
    inst.run_integration(n_iter)
    want = [(h[0], h[1]) for h in inst.history[-n_iter:]] if hasattr(inst, "history") else None
    if want:
        for (r, s), (wr, ws_) in zip(got, want):
            assert abs(r - float(wr)) <= 1e-9 * abs(r) and abs(s - float(ws_)) <= 1e-7 * s
    np.testing.assert_allclose(grid.cpu().numpy(), inst.divisions.cpu().numpy(), rtol=0, atol=1e-10)
