"""
ORACLE / CPU BASELINE (test infrastructure, NOT product code).

Op-for-op torch-CPU restatement of the reference's VEGAS iteration *in the
reference's shape*: materialised [n, d] / [d, n] tensors, two gathers, a
reduce_prod, the one-hot histogram (equal -> where -> reduce_sum, O(50*d*n)),
chunks of `events_limit` events accumulated on the host, per-dimension python
refine.  It stands in for "the reference's TensorFlow CPU path" in
`bench.py --impl reference` and in the `cpu_baseline` leg, because TensorFlow
cannot be installed in this image (no wheel, no network).  PARITY UNPINNED
(see oracle/vegas_ref.py).  Citations are file:line relative to /root/reference.
"""
import time

import numpy as np
import torch

from oracle import vegas_ref as R

BINS_MAX = 50
FBINS = 50.0


def importance_sampling_digest(xn, divisions):
    """src/vegasflow/vflow.py:39-83."""
    ind_i = xn.to(torch.int32)
    ind_f = ind_i + 1
    x_ini = torch.gather(divisions, 1, ind_i.long())
    x_fin = torch.gather(divisions, 1, ind_f.long())
    xdelta = x_fin - x_ini
    aux_rand = xn - torch.floor(xn)
    x = x_ini + xdelta * aux_rand
    weights = torch.prod(xdelta * FBINS, dim=0)
    return ind_i.T.contiguous(), x.T.contiguous(), weights


def generate_random_array(rnds, divisions):
    """src/vegasflow/vflow.py:93-126."""
    xn = FBINS * (1.0 - rnds.T)
    ind, x, w = importance_sampling_digest(xn, divisions)
    return x, w, ind


def consume_array_into_indices(input_arr, indices, result_size):
    """src/vegasflow/utils.py:17-44 (one-hot)."""
    all_bins = torch.arange(result_size, dtype=torch.int32)
    eq = torch.eq(indices, all_bins).T
    res_tmp = torch.where(eq, input_arr, torch.zeros((), dtype=input_arr.dtype))
    return res_tmp.sum(dim=1)


def symgauss(xarr):
    """examples/simgauss_tf.py:22-32."""
    n_dim = xarr.shape[-1]
    a = 0.1
    n100 = float(100 * n_dim)
    pref = pow(1.0 / a / np.sqrt(np.pi), n_dim)
    coef = torch.sum(torch.arange(n100 + 1, dtype=torch.float64))
    coef = coef + torch.sum(torch.square((xarr - 1.0 / 2.0) / a), dim=1)
    coef = coef - (n100 + 1) * n100 / 2.0
    return pref * torch.exp(-coef)


def product(xarr):
    """README.md:63-68."""
    return torch.prod(xarr, dim=1)


INTEGRANDS = {"symgauss": symgauss, "product": product}


def run_event(integrand, divisions, ncalls, n_total, gen, train=True):
    """VegasFlow._run_event, src/vegasflow/vflow.py:389-430 + monte_carlo.py:249-275."""
    n_dim = divisions.shape[0]
    rnds = torch.rand((ncalls, n_dim), dtype=torch.float64, generator=gen)
    rnds = rnds * (1.0 - 2 * R.TECH_CUT) + R.TECH_CUT
    x, w, ind = generate_random_array(rnds, divisions)
    xjac = w * (1.0 / n_total)
    tmp = xjac * integrand(x)
    tmp2 = torch.square(tmp)
    res = torch.sum(tmp)
    res2 = torch.sum(tmp2)
    arr_res2 = None
    if train:
        arr_res2 = torch.stack(
            [consume_array_into_indices(tmp2, ind[:, j : j + 1], BINS_MAX) for j in range(n_dim)]
        )
    return res, res2, arr_res2


def iteration(integrand, divisions, n_events, gen, events_limit=R.MAX_EVENTS_LIMIT, train=True):
    """One VegasFlow iteration: run_event chunk loop + _accumulate
    (monte_carlo.py:420-480, 72-92), sigma (vflow.py:437-438), refine (vflow.py:349-362)."""
    n_dim = divisions.shape[0]
    res = 0.0
    res2 = 0.0
    arr = torch.zeros((n_dim, BINS_MAX), dtype=torch.float64)
    left = n_events
    while left > 0:
        ncalls = min(left, events_limit)
        a, b, c = run_event(integrand, divisions, ncalls, n_events, gen, train)
        res, res2 = res + a, res2 + b
        if train:
            arr = arr + c
        left -= ncalls
    res, res2 = float(res), float(res2)
    sigma = float(R.vegas_sigma(res, res2, n_events))
    if train:
        new = R.refine_grid(arr.numpy(), divisions.numpy())
        divisions = torch.from_numpy(new)
    return res, sigma, divisions


def time_iterations(name, n_dim, n_events, n_iter, warmup=1, seed=0, threads=None):
    """Events/s of `n_iter` timed iterations after `warmup` untimed ones."""
    if threads:
        torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(seed)
    div = torch.from_numpy(R.initial_divisions(n_dim))
    f = INTEGRANDS[name]
    results = []
    for _ in range(warmup):
        _, _, div = iteration(f, div, n_events, gen)
    t0 = time.perf_counter()
    for _ in range(n_iter):
        res, sigma, div = iteration(f, div, n_events, gen)
        results.append((res, sigma))
    dt = time.perf_counter() - t0
    return n_events * n_iter / dt, dt, results
