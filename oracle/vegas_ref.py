"""
ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
N3PDF/vegasflow v1.4.0 VEGAS hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
``vegasflow_b200`` never does.

PARITY UNPINNED.  The reference is Python-on-TensorFlow; TensorFlow is not
installable in this image (no wheel, no network), so the reference itself can
not be run here, and its own tests hold no golden vectors for this path apart
from the histogram scatter (``tests/test_utils.py:11-30``), which is reproduced
in ``tests/test_oracle.py``.  Everything else below is pinned only against
analytic answers (integral of symgauss = 1, product = 2^-d, ...) and against the
independent C restatement in ``oracle/vegas_oracle.c``.

Conventions adopted where TensorFlow/Eigen leaves the result build-dependent
(summation order of ``reduce_sum``/``reduce_prod``, FMA contraction):
  * reductions over the dimension index run sequentially, left to right;
  * every multiply and add is a separately rounded IEEE operation (no FMA);
  * float -> int casts truncate (C semantics), like ``tf.cast``.

Each function cites the reference file:line (relative to /root/reference) that
it restates.  Array layouts follow the reference: ``rnds[n, d]``,
``divisions[d, 51]``.
"""

from itertools import product as _iproduct

import numpy as np

# src/vegasflow/configflow.py:13-24
BINS_MAX = 50
ALPHA = 1.5
BETA = 0.75
TECH_CUT = 1e-8
MAX_EVENTS_LIMIT = int(1e6)
MAX_NEVAL_HCUBE = int(1e4)

FBINS = np.float64(BINS_MAX)


# --------------------------------------------------------------------------
# a5: histogram scatter, src/vegasflow/utils.py:17-44
# --------------------------------------------------------------------------
def consume_array_into_indices(input_arr, indices, result_size):
    """One-hot scatter-add exactly as utils.py:40-43 (equal -> where -> reduce_sum).

    ``indices`` has shape [n, 1] (a column), ``input_arr`` [n].
    The sum over events runs sequentially (np.add.reduce on a strided axis is
    not guaranteed sequential, so use a cumulative loop-free exact alternative:
    ``np.add.at`` adds in index order, which is the sequential order).
    """
    input_arr = np.asarray(input_arr, dtype=np.float64)
    indices = np.asarray(indices).reshape(-1)
    out = np.zeros(int(result_size), dtype=np.float64)
    np.add.at(out, indices, input_arr)
    return out


def consume_array_into_indices_onehot(input_arr, indices, result_size):
    """Literal one-hot form of utils.py:40-43 (O(n*bins)), for small n."""
    input_arr = np.asarray(input_arr, dtype=np.float64)
    all_bins = np.arange(int(result_size), dtype=np.int32)
    eq = np.equal(np.asarray(indices).reshape(-1, 1), all_bins).T  # [bins, n]
    res_tmp = np.where(eq, input_arr, 0.0)
    return res_tmp.sum(axis=1)


# --------------------------------------------------------------------------
# a3: importance_sampling_digest, src/vegasflow/vflow.py:39-83
# --------------------------------------------------------------------------
def importance_sampling_digest(xn, divisions):
    """xn [d, n] -> ind [n, d] int32, x [n, d], weights [n]."""
    ind_i = xn.astype(np.int32)  # vflow.py:67, truncation
    ind_f = ind_i + 1  # :69
    x_ini = np.take_along_axis(divisions, ind_i, axis=1)  # :70 gather(batch_dims=1)
    x_fin = np.take_along_axis(divisions, ind_f, axis=1)  # :71
    xdelta = x_fin - x_ini  # :73
    aux_rand = xn - np.floor(xn)  # :75
    x = x_ini + xdelta * aux_rand  # :76 (mul, then add)
    t = xdelta * FBINS  # :78
    weights = t[0].copy()
    for j in range(1, t.shape[0]):  # reduce_prod, left to right
        weights = weights * t[j]
    return ind_i.T.copy(), x.T.copy(), weights  # :81-83


# a2: _generate_random_array(rnds, divisions), src/vegasflow/vflow.py:93-126
def vegas_digest(rnds, divisions):
    """rnds [n, d] in [TECH_CUT, 1-TECH_CUT) -> x [n,d], w_raw [n], ind [n,d]."""
    xn = FBINS * (1.0 - rnds.T)  # vflow.py:117
    ind, x, w = importance_sampling_digest(xn, divisions)
    return x, w, ind


# a1: MonteCarloFlow._generate_random_array, src/vegasflow/monte_carlo.py:249-275
def apply_jacobians(x, w_raw, xjac, xmin=None, xdelta=None):
    """wgts = wgts_raw * xjac (:270); limits (:271-274)."""
    w = w_raw * np.float64(xjac)
    if xdelta is not None:
        xmin = np.asarray(xmin, dtype=np.float64)
        xdelta = np.asarray(xdelta, dtype=np.float64)
        x = xmin + x * xdelta  # :273
        xdeltajac = xdelta[0]
        for j in range(1, len(xdelta)):  # reduce_prod :171
            xdeltajac = xdeltajac * xdelta[j]
        w = w * xdeltajac  # :274
    return x, w


def initial_divisions(n_dim):
    """src/vegasflow/vflow.py:239-242."""
    sub = np.linspace(0, 1, BINS_MAX + 1)
    return sub.repeat(n_dim).reshape(-1, n_dim).T.copy()


# --------------------------------------------------------------------------
# a4/a5: VegasFlow._run_event, src/vegasflow/vflow.py:389-430, 370-387
# --------------------------------------------------------------------------
def vegas_run_event(rnds, divisions, integrand, n_total, xmin=None, xdelta=None, train=True):
    """Returns (res, res2, arr_res2[d,50]) for one chunk, plus per-event detail."""
    x, w_raw, ind = vegas_digest(rnds, divisions)
    x, w = apply_jacobians(x, w_raw, 1.0 / n_total, xmin, xdelta)
    f = integrand(x)
    tmp = w * f  # vflow.py:416
    tmp2 = tmp * tmp  # :417
    res = seq_sum(tmp)
    res2 = seq_sum(tmp2)
    arr_res2 = None
    if train:
        n_dim = rnds.shape[1]
        arr_res2 = np.stack(
            [consume_array_into_indices(tmp2, ind[:, j : j + 1], BINS_MAX) for j in range(n_dim)]
        )
    return res, res2, arr_res2, dict(x=x, w=w, ind=ind, f=f, wf=tmp)


def seq_sum(a):
    """Sequential-order sum is not what numpy's pairwise sum does; for scalar
    totals of n events the order is a convention anyway -- use math.fsum-free
    pairwise np.sum (accuracy ~1e-16*log n), documented as order-free."""
    return np.sum(a)


# a7: VegasFlow._iteration_content error formula, src/vegasflow/vflow.py:437-438
def vegas_sigma(res, res2, n_events):
    err_tmp2 = (n_events * res2 - res * res) / (n_events - 1.0)
    return np.sqrt(np.maximum(err_tmp2, 0.0))


# a13: PlainFlow, src/vegasflow/plain.py:18-43
def plain_run_event(rnds, integrand, n_total, xmin=None, xdelta=None):
    x, w = apply_jacobians(rnds, np.float64(1.0), 1.0 / n_total, xmin, xdelta)
    w = np.broadcast_to(w, (rnds.shape[0],))
    tmp = integrand(x) * w  # plain.py:26
    tmp2 = tmp * tmp
    return np.sum(tmp), np.sum(tmp2), dict(x=x, w=w, wf=tmp)


def plain_sigma(res, raw_res2, n_events):
    res2 = raw_res2 * n_events  # plain.py:39
    err_tmp2 = (res2 - res * res) / (n_events - 1.0)
    return np.sqrt(np.maximum(err_tmp2, 0.0))


# --------------------------------------------------------------------------
# a8: refine_grid_per_dimension, src/vegasflow/vflow.py:135-211
# --------------------------------------------------------------------------
def refine_grid_per_dimension(t_res_sq, subdivisions):
    t_res_sq = np.asarray(t_res_sq, dtype=np.float64)
    subdivisions = np.asarray(subdivisions, dtype=np.float64)
    meaner = np.full(BINS_MAX, 3.0)
    meaner[0] = 2.0
    meaner[-1] = 2.0  # vflow.py:153-154
    res_padded = np.concatenate([[0.0], t_res_sq, [0.0]])  # :156
    smeared_tmp = (res_padded[1:-1] + res_padded[2:]) + res_padded[:-2]  # :158
    smeared = np.maximum(smeared_tmp / meaner, 1e-30)  # :159
    sum_t = 0.0
    for v in smeared:  # :162, sequential
        sum_t = sum_t + v
    sum_t = np.float64(sum_t)
    log_t = np.log(smeared)  # :163
    aux_t = (1.0 - smeared / sum_t) / (np.log(sum_t) - log_t)  # :164
    wei_t = np.power(aux_t, ALPHA)  # :165
    s = 0.0
    for v in wei_t:
        s = s + v
    ave_t = np.float64(s) / BINS_MAX  # :166

    new_bins = [0.0]  # :195
    bin_weight = np.float64(0.0)
    n_bin = -1
    cur = np.float64(0.0)
    prev = np.float64(0.0)
    for _ in range(BINS_MAX - 1):  # :201
        while bin_weight < ave_t:  # :170-190
            n_bin += 1
            if n_bin > BINS_MAX - 1:  # guard; never hit in practice
                n_bin = BINS_MAX - 1
                break
            bin_weight = bin_weight + wei_t[n_bin]
            prev = cur
            cur = subdivisions[n_bin + 1]
        bin_weight = bin_weight - ave_t  # :205
        delta = (cur - prev) * bin_weight / wei_t[n_bin]  # :206
        new_bins.append(cur - delta)  # :207
    new_bins.append(1.0)  # :208
    return np.array(new_bins, dtype=np.float64)


# VegasFlow.refine_grid, src/vegasflow/vflow.py:349-362
def refine_grid(arr_res2, divisions):
    return np.stack(
        [refine_grid_per_dimension(arr_res2[j], divisions[j]) for j in range(divisions.shape[0])]
    )


# a9: run_integration combination, src/vegasflow/monte_carlo.py:713-732
def combine_iterations(results):
    aux_res = 0.0
    weight_sum = 0.0
    for res, sigma in results:
        wgt_tmp = 1.0 / pow(sigma, 2)
        aux_res += res * wgt_tmp
        weight_sum += wgt_tmp
    return aux_res / weight_sum, np.sqrt(1.0 / weight_sum)


# --------------------------------------------------------------------------
# VEGAS+: src/vegasflow/vflowplus.py
# --------------------------------------------------------------------------
def plus_setup(n_dim, n_events, adaptive=False):
    """VegasFlowPlus.__init__, vflowplus.py:88-142.  n_strat uses float32 pow
    like tf.math.pow on python floats (:113-123)."""
    if n_dim > 13 and adaptive:
        adaptive = False  # :106-110
    if adaptive:
        neval_eff = int(n_events / 2)  # :114
    else:
        neval_eff = n_events  # :117
    n_strat = np.floor(np.power(np.float32(neval_eff / 2), np.float32(1 / n_dim)))  # :115/118
    if np.power(n_strat, np.float32(n_dim)) > MAX_NEVAL_HCUBE:  # :120
        n_strat = np.floor(np.power(np.float32(1e4), np.float32(1 / n_dim)))  # :121
    n_strat = int(n_strat)
    n_cubes = n_strat**n_dim
    min_neval_hcube = max(int(neval_eff // n_cubes), 2)  # :133-134
    n_ev = np.full(n_cubes, min_neval_hcube, dtype=np.int32)  # :136-137
    return dict(
        n_strat=n_strat,
        n_cubes=n_cubes,
        min_neval_hcube=min_neval_hcube,
        n_ev=n_ev,
        n_events=int(n_ev.sum()),  # :138
        xjac=1.0 / n_cubes,  # :139
        adaptive=adaptive,
        init_calls=n_events,  # :103
    )


def hypercube_coords(n_strat, n_dim):
    """vflowplus.py:126-128: itertools.product, dim 0 most significant."""
    return np.array(list(_iproduct(range(n_strat), repeat=n_dim)), dtype=np.int32).reshape(
        -1, n_dim
    )


# a10: generate_samples_in_hypercubes, vflowplus.py:46-80
def plus_digest(rnds, n_strat, n_ev, hypercubes, divisions):
    indices = np.repeat(np.arange(hypercubes.shape[0], dtype=np.int32), n_ev)  # :67
    points = hypercubes[indices].astype(np.float64)  # :68
    n_evs = n_ev[indices].astype(np.float64)  # :69
    xn = (points + rnds).T * FBINS / np.float64(n_strat)  # :72 (mul then div)
    ind, x, weights = importance_sampling_digest(xn, divisions)  # :74
    final_weights = weights / n_evs  # :77
    return x, final_weights, ind, indices


# a11: VegasFlowPlus._run_event, vflowplus.py:187-220
def plus_run_event(rnds, n_strat, n_ev, hypercubes, divisions, integrand, xjac, xmin=None,
                   xdelta=None, train=True):
    x, w_raw, ind, segm = plus_digest(rnds, n_strat, n_ev, hypercubes, divisions)
    x, w = apply_jacobians(x, w_raw, xjac, xmin, xdelta)
    tmp = w * integrand(x)  # :209
    tmp2 = tmp * tmp  # :210
    n_cubes = hypercubes.shape[0]
    ress = np.zeros(n_cubes)
    ress2 = np.zeros(n_cubes)
    np.add.at(ress, segm, tmp)  # segment_sum :213
    np.add.at(ress2, segm, tmp2)  # :214
    fn_ev = n_ev.astype(np.float64)
    arr_var = ress2 * fn_ev - ress * ress  # :216-217
    arr_res2 = None
    if train:
        arr_res2 = np.stack(
            [consume_array_into_indices(tmp2, ind[:, j : j + 1], BINS_MAX)
             for j in range(rnds.shape[1])]
        )
    return ress, arr_var, arr_res2, dict(x=x, w=w, ind=ind, wf=tmp, segm=segm)


# a12: VegasFlowPlus._iteration_content, vflowplus.py:222-242
def plus_result(ress, arr_var, n_ev):
    sigmas2 = np.maximum(arr_var, 0.0)  # :230
    res = np.sum(ress)  # :231
    sigma2 = np.sum(sigmas2 / (n_ev.astype(np.float64) - 1.0))  # :232
    return res, np.sqrt(sigma2)


# a12: redistribute_samples, vflowplus.py:153-163
def plus_redistribute(arr_var, min_neval_hcube, init_calls, clamp=True):
    """``clamp=True`` applies the documented divergence: clamp
    arr_var at 0 before the fractional power (reference yields NaN)."""
    v = np.maximum(arr_var, 0.0) if clamp else arr_var
    damped = np.power(v, BETA / 2)  # :157
    ssum = np.sum(damped)
    new_n_ev = np.maximum(np.float64(min_neval_hcube), damped * init_calls / 2 / ssum)  # :158-161
    n_ev = new_n_ev.astype(np.int32)  # :162 truncation
    return n_ev, int(n_ev.sum())  # :163


# --------------------------------------------------------------------------
# Integrands
# --------------------------------------------------------------------------
def symgauss(xarr):
    """examples/simgauss_tf.py:22-32 (same body tests/test_algs.py:27-36)."""
    n_dim = xarr.shape[-1]
    a = np.float64(0.1)
    n100 = np.float64(100 * n_dim)
    pref = np.power(1.0 / a / np.sqrt(np.pi), np.float64(n_dim))  # :28
    coef = np.float64(np.sum(np.arange(n100 + 1)))  # :29 exact integer
    t = (xarr - 1.0 / 2.0) / a
    sq = t * t
    s = sq[:, 0].copy()
    for j in range(1, n_dim):  # reduce_sum axis=1, left to right
        s = s + sq[:, j]
    coef = coef + s  # :30
    coef = coef - (n100 + 1) * n100 / 2.0  # :31
    return pref * np.exp(-coef)  # :32


def symgauss_constants(n_dim):
    """Host constants (pref, C) of symgauss for a given n_dim."""
    a = np.float64(0.1)
    n100 = np.float64(100 * n_dim)
    pref = np.power(1.0 / a / np.sqrt(np.pi), np.float64(n_dim))
    return float(pref), float((n100 + 1) * n100 / 2.0)


def product(xarr):
    """README.md:63-68, tests/test_misc.py:24-26: reduce_prod(x, axis=1)."""
    p = xarr[:, 0].copy()
    for j in range(1, xarr.shape[1]):
        p = p * xarr[:, j]
    return p


# ---- Drell-Yan LO, examples/drellyan_lo_tf.py ------------------------------
DY_SQRTS = np.float64(14000)  # :18
DY_S = DY_SQRTS * DY_SQRTS  # :22
CONV = np.float64(0.3893793e9)  # :24


def _c(re, im=None):
    if im is None:
        im = np.zeros_like(re)
    return re + 1j * im


def _dy_u0(p, i):
    """drellyan_lo_tf.py:88-132."""
    zeros = np.zeros_like(p[0])
    ones = np.ones_like(p[0])
    rz = p[3] / p[0]
    theta1 = np.where(rz > 0, zeros, rz)
    theta1 = np.where(rz < 0, np.pi * ones, theta1)
    phi1 = zeros
    rrr = np.where(rz < -1, -ones, rz)
    rrr = np.where(rz > 1, ones, rrr)
    theta2 = np.arccos(rrr)
    with np.errstate(divide="ignore", invalid="ignore"):
        rx = p[1] / p[0] / np.sin(theta2)
    rrr = np.where(rx < -1, -ones, rx)
    rrr = np.where(rx > 1, ones, rrr)
    with np.errstate(invalid="ignore"):
        phi2 = np.arccos(rrr)
    ry = p[2] / p[0]
    phi2 = np.where(ry < 0, -phi2, phi2)
    theta = np.where(p[1] == 0, theta1, theta2)
    phi = np.where(p[1] == 0, phi1, phi2)
    prefact = _c(np.sqrt(2) * ones, zeros) * np.sqrt(_c(p[0], zeros))
    cz = _c(zeros, zeros)
    if i == 1:
        a = _c(np.cos(theta / 2), zeros)
        b = _c(np.sin(theta / 2), zeros)
        return [prefact * a, prefact * b * _c(np.cos(phi), np.sin(phi)), cz, cz]
    a = _c(np.sin(theta / 2), zeros)
    b = _c(np.cos(theta / 2), zeros)
    return [cz, cz, prefact * a * _c(np.cos(phi), -np.sin(phi)), -prefact * b]


def _ubar0(p, i):
    """drellyan_lo_tf.py:135-183 == singletop_lo_tf.py:151-197."""
    zeros = np.zeros_like(p[0])
    ones = np.ones_like(p[0])
    rz = p[3] / p[0]
    theta1 = np.where(rz > 0, zeros, rz)
    theta1 = np.where(rz < 0, np.pi * ones, theta1)
    phi1 = zeros
    rrr = rz
    rrr = np.where(rz < -1, -ones, rrr)
    rrr = np.where(rz > 1, ones, rrr)
    theta2 = np.arccos(rrr)
    with np.errstate(divide="ignore", invalid="ignore"):
        rrr = p[1] / p[0] / np.sin(theta2)
    rrr = np.where(rrr < -1, -ones, rrr)
    rrr = np.where(rrr > 1, ones, rrr)
    with np.errstate(invalid="ignore"):
        phi2 = np.arccos(rrr)
    ry = p[2] / p[0]
    phi2 = np.where(ry < 0, -phi2, phi2)
    theta = np.where(p[1] == 0, theta1, theta2)
    phi = np.where(p[1] == 0, phi1, phi2)
    prefact = _c(np.sqrt(2) * ones, zeros) * np.sqrt(_c(p[0], zeros))
    cz = _c(zeros, zeros)
    if i == -1:
        a = _c(np.sin(theta / 2), zeros)
        b = _c(np.abs(np.cos(theta / 2)), zeros)
        return [prefact * a * _c(np.cos(phi), np.sin(phi)), -prefact * b, cz, cz]
    a = _c(np.cos(theta / 2), zeros)
    b = _c(np.sin(theta / 2), zeros)
    return [cz, cz, prefact * a, prefact * b * _c(np.cos(phi), -np.sin(phi))]


def _spinor_sum(bra, ket):
    acc = bra[0] * ket[0]
    for k in range(1, 4):  # reduce_sum axis 0, left to right
        acc = acc + bra[k] * ket[k]
    return acc


def _make_spinor_algebra(u0):
    def za(p1, p2):
        return _spinor_sum(_ubar0(p1, -1), u0(p2, 1))

    def zb(p1, p2):
        return _spinor_sum(_ubar0(p1, 1), u0(p2, -1))

    def sprod(p1, p2):
        return np.real(za(p1, p2) * zb(p2, p1))

    return za, zb, sprod


_dy_za, _dy_zb, _dy_sprod = _make_spinor_algebra(_dy_u0)


def drellyan_lo(xarr):
    """examples/drellyan_lo_tf.py:27-249 (n_dim = 4)."""
    s = DY_S
    kappa = xarr[:, 0]
    y = xarr[:, 1]
    logkappa = np.log(kappa)
    sqrtkappa = np.sqrt(kappa)
    Ycm = np.exp(logkappa * (y - 0.5))
    shat = s
    x1 = sqrtkappa * Ycm
    x2 = sqrtkappa / Ycm
    jac = np.abs(logkappa)
    # make_event :44-75
    mV = np.sqrt(shat * x1 * x2)
    mV2 = mV * mV
    ecmo2 = mV / 2
    zeros = np.zeros_like(ecmo2)
    p0 = [ecmo2, zeros, zeros, ecmo2]
    p1 = [ecmo2, zeros, zeros, -ecmo2]
    pV = [a + b for a, b in zip(p0, p1)]
    YV = 0.5 * np.log(np.abs((pV[0] + pV[3]) / (pV[0] - pV[3])))
    pVt2 = pV[1] * pV[1] + pV[2] * pV[2]
    phi = 2 * np.pi * xarr[:, 3]
    ptmax = 0.5 * mV2 / (np.sqrt(mV2 + pVt2) - (pV[1] * np.cos(phi) + pV[2] * np.sin(phi)))
    pta = ptmax * xarr[:, 2]
    pt = [zeros, pta * np.cos(phi), pta * np.sin(phi), zeros]
    Delta = (mV2 + 2 * (pV[1] * pt[1] + pV[2] * pt[2])) / 2.0 / pta / np.sqrt(mV2 + pVt2)
    yy = YV - np.arccosh(Delta)
    kallenF = 2.0 * ptmax / np.sqrt(mV2 + pVt2) / np.abs(np.sinh(YV - yy))
    p2 = [pta * np.cosh(yy), pta * np.cos(phi), pta * np.sin(phi), pta * np.sinh(yy)]
    p3 = [a - b for a, b in zip(pV, p2)]
    psw = 1 / (8 * np.pi) * kallenF
    psw = psw * jac
    flux = 1 / (2 * mV2)
    # qqxllx(-p1, -p0, p2, p3) :207-224
    q0 = [-c for c in p1]
    q1 = [-c for c in p0]
    lsprod = _dy_sprod(q0, q1)
    a = 2 * np.abs(_dy_za(q0, p2) * _dy_zb(p3, q1)) / lsprod
    b = 2 * np.abs(_dy_za(q0, p3) * _dy_zb(p2, q1)) / lsprod
    wgts = 6.0 * (a * a + b * b) / 36.0
    # luminosity :232-239 (toy pdf = x1*x2, four flavours)
    pdf = x1 * x2
    lumis = (pdf + pdf + pdf + pdf) / x1 / x2
    lumi_me2 = 2 * lumis * wgts
    return lumi_me2 * psw * flux * CONV


def drellyan_closed_form(xarr):
    """Known-answer check: the spinor chain collapses to this."""
    s = DY_S
    x0 = xarr[:, 0]
    x2 = xarr[:, 2]
    return CONV / (2 * np.pi * s) * (2 - x2 * x2) / 3 * x2 / np.sqrt(1 - x2 * x2) * np.abs(
        np.log(x0)) / x0


# ---- single-top LO, examples/singletop_lo_tf.py ----------------------------
ST_MT = np.float64(173.2)
ST_SQRTS = np.float64(8000)
ST_SQRTSMIN = np.float64(173.2)
ST_MW = np.float64(80.419)
ST_GAW = np.float64(2.1054)
ST_GF = np.float64(1.16639e-5)
ST_COLF = np.float64(9)
ST_MT2 = ST_MT * ST_MT
ST_S = ST_SQRTS * ST_SQRTS
ST_SMIN = ST_SQRTSMIN * ST_SQRTSMIN
ST_BMAX = np.sqrt(1 - ST_SMIN / ST_S)
ST_GAW2 = ST_GAW * ST_GAW
ST_MW2 = ST_MW * ST_MW
_g = 4 * np.sqrt(2) * ST_MW2 * ST_GF
ST_GW4 = _g * _g


def _st_u0(p, i):
    """singletop_lo_tf.py:105-148 (phi in {0, pi})."""
    zeros = np.zeros_like(p[0])
    ones = np.ones_like(p[0])
    rz = p[3] / p[0]
    theta1 = np.where(rz > 0, zeros, rz)
    theta1 = np.where(rz < 0, np.pi * ones, theta1)
    phi1 = zeros
    rrr = np.where(rz < -1, -ones, rz)
    rrr = np.where(rz > 1, ones, rrr)
    theta2 = np.arccos(rrr)
    rx = p[1] / p[0]
    phi2 = np.where(rx < 0, np.pi * ones, zeros)
    theta = np.where(p[1] == 0, theta1, theta2)
    phi = np.where(p[1] == 0, phi1, phi2)
    prefact = _c(np.sqrt(2) * ones, zeros) * np.sqrt(_c(p[0], zeros))
    cz = _c(zeros, zeros)
    if i == 1:
        a = _c(np.cos(theta / 2), zeros)
        b = _c(np.sin(theta / 2), zeros)
        return [prefact * a, prefact * b * _c(np.cos(phi), np.sin(phi)), cz, cz]
    a = _c(np.sin(theta / 2), zeros)
    b = _c(np.cos(theta / 2), zeros)
    return [cz, cz, prefact * a * _c(np.cos(phi), -np.sin(phi)), -prefact * b]


_st_za, _st_zb, _st_sprod = _make_spinor_algebra(_st_u0)


def _dot(p1, p2):
    return p1[0] * p2[0] - p1[1] * p2[1] - p1[2] * p2[2] - p1[3] * p2[3]


def _qqxtbx(p0, p1, p2, p3):
    """singletop_lo_tf.py:221-230."""
    pw2 = _st_sprod(p0, p1)
    d0 = pw2 - ST_MW2
    wprop = d0 * d0 + ST_MW2 * ST_GAW2
    a = _st_sprod(p0, p2)
    b = _st_sprod(p0, p3)
    c = _st_sprod(p2, p3)
    d = _st_sprod(p3, p1)
    return np.abs((a + ST_MT2 * b / c) * d) * ST_COLF / wprop * ST_GW4 / 36


def singletop_lo(xarr):
    """examples/singletop_lo_tf.py:45-270 (n_dim = 3)."""
    s = ST_S
    b = ST_BMAX * xarr[:, 0]
    onemb2 = 1 - b * b
    shat = ST_SMIN / onemb2
    tau = shat / s
    ymax = -0.5 * np.log(tau)
    y = ymax * (2 * xarr[:, 1] - 1)
    jac = 2 * tau * b * ST_BMAX / onemb2
    jac = jac * (2 * ymax)
    sqrttau = np.sqrt(tau)
    expy = np.exp(y)
    x1 = sqrttau * expy
    x2 = sqrttau / expy
    # make_event :71-92
    ecmo2 = np.sqrt(shat) / 2
    cc = ecmo2 * (1 - ST_MT2 / shat)
    cos = 1 - 2 * xarr[:, 2]
    sinxi = cc * np.sqrt(1 - cos * cos)
    cosxi = cc * cos
    zeros = np.zeros_like(ecmo2)
    p0 = [ecmo2, zeros, zeros, ecmo2]
    p1 = [ecmo2, zeros, zeros, -ecmo2]
    p2 = [cc, sinxi, zeros, cosxi]
    p3 = [np.sqrt(cc * cc + ST_MT2), -sinxi, zeros, -cosxi]
    psw = (1 - ST_MT2 / shat) / (8 * np.pi)
    psw = psw * jac
    flux = 1 / (2 * shat)
    # evaluate_matrix_element_square :233-247
    k = ST_MT2 / _dot(p3, p0) / 2
    p3 = [a - b_ * k for a, b_ in zip(p3, p0)]
    mp0 = [-c for c in p0]
    mp1 = [-c for c in p1]
    c1 = _qqxtbx(p2, mp1, p3, mp0)
    c2 = _qqxtbx(mp1, p2, p3, mp0)
    # luminosities :254-260
    pdf = x1 * x2
    lumi1 = (pdf + pdf) / x1 / x2
    lumi2 = (pdf + pdf) / x1 / x2
    lumi_me2 = 2 * lumi1 * c1 + 2 * lumi2 * c2  # reduce_sum over 2 channels
    return lumi_me2 * psw * flux * CONV


INTEGRANDS = {
    "symgauss": symgauss,
    "product": product,
    "drellyan_lo": drellyan_lo,
    "singletop_lo": singletop_lo,
}


# --------------------------------------------------------------------------
# Whole-integration drivers on an arbitrary uniform source (statistical checks)
# --------------------------------------------------------------------------
def uniform_source_numpy(seed):
    """Independent stream: numpy PCG64 mapped to [TECH_CUT, 1-TECH_CUT) like
    tf.random.uniform(minval, maxval) (monte_carlo.py:264-266)."""
    rng = np.random.default_rng(seed)

    def draw(n, d, iteration=0, offset=0):
        return TECH_CUT + rng.random((n, d)) * (1.0 - 2 * TECH_CUT)

    return draw


def vegas_integrate(integrand, n_dim, n_events, n_iter, draw, xmin=None, xmax=None, train=True,
                    divisions=None, events_limit=MAX_EVENTS_LIMIT):
    """VegasFlow: run_integration (monte_carlo.py:645-741) over _iteration_content
    (vflow.py:432-442) with run_event chunking (monte_carlo.py:438-480)."""
    divisions = initial_divisions(n_dim) if divisions is None else divisions.copy()
    xdelta = None
    if xmin is not None:
        xmin = np.asarray(xmin, dtype=np.float64)
        xdelta = np.asarray(xmax, dtype=np.float64) - xmin
    results = []
    for it in range(n_iter):
        res = 0.0
        res2 = 0.0
        arr = np.zeros((n_dim, BINS_MAX))
        done = 0
        while done < n_events:
            n = min(events_limit, n_events - done)
            r = draw(n, n_dim, iteration=it, offset=done)
            a, b, c, _ = vegas_run_event(r, divisions, integrand, n_events, xmin, xdelta, train)
            res += a
            res2 += b
            if train:
                arr += c
            done += n
        sigma = vegas_sigma(res, res2, n_events)
        if train:
            divisions = refine_grid(arr, divisions)
        results.append((res, sigma))
    final, err = combine_iterations(results)
    return final, err, results, divisions


def plain_integrate(integrand, n_dim, n_events, n_iter, draw, xmin=None, xmax=None):
    xdelta = None
    if xmin is not None:
        xmin = np.asarray(xmin, dtype=np.float64)
        xdelta = np.asarray(xmax, dtype=np.float64) - xmin
    results = []
    for it in range(n_iter):
        r = draw(n_events, n_dim, iteration=it, offset=0)
        res, res2, _ = plain_run_event(r, integrand, n_events, xmin, xdelta)
        results.append((res, plain_sigma(res, res2, n_events)))
    final, err = combine_iterations(results)
    return final, err, results


def plus_integrate(integrand, n_dim, n_events, n_iter, draw, adaptive=False, xmin=None, xmax=None,
                   train=True):
    st = plus_setup(n_dim, n_events, adaptive)
    cubes = hypercube_coords(st["n_strat"], n_dim)
    divisions = initial_divisions(n_dim)
    n_ev = st["n_ev"]
    xdelta = None
    if xmin is not None:
        xmin = np.asarray(xmin, dtype=np.float64)
        xdelta = np.asarray(xmax, dtype=np.float64) - xmin
    results = []
    for it in range(n_iter):
        n = int(n_ev.sum())
        r = draw(n, n_dim, iteration=it, offset=0)
        ress, arr_var, arr_res2, _ = plus_run_event(
            r, st["n_strat"], n_ev, cubes, divisions, integrand, st["xjac"], xmin, xdelta, train)
        res, sigma = plus_result(ress, arr_var, n_ev)
        if st["adaptive"]:
            n_ev, _ = plus_redistribute(arr_var, st["min_neval_hcube"], st["init_calls"])
        if train:
            divisions = refine_grid(arr_res2, divisions)
        results.append((res, sigma))
    final, err = combine_iterations(results)
    return final, err, results, divisions, n_ev
