"""
VegasFlowPlus: VEGAS+ (adaptive importance + adaptive stratified sampling),
API of src/vegasflow/vflowplus.py (citations relative to /root/reference).

Built-in / CUDA integrands run whole iterations through vfp_run_iterations: the stratified
event kernel (generate_samples_in_hypercubes, segment sums, histogram) followed by ONE tail
kernel (histogram reduction + grid refinement, per-cube variances, iteration result,
redistribute_samples, new event offsets).  The sample allocation n_ev and the event count stay
on the device between iterations -- the reference's `n_events` setter + recompile
(monte_carlo.py:187-193, vflowplus.py:163) has no counterpart and there is no host
synchronisation per iteration.  Under torchrun the cubes are partitioned over the ranks
(the reference is single-device, vflowplus.py:88-100).  Python-callable integrands
go through vfp_run_event / vfp_iteration_epilogue call by call.  Hypercube coordinates are
derived from the cube index inside the kernel (lexicographic, dim 0 most significant,
vflowplus.py:126-128), so no [n_cubes, n_dim] table is stored.
"""
import logging

import numpy as np
import torch

from vegasflow_b200 import _lib, parallel
from vegasflow_b200.configflow import BINS_MAX, DTYPE, MAX_NEVAL_HCUBE
from vegasflow_b200.integrands import BuiltinIntegrand
from vegasflow_b200.monte_carlo import sampler, wrapper
from vegasflow_b200.vflow import VegasFlow

logger = logging.getLogger(__name__)


def _n_strat_for(neval_eff, n_dim):
    """vflowplus.py:113-123; float32 like tf.math.pow on python floats."""
    n_strat = np.floor(np.power(np.float32(neval_eff / 2), np.float32(1 / n_dim)))
    if np.power(n_strat, np.float32(n_dim)) > MAX_NEVAL_HCUBE:
        n_strat = np.floor(np.power(np.float32(1e4), np.float32(1 / n_dim)))
    return int(n_strat)


class VegasFlowPlus(VegasFlow):
    """Implementation of the VEGAS+ algorithm (vflowplus.py:83-247)."""

    _BATCHABLE = True  # the (changing) event count lives on the device: vfp_run_iterations
    _ROW = 3           # result rows: (res, sigma, n_events of the next iteration)

    def __init__(self, n_dim, n_events, train=True, adaptive=False, events_limit=None, **kwargs):
        # vflowplus.py:90-100
        if events_limit is None:
            logger.info("Events per device limit set to %d", n_events)
            events_limit = n_events
        elif events_limit < n_events:
            logger.warning(
                "VegasFlowPlus needs to hold all events in memory at once, "
                "setting the `events_limit` to be equal to `n_events=%d`",
                n_events,
            )
            events_limit = n_events
        super().__init__(n_dim, n_events, train, events_limit=events_limit, **kwargs)
        self._init_calls = int(n_events)

        # vflowplus.py:106-110
        if n_dim > 13 and adaptive:
            self._adaptive = False
            logger.warning("Disabling adaptive mode from VegasFlowPlus, too many dimensions!")
        else:
            self._adaptive = adaptive

        neval_eff = int(self.n_events / 2) if self._adaptive else self.n_events
        self._n_strat = _n_strat_for(neval_eff, n_dim)
        self._n_cubes = self._n_strat ** int(n_dim)
        self.min_neval_hcube = max(int(neval_eff // self._n_cubes), 2)  # :133-134
        self._n_ev_host = np.full(self._n_cubes, self.min_neval_hcube, dtype=np.int32)
        self._n_events = int(self._n_ev_host.sum())  # :138
        self._events_per_run = self._n_events
        self._modified_jac = 1.0 / self._n_cubes  # :139
        self._plus_state = None
        self.events_log = []  # events evaluated by every finished iteration (whole job)
        if self._adaptive:
            logger.warning("Variable number of events requires function signatures all across")

    @property
    def xjac(self):
        """vflowplus.py:144-146"""
        return self._modified_jac

    @property
    def n_ev(self):
        """Samples per hypercube (int32 tensor [n_cubes])."""
        if self._plus_state is not None:
            return self._plus_state["n_ev"]
        return torch.from_numpy(self._n_ev_host)

    # ------------------------------------------------------------ device state
    def _ensure_plus_state(self):
        if self._plus_state is not None:
            return self._plus_state
        self._ensure_device()
        dev = self._device
        n_ev = torch.from_numpy(self._n_ev_host).to(dev)
        off = torch.zeros(self._n_cubes + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(n_ev.to(torch.int64), dim=0)
        self._plus_state = dict(
            n_ev=n_ev,
            ev_offset=off,
            ress=torch.zeros(self._n_cubes, dtype=DTYPE, device=dev),
            ress2=torch.zeros(self._n_cubes, dtype=DTYPE, device=dev),
            arr_var=torch.zeros(self._n_cubes, dtype=DTYPE, device=dev),
            n_events_dev=torch.zeros(1, dtype=torch.int64, device=dev),
        )
        return self._plus_state

    def redistribute_samples(self, arr_var):
        """Recompute the samples per hypercube from the per-cube variances
        (vflowplus.py:153-163) on the device; `arr_var` is clamped at 0 before
        the fractional power (the reference yields NaN for negative round-off)."""
        st = self._ensure_plus_state()
        lib = _lib.load()
        arr_var = torch.as_tensor(arr_var, dtype=DTYPE, device=self._device)
        fn = st["n_ev"].to(DTYPE)
        # feed the epilogue (ress=0, ress2=var/n_ev) so that it reproduces arr_var
        ress = torch.zeros_like(arr_var)
        ress2 = arr_var / fn
        scratch = torch.zeros(2, dtype=DTYPE, device=self._device)
        var_out = torch.empty_like(arr_var)
        _lib.check(
            lib.vfp_iteration_epilogue(
                self._n_cubes, _lib.ptr(ress), _lib.ptr(ress2), 1, self.min_neval_hcube,
                self._init_calls, _lib.ptr(st["n_ev"]), _lib.ptr(st["ev_offset"]),
                _lib.ptr(var_out), _lib.ptr(scratch), _lib.ptr(st["n_events_dev"]),
                _lib.stream_ptr(),
            )
        )
        self._set_n_events_from_device()

    # ------------------------------------------------------- batched iterations
    def _run_batched(self, n_iter):
        """`n_iter` whole iterations with ONE C-ABI call (vfp_run_iterations), the sample
        allocation resident on the device; None for python-callable integrands."""
        if self._builtin is None or self._vectorial:
            return None
        st = self._ensure_plus_state()
        lib = _lib.load()
        rank, world = parallel.world()
        xchg = self._plus_exchange() if world > 1 else None
        if xchg is None:
            if world > 1 and not getattr(self, "_warned_unsharded", False):
                logger.warning("peer memory unavailable: every rank runs the whole VEGAS+ "
                               "iteration (results are identical on all ranks)")
                self._warned_unsharded = True
            rank, world, ptrs, first_seq = 0, 1, None, 1
        else:
            ptrs, first_seq = xchg.ptrs, xchg.seq + 1
            xchg.seq += n_iter
        rows = self._result_rows(n_iter)
        host = self._host_ring(n_iter)
        _lib.check(
            lib.vfp_run_iterations(
                self._builtin.integrand_id(), self.n_dim, self._n_strat, self._n_cubes, self._seed,
                self._iteration, n_iter, self._rng_bits, int(bool(self.train)),
                int(bool(self._adaptive)), self.min_neval_hcube, self._init_calls,
                _lib.ptr(self._grid_tensor()), self._xmin_c, self._xdelta_c, _lib.ptr(st["n_ev"]),
                _lib.ptr(st["ev_offset"]), _lib.ptr(st["ress"]), _lib.ptr(st["ress2"]),
                _lib.ptr(st["arr_var"]), _lib.ptr(self._hist), _lib.ptr(rows), _lib.ptr(host),
                _lib.ptr(self._workspace), self._workspace.numel() * 8, rank, world, ptrs,
                first_seq, _lib.stream_ptr(),
            )
        )
        self._iteration += n_iter
        return rows

    def _plus_exchange(self):
        if not self._exchange_tried:
            self._ensure_device()
            self._exchange = parallel.make_peer_exchange(self.n_dim, self._device,
                                                         n_cubes=self._n_cubes)
            self._exchange_tried = True
        return self._exchange

    def _after_batch(self, host_rows):
        # vflowplus.py:163: the event count of the next iteration, computed on the device
        self.events_log.extend([self._n_events] + [int(r[2]) for r in host_rows[:-1]])
        self._n_events = int(host_rows[-1][2])
        self._events_per_run = self._n_events

    def _set_n_events_from_device(self):
        st = self._plus_state
        self._n_events = int(st["n_events_dev"].item())  # vflowplus.py:163
        self._events_per_run = self._n_events

    # --------------------------------------------------------------- sampling
    def generate_random_array(self, n_events, *args):
        """vflowplus.py:173-185: whole iterations are tiled and cut to `n_events`."""
        st = self._ensure_plus_state()
        lib = _lib.load()
        rnds, wgts = [], []
        for _ in range(n_events // self.n_events + 1):
            n = self.n_events
            u = torch.empty((n, self.n_dim), dtype=DTYPE, device=self._device)
            _lib.check(lib.vf_uniforms(self.n_dim, 0, n, self._seed, self._iteration,
                                       self._rng_bits, _lib.ptr(u), _lib.stream_ptr()))
            self._iteration += 1
            x = torch.empty((n, self.n_dim), dtype=DTYPE, device=self._device)
            w = torch.empty((n,), dtype=DTYPE, device=self._device)
            ress = torch.zeros(self._n_cubes, dtype=DTYPE, device=self._device)
            ress2 = torch.zeros_like(ress)
            self._launch_plus(self._sampling_integrand(), n, 0, ress, ress2, rnds=u, x=x, w=w)
            # monte_carlo.py:244-247 re-weighting
            rnds.append(x)
            wgts.append(w / (self.xjac * n))
        final_r = torch.cat(rnds, dim=0)[:n_events]
        final_w = torch.cat(wgts, dim=0)[:n_events] * self.n_events / n_events
        return final_r, final_w

    def _sampling_integrand(self):
        from vegasflow_b200.integrands import product, symgauss

        # any built-in instantiated for this n_dim works: only x and w are read back
        for cand in (self._builtin, product, symgauss):
            if cand is not None and cand.supported(self.n_dim):
                return cand
        raise ValueError(f"VegasFlowPlus sampling has no kernel for n_dim={self.n_dim}")

    # ------------------------------------------------------------- event step
    def _launch_plus(self, integrand, n_events, train, ress, ress2, rnds=None, x=None, w=None,
                     ind=None, wf=None):
        st = self._ensure_plus_state()
        lib = _lib.load()
        _lib.check(
            lib.vfp_run_event(
                integrand.integrand_id(), self.n_dim, self._n_strat, self._n_cubes, int(n_events),
                _lib.ptr(st["n_ev"]), _lib.ptr(st["ev_offset"]), self.xjac, self._seed,
                self._iteration, self._rng_bits, int(bool(train)), _lib.ptr(self._grid_tensor()),
                self._xmin_c,
                self._xdelta_c, _lib.ptr(ress), _lib.ptr(ress2), _lib.ptr(self._hist), 0,
                _lib.ptr(self._workspace), self._workspace.numel() * 8, _lib.ptr(rnds),
                _lib.ptr(x), _lib.ptr(w), _lib.ptr(ind), _lib.ptr(wf), _lib.stream_ptr(),
            )
        )

    def _run_event(self, integrand, ncalls=None, n_ev=None, ev_begin=0, accumulate=0):
        """One step of VegasFlowPlus (vflowplus.py:187-220): returns
        (ress[n_cubes], arr_var[n_cubes], arr_res2[n_dim, 50])."""
        st = self._ensure_plus_state()
        if not isinstance(integrand, BuiltinIntegrand):
            return self._run_event_unfused(integrand)
        st["ress"].zero_()
        st["ress2"].zero_()
        self._launch_plus(integrand, self.n_events, self.train, st["ress"], st["ress2"])
        return st["ress"], st["ress2"], self._hist.view(self.n_dim, BINS_MAX)

    def _run_event_unfused(self, integrand):
        """Python-callable integrand: sample all events through the parity entry of
        the kernel, evaluate, and segment-sum per cube on the device."""
        st = self._plus_state
        lib = _lib.load()
        n = self.n_events
        u = torch.empty((n, self.n_dim), dtype=DTYPE, device=self._device)
        _lib.check(lib.vf_uniforms(self.n_dim, 0, n, self._seed, self._iteration, self._rng_bits,
                                   _lib.ptr(u), _lib.stream_ptr()))
        x = torch.empty((n, self.n_dim), dtype=DTYPE, device=self._device)
        w = torch.empty((n,), dtype=DTYPE, device=self._device)
        ind = torch.empty((n, self.n_dim), dtype=torch.int32, device=self._device)
        scratch = torch.zeros(2 * self._n_cubes, dtype=DTYPE, device=self._device)
        self._launch_plus(self._sampling_integrand(), n, 0, scratch[: self._n_cubes],
                          scratch[self._n_cubes :], rnds=u, x=x, w=w, ind=ind)
        f = torch.as_tensor(integrand(x, weight=w), dtype=DTYPE, device=self._device).contiguous()
        tmp = w * f  # vflowplus.py:209
        tmp2 = tmp * tmp
        segm = torch.repeat_interleave(
            torch.arange(self._n_cubes, device=self._device), st["n_ev"].to(torch.int64),
            output_size=n)  # vflowplus.py:67
        st["ress"].zero_().index_add_(0, segm, tmp)  # :213
        st["ress2"].zero_().index_add_(0, segm, tmp2)  # :214
        _lib.check(
            lib.vf_accumulate(
                self.n_dim, n, _lib.ptr(w), _lib.ptr(f), _lib.ptr(ind),
                int(bool(self.train)), _lib.ptr(self._sums), _lib.ptr(self._hist), 0,
                _lib.ptr(self._workspace), self._workspace.numel() * 8, _lib.stream_ptr(),
            )
        )
        return st["ress"], st["ress2"], self._hist.view(self.n_dim, BINS_MAX)

    def run_event(self, tensorize_events=None, **kwargs):
        """Call-by-call path (vflowplus.py:244-247): the whole iteration is one launch on this
        rank (cube sharding over ranks happens in the batched path, `_run_batched`)."""
        out = self._launch_events(**kwargs)
        self._iteration += 1
        return out

    def _launch_events(self, **kwargs):
        if not self.event:
            raise RuntimeError("Compile must be ran before running any iterations")
        self._plus_out = self.event(**kwargs)
        return self._plus_out

    def _iteration_content(self):
        """vflowplus.py:222-242"""
        rows = self._run_batched(1)
        if rows is not None:
            return rows[0, 0], rows[0, 1]
        self.run_event()
        return self._iteration_epilogue()

    def _iteration_epilogue(self):
        st = self._ensure_plus_state()
        lib = _lib.load()
        ress, ress2, arr_res2 = self._plus_out
        slot = self._result_slot()
        _lib.check(
            lib.vfp_iteration_epilogue(
                self._n_cubes, _lib.ptr(ress), _lib.ptr(ress2), int(bool(self._adaptive)),
                self.min_neval_hcube, self._init_calls, _lib.ptr(st["n_ev"]),
                _lib.ptr(st["ev_offset"]), _lib.ptr(st["arr_var"]), _lib.ptr(slot),
                _lib.ptr(st["n_events_dev"]), _lib.stream_ptr(),
            )
        )
        if self.train:
            self.refine_grid(arr_res2)
        self.events_log.append(self._n_events)
        if self._adaptive:
            self._set_n_events_from_device()
        return slot[0], slot[1]


def vegasflowplus_wrapper(integrand, n_dim, n_iter, total_n_events, **kwargs):
    """Convenience wrapper (vflowplus.py:250-265)"""
    return wrapper(VegasFlowPlus, integrand, n_dim, n_iter, total_n_events, **kwargs)


def vegasflowplus_sampler(*args, **kwargs):
    """Convenience wrapper for sampling random numbers (vflowplus.py:268-282)"""
    return sampler(VegasFlowPlus, *args, **kwargs)
