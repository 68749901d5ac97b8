"""
Driver base class: the API shell of the reference's ``MonteCarloFlow``
(src/vegasflow/monte_carlo.py:95-741; citations relative to /root/reference).

What stays in Python: integrand registration and signature adaptation
(``compile``), the per-iteration loop, the inverse-variance combination of
iterations, history/logging, integration limits, seeding.

What moved into CUDA behind the C ABI (include/vegasflow_b200.h): RNG, grid
map, integrand, reductions, histogram and the chunk loop + ``_accumulate``
(monte_carlo.py:72-92, 249-275, 420-480) -- one ``vf_run_event`` call per
iteration per GPU.  With torch.distributed initialised, events are sharded
over ranks and one all-reduce per iteration replaces the joblib device pool
(monte_carlo.py:143-157, 318-365).  dask distribution and differentiable mode
are out of scope.
"""

from abc import ABC, abstractmethod
import inspect
import itertools
import logging
import os
import time

import numpy as np
import torch

from vegasflow_b200 import _lib, parallel
from vegasflow_b200.configflow import (
    DEFAULT_ACTIVE_DEVICES,
    DTYPE,
    DTYPEINT,
    MAX_EVENTS_LIMIT,
    BINS_MAX,
)
from vegasflow_b200.integrands import BuiltinIntegrand, resolve

logger = logging.getLogger(__name__)

# Every integrator instance draws from its own Philox key so that two instances
# are statistically independent (TensorFlow's global generator keeps advancing
# across instances in the reference).
_instance_counter = itertools.count()
_BASE_SEED = int(os.environ.get("VEGASFLOW_SEED", "0"))


def print_iteration(it, res, error, extra="", threshold=0.1):
    """monte_carlo.py:61-69."""
    if res < threshold:
        return f"Result for iteration {it}: {res:.3e} +/- {error:.3e}" + extra
    return f"Result for iteration {it}: {res:.4f} +/- {error:.4f}" + extra


class MonteCarloFlow(ABC):
    """
    Parent class of the Monte Carlo integrators (monte_carlo.py:95-175).

    Parameters
    ----------
        `n_dim`: number of dimensions of the integrand
        `n_events`: number of events per iteration
        `events_limit`: maximum number of events per step of the UNFUSED path
            (python-callable integrands), to bound the memory of x[n, d].
            The fused path never materialises per-event data and runs each
            iteration in one launch regardless of this value.
        `list_devices`: accepted for compatibility; one process drives one GPU
        `xmin`, `xmax`: integration limits
        `rng_bits`: 52 (default) or 32 bits of Philox output per uniform draw
    """

    _CAN_RUN_VECTORIAL = False
    _ROW = 2  # doubles per result row: (res, sigma)
    _MODE = _lib.MODE_PLAIN
    # True when whole iterations can be enqueued by vf_run_iterations (single rank, fused)
    _BATCHABLE = False

    def __init__(
        self,
        n_dim,
        n_events,
        events_limit=MAX_EVENTS_LIMIT,
        list_devices=DEFAULT_ACTIVE_DEVICES,  # pylint: disable=dangerous-default-value
        verbose=True,
        xmin=None,
        xmax=None,
        rng_bits=None,
        **kwargs,
    ):
        if "simplify_signature" in kwargs:
            logger.warning("simplify_signature is deprecated and will be removed")
        self.n_dim = int(n_dim)
        self._integrand = None
        self._builtin = None
        self._torch_integrand = None
        self.event = None
        self._verbose = verbose
        self._history = []
        self._n_events = int(n_events)
        self._events_limit = int(events_limit)
        self._events_per_run = min(self._events_limit, self._n_events)
        self._compilation_arguments = None
        self._vectorial = False
        self.distribute = False
        self._pass_weight = False
        self.devices = None

        # monte_carlo.py:159-175
        if xmin is not None or xmax is not None:
            if xmin is None or xmax is None:
                raise ValueError(
                    "Both xmin and xmax must be provided if the integration limits are to change"
                )
            if not (len(xmin) == len(xmax) == n_dim):
                raise ValueError("The integration limits must be given for all dimensions")
            self._xmin = np.asarray(xmin, dtype=np.float64)
            self._xdelta = np.asarray(xmax, dtype=np.float64) - self._xmin
            if any(self._xdelta < 0.0):
                raise ValueError(f"No xmin ({xmin}) can be bigger than xmax ({xmax})")
            jac = self._xdelta[0]
            for v in self._xdelta[1:]:
                jac = jac * v
            self._xdeltajac = float(jac)
        else:
            self._xmin = None
            self._xdelta = None
            self._xdeltajac = None
        self._xmin_c = _lib.host_doubles(self._xmin)
        self._xdelta_c = _lib.host_doubles(self._xdelta)

        # Random stream.  rng_bits = 52 (default): two Philox words per uniform, all 52 mantissa
        # bits like tf.random.uniform; 32: one word per uniform (half the integer work).
        if rng_bits is None:
            rng_bits = int(os.environ.get("VEGASFLOW_B200_RNG_BITS", "52"))
        if rng_bits not in (52, 32):
            raise ValueError(f"rng_bits must be 52 or 32, got {rng_bits}")
        self._rng_bits = rng_bits
        # Philox key + iteration counter
        self._seed = (_BASE_SEED + next(_instance_counter)) & 0xFFFFFFFFFFFFFFFF
        self._iteration = 0
        # Device state is allocated lazily (construction works without a GPU so
        # argument validation can be tested on CPU; computing does not).
        self._device = None
        self._workspace = None
        self._packed = None
        self._results = None
        self._exchange = None
        self._exchange_tried = False
        self._last_row = None

    # ------------------------------------------------------------------ state
    def _ensure_device(self):
        if self._device is not None:
            return
        _lib.require_cuda()
        self._device = torch.device("cuda", torch.cuda.current_device())
        nbytes = _lib.load().vf_workspace_bytes(self.n_dim)
        # zero-initialised once: the library keeps the histogram accumulator zeroed afterwards
        self._workspace = torch.zeros(nbytes // 8, dtype=DTYPE, device=self._device)
        # packed per-iteration buffer: histogram [n_dim*50] then (sum wf, sum wf^2)
        self._packed = torch.zeros(self.n_dim * BINS_MAX + 2, dtype=DTYPE, device=self._device)
        self._results = torch.zeros((64, self._ROW), dtype=DTYPE, device=self._device)
        self._results_used = 0
        # page-locked, device-mapped mirror of the rows of the LAST batched call: the tail
        # kernel of every iteration stores its row there, the host reads it after ONE stream
        # synchronisation (no blocking device->host copy per iteration)
        self._host_rows = torch.zeros((64, self._ROW), dtype=DTYPE).pin_memory()
        self._host_rows_n = 0

    def _result_rows(self, n):
        """`n` consecutive device rows that receive (res, sigma[, ...]) of the next iterations."""
        if self._results_used + n > self._results.shape[0]:
            size = max(2 * self._results.shape[0], self._results_used + n)
            grown = torch.zeros((size, self._ROW), dtype=DTYPE, device=self._device)
            grown[: self._results.shape[0]] = self._results
            self._results = grown
        rows = self._results[self._results_used : self._results_used + n]
        self._results_used += n
        self._last_row = rows[n - 1]
        return rows

    def _host_ring(self, n):
        """Pinned host rows for the next `n` iterations of a batched call."""
        if n > self._host_rows.shape[0]:
            torch.cuda.current_stream().synchronize()  # nothing in flight may write the old one
            self._host_rows = torch.zeros((max(n, 2 * self._host_rows.shape[0]), self._ROW),
                                          dtype=DTYPE).pin_memory()
        self._host_rows_n = n
        return self._host_rows

    def _fetch_rows(self):
        """Host copies [(res, sigma[, ...])] of the last batched call: one stream sync."""
        torch.cuda.current_stream().synchronize()
        rows = self._host_rows[: self._host_rows_n].tolist()
        if any(r[0] != r[0] for r in rows) and self._exchange is not None:
            self._exchange.check()
        return rows

    def _result_slot(self):
        """Device row that receives (res, sigma) of the next iteration."""
        return self._result_rows(1)[0]

    def _fused_single_rank(self):
        return (self._BATCHABLE and self._builtin is not None and not self._vectorial
                and parallel.world()[1] == 1)

    def _fused_peer_exchange(self):
        """Multi-rank fused iteration over peer memory (None -> NCCL all-reduce path)."""
        if not (self._BATCHABLE and self._builtin is not None and not self._vectorial
                and parallel.world()[1] > 1):
            return None
        if not self._exchange_tried:
            self._ensure_device()
            self._exchange = parallel.make_peer_exchange(self.n_dim, self._device)
            self._exchange_tried = True
        return self._exchange

    def _run_sharded_iterations(self, xchg, n_iter):
        """`n_iter` iterations of this rank with ONE C-ABI call: per iteration the event kernel
        over its shard, then one kernel doing block reduction + NVLink exchange + sigma + refine
        (vf_run_iterations_sharded).  Returns the device rows [(res, sigma)] * n_iter."""
        lib = _lib.load()
        begin, end = parallel.shard_range(self.n_events)
        rows = self._result_rows(n_iter)
        host = self._host_ring(n_iter)
        first_seq = xchg.seq + 1
        xchg.seq += n_iter
        _lib.check(
            lib.vf_run_iterations_sharded(
                self._mode_word, self._builtin.integrand_id(), self.n_dim, begin, end - begin,
                self.n_events, self._seed, self._iteration, n_iter,
                int(bool(getattr(self, "train", False))), _lib.ptr(self._grid_tensor()),
                self._xmin_c, self._xdelta_c, _lib.ptr(self._packed), _lib.ptr(rows),
                _lib.ptr(host), _lib.ptr(self._workspace), self._workspace.numel() * 8, xchg.rank,
                xchg.world_size, xchg.ptrs, first_seq, _lib.stream_ptr(),
            )
        )
        self._iteration += n_iter
        return rows

    def _run_batched(self, n_iter):
        """All `n_iter` iterations enqueued by one call, or None if this configuration has to
        go iteration by iteration (python integrand, VEGAS+, NCCL collective)."""
        if self._fused_single_rank():
            return self._run_fused_iterations(n_iter)
        xchg = self._fused_peer_exchange()
        if xchg is not None:
            return self._run_sharded_iterations(xchg, n_iter)
        return None

    def _run_fused_iterations(self, n_iter):
        """Enqueue `n_iter` whole iterations with ONE C-ABI call (vf_run_iterations): per
        iteration the fused event kernel and the reduce+sigma+refine kernel, no host sync.
        Returns the device rows [(res, sigma)] * n_iter."""
        self._ensure_device()
        lib = _lib.load()
        rows = self._result_rows(n_iter)
        host = self._host_ring(n_iter)
        _lib.check(
            lib.vf_run_iterations(
                self._mode_word, self._builtin.integrand_id(), self.n_dim, self.n_events, self._seed,
                self._iteration, n_iter, int(bool(getattr(self, "train", False))),
                _lib.ptr(self._grid_tensor()), self._xmin_c, self._xdelta_c,
                _lib.ptr(self._packed), _lib.ptr(rows), _lib.ptr(host), _lib.ptr(self._workspace),
                self._workspace.numel() * 8, _lib.stream_ptr(),
            )
        )
        self._iteration += n_iter
        return rows

    @property
    def _mode_word(self):
        """Sampling mode plus stream options, as the C ABI's `mode` argument."""
        return self._MODE | (_lib.MODE_RNG32 if self._rng_bits == 32 else 0)

    @property
    def _hist(self):
        return self._packed[: self.n_dim * BINS_MAX]

    @property
    def _sums(self):
        return self._packed[self.n_dim * BINS_MAX :]

    # --------------------------------------------------------------- properties
    @property
    def n_events(self):
        """Number of events to run in a single iteration"""
        return self._n_events

    @n_events.setter
    def n_events(self, val):
        """monte_carlo.py:187-193"""
        self._n_events = int(val)
        self.events_per_run = self._events_limit
        self._recompile()

    @property
    def events_per_run(self):
        """Events per step of the unfused path (monte_carlo.py:195-199)"""
        return self._events_per_run

    @events_per_run.setter
    def events_per_run(self, val):
        self._events_per_run = min(int(val), self.n_events)
        if self.n_events % self._events_per_run != 0:
            logger.warning(
                f"The number of events per run step {self._events_per_run} doesn't perfectly"
                f"divide the number of events {self.n_events}, which can harm performance"
            )

    @property
    def history(self):
        """List of (result, error, histograms) per iteration (monte_carlo.py:211-222)"""
        return self._history

    @property
    def xjac(self):
        """The default jacobian is 1 / total number of events (monte_carlo.py:224-227)"""
        return 1.0 / self.n_events

    # ----------------------------------------------------------------- sampling
    def generate_random_array(self, n_events, *args):
        """(x[n, n_dim], p(x)) like monte_carlo.py:229-247."""
        rnds, xjac_raw, *_ = self._generate_random_array(n_events, *args)
        self._iteration += 1  # never hand out the same stream twice
        xjac = xjac_raw / (self.xjac * n_events)
        return rnds, xjac

    def _generate_random_array(self, n_events, *args, ev_begin=0, want_ind=True):
        """Unfused sampling through vf_sample (monte_carlo.py:249-275)."""
        self._ensure_device()
        lib = _lib.load()
        n = int(n_events)
        x = torch.empty((n, self.n_dim), dtype=DTYPE, device=self._device)
        w = torch.empty((n,), dtype=DTYPE, device=self._device)
        ind = None
        if want_ind and self._MODE == _lib.MODE_VEGAS:
            ind = torch.empty((n, self.n_dim), dtype=torch.int32, device=self._device)
        _lib.check(
            lib.vf_sample(
                self._mode_word, self.n_dim, ev_begin, n, self.xjac, self._seed, self._iteration,
                _lib.ptr(self._grid_tensor()), self._xmin_c, self._xdelta_c, _lib.ptr(x),
                _lib.ptr(w), _lib.ptr(ind), _lib.stream_ptr(),
            )
        )
        return x, w, ind

    def _grid_tensor(self):
        """Importance-sampling grid for vf_sample / vf_run_event (None for PlainFlow)."""
        return None

    # ----------------------------------------------------------------- abstract
    @abstractmethod
    def _run_iteration(self):
        """Run one iteration; returns device scalars (res, sigma)."""

    @abstractmethod
    def _run_event(self, integrand, ncalls=None):
        """Run one batch of events (monte_carlo.py:283-288)."""

    def _can_run_vectorial(self, expected_shape=None):
        return self._CAN_RUN_VECTORIAL

    # --------------------------------------------------------------- management
    def set_seed(self, seed):
        """Sets the random seed (monte_carlo.py:313-315): Philox key, counters restart."""
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._iteration = 0

    def set_distribute(self, queue_object):
        raise NotImplementedError(
            "dask distribution is out of scope of the B200 engine; launch one process per GPU "
            "with torchrun instead (events are sharded over ranks automatically)"
        )

    def make_differentiable(self):
        raise NotImplementedError(
            "differentiable mode relies on TensorFlow autograph and is not provided by the "
            "fused CUDA path"
        )

    # ---------------------------------------------------------------- run_event
    def run_event(self, tensorize_events=False, **kwargs):
        """
        Run one full iteration's worth of events and return the accumulated
        tuple of `_run_event` (monte_carlo.py:420-480).

        Fused path: ONE launch covers this rank's shard of the iteration (the
        chunk loop and `_accumulate` live inside the kernel); the packed
        buffer is then all-reduced over ranks.
        Unfused path: chunks of `events_per_run` as in the reference.
        """
        out = self._launch_events(**kwargs)
        self._allreduce(out)
        self._iteration += 1
        return out

    def _launch_events(self, **kwargs):
        """Enqueue this rank's share of the iteration's events (no collective)."""
        if not self.event:
            raise RuntimeError("Compile must be ran before running any iterations")
        self._ensure_device()
        begin, end = parallel.shard_range(self.n_events)
        if self._builtin is not None:
            out = self.event(ev_begin=begin, ncalls=end - begin, accumulate=0, **kwargs)
        else:
            out = None
            done = begin
            first = True
            while done < end:
                ncalls = min(end - done, self.events_per_run)
                if self._verbose:
                    pc = (done + ncalls - begin) / max(end - begin, 1) * 100
                    print(f"Events sent to the computing device: {pc:.1f} %", end="\r")
                out = self.event(ev_begin=done, ncalls=ncalls, accumulate=0 if first else 1,
                                 **kwargs)
                first = False
                done += ncalls
        return out

    def _allreduce(self, out):
        """NCCL path: all-reduce what this iteration produced -- the whole packed buffer when a
        histogram was filled, only the two sums otherwise (frozen grid, PlainFlow: the histogram
        section is stale there and must not be multiplied by the world size every iteration)."""
        with_hist = self._MODE == _lib.MODE_VEGAS and bool(getattr(self, "train", False))
        parallel.allreduce_sum_(self._packed if with_hist else self._sums)

    # ------------------------------------------------------------------ compile
    def compile(self, integrand, compilable=True, signature=None, trace=False, check=True):
        """
        Register the integrand (monte_carlo.py:489-636).

        `integrand` is either a built-in handle / name from
        `vegasflow_b200.integrands` (fused CUDA path) or any python callable
        `f(x[n, n_dim], [n_dim], [weight[n]]) -> [n] | [n, k]` working on CUDA
        torch tensors (unfused path).  `compilable`, `signature` and `trace`
        are accepted for compatibility and ignored (nothing is traced).
        """
        kwargs = {"compilable": compilable, "signature": signature, "trace": trace}
        self._compilation_arguments = (integrand, kwargs)
        self._vectorial = False
        builtin = resolve(integrand)
        self._integrand = builtin if builtin is not None else integrand

        if builtin is not None:
            if builtin.fixed_dim is not None and builtin.fixed_dim != self.n_dim:
                raise ValueError(
                    f"The integrand {builtin.name} is {builtin.fixed_dim}-dimensional, "
                    f"the integrator was instantiated with n_dim={self.n_dim}"
                )
            _lib.require_cuda()
            if builtin.supported(self.n_dim):
                self._builtin = builtin
                self._torch_integrand = None

                def fused_event(**kw):
                    return self._run_event(builtin, **kw)

                self.event = fused_event
                return
            logger.warning(
                "No fused kernel for %s with n_dim=%d, using the unfused path",
                builtin.name, self.n_dim,
            )
            integrand = builtin

        self._builtin = None
        target = integrand.__call__ if isinstance(integrand, BuiltinIntegrand) else integrand
        try:
            spec = inspect.getfullargspec(target)
            args = [a for a in spec.args if a != "self"][1:]
        except TypeError:
            args = []

        # monte_carlo.py:583-591
        def new_integrand(xarr, weight=None, **kw):
            if "weight" in args and "n_dim" in args:
                return integrand(xarr, n_dim=self.n_dim, weight=weight, **kw)
            if "weight" in args:
                return integrand(xarr, weight=weight, **kw)
            if "n_dim" in args:
                return integrand(xarr, n_dim=self.n_dim, **kw)
            return integrand(xarr, **kw)

        def batch_events(**kw):
            return self._run_event(new_integrand, **kw)

        self.event = batch_events
        self._torch_integrand = new_integrand

        # monte_carlo.py:602-633: shape check on 23 events, vector detection
        event_size = 23
        if check:
            _lib.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device())
            test_array = torch.rand((event_size, self.n_dim), dtype=DTYPE, device=dev)
            wgt = torch.rand((event_size,), dtype=DTYPE, device=dev)
            res_tmp = torch.as_tensor(new_integrand(test_array, weight=wgt))
            res_shape = tuple(res_tmp.shape)
            expected_shape = (event_size,)
            if len(res_shape) == 2:
                self._vectorial = True
                expected_shape = tuple(res_tmp.reshape(event_size, -1).shape)
                if not self._can_run_vectorial(expected_shape):
                    raise NotImplementedError(
                        f"The {self.__class__.__name__} algorithm does not support vectorial "
                        "integrands"
                    )
            if res_shape != expected_shape:
                error_str = "the shape of the integrand output should be: (n_events,"
                if self._vectorial:
                    error_str += " output_dim,"
                logger.error(f"Wrong integrand output shape, {error_str})")
                raise ValueError(
                    "The integrand is not returning a value per event, expected shape: "
                    f"{expected_shape}, found: {res_shape}"
                )

    def _recompile(self):
        """monte_carlo.py:638-643"""
        if self._compilation_arguments is None:
            raise RuntimeError("recompile was called without ever having called compile")
        self.compile(self._compilation_arguments[0], **self._compilation_arguments[1])

    # ----------------------------------------------------------- run_integration
    def run_integration(self, n_iter, log_time=True, histograms=None):
        """
        Run `n_iter` iterations and combine them by inverse variance
        (monte_carlo.py:645-741).  Returns (final_result, sigma) as floats.

        Iterations are enqueued back to back on the current CUDA stream; with
        `verbose=False` the host synchronises once, after the last iteration.
        With `verbose=True` every iteration is logged as it finishes, like the
        reference (one small device->host read per iteration).
        `histograms`: tuple of torch tensors that a python integrand accumulates into
        (monte_carlo.py:688-729): they are copied and emptied after every iteration and end up
        holding the inverse-variance weighted average over iterations.
        """
        if not self.event:
            raise RuntimeError("Compile must be ran before running any iterations")
        self._ensure_device()
        all_results = []
        first_slot = len(self._history)
        histo_results = []
        rows = None if (self._verbose or histograms) else self._run_batched(n_iter)
        batched = rows is not None
        if batched:
            for k in range(n_iter):
                all_results.append((rows[k, 0], rows[k, 1]))
                self._history.append((rows[k, 0], rows[k, 1], None))
            host_rows = self._fetch_rows()
            self._after_batch(host_rows)
            self._host_rows_n = 0
        for i in range(0 if batched else n_iter):
            start = time.time() if log_time else None
            self._last_row = None
            self._host_rows_n = 0
            res, error = self._run_iteration()
            all_results.append((res, error))
            # monte_carlo.py:688-694: store the user histograms of this iteration and empty them
            hist_copy = None
            if histograms:
                hist_copy = tuple(h.clone() for h in histograms)
                histo_results.append(hist_copy)
                for h in histograms:
                    h.zero_()
            if self._verbose:
                if self._last_row is not None:  # the row already sits in pinned host memory
                    res_h, err_h = self._row_to_host()
                    res, error = res_h, err_h
                    all_results[-1] = (res, error)
                else:
                    res_h, err_h = self._to_host(res), self._to_host(error)
                time_str = f"(took {time.time()-start:.5f} s)" if log_time else ""
                if self._vectorial:
                    all_info = [
                        print_iteration(i, rr, ee, extra=f" [dimension {d}] {time_str}")
                        for d, (rr, ee) in enumerate(zip(res_h, err_h))
                    ]
                else:
                    all_info = [print_iteration(i, res_h, err_h, extra=time_str)]
                logger.info("\n      ".join(all_info))
            self._history.append((res, error, hist_copy))

        # One read-back for everything that is still on the device
        if batched:
            host = [(row[0], row[1]) for row in host_rows]
        else:
            host = [(self._to_host(r), self._to_host(e)) for r, e in all_results]
        for k, (r, e) in enumerate(host):
            self._history[first_slot + k] = (r, e, self._history[first_slot + k][2])

        # monte_carlo.py:713-732
        aux_res = 0.0
        weight_sum = 0.0
        for i, (res, sigma) in enumerate(host):
            wgt_tmp = 1.0 / np.power(sigma, 2)
            aux_res = aux_res + res * wgt_tmp
            weight_sum = weight_sum + wgt_tmp
            if histograms:  # monte_carlo.py:721-725
                for aux_h, curr_h in zip(histograms, histo_results[i]):
                    aux_h.add_(curr_h * float(np.mean(wgt_tmp)))
        if histograms:  # monte_carlo.py:727-729
            for histogram in histograms:
                histogram.div_(float(np.mean(weight_sum)))
        final_result = aux_res / weight_sum
        sigma = np.sqrt(1.0 / weight_sum)
        if self._verbose:
            if self._vectorial:
                final_results = [
                    f"Final results [{dim = }]: {rr:g} +/- {ee:g}"
                    for dim, (rr, ee) in enumerate(zip(final_result, sigma))
                ]
            else:
                final_results = [f" > Final results: {final_result:g} +/- {sigma:g}"]
            logger.info("\n     ".join(final_results))
        if self._vectorial:
            return np.asarray(final_result), np.asarray(sigma)
        return float(final_result), float(sigma)

    def run_iteration(self):
        """One iteration through the public path, results on the host: `(res, sigma)` as floats
        (arrays for vectorial integrands).  One small device->host read, like the per-iteration
        logging of the reference's run_integration (monte_carlo.py:685-710)."""
        if not self.event:
            raise RuntimeError("Compile must be ran before running any iterations")
        self._ensure_device()
        self._last_row = None
        self._host_rows_n = 0
        res, error = self._run_iteration()
        if self._last_row is not None:
            res, error = self._row_to_host()
        else:
            res, error = self._to_host(res), self._to_host(error)
        self._history.append((res, error, None))
        return res, error

    def _row_to_host(self):
        """(res, sigma) of the iteration just enqueued, on the host."""
        if self._host_rows_n:  # batched call: the tail kernel wrote the row to pinned memory
            rows = self._fetch_rows()
            self._after_batch(rows)
            self._host_rows_n = 0
            return rows[-1][0], rows[-1][1]
        res, error = self._last_row[:2].tolist()
        return res, error

    def _after_batch(self, host_rows):
        """Hook: host-side bookkeeping from the rows of a batched call (VEGAS+ event count)."""

    @staticmethod
    def _to_host(v):
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
            return float(v) if v.ndim == 0 else v
        return v


def wrapper(integrator_class, integrand, n_dim, n_iter, total_n_events, compilable=True):
    """Convenience wrapper (monte_carlo.py:744-762)."""
    mc_instance = integrator_class(n_dim, total_n_events)
    mc_instance.compile(integrand, compilable=compilable)
    return mc_instance.run_integration(n_iter)


def sampler(
    integrator_class,
    integrand,
    n_dim,
    total_n_events,
    training_steps=5,
    compilable=True,
    return_class=False,
):
    """Convenience wrapper for sampling random numbers (monte_carlo.py:765-794)."""
    mc_instance = integrator_class(n_dim, total_n_events, verbose=False)
    mc_instance.compile(integrand, compilable=compilable)
    _ = mc_instance.run_integration(training_steps, log_time=False)
    if return_class:
        return mc_instance
    return mc_instance.generate_random_array
