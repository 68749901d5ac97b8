"""
PlainFlow: plain Monte Carlo, API of src/vegasflow/plain.py (citations relative
to /root/reference).  Same fused kernel as VegasFlow with the identity map.
"""
import torch

from vegasflow_b200 import _lib
from vegasflow_b200.configflow import DTYPE
from vegasflow_b200.integrands import BuiltinIntegrand
from vegasflow_b200.monte_carlo import MonteCarloFlow, sampler, wrapper


class PlainFlow(MonteCarloFlow):
    """Simple Monte Carlo integrator (plain.py:10-43)."""

    _CAN_RUN_VECTORIAL = True
    _MODE = _lib.MODE_PLAIN
    _BATCHABLE = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._vec_acc = None

    def _run_event(self, integrand, ncalls=None, ev_begin=0, accumulate=0):
        """plain.py:18-35: returns (res, res2)."""
        self._ensure_device()
        lib = _lib.load()
        n_events = self.n_events if ncalls is None else int(ncalls)
        if isinstance(integrand, BuiltinIntegrand):
            _lib.check(
                lib.vf_run_event(
                    self._mode_word, integrand.integrand_id(), self.n_dim, ev_begin, n_events,
                    self.xjac, self._seed, self._iteration, 0, None, self._xmin_c,
                    self._xdelta_c, _lib.ptr(self._sums), None, accumulate,
                    _lib.ptr(self._workspace), self._workspace.numel() * 8, _lib.stream_ptr(),
                )
            )
            return self._sums[0], self._sums[1]
        rnds, xjac, _ = self._generate_random_array(n_events, ev_begin=ev_begin)
        int_result = torch.as_tensor(integrand(rnds, weight=xjac), dtype=DTYPE,
                                     device=self._device)
        if self._vectorial:
            tmp = int_result * xjac.reshape(-1, 1)  # plain.py:26
            tmp2 = tmp * tmp
            res, res2 = tmp.sum(dim=0), tmp2.sum(dim=0)
            if accumulate and self._vec_acc is not None:
                res, res2 = self._vec_acc[0] + res, self._vec_acc[1] + res2
            self._vec_acc = (res, res2)
            return res, res2
        f_acc = int_result.contiguous()  # bound to a name: must outlive the enqueue below
        _lib.check(
            lib.vf_accumulate(
                self.n_dim, n_events, _lib.ptr(xjac), _lib.ptr(f_acc), None, 0,
                _lib.ptr(self._sums), None, accumulate, _lib.ptr(self._workspace),
                self._workspace.numel() * 8, _lib.stream_ptr(),
            )
        )
        return self._sums[0], self._sums[1]

    def _allreduce(self, out):
        super()._allreduce(out)
        if self._vectorial:
            from vegasflow_b200 import parallel

            parallel.allreduce_sum_(self._vec_acc[0])
            parallel.allreduce_sum_(self._vec_acc[1])

    def _run_iteration(self):
        """plain.py:37-43"""
        rows = self._run_batched(1)
        if rows is not None:
            return rows[0, 0], rows[0, 1]
        self.run_event()
        return self._iteration_epilogue()

    def _iteration_epilogue(self):
        if self._vectorial:
            res, raw_res2 = self._vec_acc
            n = float(self.n_events)
            err_tmp2 = (raw_res2 * n - res * res) / (n - 1.0)
            return res, torch.sqrt(torch.clamp(err_tmp2, min=0.0))
        lib = _lib.load()
        slot = self._result_slot()
        _lib.check(
            lib.vf_iteration_epilogue(
                self.n_dim, self.n_events, 0, _lib.ptr(self._sums), None, None, _lib.ptr(slot),
                _lib.stream_ptr(),
            )
        )
        return slot[0], slot[1]


def plain_wrapper(*args, **kwargs):
    """Wrapper around PlainFlow (plain.py:46-48)"""
    return wrapper(PlainFlow, *args, **kwargs)


def plain_sampler(*args, **kwargs):
    """Wrapper sampler around PlainFlow (plain.py:51-53)"""
    return sampler(PlainFlow, *args, **kwargs)
