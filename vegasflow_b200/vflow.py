"""
VegasFlow: VEGAS importance sampling, API of src/vegasflow/vflow.py:215-481
(citations relative to /root/reference).  The per-event work and the grid
refinement run in libvegasflow_b200.so (vf_run_event, vf_iteration_epilogue).
"""
import json
import logging

import numpy as np
import torch

from vegasflow_b200 import _lib
from vegasflow_b200.configflow import ALPHA, BINS_MAX, DTYPE
from vegasflow_b200.integrands import BuiltinIntegrand
from vegasflow_b200.monte_carlo import MonteCarloFlow, sampler, wrapper

logger = logging.getLogger(__name__)


def refine_grid_per_dimension(t_res_sq, subdivisions):
    """Refine one dimension of the grid on the device (vflow.py:135-211)."""
    lib = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    hist = torch.as_tensor(t_res_sq, dtype=DTYPE, device=dev).reshape(1, BINS_MAX).contiguous()
    div = torch.as_tensor(subdivisions, dtype=DTYPE, device=dev).reshape(1, BINS_MAX + 1).clone()
    _lib.check(lib.vf_refine_grid(1, _lib.ptr(hist), _lib.ptr(div), _lib.stream_ptr()))
    return div.reshape(-1)


class VegasFlow(MonteCarloFlow):
    """
    Importance sampling algorithm from Vegas (vflow.py:215-446).

    Parameters
    ----------
        n_dim: int
            number of dimensions to be integrated
        n_events: int
            number of events per iteration
        train: bool
            whether to train the grid
        main_dimension: int
            in case of vectorial output, main dimension in which to train
    """

    _MODE = _lib.MODE_VEGAS
    _BATCHABLE = True

    def __init__(self, n_dim, n_events, train=True, main_dimension=0, **kwargs):
        super().__init__(n_dim, n_events, **kwargs)
        self.train = train
        # vflow.py:239-242 -- the grid lives on the host until the first device use
        self.grid_bins = BINS_MAX + 1
        subdivision_np = np.linspace(0, 1, self.grid_bins)
        self._divisions_host = subdivision_np.repeat(n_dim).reshape(-1, n_dim).T.copy()
        self._divisions_dev = None
        self._main_dimension = main_dimension
        self._vec_acc = None

    # ------------------------------------------------------------------ grid
    @property
    def divisions(self):
        """Grid tensor [n_dim, 51] (float64, on the GPU once one is in use)."""
        if self._divisions_dev is not None:
            return self._divisions_dev
        return torch.from_numpy(self._divisions_host)

    def _grid_tensor(self):
        if self._divisions_dev is None:
            self._ensure_device()
            self._divisions_dev = torch.from_numpy(self._divisions_host).to(self._device)
        return self._divisions_dev

    def _can_run_vectorial(self, expected_shape):
        # vflow.py:245-252
        if self._main_dimension >= expected_shape[-1]:
            raise ValueError(
                f"The main dimension index ({self._main_dimension}) is greater than the "
                f"dimensionality of the output ({expected_shape[-1]}). "
                "Remember that arrays in python are 0-indexed!"
            )
        return self.__class__.__name__ == "VegasFlow"

    def freeze_grid(self):
        """Stops the grid from refining any more (vflow.py:261-264)"""
        self.train = False
        self._recompile()

    def unfreeze_grid(self):
        """Enable the refining of the grid (vflow.py:266-269)"""
        self.train = True
        self._recompile()

    # -- grid persistence -------------------------------------------------------------------
    # File format of the reference (vflow.py:271-347), kept key for key so that grids written by
    # either implementation load in the other: {"dimensions", "ALPHA", "BINS", "integrand", "grid"}.
    def _integrand_label(self):
        return getattr(self._integrand, "__name__", "") if self._integrand else ""

    def save_grid(self, file_name):
        """Write the grid to `file_name` as json (same schema as the reference, vflow.py:271-292)."""
        payload = dict(dimensions=self.n_dim, ALPHA=ALPHA, BINS=self.grid_bins,
                       integrand=self._integrand_label(),
                       grid=self.divisions.detach().cpu().numpy().tolist())
        with open(file_name, "w") as fh:
            json.dump(payload, fh, indent=True)

    def _check_grid_shape(self, n_dim, n_bins):
        """ValueError on a grid of the wrong shape (the reference's checks, vflow.py:330-341)."""
        for what, got, have in (("dimensions", n_dim, self.n_dim), ("bins", n_bins, self.grid_bins)):
            if got is not None and got != have:
                raise ValueError(f"The grid to load has {got} {what}, this integrator was "
                                 f"instantiated with {have}")

    def load_grid(self, file_name=None, numpy_grid=None):
        """Take the grid from a json file written by `save_grid` or from a `(n_dim, bins)` array
        (vflow.py:294-347: exactly one of the two, shape checked against the instance)."""
        if (file_name is None) == (numpy_grid is None):
            if file_name is None:
                raise ValueError("load_grid was called but no grid was provided!")
            raise ValueError("load_grid takes either a file name or a numpy grid, not both")
        if file_name is not None:
            with open(file_name, "r") as fh:
                stored = json.load(fh)
            grid = np.array(stored["grid"])
            self._check_grid_shape(stored.get("dimensions"), stored.get("BINS"))
            written_for = stored.get("integrand")
            if self._integrand and written_for != self._integrand_label():
                logger.warning(f"The grid was written for the integrand {written_for!r}, "
                               f"the current one is {self._integrand_label()!r}")
            logger.info(f" > SUCCESS: Loaded grid from {file_name}")
        else:
            grid = np.asarray(numpy_grid)
            self._check_grid_shape(grid.shape[0], grid.shape[1])
        self._divisions_host = np.ascontiguousarray(grid, dtype=np.float64)
        if self._divisions_dev is not None:
            self._divisions_dev.copy_(torch.from_numpy(self._divisions_host))

    def refine_grid(self, arr_res2):
        """Refine every dimension from `arr_res2[n_dim, 50]` (vflow.py:349-362): one launch."""
        lib = _lib.require_cuda()
        grid = self._grid_tensor()
        hist = torch.as_tensor(arr_res2, dtype=DTYPE, device=self._device).contiguous()
        _lib.check(lib.vf_refine_grid(self.n_dim, _lib.ptr(hist), _lib.ptr(grid),
                                      _lib.stream_ptr()))

    # ------------------------------------------------------------- event step
    def _run_event(self, integrand, ncalls=None, ev_begin=0, accumulate=0):
        """One step of Vegas (vflow.py:389-430): returns (res, res2, arr_res2)."""
        self._ensure_device()
        lib = _lib.load()
        n_events = self.n_events if ncalls is None else int(ncalls)
        grid = self._grid_tensor()
        hist2d = self._hist.view(self.n_dim, BINS_MAX)
        if isinstance(integrand, BuiltinIntegrand):
            _lib.check(
                lib.vf_run_event(
                    self._mode_word, integrand.integrand_id(), self.n_dim, ev_begin, n_events,
                    self.xjac, self._seed, self._iteration, int(bool(self.train)),
                    _lib.ptr(grid), self._xmin_c, self._xdelta_c, _lib.ptr(self._sums),
                    _lib.ptr(self._hist), accumulate, _lib.ptr(self._workspace),
                    self._workspace.numel() * 8, _lib.stream_ptr(),
                )
            )
            return self._sums[0], self._sums[1], hist2d

        # unfused: sample -> python integrand -> accumulate
        x, xjac, ind = self._generate_random_array(n_events, ev_begin=ev_begin)
        int_result = torch.as_tensor(integrand(x, weight=xjac), dtype=DTYPE, device=self._device)
        if self._vectorial:
            tmp = xjac.reshape(-1, 1) * int_result  # vflow.py:414-417
            tmp2 = tmp * tmp
            res, res2 = tmp.sum(dim=0), tmp2.sum(dim=0)
            if accumulate and self._vec_acc is not None:
                res, res2 = self._vec_acc[0] + res, self._vec_acc[1] + res2
            self._vec_acc = (res, res2)
            f_hist = int_result[:, self._main_dimension].contiguous()  # vflow.py:425-426
        else:
            f_hist = int_result.contiguous()
        _lib.check(
            lib.vf_accumulate(
                self.n_dim, n_events, _lib.ptr(xjac), _lib.ptr(f_hist), _lib.ptr(ind),
                int(bool(self.train)), _lib.ptr(self._sums), _lib.ptr(self._hist), accumulate,
                _lib.ptr(self._workspace), self._workspace.numel() * 8, _lib.stream_ptr(),
            )
        )
        if self._vectorial:
            return self._vec_acc[0], self._vec_acc[1], hist2d
        return self._sums[0], self._sums[1], hist2d

    def _allreduce(self, out):
        super()._allreduce(out)
        if self._vectorial:
            from vegasflow_b200 import parallel

            parallel.allreduce_sum_(self._vec_acc[0])
            parallel.allreduce_sum_(self._vec_acc[1])

    def _iteration_content(self):
        """Steps to follow per iteration (vflow.py:432-442)"""
        rows = self._run_batched(1)
        if rows is not None:
            return rows[0, 0], rows[0, 1]
        self.run_event()
        return self._iteration_epilogue()

    def _iteration_epilogue(self):
        """sigma (vflow.py:437-438) and grid refinement (vflow.py:440-441) of the iteration
        whose reduced sums/histogram sit in the packed buffer."""
        if self._vectorial:
            res, res2 = self._vec_acc
            n = float(self.n_events)
            err_tmp2 = (n * res2 - res * res) / (n - 1.0)
            sigma = torch.sqrt(torch.clamp(err_tmp2, min=0.0))
            if self.train:
                self.refine_grid(self._hist.view(self.n_dim, BINS_MAX))
            return res, sigma
        lib = _lib.load()
        slot = self._result_slot()
        _lib.check(
            lib.vf_iteration_epilogue(
                self.n_dim, self.n_events, int(bool(self.train)), _lib.ptr(self._sums),
                _lib.ptr(self._hist), _lib.ptr(self._grid_tensor()), _lib.ptr(slot),
                _lib.stream_ptr(),
            )
        )
        return slot[0], slot[1]

    def _run_iteration(self):
        """Runs one iteration of the Vegas integrator (vflow.py:444-446)"""
        return self._iteration_content()


def vegas_wrapper(integrand, n_dim, n_iter, total_n_events, **kwargs):
    """Convenience wrapper (vflow.py:449-464): returns (final_result, sigma)."""
    return wrapper(VegasFlow, integrand, n_dim, n_iter, total_n_events, **kwargs)


def vegas_sampler(*args, **kwargs):
    """Convenience wrapper for sampling random numbers (vflow.py:467-481)."""
    return sampler(VegasFlow, *args, **kwargs)
