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

    def save_grid(self, file_name):
        """Save the `divisions` array in a json file (vflow.py:271-292, same schema)."""
        div_np = self.divisions.detach().cpu().numpy()
        int_name = self._integrand.__name__ if self._integrand else ""
        json_dict = {
            "dimensions": self.n_dim,
            "ALPHA": ALPHA,
            "BINS": self.grid_bins,
            "integrand": int_name,
            "grid": div_np.tolist(),
        }
        with open(file_name, "w") as f:
            json.dump(json_dict, f, indent=True)

    def load_grid(self, file_name=None, numpy_grid=None):
        """Load the `divisions` array from a json file or a numpy array (vflow.py:294-347)."""
        if file_name is not None and numpy_grid is not None:
            raise ValueError(
                "Received both a numpy grid and a file_name to load the grid from."
                "Ambiguous call to `load_grid`"
            )
        if file_name:
            with open(file_name, "r") as f:
                json_dict = json.load(f)
            grid_dim = json_dict.get("dimensions")
            grid_bins = json_dict.get("BINS")
            if self._integrand:
                integrand_name = self._integrand.__name__
                integrand_grid = json_dict.get("integrand")
                if integrand_name != integrand_grid:
                    logger.warning(
                        f"The grid was written for the integrand: {integrand_grid}"
                        f"which is different from {integrand_name}"
                    )
            numpy_grid = np.array(json_dict["grid"])
        elif numpy_grid is not None:
            grid_dim = numpy_grid.shape[0]
            grid_bins = numpy_grid.shape[1]
        else:
            raise ValueError("load_grid was called but no grid was provided!")
        if grid_dim is not None and self.n_dim != grid_dim:
            raise ValueError(
                f"Received a {grid_dim}-dimensional grid while VegasFlow"
                f"was instantiated with {self.n_dim} dimensions"
            )
        if grid_bins is not None and self.grid_bins != grid_bins:
            raise ValueError(
                f"The received grid contains {grid_bins} bins while the"
                f"current settings is of {self.grid_bins} bins"
            )
        if file_name:
            logger.info(f" > SUCCESS: Loaded grid from {file_name}")
        self._divisions_host = np.ascontiguousarray(numpy_grid, dtype=np.float64)
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
