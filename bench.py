#!/usr/bin/env python
"""
bench.py -- events/s of the fused fp64 VEGAS iteration on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

A "step" is ONE VEGAS iteration with train=True over one batch of N events per
GPU: fused event kernel (Philox -> map -> integrand -> sums + histograms), the
deterministic block reduction, the all-reduce of the packed [d*50+2] buffer
when N>1, and the epilogue (sigma + grid refinement).  Inputs are synthetic by
construction (the uniforms are generated in-kernel from the Philox counter).

Workloads (BASELINE.json configs; events are PER GPU, i.e. weak scaling):
  c1  symgauss d=4,  1e6 events/iter          (configs[0])
  c2  product  d=8,  1e7 events/iter          (configs[1])
  c3  VegasFlowPlus adaptive symgauss d=8, 1e8 events/iter (configs[2], 1 GPU)
  c4dy / c4st  Drell-Yan d=4 / single-top d=3, 1e8 events/iter (configs[3])
  c5  symgauss d=20, 1.25e8 events/iter/GPU   (configs[4]: 1e9 over 8 GPUs)
  sg8 symgauss d=8,  1e8 events/iter          (north_star target line, DEFAULT)

The default line is sg8 -- the workload the north_star's ">= 50 % of fp64 peak" and ">= 7x at 8
GPUs" targets are quoted on (it fits one GPU); the same JSON line carries a `workloads`
sub-table with one short measurement of every other BASELINE.json config.

Prints ONE JSON line on rank 0 (see the driver contract in the task statement).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(alg="vegas", integrand="symgauss", n_dim=4, n_events=10**6,
               name="symgauss d=4, 1e6 events/iter (configs[0])"),
    "c2": dict(alg="vegas", integrand="product", n_dim=8, n_events=10**7,
               name="product x1..x8 d=8, 1e7 events/iter (configs[1])"),
    "c3": dict(alg="plus", integrand="symgauss", n_dim=8, n_events=10**8,
               name="VegasFlowPlus adaptive symgauss d=8, 1e8 events/iter (configs[2])"),
    "c4dy": dict(alg="vegas", integrand="drellyan_lo", n_dim=4, n_events=10**8,
                 name="Drell-Yan LO d=4, 1e8 events/iter (configs[3])"),
    "c4st": dict(alg="vegas", integrand="singletop_lo", n_dim=3, n_events=10**8,
                 name="single-top LO d=3, 1e8 events/iter (configs[3])"),
    "c5": dict(alg="vegas", integrand="symgauss", n_dim=20, n_events=125 * 10**6,
               name="symgauss d=20, 1.25e8 events/iter/GPU = 1e9 over 8 GPUs (configs[4])"),
    "sg8": dict(alg="vegas", integrand="symgauss", n_dim=8, n_events=10**8,
                name="symgauss d=8, 1e8 events/iter (north_star target)"),
    # labelled second line, never the headline: the OPTIONAL stream definition with one 32-bit
    # word per uniform (four uniforms per Philox block instead of two)
    "sg8_rng32": dict(alg="vegas", integrand="symgauss", n_dim=8, n_events=10**8, rng_bits=32,
                      name="symgauss d=8, 1e8 events/iter, OPTIONAL 32-bit stream (rng_bits=32: "
                           "not the default stream, not the headline)"),
}
METRIC = "events/sec (fused fp64 VEGAS iteration)"
UNIT = "events/s"
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz


# ----------------------------------------------------------------------------
# clock sampling during the timed region (NVML; nvidia-smi fallback)
# ----------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index, period=0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # pragma: no cover
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------
# reference arm / CPU baseline (oracle/, test infrastructure)
# ----------------------------------------------------------------------------
def cpu_reference_run(wl, steps, warmup, sample_events):
    """Time the reference-shaped CPU restatement (oracle/ref_shaped_torch.py) of the same
    workload on a bounded sample: `steps` iterations of `sample_events` events."""
    import torch

    from oracle import ref_shaped_torch as T

    if wl["integrand"] not in T.INTEGRANDS or wl["alg"] != "vegas":
        return None
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    torch.set_num_threads(os.cpu_count() or 1)
    evs, dt, _ = T.time_iterations(wl["integrand"], wl["n_dim"], sample_events, steps,
                                   warmup=warmup)
    return dict(value=evs, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} iterations x {sample_events} events/iter of the same workload, "
                       f"reference-shaped torch-CPU restatement (one-hot histogram, 1e6-event "
                       f"chunks), {dt:.1f} s; os.cpu_count()={os.cpu_count()}",
                seconds=dt)


def cpu_best_effort(wl, sample_events):
    """Fused C/OpenMP restatement (oracle/vegas_oracle.c): a non-strawman CPU number."""
    from oracle import c_oracle as co
    from oracle import vegas_ref as R

    if wl["integrand"] not in co.INTEGRAND_IDS or wl["alg"] != "vegas":
        return None
    grid = R.initial_divisions(wl["n_dim"])
    nthreads = os.cpu_count() or 1
    co.run_event(1, wl["integrand"], wl["n_dim"], 0, sample_events // 10, 1.0, 1, 0, True, grid,
                 nthreads=nthreads)
    t0 = time.perf_counter()
    co.run_event(1, wl["integrand"], wl["n_dim"], 0, sample_events, 1.0 / sample_events, 1, 1, True,
                 grid, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return dict(value=sample_events / dt, unit=UNIT, cores=os.cpu_count(), kind="port",
                sample=f"1 fused iteration x {sample_events} events, C/OpenMP restatement, {dt:.2f} s")


def run_reference_arm(args, wl):
    """`--impl reference`: the reference's CPU path for the SAME workload, timed on the host
    cores.  TensorFlow cannot be installed here, so this is the op-for-op reference-shaped
    torch-CPU port (kind "port"); every step is a bounded sample of the workload -- ONE chunk of
    MAX_EVENTS_LIMIT = 1e6 events, the unit the reference itself processes per `self.event` call
    (monte_carlo.py:438-452, configflow.py:22) -- same K and W as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(wl["n_events"], 10**6)
    steps, warmup = max(1, min(args.steps, 50)), max(1, min(args.warmup, 5))
    base = cpu_reference_run(wl, steps, warmup, sample)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable":
                          f"no CPU restatement for workload {args.workload}"}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": base["seconds"] / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"],
                   "sample": f"each step = one {sample}-event chunk of that workload (the "
                             "reference's own MAX_EVENTS_LIMIT chunk); CPU throughput is "
                             "size-independent above 1e6 events",
                   "note": "TensorFlow is not installable here; this is the reference-shaped CPU "
                           "restatement (oracle/ref_shaped_torch.py), all host cores"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def _profiled_traffic(workload):
    """DRAM bytes per event-kernel launch from the committed ncu --set full capture, or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(workload)
            if v is not None:
                return v
        except Exception:
            pass
    return None


RNG_BITS = 52


def make_instance(wl, world):
    import vegasflow_b200 as vf

    # weak scaling: per-GPU events fixed.  VEGAS+ shards cubes: the requested total grows with
    # the world size as well (the stratification is recomputed from it).
    n_total = wl["n_events"] * world
    bits = wl.get("rng_bits", RNG_BITS)
    if wl["alg"] == "vegas":
        inst = vf.VegasFlow(wl["n_dim"], n_total, verbose=False, rng_bits=bits)
    elif wl["alg"] == "plus":
        inst = vf.VegasFlowPlus(wl["n_dim"], n_total, adaptive=True, verbose=False,
                                rng_bits=bits)
    else:
        inst = vf.PlainFlow(wl["n_dim"], n_total, verbose=False, rng_bits=bits)
    inst.set_seed(2024)
    inst.compile(getattr(vf.integrands, wl["integrand"]))
    return inst


class Bench:
    """Shared state of the GPU arm: library handle, process group, barrier."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import __graft_entry__ as graft

        graft.build()
        from vegasflow_b200 import _lib

        self.torch, self.dist, self._lib = torch, dist, _lib
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import datetime

            # short collective timeout: a hang must fail fast, not hold 8 GPUs for 10 minutes
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                    timeout=datetime.timedelta(seconds=90))
        self.lib = _lib.require_cuda()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def agree_max(self, x):
        """Same value on every rank (max), so loop counts derived from it cannot diverge --
        a rank-dependent iteration count would deadlock the per-iteration collective."""
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    @staticmethod
    def run_n(inst, n):
        """n steps of the product path (identical count on every rank); events processed by
        the WHOLE job (all ranks)."""
        if n <= 0:
            return 0
        if inst._run_batched(n) is not None:
            if inst._ROW == 3:  # VEGAS+: the event count changes, read it from the result rows
                inst._after_batch(inst._fetch_rows())
                inst._host_rows_n = 0
                return sum(inst.events_log[-n:])
            return inst.n_events * n
        done = 0
        for _ in range(n):
            done += inst.n_events
            inst._run_iteration()
        return done

    def run_steps(self, inst, n):
        return self.run_n(inst, n)

    def measure(self, wl, K, W, clocks=False):
        """W warm-up steps, ~0.3 s more to ramp the clocks, then EXACTLY K timed steps
        (barrier + synchronize on both sides, CUDA events, max over ranks), then an identical
        K-step pass with per-kernel CUDA events for the roofline."""
        import ctypes

        torch, lib, _lib = self.torch, self.lib, self._lib
        inst = make_instance(wl, self.world)
        self.barrier()
        t0 = time.perf_counter()
        self.run_steps(inst, W)
        self.barrier()
        step_s = self.agree_max((time.perf_counter() - t0) / W)
        self.run_steps(inst, int(min(2000, max(0, 0.3 / max(step_s, 1e-6)))))
        self.barrier()
        repeats = 1
        if K is None:  # sub-table entries: enough steps for ~5 ms of timed work, 10 ... 200, and
            # the MEDIAN of three such timed regions: a 1e6-event workload times ~1.5 ms per
            # region, where one host hiccup while the launches are enqueued is a 10-25 % error
            K = int(min(200, max(10, 5e-3 / max(step_s, 1e-6))))
            repeats = 3

        sampler = ClockSampler(torch.cuda.current_device()) if clocks else None
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        samples = []
        for _ in range(repeats):
            self.barrier()
            lib.vf_launch_count(1)
            if sampler:
                sampler.start()
            ev0.record()
            events_done = self.run_steps(inst, K)
            ev1.record()
            self.barrier()
            launches = int(lib.vf_launch_count(0))
            samples.append((self.agree_max(ev0.elapsed_time(ev1)), events_done))  # max over ranks
        ms, events_done = sorted(samples)[len(samples) // 2]
        # Second, identical K-step pass with the library's CUDA-event bracket around every
        # event-kernel / tail-kernel launch (same stream): per-kernel durations for the
        # roofline.  Kept out of the timed region because the event records break the
        # programmatic-dependent-launch overlap and widen the gaps by a few microseconds.
        lib.vf_kernel_timing(1)
        events_pass2 = self.run_steps(inst, K)
        self.barrier()
        kt, kn = ctypes.c_double(0.0), ctypes.c_int(0)
        _lib.check(lib.vf_kernel_time_ms(ctypes.byref(kt), ctypes.byref(kn)))
        et, en = ctypes.c_double(0.0), ctypes.c_int(0)
        _lib.check(lib.vf_epilogue_time_ms(ctypes.byref(et), ctypes.byref(en)))
        lib.vf_kernel_timing(0)
        kern_ms = kt.value / max(kn.value, 1)
        epi_ms = et.value / max(en.value, 1) if en.value else None
        # a step that produced NaN / Inf is not a measurement: the trained grid is refined from the
        # histogram of every timed iteration, so one non-finite event weight would show up here
        if hasattr(inst, "divisions"):
            g = inst.divisions
            if not (bool(torch.isfinite(g).all()) and bool((g[:, 1:] >= g[:, :-1]).all())):
                raise RuntimeError(f"workload {wl['name']}: the trained grid is not finite/monotone")
        clk = None
        if sampler:
            # keep the same steps running ~1 s more so the sampler sees the kernel under load
            # (count derived from the agreed step time: identical on every rank)
            if ms < 1000.0 and self.agree_max(1.0 if sampler.nv is not None else 0.0) > 0.5:
                self.run_steps(inst, int(min(20000, 1000.0 / max(ms / K, 1e-3))))
                self.barrier()
            clk = sampler.stop()
            clk["note"] = ("NVML samples every 2 ms over the timed region plus a ~1 s repeat of "
                           "the same steps right after it" if ms < 1000.0
                           else "NVML samples over the timed region")
        plus = 1 if wl["alg"] == "plus" else 0
        mode = 0 if wl["alg"] == "plain" else 1
        iid = lib.vf_integrand_id(wl["integrand"].encode())
        f_alg = lib.vf_flops_per_event(mode, iid, wl["n_dim"], plus)
        per_gpu_events = events_done / K / self.world
        per_gpu_events_pass2 = events_pass2 / K / self.world
        achieved = per_gpu_events_pass2 * f_alg / (kern_ms * 1e-3) / 1e12
        return dict(inst=inst, ms=ms, events_done=events_done, launches=launches,
                    kern_ms=kern_ms, kern_n=kn.value, epi_ms=epi_ms, clocks=clk, f_alg=f_alg,
                    per_gpu_events=per_gpu_events, achieved=achieved, plus=plus, steps=K,
                    value=events_done / (ms * 1e-3))

    def measure_e2e(self, wl, K):
        """Public API end to end: grid uploaded from PINNED host memory, run_integration(K)
        (every iteration's (res, sigma) lands in pinned host memory, written by that
        iteration's tail kernel), trained grid downloaded -- all inside the timed region."""
        import numpy as np

        torch = self.torch
        inst = make_instance(wl, self.world)
        inst.run_integration(3)  # warm-up through the same public call
        has_grid = hasattr(inst, "load_grid")
        host_grid = None
        if has_grid:
            host_grid = torch.from_numpy(
                np.ascontiguousarray(inst.divisions.cpu().numpy())).pin_memory()
        before = len(inst.history)
        self.barrier()
        t0 = time.perf_counter()
        if has_grid:
            inst.load_grid(numpy_grid=host_grid.numpy())  # H2D of the grid (n_dim*51*8 B)
        inst.run_integration(K)  # public API; host receives (res, sigma) of every iteration
        final_grid = inst.divisions.cpu() if has_grid else None  # D2H of the trained grid
        self.barrier()
        e2e_s = self.agree_max(time.perf_counter() - t0)
        rows = inst.history[before:]
        assert len(rows) == K and all(isinstance(r[0], float) for r in rows)
        events = sum(inst.events_log[-K:]) if inst._ROW == 3 else inst.n_events * K
        grid_bytes = host_grid.numel() * 8 if has_grid else 0
        return dict(value=events / e2e_s, seconds=e2e_s, h2d=grid_bytes / K,
                    d2h=8 * inst._ROW + (final_grid.numel() * 8 / K if has_grid else 0))


# Matrix elements: `flops_per_event` (vf_flops_per_event) is SURVEY 8(d)'s figure -- the device
# header compiled with an op-counting scalar, i.e. the operations of the chain AS IMPLEMENTED
# (128 / 245 per event on top of the 12d+5 of the VEGAS step).  The reference's LITERAL chain --
# zero-padded complex arithmetic, acos / sincos round trips, staged quotients -- is 448 / 1354
# operations (oracle/count_flops.py; tests/test_oracle.py keeps these in sync): the rows carry
# that figure too, as the rate of reference work delivered, NOT as pipe utilisation.
REFERENCE_CHAIN_OPS = {"drellyan_lo": 448.0, "singletop_lo": 1354.0}


def reference_chain_fields(wl, t, peak):
    ops = REFERENCE_CHAIN_OPS.get(wl["integrand"])
    if ops is None:
        return {}
    f_ref = 12.0 * wl["n_dim"] + 5.0 + ops
    return {"flops_per_event_reference_chain": f_ref,
            "frac_reference_chain": t["achieved"] * f_ref / t["f_alg"] / peak}


TABLE_1GPU = ["c1", "c2", "c3", "c4dy", "c4st", "c5", "sg8_rng32"]
TABLE_NGPU = ["c2", "c3", "c5"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="sg8", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-table", action="store_true",
                    help="skip the per-workload sub-table (c1, c2, c3, c4dy, c4st, c5)")
    ap.add_argument("--rng-bits", type=int, default=52, choices=[52, 32],
                    help="Philox bits per uniform (52 = default stream)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    global RNG_BITS
    RNG_BITS = args.rng_bits
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import ctypes

    B = Bench()
    lib, _lib, world, rank = B.lib, B._lib, B.world, B.rank
    K, W = args.steps, max(args.warmup, 3)

    # measured fp64 DFMA peak (MEASURED_PEAKS.json has no fp64 entry)
    peak = ctypes.c_double(0.0)
    _lib.check(lib.vf_fp64_peak_probe(20000, ctypes.byref(peak)))

    m = B.measure(wl, K, W, clocks=True)
    e2e = B.measure_e2e(wl, K)

    # ---- every BASELINE.json config as a short line of its own (same protocol, fewer steps)
    table = {}
    if not args.no_table:
        for name in (TABLE_1GPU if world == 1 else TABLE_NGPU):
            if name == args.workload:
                continue
            t = B.measure(WORKLOADS[name], None, 3)
            table[name] = {
                "workload": WORKLOADS[name]["name"], "value": t["value"], "unit": UNIT,
                "steps": t["steps"], "timed_regions": "median of 3 regions of `steps` steps",
                "ms_per_step": t["ms"] / t["steps"],
                "events_per_step_per_gpu": t["per_gpu_events"],
                "kernel_ms": t["kern_ms"], "epilogue_kernel_ms": t["epi_ms"],
                "flops_per_event": t["f_alg"], "achieved_tflops": t["achieved"],
                "frac": t["achieved"] / peak.value,
                **reference_chain_fields(WORKLOADS[name], t, peak.value),
            }

    if rank == 0:
        inst = m["inst"]
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": m["ms"] / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl["name"], "events_per_step_per_gpu": m["per_gpu_events"],
                "train": True, "rng": f"philox4x32-10, {RNG_BITS}-bit uniforms, generated in-kernel",
                "collective": ("none (1 GPU)" if world == 1 else
                               "fused block-reduce + NVLink peer-memory all-reduce + refine kernel"
                               if getattr(inst, "_exchange", None) is not None else
                               "NCCL all_reduce of the packed [d*50+2] buffer"),
                "l2": "n/a: no per-event input or output touches HBM (inputs are Philox counters);"
                      " grid + partials are <1 MB",
            },
            "roofline": {
                "bound": "fp64", "achieved": m["achieved"], "peak": peak.value, "unit": "TFLOP/s",
                "frac": m["achieved"] / peak.value, "traffic": _profiled_traffic(args.workload),
                "peak_source": "DFMA-chain probe run in this process (MEASURED_PEAKS.json has no "
                               "fp64 entry)",
                "peak_nominal": FP64_NOMINAL_TFLOPS,
                "frac_of_nominal": m["achieved"] / FP64_NOMINAL_TFLOPS,
                "flops_per_event": m["f_alg"],
                "kernel": "plus_event_kernel" if m["plus"] else "event_kernel",
                "kernel_ms": m["kern_ms"], "kernel_launches_timed": m["kern_n"],
                "kernel_timing": "CUDA events around each launch on its stream, over an identical "
                                 "K-step pass run immediately after the timed region",
                "kernel_share_of_step": m["kern_ms"] / (m["ms"] / K),
                "epilogue_kernel_ms": m["epi_ms"],
                **reference_chain_fields(wl, m, peak.value),
            },
            "clocks": m["clocks"],
            "e2e": {"value": e2e["value"], "unit": UNIT,
                    "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "note": "public API: grid uploaded from pinned host memory, "
                            "run_integration(K) with (res, sigma) of EVERY iteration delivered to "
                            "pinned host memory by that iteration's tail kernel, trained grid "
                            "downloaded; per-step inputs are the Philox (seed, iteration) launch "
                            "parameters -- the path has no per-event host data by construction"},
            "gpu_launches": m["launches"],
            "workloads": table,
        }
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_reference_run(wl, 3, 1, min(wl["n_events"], 10**6))
            if base is not None:
                line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind",
                                                             "sample")}
                best = cpu_best_effort(wl, min(wl["n_events"], 10**7))
                if best is not None:
                    line["cpu_best_effort"] = best
        print(json.dumps(line))
    if world > 1:
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
