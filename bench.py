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
  c2  product  d=8,  1e7 events/iter          (configs[1], DEFAULT)
  c3  VegasFlowPlus adaptive symgauss d=8, 1e8 events/iter (configs[2], 1 GPU)
  c4dy / c4st  Drell-Yan d=4 / single-top d=3, 1e8 events/iter (configs[3])
  c5  symgauss d=20, 1.25e8 events/iter/GPU   (configs[4]: 1e9 over 8 GPUs)
  sg8 symgauss d=8,  1e8 events/iter          (north_star target line)

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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(wl["n_events"], 10**6)
    steps = max(1, min(args.steps, 20))
    base = cpu_reference_run(wl, steps, min(args.warmup, 1), sample)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable":
                          f"no CPU restatement for workload {args.workload}"}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": base["seconds"] / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "note": "TensorFlow is not installable here; this is "
                   "the reference-shaped CPU restatement (oracle/ref_shaped_torch.py)"},
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
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


RNG_BITS = 52


def make_instance(wl, world):
    import vegasflow_b200 as vf

    n_total = wl["n_events"] * (world if wl["alg"] != "plus" else 1)
    if wl["alg"] == "vegas":
        inst = vf.VegasFlow(wl["n_dim"], n_total, verbose=False, rng_bits=RNG_BITS)
    elif wl["alg"] == "plus":
        inst = vf.VegasFlowPlus(wl["n_dim"], n_total, adaptive=True, verbose=False,
                                rng_bits=RNG_BITS)
    else:
        inst = vf.PlainFlow(wl["n_dim"], n_total, verbose=False, rng_bits=RNG_BITS)
    inst.set_seed(2024)
    inst.compile(getattr(vf.integrands, wl["integrand"]))
    return inst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rng-bits", type=int, default=52, choices=[52, 32],
                    help="Philox bits per uniform (52 = default stream)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    global RNG_BITS
    RNG_BITS = args.rng_bits
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as graft

    graft.build()
    from vegasflow_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # short collective timeout: a hang must fail fast, not hold 8 GPUs for 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=90))
    lib = _lib.require_cuda()
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # measured fp64 DFMA peak (MEASURED_PEAKS.json has no fp64 entry)
    import ctypes

    peak = ctypes.c_double(0.0)
    _lib.check(lib.vf_fp64_peak_probe(20000, ctypes.byref(peak)))

    inst = make_instance(wl, world)

    def run_n(n):
        """n steps of the product path (identical count on every rank)."""
        n_per_step = inst.n_events
        if n <= 0:
            return 0
        if inst._run_batched(n) is not None:
            return n_per_step * n
        done = 0
        for _ in range(n):
            done += inst.n_events
            inst._run_iteration()
        return done

    def agree_max(x):
        """Same value on every rank (max), so loop counts derived from it cannot diverge --
        a rank-dependent iteration count would deadlock the per-iteration collective."""
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up: W steps (timed to size the clock ramp), then ~0.3 s more of the same steps
    barrier()
    t0 = time.perf_counter()
    run_n(W)
    barrier()
    step_s = agree_max((time.perf_counter() - t0) / W)
    run_n(int(min(2000, max(0, 0.3 / max(step_s, 1e-6)))))
    barrier()

    # ---- timed region: exactly K steps of the product path, device-timed, max over ranks.
    # Single rank: ONE vf_run_iterations call enqueues all K iterations (2 launches each).
    # Multi rank: per iteration vf_run_event, the NCCL all-reduce, vf_iteration_epilogue.
    sampler = ClockSampler(torch.cuda.current_device())
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    def run_steps():
        return run_n(K)

    barrier()
    lib.vf_launch_count(1)
    sampler.start()
    ev0.record()
    events_done = run_steps()
    ev1.record()
    barrier()
    launches = int(lib.vf_launch_count(0))
    ms = agree_max(ev0.elapsed_time(ev1))  # max over ranks
    value = events_done / (ms * 1e-3)
    # Second, identical K-step pass with the library's CUDA-event bracket around every
    # event-kernel / epilogue launch (same stream): per-kernel durations for the roofline.
    # Kept out of the timed region above because the extra event records widen the gaps
    # between the back-to-back launches by a few microseconds.
    lib.vf_kernel_timing(1)
    run_steps()
    barrier()
    kt, kn = ctypes.c_double(0.0), ctypes.c_int(0)
    _lib.check(lib.vf_kernel_time_ms(ctypes.byref(kt), ctypes.byref(kn)))
    et, en = ctypes.c_double(0.0), ctypes.c_int(0)
    _lib.check(lib.vf_epilogue_time_ms(ctypes.byref(et), ctypes.byref(en)))
    lib.vf_kernel_timing(0)
    kern_ms = kt.value / max(kn.value, 1)
    epi_ms = et.value / max(en.value, 1) if en.value else None
    # keep the same steps running ~1 s more so the clock sampler sees the kernel under load
    # (count derived from the agreed step time: identical on every rank)
    if ms < 1000.0 and agree_max(1.0 if sampler.nv is not None else 0.0) > 0.5:
        run_n(int(min(20000, 1000.0 / max(ms / K, 1e-3))))
        barrier()
    clocks = sampler.stop()
    clocks["note"] = ("NVML samples every 2 ms over the timed region plus a ~1 s repeat of the same "
                      "steps right after it" if ms < 1000.0 else "NVML samples over the timed region")

    # ---- e2e: public API, one D2H read of (res, sigma) per step like the reference's logging
    inst2 = make_instance(wl, world)
    import numpy as np

    host_grid = np.ascontiguousarray(inst2.divisions.cpu().numpy())
    for _ in range(3):
        inst2.run_iteration()
    barrier()
    e2e_events = 0
    t0 = time.perf_counter()
    inst2.load_grid(numpy_grid=host_grid)  # H2D of the grid (n_dim*51*8 B)
    for _ in range(K):
        e2e_events += inst2.n_events
        inst2.run_iteration()  # public API: one 16-byte D2H read + sync, every step
    final_grid = inst2.divisions.cpu()  # D2H of the trained grid
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    grid_bytes = host_grid.nbytes

    if rank == 0:
        plus = 1 if wl["alg"] == "plus" else 0
        mode = 0 if wl["alg"] == "plain" else 1
        iid = lib.vf_integrand_id(wl["integrand"].encode())
        f_alg = lib.vf_flops_per_event(mode, iid, wl["n_dim"], plus)
        per_gpu_events = events_done / K / (world if wl["alg"] != "plus" else 1)
        achieved = per_gpu_events * f_alg / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl["name"], "events_per_step_per_gpu": per_gpu_events,
                "train": True, "rng": f"philox4x32-10, {RNG_BITS}-bit uniforms, generated in-kernel",
                "collective": ("none (1 GPU)" if world == 1 else
                               "fused block-reduce + NVLink peer-memory all-reduce + refine kernel"
                               if getattr(inst, "_exchange", None) is not None else
                               "NCCL all_reduce of the packed [d*50+2] buffer"),
                "l2": "n/a: no per-event input or output touches HBM (inputs are Philox counters);"
                      " grid + partials are <1 MB",
            },
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value, "traffic": _profiled_traffic(args.workload),
                "peak_source": "DFMA-chain probe run in this process (MEASURED_PEAKS.json has no "
                               "fp64 entry)",
                "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
                "flops_per_event": f_alg,
                "kernel": "plus_event_kernel" if plus else "event_kernel",
                "kernel_ms": kern_ms, "kernel_launches_timed": kn.value,
                "kernel_timing": "CUDA events around each launch on its stream, over an identical "
                                 "K-step pass run immediately after the timed region",
                "kernel_share_of_step": kern_ms / (ms / K),
                "epilogue_kernel_ms": epi_ms,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_events / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": grid_bytes / K,
                    "d2h_bytes_per_step": 16 + final_grid.numel() * 8 / K,
                    "note": "VegasFlow public API: grid uploaded from host, (res, sigma) read back "
                            "to the host after every iteration, trained grid downloaded at the end"},
            "gpu_launches": launches,
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
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
