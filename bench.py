#!/usr/bin/env python
"""bench.py -- fused indicator suite throughput (symbol.bars/s) on B200, one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2]

A "step" is one pass of the fused 15-indicator / 21-output suite over one synthetic random-walk
OHLCV panel already resident in HBM.  N > 1 is launched by torchrun, one process per GPU; symbols
are independent so every rank runs its own panel of the full workload shape (weak scaling, no
collective on the data path); torch.distributed (NCCL) is used only for the barrier and the
max-over-ranks of the device-timed duration.

Keys beyond the base contract:
  roofline      dominant kernel (suite_fused_kernel) vs the measured HBM copy peak
  cpu_baseline  the C oracle (a port of the reference's Rust loops) on this box's host cores
  e2e           same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H
`--impl reference` times the reference's own CPU path -- the C oracle port, since the Rust crate
cannot be built in this image -- on all host cores, same config/metric/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    # name: (symbols, bars, description)
    "c4": (50_000, 5_040, "BASELINE config 4 / north_star target: 50,000 symbols x 5,040 bars f64 OHLCV, full 15-indicator suite"),
    "c2": (5_000, 2_520, "BASELINE config 2: 5,000 symbols x 2,520 bars f64 OHLCV, full 15-indicator suite"),
}
N_IN, N_OUT = 4, 21
ALGO_BYTES_PER_SYMBOL_BAR = 8 * (N_IN + N_OUT)          # 200 B; validity bits (+2.6 B) not counted
METRIC = "symbol_bars_per_sec_fused_indicator_suite"
UNIT = "symbol*bars/s"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML in-process every ~5 ms
    (nvidia_ml_py), falling back to polling nvidia-smi."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _run_nvml(self):
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                if mx:
                    self.mx.append(mx)
                r = int(get_reasons(self._h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.005)

    def _run_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    r = [c.strip() for c in line.split(",")]
                    self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml else "nvidia-smi"}


def cpu_suite_rate(c, h, l, v, threads: int, budget_s: float):
    """Times the C oracle (reference port) over [symbols, bars] host arrays; repeats to ~budget."""
    from oracle import pqo
    S, N = c.shape
    t0 = time.perf_counter()
    _, _, used = pqo.suite_panel(c, h, l, v, threads=threads)
    dt = time.perf_counter() - t0
    reps, total = 1, dt
    while total < budget_s and reps < 50:
        t0 = time.perf_counter()
        pqo.suite_panel(c, h, l, v, threads=threads)
        total += time.perf_counter() - t0
        reps += 1
    return S * N * reps / total, used, reps


def synth_host_sample(n_symbols, n_bars, seed=0xC0FFEE):
    import synth
    d = synth.ohlcv(n_symbols, n_bars, seed=seed)
    return d["close"], d["high"], d["low"], d["volume"]


def run_reference(args, shape, rank, world):
    """--impl reference: the reference's CPU path (oracle port; oracle/_ref does not exist because
    the Rust reference cannot be compiled here) on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    S, N, desc = shape
    cores = os.cpu_count() or 1
    # bounded sample of the workload: enough symbols for ~1-2 s per step on this box
    sample_symbols = min(S, max(64, 16 * cores))
    c, h, l, v = synth_host_sample(sample_symbols, N)
    from oracle import pqo
    for _ in range(max(args.warmup, 1)):
        pqo.suite_panel(c, h, l, v, threads=cores)
    t0 = time.perf_counter()
    used = cores
    for _ in range(args.steps):
        _, _, used = pqo.suite_panel(c, h, l, v, threads=cores)
    dt = time.perf_counter() - t0
    value = sample_symbols * N * args.steps / dt
    sample = f"{sample_symbols} of {S} symbols x {N} bars per step, numpy random-walk OHLCV, {used} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "symbols": S, "bars": N, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--symbols", type=int, default=0, help="override symbols per GPU")
    ap.add_argument("--bars", type=int, default=0)
    ap.add_argument("--no-extra", action="store_true", help="skip the kernel-only lines of the other shapes (config 2, candles)")
    ap.add_argument("--e2e-symbols", type=int, default=8192, help="symbols of the workload pushed through the host path")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    S, N, desc = WORKLOADS[args.workload]
    if args.symbols:
        S = args.symbols
    if args.bars:
        N = args.bars

    if args.impl == "reference":
        run_reference(args, (S, N, desc), rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner (NCCL_DEBUG=VERSION/INFO in some images) on stdout: keep stdout
        # to the one JSON line of the contract
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # ... and some builds print it with a plain printf when the communicator is created: point fd 1 at stderr
        # until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as NV

    def barrier():
        if dist is not None:
            dist.barrier()

    engine = pq.get_engine(local_rank)
    params = NV.default_params()
    panel = pq.Panel(S, N, engine=engine, host_staging=False)
    panel.fill_synthetic(seed=0xC0FFEE + 1_000_003 * rank, sigma=0.02)

    # ---- device-resident throughput: W warm-up + K timed steps, CUDA events on the engine stream ----
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms_total, ms_fused, launches = panel.time_device(params, warmup=args.warmup, iters=args.steps)
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([ms_total, ms_fused], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_fused = float(t[0]), float(t[1])
    units = S * N * world
    value = units * args.steps / (ms_total * 1e-3)
    peak, peak_src = measured_peak()
    fused_ms_avg = ms_fused / args.steps
    achieved = ALGO_BYTES_PER_SYMBOL_BAR * S * N / (fused_ms_avg * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    e2e_launches = 0
    if not args.no_e2e:
        Se = min(S, args.e2e_symbols)
        hp = pq.Panel(Se, N, engine=engine, host_staging=True)
        hp.fill_synthetic(seed=0xC0FFEE + 1_000_003 * rank, sigma=0.02, to_host=True)
        barrier()
        ms_host = hp.time_host(params, chunk_symbols=0, warmup=1, iters=args.e2e_steps)
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([ms_host], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_host = float(t[0])
        pitch = hp.pitch
        e2e = {"value": Se * N * world * args.e2e_steps / (ms_host * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(N_IN * Se * pitch * 8 + Se * 4),
               "d2h_bytes_per_step": int(N_OUT * Se * pitch * 8 + N_OUT * Se * hp.validity_pitch),
               "symbols": Se, "bars": N, "ms_per_step": ms_host / args.e2e_steps,
               "note": "row-major pinned host panel -> chunked H2D || pack + fused suite + unpack || D2H on 3 streams; PCIe-bound"}
        e2e_launches = hp.last_launches() * args.e2e_steps

        # ---- CPU baseline on rank 0: the oracle port over a bounded sample of the same panel ----
        cpu = None
        if rank == 0 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ns = min(Se, max(64, 16 * cores))
            c, h, l, v = (np.ascontiguousarray(hp.host_field(f)[:ns, :N]) for f in ("close", "high", "low", "volume"))
            rate, used, reps = cpu_suite_rate(c, h, l, v, cores, args.cpu_seconds)
            cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
                   "sample": f"{ns} symbols x {N} bars of the same synthetic panel, {reps} passes, C oracle (oracle/pq_oracle.c)"}
        hp.close()
    else:
        cpu = None

    # ---- the other measured shapes of BASELINE.json / SURVEY.md 8f, kernel-only, reported beside the headline
    #      (N=1 only; a few seconds): config 2 on the same suite kernel, and the fused candle kernel ----
    other = None
    if world == 1 and not args.no_extra:
        other = {}
        try:
            panel.close()
            if (S, N) != WORKLOADS["c2"][:2]:
                S2, N2, d2 = WORKLOADS["c2"]
                p2 = pq.Panel(S2, N2, engine=engine, host_staging=False)
                p2.fill_synthetic(seed=0xC0FFEE, sigma=0.02)
                _, f2, _ = p2.time_device(params, warmup=3, iters=20)
                p2.close()
                g2 = ALGO_BYTES_PER_SYMBOL_BAR * S2 * N2 / (f2 / 20 * 1e-3) / 1e9
                other["c2"] = {"workload": d2, "kernel": "suite_fused_kernel<true,false>", "kernel_ms": f2 / 20,
                               "value": S2 * N2 / (f2 / 20 * 1e-3), "unit": UNIT, "achieved_gbs": g2, "frac": g2 / peak}
            from polars_quant_b200 import candles
            Sc, Nc = 20_000, 5_040
            cp = candles.CandlePanel(Sc, Nc, engine=engine, host_staging=False)
            cp.fill_random_walk(seed=0xC0FFEE, sigma=0.02)
            msc = cp.time_device(warmup=3, iters=10) / 10
            cp.close()
            gc = 316 * Sc * Nc / (msc * 1e-3) / 1e9
            other["candles"] = {"workload": "SURVEY 8f.1: 61 cdl* patterns + 4 price transforms + bop, 20,000 x 5,040 random-walk OHLC",
                                "kernel": "candle_kernel<true>", "kernel_ms": msc, "value": Sc * Nc / (msc * 1e-3),
                                "unit": "symbol*bars/s", "algorithmic_bytes_per_symbol_bar": 316, "achieved_gbs": gc,
                                "frac": gc / peak}
            # what HBM delivers for the suite's own access mix (4 planes read, 21 written, 256-byte warp rows) with no
            # arithmetic at all: the copy bandwidth used as `peak` is a 1 : 1 mix, a write-heavy stream gets less
            import ctypes as C
            from polars_quant_b200 import _native as NN
            msx = C.c_float()
            NN.check(NN.lib().pqb_stream_mix(engine._h, N_IN, N_OUT, 20_000 * 5_040, 2, 5, C.byref(msx)))
            gx = 8 * (N_IN + N_OUT) * 20_000 * 5_040 / (msx.value * 1e-3) / 1e9
            other["access_mix_ceiling"] = {"workload": "streaming kernel, %d planes read : %d written, 20,000 x 5,040 doubles each, no arithmetic" % (N_IN, N_OUT),
                                           "kernel": "stream_mix_kernel", "kernel_ms": msx.value, "achieved_gbs": gx, "frac": gx / peak,
                                           "suite_over_mix": achieved / gx}
        except Exception as ex:          # never lose the headline line to an extra
            other["error"] = repr(ex)

    if rank == 0:
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                if tj.get("symbols") == S and tj.get("bars") == N:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "symbols_per_gpu": S, "bars": N, "indicators": 15, "outputs": N_OUT,
                       "inputs": N_IN, "l2": "inputs+outputs per step (%.1f GB) are far larger than the 126 MB L2; no flush needed"
                       % (ALGO_BYTES_PER_SYMBOL_BAR * S * N / 1e9),
                       "parallelism": "symbols sharded per GPU, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "suite_fused_kernel<true,false>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SYMBOL_BAR * S * N,
                         "kernel_ms": fused_ms_avg},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches * args.steps + e2e_launches,
            "clocks": clocks.summary(),
        }
        if other:
            line["other_workloads"] = other
        print(json.dumps(line), flush=True)
    panel.close()
    if dist is not None:
        dist.destroy_process_group()



if __name__ == "__main__":
    main()
