#!/usr/bin/env python
"""Candle engine throughput (GPU box): kernel-only on a device-resident synthetic panel + the host path.
usage: python scripts/bench_candles.py [S1xN1 S2xN2 ...] [--json out.json]
One launch = 61 Int32 pattern planes + 5 f64 price planes from 4 f64 input planes:
algorithmic bytes = 4*8 + 61*4 + 5*8 (+ 5/8 validity) = 316 B per symbol-bar."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import polars_quant_b200 as pq
from polars_quant_b200 import candles

ALGO = 4 * 8 + 61 * 4 + 5 * 8
peak = 6550.0
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
specs = [a for a in sys.argv[1:] if "x" in a and not a.startswith("--")] or ["5000x2520", "20000x5040"]
out = []
eng = pq.get_engine(0)
for spec in specs:
    S, N = (int(x) for x in spec.split("x"))
    host = S * N <= 5000 * 2520 * 4
    p = candles.CandlePanel(S, N, engine=eng, host_staging=host)
    iters = 10
    p.fill_synthetic(seed=3, to_host=False)
    ms_busy = p.time_device(warmup=3, iters=iters) / iters
    p.fill_random_walk(seed=0xC0FFEE, sigma=0.02, to_host=host)      # the SURVEY 8d panel: the headline data
    ms = p.time_device(warmup=3, iters=iters) / iters
    rec = {"symbols": S, "bars": N, "data": "random-walk OHLC sigma 0.02 (SURVEY 8d)", "kernel_ms": ms,
           "kernel_ms_busy_candles": ms_busy, "frac_busy_candles": ALGO * S * N / ms_busy / 1e6 / peak, "achieved_gbs": ALGO * S * N / ms / 1e6,
           "frac_of_measured_hbm_peak": ALGO * S * N / ms / 1e6 / peak, "symbol_bars_per_s": S * N / ms * 1e3,
           "algorithmic_bytes_per_symbol_bar": ALGO, "peak_gbs": peak}
    if host:
        import time
        p.run_host()
        t0 = time.perf_counter()
        for _ in range(3):
            p.run_host()
        dt = (time.perf_counter() - t0) / 3
        fired = int(sum((p.pattern(k) != 0).sum() for k in range(61)))
        rec.update({"e2e_ms": dt * 1e3, "e2e_symbol_bars_per_s": S * N / dt, "h2d_bytes": 32 * S * N,
                    "d2h_bytes": (61 * 4 + 5 * 8) * S * N, "pattern_hits": fired})
    if host and "--no-cpu" not in sys.argv:
        # CPU baseline beside it: oracle/pq_candles.c (the reference's per-function loops, 66 passes per symbol) on a
        # bounded sample of the same panel, all host cores
        import os, time
        sys.path.insert(0, str(ROOT))
        from oracle import pqo
        ns = min(S, 16 * (os.cpu_count() or 1))
        cols = [np.ascontiguousarray(p.host_field(f)[:ns]) for f in range(4)]
        pqo.candles_panel(*cols)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 10.0:
            _, _, used = pqo.candles_panel(*cols); reps += 1
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": ns * N * reps / dt, "unit": "symbol*bars/s", "cores": used, "kind": "port",
                               "sample": f"{ns} symbols x {N} bars of the same panel, {reps} passes, oracle/pq_candles.c"}
    print(json.dumps(rec))
    out.append(rec)
    p.close()
if "--json" in sys.argv:
    Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1))
