#!/usr/bin/env python
"""The public panel API end to end: a wide `date` + `{symbol}_{column}` pyarrow table (what load() returns, README.md:88-161)
-> WidePanel.suite() -> `date` + `{symbol}_{output}` table.  Prints one JSON line per shape; beside it the C oracle on the
same columns (all host cores) so the API's speed-up over the CPU path is visible."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import pyarrow as pa

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import synth  # noqa: E402
from polars_quant_b200.wide import WidePanel  # noqa: E402


def table(S, N, seed=1):
    d = synth.ohlcv(S, N, seed=seed)
    cols, names = [pa.array(np.arange(N, dtype=np.int64))], ["date"]
    for s in range(S):
        for f in ("open", "high", "low", "close", "volume"):
            cols.append(pa.array(d[f][s]))
            names.append("S%05d_%s" % (s, f))
    return pa.table(cols, names=names), d


def main():
    shapes = [(2_000, 2_520)] + ([(20_000, 5_040)] if "--big" in sys.argv else [])
    devices = None
    if "--devices" in sys.argv:
        devices = [int(x) for x in sys.argv[sys.argv.index("--devices") + 1].split(",")]
    for S, N in shapes:
        t, d = table(S, N)
        wp = WidePanel(t)
        times = []
        for it in range(4):
            t0 = time.perf_counter()
            out = wp.suite(devices=devices)
            times.append(time.perf_counter() - t0)
            assert out.num_columns == 21 * S + 1
            del out
        lazy_times = []
        for it in range(4):
            t0 = time.perf_counter()
            res = wp.suite(devices=devices, lazy=True)
            col = res["S%05d_rsi" % (S // 2)]                 # one column wrapped on demand
            lazy_times.append(time.perf_counter() - t0)
            assert len(res) == 21 * S and len(col) == N
            del res, col
        lazy_stages = {k: round(v, 2) for k, v in wp.last_timings.items()}
        out = wp.suite(devices=devices); del out              # (last_timings below: the full-table call)
        rec = {"workload": "WidePanel.suite(): %d symbols x %d bars, %d columns in, %d out" % (S, N, t.num_columns, 21 * S + 1),
               "devices": devices or [0], "first_call_ms": times[0] * 1e3, "ms": min(times[1:]) * 1e3,
               "symbol_bars_per_s": S * N / min(times[1:]), "last_call_stages_ms": {k: round(v, 2) for k, v in wp.last_timings.items()},
               "lazy_ms": min(lazy_times[1:]) * 1e3, "lazy_symbol_bars_per_s": S * N / min(lazy_times[1:]), "lazy_stages_ms": lazy_stages}
        if "--cpu" in sys.argv:
            from oracle import pqo
            ns = min(S, 512)
            cores = os.cpu_count() or 1
            pqo.suite_panel(d["close"][:ns], d["high"][:ns], d["low"][:ns], d["volume"][:ns], threads=cores)
            t0 = time.perf_counter()
            pqo.suite_panel(d["close"][:ns], d["high"][:ns], d["low"][:ns], d["volume"][:ns], threads=cores)
            rec["cpu_oracle_symbol_bars_per_s"] = ns * N / (time.perf_counter() - t0)
            rec["cpu_cores"] = cores
            rec["speedup_over_cpu_oracle"] = rec["symbol_bars_per_s"] / rec["cpu_oracle_symbol_bars_per_s"]
            rec["lazy_speedup_over_cpu_oracle"] = rec["lazy_symbol_bars_per_s"] / rec["cpu_oracle_symbol_bars_per_s"]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
