#!/usr/bin/env python
"""Where the end-to-end time goes (GPU box): raw pinned <-> device copy rates (torch, large buffers, alone and both directions
at once), pqb_suite_run_host from the panel's own pinned staging, and pqb_suite_run_columns from caller-owned pageable columns
at several intake thread counts.  One JSON line per measurement."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import polars_quant_b200 as pq
from polars_quant_b200 import _native as NV

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = 5040
print(json.dumps({"cpus": os.cpu_count(), "symbols": S, "bars": N}), flush=True)

# ---- raw copies -------------------------------------------------------------------------------------------------------
nbytes = 4 << 30
hbuf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
hbuf2 = torch.empty(nbytes // 4, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
dbuf2 = torch.empty(nbytes // 4, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def d2h():
    with torch.cuda.stream(s1): hbuf.copy_(dbuf, non_blocking=True)
def h2d():
    with torch.cuda.stream(s2): dbuf2.copy_(hbuf2, non_blocking=True)
def both():
    d2h(); h2d()
t = timed(d2h); print(json.dumps({"raw_d2h_gbs": nbytes / t / 1e9}), flush=True)
t = timed(h2d); print(json.dumps({"raw_h2d_gbs": nbytes / 4 / t / 1e9}), flush=True)
t = timed(both); print(json.dumps({"raw_both_d2h_gbs": nbytes / t / 1e9, "raw_both_h2d_gbs": nbytes / 4 / t / 1e9}), flush=True)
del hbuf, hbuf2, dbuf, dbuf2

# ---- the engine's paths -----------------------------------------------------------------------------------------------
eng = pq.get_engine(0)
prm = NV.default_params()
hp = pq.Panel(S, N, engine=eng, host_staging=True)
hp.fill_synthetic(seed=1, sigma=0.02, to_host=True)
out_gb = 21 * S * hp.pitch * 8 / 1e9
def t_host():
    hp.run_host(prm)
    t0 = time.perf_counter()
    for _ in range(3): hp.run_host(prm)
    return (time.perf_counter() - t0) / 3
t = t_host(); print(json.dumps({"run_host_ms": t * 1e3, "d2h_gbs": out_gb / t, "symbol_bars_per_s": S * N / t}), flush=True)
mats = {f: np.array(hp.host_field(f)[:S, :N]) for f in ("close", "high", "low", "volume")}
refs, keep = pq.Panel.field_refs(**mats)
for th in (4, 8, 16, 32):
    hp.run_columns(refs, prm, threads=th)
    t0 = time.perf_counter()
    for _ in range(3): hp.run_columns(refs, prm, threads=th)
    t = (time.perf_counter() - t0) / 3
    print(json.dumps({"run_columns_threads": th, "ms": t * 1e3, "d2h_gbs": out_gb / t, "symbol_bars_per_s": S * N / t}), flush=True)
# intake alone (no GPU work)
for th in (8, 16):
    t0 = time.perf_counter()
    NV.check(NV.lib().pqb_panel_set_columns(hp._h, refs, len(refs), th))
    t = time.perf_counter() - t0
    print(json.dumps({"set_columns_threads": th, "ms": t * 1e3, "gbs": 4 * S * N * 8 / t / 1e9}), flush=True)
