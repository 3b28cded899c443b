#!/usr/bin/env python
"""Kernel-only timing of BASELINE configs 3 and 5 at full size (GPU box), device-resident synthetic panels.
C3: 500 symbols x 1,000,000 minute bars, EMA(p) + MACD(12,26,9) per launch, p in 12/26/200/5000 (algorithmic
    bytes per launch: 1 in + 4 out = 40 B per symbol-bar).
C5: 10,000 x 5,040, one launch per window: KDJ(k)+ATR(14) for k in 5/9/14/60/250 (3 in + 4 out = 56 B) and
    WILLR(p)+MIDPRICE(p) for p in 5/20/55/250 (3 in + 2 out = 40 B), and the same + the Donchian channel's upper / lower."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
sys.path.insert(0, str(ROOT / "oracle"))
NAMES = N.OUTPUT_NAMES
peak = 6550.0
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
eng = pq.get_engine(0)
out = []

def run(tag, panel, params, bytes_per_sb, S, NB, iters=5):
    tot, fused, nl = panel.time_device(params, warmup=2, iters=iters)
    ms = fused / iters
    rec = {"config": tag, "symbols": S, "bars": NB, "kernel_ms": ms, "algorithmic_bytes_per_symbol_bar": bytes_per_sb,
           "achieved_gbs": bytes_per_sb * S * NB / ms / 1e6, "frac": bytes_per_sb * S * NB / ms / 1e6 / peak,
           "symbol_bars_per_s": S * NB / ms * 1e3}
    print(json.dumps(rec)); out.append(rec)

which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "c3,c5"
if "c3" in which:
    S, NB = 500, 1_000_000
    om = sum(1 << NAMES.index(o) for o in ("ema", "macd", "macd_signal", "macd_hist"))
    p = pq.Panel(S, NB, engine=eng, fields_mask=1, outputs_mask=om, host_staging=False)
    p.fill_synthetic(seed=3, sigma=0.0005)
    for period in (12, 26, 200, 5000):
        run(f"c3 ema({period})+macd(12,26,9)", p, N.default_params(indicators=N.IND["ema"] | N.IND["macd"], ema_period=period), 40, S, NB, iters=3)
    p.close()
if "c3" in which:
    # the same config through a time-split panel (include/pqb200.h "time-split panels"): rows cut into chunks that run
    # as virtual symbols with the warm-up the periods need (results within 1e-12 relative of the serial walk)
    S, NB = 500, 1_000_000
    om = sum(1 << NAMES.index(o) for o in ("ema", "macd", "macd_signal", "macd_hist"))
    for period, chunks in ((12, 64), (26, 64), (200, 64), (5000, 16), (5000, 32)):
        prm = N.default_params(indicators=N.IND["ema"] | N.IND["macd"], ema_period=period)
        W = pq.SplitPanel.required_warmup(prm)
        sp = pq.SplitPanel(S, NB, chunks=chunks, warmup=W, engine=eng, fields_mask=1, outputs_mask=om, host_staging=False)
        sp.fill_synthetic(seed=3, sigma=0.0005)
        tot, fused, nl = sp.time_device(prm, warmup=2, iters=5)
        ms = fused / 5
        rec = {"config": f"c3 split ema({period})+macd(12,26,9)", "symbols": S, "bars": NB, "chunks": chunks, "warmup": sp.warmup,
               "virtual_symbols": sp.virtual_symbols, "virtual_bars": sp.virtual_bars, "kernel_ms": ms,
               "algorithmic_bytes_per_symbol_bar": 40, "achieved_gbs": 40 * S * NB / ms / 1e6, "frac": 40 * S * NB / ms / 1e6 / peak,
               "traffic_bytes_per_symbol_bar": 40 * sp.virtual_symbols * sp.virtual_bars / (S * NB),
               "symbol_bars_per_s": S * NB / ms * 1e3}
        print(json.dumps(rec)); out.append(rec)
        sp.close()
if "c5" in which:
    S, NB = 10_000, 5_040
    om = sum(1 << NAMES.index(o) for o in ("atr", "kdj_k", "kdj_d", "kdj_j", "willr", "midprice", "donchian_upper", "donchian_lower"))
    p = pq.Panel(S, NB, engine=eng, outputs_mask=om, host_staging=False)
    p.fill_synthetic(seed=55, sigma=0.02)
    for k in (5, 9, 14, 60, 250):
        run(f"c5 kdj({k},3,3)+atr(14)", p, N.default_params(indicators=N.IND["kdj"] | N.IND["atr"], kdj_fastk=k), 56, S, NB)
    for w in (5, 20, 55, 250):
        run(f"c5 willr({w})+midprice({w})", p, N.default_params(indicators=N.IND["willr"] | N.IND["midprice"], willr_period=w, midprice_period=w), 40, S, NB)
    for w in (5, 20, 55, 250):      # + the Donchian channel's upper / lower lines (a second launch: optional group; 2 in + 2 out more)
        run(f"c5 willr({w})+midprice({w})+donchian({w})", p, N.default_params(indicators=N.IND["willr"] | N.IND["midprice"] | N.IND_EXTRA["donchian"],
            willr_period=w, midprice_period=w, donchian_period=w), 40 + 32, S, NB)
    p.close()
if "--json" in sys.argv:
    Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1))
