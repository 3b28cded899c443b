import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
S, NB = 8192, 5040
p = pq.Panel(S, NB, engine=pq.get_engine(0))
p.fill_synthetic(seed=5, to_host=True)
ok = np.ones(NB, dtype=bool); ok[2000:2003] = False
bits = np.packbits(ok, bitorder="little")
for s in range(0, S, 99):
    p.set_column(s, "close", np.ascontiguousarray(p.host_field("close")[s]), validity=bits)
p.upload()
for name, mask in (("all", N.IND_ALL), ("sma", 1), ("ema", 2), ("tema", 4), ("macd", 1 << 5), ("sma+ema", 3), ("bbands", 1 << 4), ("rsi", 1 << 6)):
    print("==", name, flush=True)
    sys.stderr.flush()
    prm = N.default_params(indicators=mask)
    p.run(prm); p.sync()
    tot, fused, nl = p.time_device(prm, warmup=1, iters=3)
    print("   kernel ms", fused / 3, flush=True)
