import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
S, NB = int(os.environ.get("PQB_BENCH_SYMBOLS", 20000)), 5040
n = int(os.environ.get("PQB_HALTED", 200))
p = pq.Panel(S, NB, engine=pq.get_engine(0))
p.fill_synthetic(seed=5, to_host=True)
ok = np.ones(NB, dtype=bool); ok[2000:2003] = False
bits = np.packbits(ok, bitorder="little")
for s in np.linspace(0, S - 1, n).astype(int):
    p.set_column(int(s), "close", np.ascontiguousarray(p.host_field("close")[int(s)]), validity=bits)
p.upload()
p.run(); p.sync(); p.run(); p.sync()
