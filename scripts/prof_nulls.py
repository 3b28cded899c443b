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
p.run(); p.run(); p.sync()
