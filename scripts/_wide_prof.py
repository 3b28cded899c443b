import sys, time, cProfile, pstats
import numpy as np, pyarrow as pa
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import synth
import polars_quant_b200 as pq
from polars_quant_b200 import wide
S, NB = 2000, 2520
d = synth.ohlcv(S, NB, seed=17)
cols, names = [pa.array(np.arange(NB, dtype=np.int32))], ["date"]
for s in range(S):
    for f in ("open", "high", "low", "close", "volume"):
        cols.append(pa.array(d[f][s])); names.append("S%05d_%s" % (s, f))
t = pa.table(cols, names=names)
wp = wide.WidePanel(t, engine=pq.get_engine(0))
wp.suite()
pr = cProfile.Profile(); pr.enable(); r = wp.suite(); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(18)
