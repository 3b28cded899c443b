import sys, json
sys.path.insert(0, "/root/repo")
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
eng = pq.get_engine(0)
S, NB = (int(sys.argv[1]), 5_040) if len(sys.argv) > 1 else (10_000, 5_040)
p = pq.Panel(S, NB, engine=eng, outputs_mask=(1 << N.N_OUTPUTS) - 1, host_staging=False)
p.fill_synthetic(seed=7)
for name, bit in list(N.IND_EXTRA.items()) + [("base15", N.IND_ALL), ("base15+mom", N.IND_ALL | N.IND_EXTRA["mom"]), ("all", N.IND_ALL | sum(N.IND_EXTRA.values()))]:
    tot, fused, nl = p.time_device(N.default_params(indicators=bit), warmup=1, iters=3)
    print(S, name, "fused %.3f ms" % (fused / 3), flush=True)
