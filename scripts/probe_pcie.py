#!/usr/bin/env python
"""Host <-> device bandwidth of every visible GPU, alone and all together, from pinned staging allocated by the engine
(bound to the GPU-local CPUs): the ceiling of the end-to-end numbers.  One JSON line."""
import ctypes as C
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq  # noqa: E402
from polars_quant_b200 import _native as N  # noqa: E402


def main():
    n = N.lib().pqb_device_count()
    S, NB = 2048, 5040
    panels, rec = [], {"gpus": n, "panel": "%d x %d per GPU" % (S, NB), "local_cpus": {}}
    for g in range(n):
        eng = pq.get_engine(g)
        rec["local_cpus"][g] = len(eng.local_cpus())
        p = pq.Panel(S, NB, engine=eng)
        p.fill_synthetic(to_host=True)
        panels.append(p)
    prm = N.default_params()
    moved = lambda p: (4 + 21) * S * p.pitch * 8 / 1e9

    def one(p, out, i):
        p.run_host(prm)
        t0 = time.perf_counter()
        for _ in range(3):
            p.run_host(prm)
        out[i] = 3 * moved(p) / (time.perf_counter() - t0)

    alone = [0.0] * n
    for i, p in enumerate(panels):
        one(p, alone, i)
    together = [0.0] * n
    ts = [threading.Thread(target=one, args=(p, together, i)) for i, p in enumerate(panels)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    rec["e2e_gb_per_s_alone"] = [round(x, 1) for x in alone]
    rec["e2e_gb_per_s_together"] = [round(x, 1) for x in together]
    rec["sum_together"] = round(sum(together), 1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
