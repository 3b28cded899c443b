import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
S, N = int(sys.argv[1]), 5040
eng = pq.get_engine(0); prm = NV.default_params()
hp = pq.Panel(S, N, engine=eng, host_staging=True)
hp.fill_synthetic(seed=1, sigma=0.02, to_host=True)
mats = {f: np.array(hp.host_field(f)[:S, :N]) for f in ("close", "high", "low", "volume")}
refs, keep = pq.Panel.field_refs(**mats)
arr, sch = NV.ArrowArray(), NV.ArrowSchema()
rel_a = C.CFUNCTYPE(None, C.POINTER(NV.ArrowArray)); rel_s = C.CFUNCTYPE(None, C.POINTER(NV.ArrowSchema))
for it in range(3):
    t0 = time.perf_counter(); hp.run_columns(refs, prm, threads=16); t1 = time.perf_counter()
    NV.check(NV.lib().pqb_panel_export_arrow(hp._h, 0, None, C.byref(arr), C.byref(sch))); t2 = time.perf_counter()
    rel_a(arr.release)(C.pointer(arr)); rel_s(sch.release)(C.pointer(sch)); t3 = time.perf_counter()
    print("S=%d run_columns %.1f ms (%.1f GB/s d2h)  export %.1f ms  release %.1f ms" % (S, (t1-t0)*1e3, 21*S*hp.pitch*8/(t1-t0)/1e9, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
t0 = time.perf_counter(); hp.run_host(prm); t1 = time.perf_counter(); print("run_host %.1f ms (%.1f GB/s)" % ((t1-t0)*1e3, 21*S*hp.pitch*8/(t1-t0)/1e9))
for th in (4, 8):
    t0 = time.perf_counter(); hp.run_columns(refs, prm, threads=th); t1 = time.perf_counter(); print("threads", th, "run_columns %.1f ms" % ((t1-t0)*1e3))
