"""Random-access microbenchmarks = the measured denominators of the insert kernels.
python scripts/microbench.py  -> ops/s for HBM-sized (6 GiB) and L2-resident (16 MiB) tables"""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbg_assembly_b200 import capi
L = capi.load()
names = {0: "ld32+st", 1: "ld32+cas64", 2: "red.u32", 3: "red.f16x8", 4: "ld32", 5: "ld32+red.f16x8"}
out = {}
for tag, nbytes in (("hbm_6GiB", 6 << 30), ("l2_16MiB", 16 << 20), ("l2_64MiB", 64 << 20)):
    for mode in range(6):
        ms = C.c_float(0)
        n = 1 << 28
        capi.check(L.dbg_measure_random_rmw(0, nbytes, n, mode, C.byref(ms)), "rmw")
        out[f"{tag}:{names[mode]}"] = round(n / (ms.value * 1e-3) / 1e9, 2)
print(json.dumps({"random_32B_sector_ops_Gops_per_s": out}))
