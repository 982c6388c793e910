"""timestamps of the NIW MMA issuer (build with -DMSB_NIW_TRACE): per tile of CTA 0, cycles relative to the tile's start"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import common_b200 as cb
from common_b200 import _lib
cfg = cb.synth.config("C4")
descs = cfg["models"]; n, k = 250_000, cfg["k"]
arr, z = cb.synth.make_dataset(descs, n, k, seed=73)
ctx = cb.Context(0)
st = cb.state(ctx, descs, max_groups=k + 8, cluster_hp={"alpha": 1.0})
st.bind(cb.numpy_dataview(arr))
g = np.asarray([st.create_group() for _ in range(k)])
st.add_values(g[z])
for i in range(3):
    st.sweep(seed=1, sweep=i)
lib = _lib.load()
buf = (ctypes.c_longlong * 256)()
lib.msb_debug_niw_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
print("rc", lib.msb_debug_niw_trace(buf, 256))
a = np.array(buf[:256]).reshape(32, 8)
names = ["top", "acc_empty", "a_full0", "issued0", "a_full1", "issued1", "epi_sees_full", "epi_released"]
print("tile  " + " ".join("%13s" % x for x in names) + "   period")
prev = None
for i, r in enumerate(a):
    print("%4d  " % (i + 8) + " ".join("%13d" % (x - r[0]) for x in r) + ("   %6d" % (r[0] - prev) if prev is not None else ""))
    prev = r[0]
