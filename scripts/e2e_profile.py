#!/usr/bin/env python
"""Times each phase of the host-buffer (e2e) step with a sync after each, to see where the time goes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import common_b200 as cb
from common_b200.dataview import device_dataview

cfg = cb.synth.config("C2")
n, k, descs = cfg["n"], cfg["k"], cfg["models"]
arr, z = cb.synth.make_dataset(descs, n, k, seed=73, storage=cfg.get("storage"))
ctx = cb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
view = cb.numpy_dataview(arr)
raw, _ = view.raw()
pinned = torch.from_numpy(raw).pin_memory()
types = view.types()
gz = None
for it in range(3):
    T = {}
    def lap(name, t0):
        torch.cuda.synchronize(); T[name] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); dv = device_dataview(ctx, data=pinned.data_ptr(), n=n, types=types); lap("dataview_create(H2D)", t0)
    t0 = time.perf_counter(); st = cb.state(ctx, descs, max_groups=k + 8, cluster_hp={"alpha": 1.0}); lap("state_create", t0)
    t0 = time.perf_counter(); st.bind(dv); lap("bind(pack+scorecol)", t0)
    t0 = time.perf_counter(); gids = np.asarray([st.create_group() for _ in range(k)]); lap("create_groups", t0)
    if gz is None: gz = gids[z].astype(np.int64)
    t0 = time.perf_counter(); st.add_values(gz); lap("add_values", t0)
    t0 = time.perf_counter(); r = st.sweep(seed=73, sweep=it); lap("sweep", t0)
    t0 = time.perf_counter(); a = st.assignments(); lap("assignments(D2H)", t0)
    t0 = time.perf_counter(); st.close(); dv.close(); lap("close", t0)
    print(it, {k_: round(v, 2) for k_, v in T.items()}, "total", round(sum(T.values()), 1))
