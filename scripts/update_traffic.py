#!/usr/bin/env python
"""profiles/traffic.json from full-size ncu digests (scripts/gpu_round.sh fullsum legs).
usage: update_traffic.py TAG   (reads gpurun_out/TAG_prof_{C2..C5}_full.txt, copies them to profiles/)"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
ROWS = {"C2": 1000000, "C3": 4000000, "C4": 1000000, "C5": 1000000}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
path = os.path.join(ROOT, "profiles", "traffic.json")
tr = json.load(open(path))
tr["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from ncu --set full captures "
                  "(one launch, same workload and size as bench.py runs; scripts/gpu_round.sh fullsum legs, scripts/update_traffic.py). "
                  "bench.py copies the entry of its workload into roofline.traffic.")
for wl, rows in ROWS.items():
    src = os.path.join(ROOT, "gpurun_out", "%s_prof_%s_full.txt" % (tag, wl))
    if not os.path.exists(src):
        print("missing", src); continue
    txt = open(src).read()
    def metric(name):
        m = re.search(r"^%s\s+([0-9.,]+)\s*(\S*)" % re.escape(name), txt, re.M)
        if not m:
            return None, None
        return float(m.group(1).replace(",", "")), m.group(2)
    rd, ru = metric("dram__bytes_read.sum")
    wr, wu = metric("dram__bytes_write.sum")
    if rd is None or wr is None:
        print("no dram metrics in", src); continue
    rd *= UNIT.get(ru, 1); wr *= UNIT.get(wu, 1)
    kern = re.search(r"^== (.*)$", txt, re.M).group(1)
    dst = "profiles/%s_prof_%s_full.txt" % (tag, wl)
    shutil.copy(src, os.path.join(ROOT, dst))
    e = {"kernel": kern[:90], "rows": rows, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
         "dram_bytes_per_launch": int(rd + wr), "capture": dst}
    for key, name in ((("smem_wavefronts_per_launch", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),) if wl == "C2" else ()):
        v, _ = metric(name)
        if v is not None:
            e[key] = int(v)
    for key, name in (("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                      ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                      ("lsu_pipe_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                      ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                      ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active")):
        v, _ = metric(name)
        if v is not None:
            e[key] = round(v, 2)
    tr[wl] = e
    print(wl, e)
json.dump(tr, open(path, "w"), indent=1)
