#!/usr/bin/env python
"""One readable block per configuration of a bench.py JSON line. usage: bench_digest.py file.json"""
import json
import sys

txt = open(sys.argv[1]).read().strip().splitlines()
line = None
for t in reversed(txt):
    if t.startswith("{"):
        line = json.loads(t)
        break
if line is None:
    print("no JSON line in", sys.argv[1]); print("\n".join(txt[-15:])); sys.exit(1)
print("headline %.4g %s  ms/step %.3f  e2e %s  cpu %s" % (
    line["value"], line["unit"], line["ms_per_step"],
    ("%.4g (%.3f ms)" % (line["e2e"]["value"], line["e2e"]["ms_per_step"])) if "e2e" in line else None,
    ("%.3g on %d cores" % (line["cpu_baseline"]["value"], line["cpu_baseline"]["cores"])) if "cpu_baseline" in line else None))
for r in line.get("configs", []):
    ph = {k: round(v, 3) for k, v in r["phase_ms_per_step"].items()}
    rf = r.get("roofline", {})
    print("%s/%s n_gpus %d  step %.3f ms  %.4g u/s  phases %s" % (r["name"], r["scaling"], r["n_gpus"], r["ms_per_step"], r["value"], ph))
    if rf:
        print("   roofline %s %s frac %.4f (%.1f %s of %.0f) launch %.3f ms" % (rf["kernel"], rf["bound"], rf["frac"], rf["achieved"], rf["unit"], rf["peak"], rf["launch_ms"]))
    if "e2e" in r:
        print("   e2e %.3f ms/step %.4g u/s" % (r["e2e"]["ms_per_step"], r["e2e"]["value"]))
    if "parity" in r:
        p = r["parity"]
        print("   parity", {k: p.get(k) for k in ("ok", "max_rel_err", "draws_bit_exact", "group_counts_exact", "draws_differing_under_glibc_expf", "error") if k in p})
    if "replica_check" in r:
        print("   replicas", r["replica_check"], r.get("nvlink_bytes_per_step", {}).get("message_bytes"))
    if r.get("clocks"):
        print("   clocks", r["clocks"].get("sm_mhz"), r["clocks"].get("reasons"))
    for e in r.get("kernels_one_step", []):
        print("   %-60s x%-3d %8.4f ms %s" % (e["kernel"][:60], e["launches"], e["ms"],
                                              ("%.0f GB/s frac %.3f" % (e["gbs"], e["frac"])) if e.get("gbs") else ""))
