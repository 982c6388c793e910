#!/usr/bin/env python
"""bench.py -- row x group x feature scores/sec of the hot path on N B200s of one node.

A "step" is one full pass of the hot path over one batch of synthetic rows:
parameter build -> score (N x K matrix materialised) -> categorical draw ->
suffstat update (-> one all-reduce of the suffstat deltas when N > 1).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the CPU path, all host threads

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "row_x_group_x_feature_scores_per_sec"
UNIT = "scores/s"

# per-unit byte sizes of SURVEY.md section 8(d) / BASELINE.md section 5
B_X = {"bb": 1, "dd": 1, "gp": 4, "nich": 4}
B_SS = {"bb": 8, "gp": 12, "nich": 12}
B_HP = {"bb": 8, "gp": 8, "nich": 16}


def algorithmic_bytes_score(descs, storage, n, k):
    """B = N sum b_x + K sum b_ss + sum b_hp + 4 N K   (score mode)"""
    bx = bss = bhp = 0
    for d, desc in enumerate(descs):
        m = desc()
        nm = m.name()
        if nm == "dd":
            C = m._param()
            bx += np.dtype(storage[d]).itemsize if storage and storage[d] is not None else 4
            bss += 4 * (C + 1); bhp += 4 * C
        elif nm == "niw":
            dim = m._param()
            bx += 4 * dim; bss += 4 * (1 + dim + dim * dim); bhp += 4 * (2 + dim + dim * dim)
        else:
            bx += B_X[nm]; bss += B_SS[nm]; bhp += B_HP[nm]
    return n * bx + k * bss + bhp + 4 * n * k


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def usable_cores():
    """host threads this process may really use: the affinity mask, capped by the cgroup CPU quota
    (os.cpu_count() reports the machine, not the container)"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def build_workload(args, rank, world):
    import common_b200 as cb
    cfg = cb.synth.config(args.workload)
    n = args.rows or cfg["n"]
    k = args.groups or cfg["k"]
    descs = cfg["models"]
    storage = cfg.get("storage")
    # weak scaling: every rank holds its own n rows of the same planted mixture (stream id = rank)
    arr, z = cb.synth.make_dataset(descs, n, k, seed=73, stream=rank, storage=storage)
    return cfg, descs, storage, n, k, arr, z


def cpu_reference_rate(descs, hp_by_feature, arr, z, k, budget_s, nthreads, f32=True):
    """times the CPU path (the reference's per-(row, group, feature) loop) on a bounded row sample.
    Uses oracle/_ref (the reference's own headers) when it was built, else the C port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common_b200 as cb
    import oracle_lib as ol
    import ref_lib
    orc = ol.load()
    hp = np.concatenate([orc.flat_hp(d, hp_by_feature) for d in descs])
    n_build = min(arr.shape[0], 200_000)
    view_all = cb.numpy_dataview(arr[:n_build])
    ss, counts = ol.build_suffstats(orc, descs, hp, view_all, z[:n_build], k)
    lp = ol.logprior(counts, 1.0)
    ref = ref_lib.load()
    kind = "reference-api" if ref is not None else "port"
    D = len(descs)

    def run(rows):
        view = cb.numpy_dataview(arr[:rows])
        t0 = time.perf_counter()
        if ref is not None:
            ref.score_rows(descs, hp, ss, lp, view, nthreads)
        else:
            orc.score_rows(descs, hp, ss, lp, view, nthreads=nthreads, f32=f32)
        return time.perf_counter() - t0

    rows = max(8, min(256, arr.shape[0]))
    t = run(rows)
    # grow the sample until it fills the budget (bounded by the data we have)
    while t < budget_s / 4 and rows < arr.shape[0]:
        rows = int(min(arr.shape[0], max(rows * 2, rows * (budget_s / 2) / max(t, 1e-6))))
        t = run(rows)
    noop_ns = None
    if ref is not None and hasattr(ref.lib, "ref_noop_overhead_ns"):
        noop_ns = ref.noop_overhead_ns()     # the reference's own noop model: what the plugin API costs per call
    return {"value": rows * k * D / t, "seconds": t, "rows": rows, "kind": kind, "cores": nthreads, "noop_ns": noop_ns,
            "sample": "%d of the workload's rows x K=%d groups x D=%d features, %s" % (
                rows, k, D, "reference headers (models/base.hpp, recarray/dataview.hpp) + restated family maths, libm logf/lgammaf"
                if ref is not None else "C port of the reference loop, libm logf/lgammaf")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    import common_b200 as cb

    if args.impl == "reference":
        if rank != 0:
            return 0
        cfg, descs, storage, n, k, arr, z = build_workload(args, 0, 1)
        name = workload_name(args, cfg, descs, n, k)   # the same workload as the b200 arm; each step times a bounded sample of its rows
        n = min(n, 200_000)
        arr, z = arr[:n], z[:n]
        ncores = usable_cores()
        hpx = cfg.get("hp")
        rates = []
        for i in range(args.warmup + args.steps):
            r = cpu_reference_rate(descs, hpx, arr, z, k, max(2.0, args.cpu_seconds / max(1, args.steps)), ncores)
            if i >= args.warmup:
                rates.append(r)
        val = float(np.mean([r["value"] for r in rates]))
        ms = float(np.mean([r["seconds"] for r in rates]) * 1e3)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": name, "rows_per_step": rates[-1]["rows"], "groups": k, "features": len(descs),
                           "sample": "each step scores rows_per_step of the workload's rows against all K groups x D features"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": rates[-1]["cores"], "kind": "port", "api": rates[-1]["kind"], "sample": rates[-1]["sample"]},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from common_b200 import dist as cbd

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg, descs, storage, n, k, arr, z = build_workload(args, rank, world)
    D = len(descs)

    # the library's own (non-blocking) stream becomes torch's current stream: the timing events, the NCCL
    # collectives and the library's kernels are all ordered on it
    ctx = cb.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    torch.cuda.set_stream(stream)
    view = cb.numpy_dataview(arr)
    st = cb.state(ctx, descs, max_groups=k + 8, cluster_hp={"alpha": 1.0})
    if cfg.get("hp"):
        for d in range(D):
            st.set_component_hp(d, cfg["hp"])
    st.bind(view)
    gids = np.asarray([st.create_group() for _ in range(k)])
    if world > 1:
        cbd.add_values_sharded(st, gids[z], device)   # local rows -> delta buffer -> all-reduce -> every replica applies the sum
    else:
        st.add_values(gids[z])

    def step(i, s_=None):
        # everything is enqueued on the stream; nothing in a step waits for the device
        s_ = s_ or st
        if world > 1:
            r = s_.sweep(seed=73, sweep=i, row_id_offset=rank * n, defer_apply=True, wait=False)
            cbd.allreduce_deltas(s_, device)
        else:
            r = s_.sweep(seed=73, sweep=i, wait=False)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for i in range(warmup):
        step(i)
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"build": 0.0, "score": 0.0, "sample": 0.0, "update": 0.0, "apply": 0.0}
    units = 0
    e0.record(stream)
    for i in range(args.steps):
        r = step(warmup + i)
        units += r["units"]
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    timed = min(args.steps, 64)   # the library keeps the phase events of the last 64 sweeps
    for back in range(timed):
        for kk, v in st.last_timings(back).items():
            phase[kk] += v * args.steps / timed
    launches = ctx.launch_count() - launches0
    clk = None
    if rank == 0:
        # nvidia-smi cannot sample faster than ~100 ms: when the timed region is shorter than that,
        # keep the same step running (untimed) until the sampler has seen the GPU under this load
        extra_t0 = time.perf_counter()
        extra = 0
        while len(clocks.rows) < 5 and time.perf_counter() - extra_t0 < 3.0 and world == 1:
            step(warmup + args.steps + extra)
            torch.cuda.synchronize(device)
            extra += 1
        torch.cuda.synchronize(device)
        clk = clocks.stop()
        clk["sampled_over"] = "timed region" if extra == 0 else "timed region + %d more untimed steps of the same workload" % extra
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    u = torch.tensor([float(units)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    ms_total = float(t.item()); units_all = float(u.item())
    value = units_all / (ms_total * 1e-3)

    # ---- e2e: the same step through the public API from HOST buffers -------------------
    # The reference re-reads its borrowed host rows on every pass (recarray/dataview.hpp:194-217); here one
    # pass over host rows is: H2D of the records (pinned) -> AoS->SoA conversion -> sweep (score + draw +
    # suffstat update, + all-reduce) -> assignments back to pinned host memory.  Groups, hypers and
    # suffstats stay resident in HBM between passes, as they stay resident in the reference's state object.
    e2e = None
    if not args.no_e2e:
        raw, mraw = view.raw()
        pinned = torch.from_numpy(raw).pin_memory()
        out_host = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(2)]
        out_np = [t_.numpy() for t_ in out_host]
        from common_b200.dataview import device_dataview
        dv2 = device_dataview(ctx, data=pinned.data_ptr(), n=n, types=view.types())
        s2 = cb.state(ctx, descs, max_groups=k + 8, cluster_hp={"alpha": 1.0})
        if cfg.get("hp"):
            for d in range(D):
                s2.set_component_hp(d, cfg["hp"])
        s2.bind(dv2)
        g2 = np.asarray([s2.create_group() for _ in range(k)])
        if world > 1:
            cbd.add_values_sharded(s2, g2[z], device)
        else:
            s2.add_values(g2[z])

        dv2.upload(pinned.data_ptr())                # the first pass's records
        s2.prefetch()

        def e2e_step(i):
            s2.refresh()                             # this pass's columns (converted on the copy stream) become current
            dv2.upload(pinned.data_ptr())            # H2D of the NEXT pass's records and their AoS -> SoA conversion into
            s2.prefetch()                            # the second column buffer: copy stream, under this pass's sweep
            rr = step(i, s2)
            s2.assignments_wait()                    # the previous pass's result has landed in pinned memory
            s2.assignments_async(out_np[i & 1])      # D2H of this pass's result, on the copy stream
            return rr["units"]

        for i in range(3):
            e2e_step(1000 + i)
        barrier()
        t0 = time.perf_counter()
        eu = 0
        esteps = max(3, args.steps)
        for i in range(esteps):
            eu += e2e_step(2000 + i)
        s2.assignments_wait()                        # the last pass's result is on the host
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=device, dtype=torch.float64)
        uu = torch.tensor([float(eu)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(uu, op=dist.ReduceOp.SUM)
        e2e = {"value": float(uu.item()) / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(raw.nbytes), "d2h_bytes_per_step": int(n * 8),
               "ms_per_step": float(tt.item()) * 1e3 / esteps, "steps": esteps,
               "what": "per pass: one H2D of the host AoS records (pinned) + AoS->SoA conversion, both for the NEXT pass on a copy stream under this pass's sweep (double-buffered columns) -> sweep (score, draw, suffstat update"
                       + (", all-reduce" if world > 1 else "") + ") -> int64 assignments to pinned host memory (copied on the copy stream, waited for during the next pass; the last one inside the timed region); groups/hypers/suffstats resident in HBM"}
        s2.close(); dv2.close()

    tf32_peak = None
    if rank == 0 and any(d().name() == "niw" for d in descs):
        # TF32 dense peak is not in MEASURED_PEAKS.json: measure it here with a library GEMM (8192^3)
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=device); b = torch.randn(8192, 8192, device=device)
        for _ in range(3):
            a @ b
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            t0.record(); a @ b; t1.record(); torch.cuda.synchronize(device)
            best = min(best, t0.elapsed_time(t1))
        tf32_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        score_ms = phase["score"] / args.steps
        abytes = algorithmic_bytes_score(descs, storage, n, k)
        achieved = abytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
        traffic, wf = None, None
        ctx_sm_count = torch.cuda.get_device_properties(device).multi_processor_count
        try:  # DRAM bytes per launch of this kernel from the committed ncu capture (same workload, same size)
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload.upper())
            if tr and tr.get("rows") == n and world >= 1:
                traffic = tr["dram_bytes_per_launch"]
                wf = tr.get("smem_wavefronts_per_launch")
        except Exception:
            pass
        binding = None
        try:
            if wf and score_ms > 0 and clk and clk.get("sm_mhz"):
                # the pipe that really binds this kernel: one 128-byte shared-memory wavefront per SM per clock
                per_clk = wf / (score_ms * 1e-3 * clk["sm_mhz"] * 1e6 * ctx_sm_count)
                binding = {"resource": "shared-memory wavefronts (LSU data pipe, 1 per SM per clock)",
                           "wavefronts_per_launch": wf, "frac": per_clk, "source": tr["capture"]}
        except Exception:
            pass
        roof = {"kernel": "score_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": abytes,
                "launch_ms": score_ms, "peak_source": peak_src,
                "note": "score kernels reuse every loaded value K times: the binding roof is shared-memory lookup / FP32 issue rate, see DESIGN.md"}
        if binding:
            roof["binding"] = binding
        if tf32_peak:
            dims = [d()._param() for d in descs if d().name() == "niw"]
            flops = sum(2.0 * dd_ * dd_ * n * k for dd_ in dims)   # whitened-GEMM form, SURVEY.md section 8(d)
            ach = flops / (score_ms * 1e-3) / 1e12
            if os.environ.get("MSB_NIW_TF32"):
                roof = {"kernel": "niw_tc_kernel (+ pack, fill)", "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": ach / tf32_peak, "traffic": None, "algorithmic_flops_per_launch": flops, "launch_ms": score_ms,
                        "peak_source": "torch.matmul TF32 8192^3 measured in this run (no TF32 figure in MEASURED_PEAKS.json)",
                        "note": "the kernel issues 3 tf32 products per algorithmic product (hi*hi + hi*lo + lo*hi) for fp32-class accuracy: tensor-pipe work is 3x the algorithmic FLOPs"}
            else:
                f16_peak = float(peaks.get("bf16_tflops", 0.0)) or 2.0 * tf32_peak
                roof = {"kernel": "niw_tc16_kernel (+ colmax, convert, pack)", "bound": "tensor", "achieved": ach, "peak": f16_peak, "unit": "TFLOP/s",
                        "frac": ach / f16_peak, "traffic": None, "algorithmic_flops_per_launch": flops, "launch_ms": score_ms,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; fp16 runs at the bf16 rate)" if peaks.get("bf16_tflops") else "2 x the TF32 rate measured in this run",
                        "tf32_tflops_measured_here": tf32_peak,
                        "note": "the kernel issues 3 fp16 products per algorithmic product (hi*hi + hi*lo + lo*hi, scaled operands) for fp32-class accuracy: tensor-pipe work is 3x the algorithmic FLOPs, i.e. pipe utilisation = 3 x frac"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args, cfg, descs, n, k), "rows_per_gpu": n, "groups": k, "features": D,
                           "step": "param build + score (N x K fp32 materialised) + sample + suffstat update" + (" + all-reduce" if world > 1 else ""),
                           "l2": "no explicit flush: each step streams the %.0f MB score matrix (> 126 MB L2)" % (4.0 * n * st.last_scores()[1] / 1e6)},
                "phase_ms_per_step": {kk: v / args.steps for kk, v in phase.items()},
                "roofline": roof,
                "gpu_launches": int(launches), "clocks": clk}
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu and world == 1:
            c = cpu_reference_rate(descs, cfg.get("hp"), arr, z, k, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": c["value"], "unit": UNIT, "cores": c["cores"],
                                    "kind": "port", "api": c["kind"], "sample": c["sample"],
                                    "api_overhead_ns_per_call": c["noop_ns"],
                                    "api_overhead_note": "the reference's own noop model (models/noop.hpp) in the perf_group.cpp loop: "
                                                         "the floor the plugin API itself puts under any per-value implementation"}
        print(json.dumps(line))
    st.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def workload_name(args, cfg, descs, n, k):
    names = {}
    for d in descs:
        m = d()
        key = m.name() + ("(%d)" % m._param() if m._param() else "")
        names[key] = names.get(key, 0) + 1
    return "%s: %s, N=%d rows x K=%d groups x D=%d features" % (
        args.workload.upper(), " + ".join("%d x %s" % (c, nm) for nm, c in names.items()), n, k, len(descs))


if __name__ == "__main__":
    sys.exit(main())
