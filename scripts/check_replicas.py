#!/usr/bin/env python
"""Multi-GPU correctness on real devices (run with torchrun, one rank per GPU).

After a few row-sharded sweeps with the delta all-reduce
  * every rank must hold bit-identical suffstats (replicas of one run apply the same reduced buffer);
  * the all-reduce inside the C ABI (msb_state_allreduce_deltas, the library's own ncclComm_t) and the one carried by
    torch.distributed must give the same state, and so must the int32 and the fp64 delta paths;
  * the group sizes must add up to the global row count;
  * the sharded run must equal ONE GPU sweeping all the rows: draws are keyed by the global row id, so the
    assignments are identical, and for count-valued states the suffstat buffer is identical bit for bit.
Prints one line per case and "all ok" / "FAILED"; exit status 0 / 1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import common_b200 as cb  # noqa: E402
from common_b200 import dist as cbd  # noqa: E402

N_PER_RANK, K, SWEEPS = 40000, 12, 3


def run(descs, rank, world, device, ctx, comm):
    """comm: NcclComm -> collective inside the library; None -> torch.distributed"""
    n = N_PER_RANK
    arr, z = cb.synth.make_dataset(descs, n, K, seed=5, stream=rank)
    st = cb.state(ctx, descs, max_groups=K + 2, cluster_hp={"alpha": 1.0})
    st.bind(cb.numpy_dataview(arr))
    gids = np.asarray([st.create_group() for _ in range(K)])
    cbd.add_values_sharded(st, gids[z], device, comm, n * world)
    for it in range(SWEEPS):
        st.sweep(seed=9, sweep=it, row_id_offset=rank * n, defer_apply=True, wait=False)
        cbd.allreduce_deltas(st, device, comm, n * world)
    ptr, cnt = st.suffstat_buffer()
    t = cbd.as_tensor(ptr, cnt, device).clone()
    torch.cuda.synchronize(device)
    all_t = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    same = all(torch.equal(all_t[0], x) for x in all_t)
    sizes = [st.groupsize(int(g)) for g in gids]
    assign = np.searchsorted(gids, st.assignments())
    st.close()
    return same, t.cpu().numpy(), sizes, assign


def run_single(descs, world, ctx):
    """the same rows on one GPU: rank r's shard is stream r of the generator"""
    parts = [cb.synth.make_dataset(descs, N_PER_RANK, K, seed=5, stream=r) for r in range(world)]
    arr = np.concatenate([p[0] for p in parts])
    z = np.concatenate([p[1] for p in parts])
    st = cb.state(ctx, descs, max_groups=K + 2, cluster_hp={"alpha": 1.0})
    st.bind(cb.numpy_dataview(arr))
    gids = np.asarray([st.create_group() for _ in range(K)])
    st.add_values(gids[z])
    for it in range(SWEEPS):
        st.sweep(seed=9, sweep=it)
    ptr, cnt = st.suffstat_buffer()
    torch.cuda.synchronize()
    buf = cbd.as_tensor(ptr, cnt, torch.device("cuda", ctx.device)).cpu().numpy().copy()
    assign = np.searchsorted(gids, st.assignments())
    st.close()
    return buf, assign


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    ctx = cb.Context(local)
    torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream(), device=device))
    comm = cbd.NcclComm(ctx, rank, world)
    if rank == 0:
        print("NCCL %d through the C ABI, %d ranks" % (cbd.nccl_version(), world))
    ok = True
    for name, descs in (("counts only (int32 deltas)", [cb.dd(40), cb.bb, cb.dd(7), cb.bb]),
                        ("mixed (fp64 deltas)", [cb.dd(9), cb.nich, cb.gp, cb.bb, cb.bnb, cb.bbnc])):
        counts_only = "int32" in name
        same, a, sizes, assign = run(descs, rank, world, device, ctx, comm)          # collective inside the C ABI
        same_t, a_t, sizes_t, _ = run(descs, rank, world, device, ctx, None)        # collective through torch.distributed
        os.environ["MSB_NO_I32_DELTAS"] = "1"                                       # read by the torch path's buffer call
        same2, b, sizes2, _ = run(descs, rank, world, device, ctx, None)
        del os.environ["MSB_NO_I32_DELTAS"]
        # real-valued moments are accumulated with fp64 atomics, whose order (hence last-bit rounding) differs between
        # two runs; replicas of ONE run are still bit-identical (they apply the same all-reduced buffer)
        close = (lambda x, y: np.array_equal(x, y)) if counts_only else (lambda x, y: np.allclose(x, y, rtol=1e-11, atol=1e-9))
        abi_eq_torch = close(a, a_t)
        i32_eq_f64 = close(a_t, b)
        good = same and same_t and same2 and abi_eq_torch and i32_eq_f64 and sizes == sizes_t == sizes2 and sum(sizes) == N_PER_RANK * world
        # against one GPU sweeping all the rows (bbnc draws its per-group p from the group seed: identical on every replica)
        single_ok = True
        frac = 1.0
        if rank == 0:
            s_buf, s_assign = run_single(descs, world, ctx)
        gathered = [None] * world
        dist.all_gather_object(gathered, assign)
        if rank == 0:
            sharded_assign = np.concatenate(gathered)
            frac = float((sharded_assign == s_assign).mean())
            if counts_only:
                single_ok = bool(np.array_equal(s_buf, a) and frac == 1.0)
            else:
                single_ok = bool(np.allclose(s_buf, a, rtol=1e-9, atol=1e-6) and frac >= 0.999)
        flag = torch.tensor([1 if (good and single_ok) else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(flag.item())
        if rank == 0:
            print("%-28s replicas identical: %s / %s / %s, C-ABI == torch collective: %s, int32 == fp64 path: %s, sum of group sizes %d, "
                  "sharded == 1 GPU: %s (assignments equal on %.6f of the rows)" % (
                      name, same, same_t, same2, abi_eq_torch, i32_eq_f64, sum(sizes), single_ok, frac))
    if rank == 0:
        print("all ok" if ok else "FAILED")
    comm.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
