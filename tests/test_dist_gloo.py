"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (row sharding, one
all-reduce of the flat fp64 suffstat-delta buffer, identical replicas).  The device work is
replaced by the oracle's update rule here -- this test is about the exchange, not the kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import common_b200 as cb
    import oracle_lib as ol
    from common_b200 import dist as cbd

    orc = ol.load()
    descs = [cb.bb, cb.gp, cb.nich, cb.dd(6)]
    n_total, k = 400, 5
    arr, z = cb.synth.make_dataset(descs, n_total, k, seed=3)           # every rank generates the same data ...
    lo, hi = cbd.shard_rows(n_total, rank, world)                        # ... and owns one contiguous shard
    view = cb.numpy_dataview(arr[lo:hi])
    hp = np.concatenate([orc.flat_hp(d) for d in descs])
    W = orc.ss_total(descs)
    # flat buffer layout of msb_state_delta_buffer: [group counts | suffstats]
    local = np.zeros((k, W)); cnt = np.zeros(k)
    orc.update_rows(descs, hp, local, cnt, view, None, z[lo:hi].astype(np.int32))
    flat = torch.from_numpy(np.concatenate([cnt, _to_additive(orc, descs, local).ravel()]))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                           # the ONE collective of the path
    glob = flat.numpy()
    counts, ss = glob[:k].copy(), glob[k:].reshape(k, W).copy()
    # a sweep: score local rows against the global replica, sample with GLOBAL row ids, all-reduce the deltas
    lp = ol.logprior(counts, 1.0)
    S = orc.score_rows(descs, hp, _ref_repr(orc, descs, ss), lp, view).astype(np.float32)
    u = np.array([orc.philox_u01(73, lo + i, 0) for i in range(hi - lo)], np.float32)
    new = orc.sample_rows(S, u)
    # local deltas in additive form: (state after the local moves) - (state before), per statistic
    before = _ref_repr(orc, descs, ss); after = before.copy(); d_cnt = np.zeros(k)
    orc.update_rows(descs, hp, after, d_cnt, view, z[lo:hi].astype(np.int32), new, prec=64)
    d_ss = _to_additive(orc, descs, after) - ss
    delta = torch.from_numpy(np.concatenate([d_cnt, d_ss.ravel()]))
    dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    merged = glob + delta.numpy()
    gathered = [torch.zeros_like(delta) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(merged))
    assert all(torch.equal(gathered[0], g) for g in gathered)            # replicas stay identical
    if rank == 0:
        np.save(out, np.concatenate([merged, new.astype(np.float64)]))
    else:
        np.save(out + ".r1", new.astype(np.float64))
    dist.destroy_process_group()


def _nich_offsets(orc, descs):
    offs, o = [], 0
    for d in descs:
        m = orc.model(d)
        if d().name() == "nich":
            offs.append(o)
        o += orc.ss_size(m)
    return offs


def _to_additive(orc, descs, ss):
    """(count, mean, count_times_variance) -> (count, sum x, sum x^2): what the device buffer holds and
    the only form that may be summed across ranks"""
    out = ss.copy()
    for o in _nich_offsets(orc, descs):
        n, mean, ctv = ss[:, o], ss[:, o + 1], ss[:, o + 2]
        out[:, o + 1] = n * mean
        out[:, o + 2] = ctv + n * mean * mean
    return out


def _ref_repr(orc, descs, add):
    out = add.copy()
    for o in _nich_offsets(orc, descs):
        n, s1, s2 = add[:, o], add[:, o + 1], add[:, o + 2]
        mean = np.where(n > 0, s1 / np.maximum(n, 1), 0.0)
        out[:, o + 1] = mean
        out[:, o + 2] = np.where(n > 0, np.maximum(s2 - s1 * mean, 0.0), 0.0)
    return out


@pytest.mark.timeout(300)
def test_row_sharded_sweep_equals_single_rank(tmp_path, oracle):
    import common_b200 as cb
    import oracle_lib as ol
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    new1 = np.load(out + ".r1.npy")
    # single-rank run of the same batched sweep
    descs = [cb.bb, cb.gp, cb.nich, cb.dd(6)]
    n_total, k = 400, 5
    arr, z = cb.synth.make_dataset(descs, n_total, k, seed=3)
    view = cb.numpy_dataview(arr)
    hp = np.concatenate([oracle.flat_hp(d) for d in descs])
    ss, cnt = ol.build_suffstats(oracle, descs, hp, view, z, k)
    S = oracle.score_rows(descs, hp, ss, ol.logprior(cnt, 1.0), view).astype(np.float32)
    u = np.array([oracle.philox_u01(73, i, 0) for i in range(n_total)], np.float32)
    new = oracle.sample_rows(S, u)
    W = ss.shape[1]
    # the sharded ranks accumulated float suffstats in a different order: integers exact, floats to 1e-9
    d_ss = np.zeros_like(ss); d_cnt = np.zeros(k)
    oracle.update_rows(descs, hp, d_ss, d_cnt, view, z.astype(np.int32), new)
    want_cnt = cnt + d_cnt
    assert np.array_equal(got[:k], want_cnt)
    new_sharded = np.concatenate([got[k + k * W:], new1])
    assert np.array_equal(new_sharded.astype(np.int32), new)             # draws do not depend on the number of ranks


def test_shard_rows_partition():
    from common_b200 import dist as cbd
    for n, w in [(10, 3), (1_000_000, 8), (7, 8)]:
        spans = [cbd.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
