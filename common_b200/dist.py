"""Multi-GPU plumbing: rows are sharded across ranks, every rank keeps a full replica of
the K x D suffstats, and one all-reduce per sweep sums the per-group suffstat deltas
(SURVEY.md section 8e).  torch.distributed (NCCL over NVLink on the box, gloo in the CPU
tests) is only the transport; the buffers are the library's own device memory.
"""
import numpy as np


class _DevicePtr(object):
    """__cuda_array_interface__ view of a device buffer owned by the C library"""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def as_tensor(ptr, n, device, typestr="<f8"):
    import torch
    return torch.as_tensor(_DevicePtr(ptr, n, typestr), device=device)


def shard_rows(n_total, rank, world):
    """contiguous row shard [lo, hi) of rank (rows are independent given frozen suffstats)"""
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def allreduce_deltas(st, device):
    """sum the flat fp64 delta buffer over all ranks, then apply it on every replica"""
    import torch.distributed as dist
    ptr32, n = st.delta_buffer_i32()
    if ptr32:  # every suffstat is a count of rows: exact int32 deltas, half the bytes on the wire
        t = as_tensor(ptr32, n, device, "<i4")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        st.delta_from_i32()
    else:
        ptr, n = st.delta_buffer()
        t = as_tensor(ptr, n, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    st.apply_deltas()


def add_values_sharded(st, gids, device):
    """replica initialisation through the delta buffer: every rank adds its own rows (deferred), the deltas are
    summed, every replica applies the sum.  Unlike summing the suffstat buffers this never touches per-group
    parameters that are not sums over rows (bbnc's p)."""
    st.add_values(gids, defer_apply=True)
    allreduce_deltas(st, device)


def allreduce_suffstats(st, device):
    """replica initialisation: every rank added its own rows; the sum is the global state"""
    import torch.distributed as dist
    ptr, n = st.suffstat_buffer()
    t = as_tensor(ptr, n, device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    st.apply_deltas()  # no pending deltas: refreshes the host-side group counts


def merge_deltas_host(deltas):
    """what the all-reduce computes, on host arrays (used by the gloo CPU tests)"""
    return np.sum(np.stack(deltas, 0), axis=0)
