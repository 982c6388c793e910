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


class NcclComm(object):
    """an ncclComm_t created through the C ABI (msb_nccl_comm_create): the collective of the path then runs inside the
    library (msb_state_allreduce_deltas) on the context's stream.  The 128-byte unique id travels from rank 0 to the
    other ranks through ``exchange`` (default: torch.distributed's object broadcast -- plumbing only)."""

    def __init__(self, ctx, rank, world, exchange=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        buf = C.create_string_buffer(_lib.NCCL_UNIQUE_ID_BYTES)
        if rank == 0:
            _lib.check(lib.msb_nccl_unique_id(buf))
        raw = bytes(buf.raw)
        if exchange is None:
            import torch.distributed as dist
            box = [raw]
            dist.broadcast_object_list(box, src=0)
            raw = box[0]
        else:
            raw = exchange(raw)
        h = C.c_void_p()
        _lib.check(lib.msb_nccl_comm_create(ctx.handle, int(world), int(rank), raw, C.byref(h)))
        self.handle = h
        self.rank, self.world = rank, world

    def close(self):
        if self.handle:
            from . import _lib
            _lib.load().msb_nccl_comm_destroy(self.handle)
            self.handle = None


def nccl_version():
    import ctypes as C
    from . import _lib
    v = C.c_int()
    _lib.check(_lib.load().msb_nccl_version(C.byref(v)))
    return v.value


def allreduce_deltas(st, device, comm=None, global_rows=0):
    """sum the flat delta buffer over all ranks, then apply it on every replica.  With ``comm`` (NcclComm) the
    collective runs inside the library; without, torch.distributed carries it (gloo in the CPU tests)."""
    if comm is not None:
        st.allreduce_deltas(comm, global_rows)
        return
    import torch
    import torch.distributed as dist
    ptr32, n = st.delta_buffer_i32()
    # The int32 form is the library's LOCAL offer (every suffstat a count, local rows < 2^31); the sum runs over all
    # ranks, so the GLOBAL row count has to fit as well, and every rank must make the same choice or the collective
    # mismatches: the ranks agree on the minimum of their offers first.
    use32 = torch.tensor([1 if (ptr32 and int(global_rows) < 2 ** 31) else 0], device=device, dtype=torch.int32)
    dist.all_reduce(use32, op=dist.ReduceOp.MIN)
    if int(use32.item()):  # exact int32 deltas, half the bytes on the wire
        t = as_tensor(ptr32, n, device, "<i4")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        st.delta_from_i32()
    else:
        ptr, n = st.delta_buffer()
        t = as_tensor(ptr, n, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    st.apply_deltas()


def add_values_sharded(st, gids, device, comm=None, global_rows=0):
    """replica initialisation through the delta buffer: every rank adds its own rows (deferred), the deltas are
    summed, every replica applies the sum.  Unlike summing the suffstat buffers this never touches per-group
    parameters that are not sums over rows (bbnc's p)."""
    st.add_values(gids, defer_apply=True)
    allreduce_deltas(st, device, comm, global_rows)


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
