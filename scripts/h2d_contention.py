"""Host -> device copy bandwidth from pinned memory, one rank alone and all ranks at once (torchrun): what bounds the
end-to-end pass over host rows when every GPU of the box streams its records at the same time."""
import os, time, json
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nbytes = 32 << 20                                   # C2's records of one pass: 1M rows x 32 bytes
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
def run(active, reps=50):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    return reps * nbytes / dt / 1e9 if active else 0.0
run(True, 5)
alone = run(rank == 0)
together = run(True)
t = torch.tensor([alone, together], device="cuda", dtype=torch.float64)
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    tog = [float(o[1]) for o in out]
    print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "rank0_alone_GBps": round(float(out[0][0]), 2),
                      "per_rank_together_GBps": [round(x, 2) for x in tog], "aggregate_together_GBps": round(sum(tog), 2),
                      "ms_per_32MB_copy_together": round(nbytes / (min(tog) * 1e9) * 1e3, 3)}))
dist.destroy_process_group()
