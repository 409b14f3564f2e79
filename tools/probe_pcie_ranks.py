"""Raw concurrent host<->device copy ceiling with every rank of a node copying at once (torchrun): each rank moves the
bench's 78.6 MB in and 78.6 MB out per repetition on two streams, all ranks between the same two barriers.  What the
e2e figure of an N-GPU run can reach at best on this box's host memory system."""
import os, sys, time, json, torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 8192 * 300 * 8
xh, oh = (torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2))
d1, d2 = (torch.empty(n, device=dev) for _ in range(2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d1.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2):
        oh.copy_(d2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for mode, fn in (("H2D + D2H", both), ("H2D only", lambda: d1.copy_(xh, non_blocking=True)), ("D2H only", lambda: oh.copy_(d2, non_blocking=True))):
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        per_dir = n * 4 / (ms.item() * 1e-3) / 1e9
        print(json.dumps({"ranks": world, "mode": mode, "ms_per_rep_max_over_ranks": round(ms.item(), 3),
                          "GBps_per_direction_per_rank": round(per_dir, 1), "GBps_per_direction_all_ranks": round(per_dir * world, 1)}), flush=True)
if world > 1:
    dist.destroy_process_group()
