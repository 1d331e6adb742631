"""What the box's PCIe links give when every GPU copies to pinned host memory at once (run under torchrun): the ceiling of
bench.py's end-to-end number at N GPUs. Prints per-rank and aggregate GB/s for plain cudaMemcpyAsync D2H of 1 GiB blocks."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nb = 1 << 30
    d = torch.empty(nb, dtype=torch.uint8, device="cuda")
    h = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for i in range(2):
        h[i].copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 12
    t0 = time.perf_counter()
    for i in range(reps):
        h[i & 1].copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([reps * nb / dt / 1e9], dtype=torch.float64, device="cuda")
    allg = [torch.zeros_like(gbs) for _ in range(world)]
    if world > 1:
        dist.all_gather(allg, gbs)
    else:
        allg = [gbs]
    if rank == 0:
        v = [float(x.item()) for x in allg]
        print("d2h_ceiling world=%d: per rank GB/s %s, aggregate %.1f GB/s" % (world, ["%.1f" % x for x in v], sum(v)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
