"""One process per GPU (torchrun): every rank copies its own pinned host buffer to its own GPU at the same time.
Compares with scripts/h2d_topology.py (one process drives all GPUs). For each rank the pinned buffer is allocated and first
touched while the process is bound to one CPU block after another, to see whether WHERE the pinned pages live matters
(the container reports a single NUMA node).
usage: python -m torch.distributed.run --nproc-per-node N scripts/h2d_ranks.py"""
import os, time
import torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
G = 1 << 30
devbuf = torch.empty(G, dtype=torch.uint8, device="cuda")
allcpus = sorted(os.sched_getaffinity(0))
nb = 4
blocks = [allcpus[i * len(allcpus) // nb:(i + 1) * len(allcpus) // nb] for i in range(nb)]
def measure(host, reps=6):
    devbuf.copy_(host, non_blocking=True); torch.cuda.synchronize()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): devbuf.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([reps * G / dt / 1e9], device="cuda", dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]
res = {}
for name, cpus in [("all", allcpus)] + [("cpus %d-%d" % (b[0], b[-1]), b) for b in blocks] + [("rank block", blocks[lr % nb])]:
    os.sched_setaffinity(0, cpus)
    host = torch.empty(G, dtype=torch.uint8).pin_memory()
    host.fill_(1)                       # first touch under this binding
    r = measure(host)
    if rank == 0:
        print("%-12s per rank GB/s: %s   total %.1f" % (name, " ".join("%5.1f" % x for x in r), sum(r)), flush=True)
    del host
os.sched_setaffinity(0, allcpus)
# two ranks at a time
host = torch.empty(G, dtype=torch.uint8).pin_memory(); host.fill_(1)
for active in ([0], [0, 1], [0, 1, 2], list(range(world))):
    if max(active) >= world: continue
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank in active:
        for _ in range(6): devbuf.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([6 * G / dt / 1e9 if rank in active else 0.0], device="cuda", dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        print("active ranks %s: %s" % (active, " ".join("%5.1f" % float(o.item()) for o in out)), flush=True)
dist.destroy_process_group()
