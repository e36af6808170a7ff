"""Under torchrun: concurrent pinned H2D rate per rank (plain copies, 1 and 2 streams, before / after CPU binding) and the
HostPipeline e2e rate with 1 and 2 copy streams. Explains the e2e scaling numbers of bench.py."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from memflow_b200 import estimator as ES
from memflow_b200.distributed import HostPipeline, bind_to_gpu_cpus, gpu_cpu_affinity
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def report(tag, gbps):
    t = torch.tensor([gbps], dtype=torch.float64, device=dev)
    lo, hi = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0: print("%-52s per rank min %.1f max %.1f GB/s" % (tag, lo.item(), hi.item()), flush=True)
NB = 1 << 30
def plain(tag, nstreams):
    h = torch.empty(NB, dtype=torch.uint8).pin_memory(); d = torch.empty(NB, dtype=torch.uint8, device=dev)
    ss = [torch.cuda.Stream() for _ in range(nstreams)]
    per = NB // nstreams
    def go():
        for i, s in enumerate(ss):
            with torch.cuda.stream(s): d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
    go(); barrier()
    t0 = time.perf_counter()
    for _ in range(8): go()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    barrier()
    report(tag, 8 * NB / dt / 1e9)
if rank == 0: print("affinity before:", len(os.sched_getaffinity(0)), "cpus; gpu", gpu_cpu_affinity(local)[0], flush=True)
plain("plain H2D 1 stream, no binding", 1); plain("plain H2D 2 streams, no binding", 2)
b = bind_to_gpu_cpus(local)
if rank == 0: print("bound:", b, flush=True)
plain("plain H2D 1 stream, bound", 1); plain("plain H2D 2 streams, bound", 2)
T, CH, F, PW = 2, 1 << 20, 7, 29
smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, seed=1, device=dev)
for ncs in (2, 1, 2, 1):
    pipe = HostPipeline(smp, (T, CH, F, PW), dtype=torch.float32, n_copy_streams=ncs, device=dev)
    host = pipe.pinned_chunk(); host.normal_()
    acc = ES.new_accumulator(dev); res = torch.empty(3, dtype=torch.float64).pin_memory()
    def run(n, off):
        for i in range(n):
            pipe.submit(host, off + i * CH, acc); res.copy_(acc, non_blocking=True)
        torch.cuda.synchronize()
    run(2, 0); barrier()
    for rep in range(2):
        t0 = time.perf_counter(); run(12, 2 * CH); dt = time.perf_counter() - t0
        barrier()
        report("HostPipeline %d copy stream(s) rep %d: e2e H2D" % (ncs, rep), 12 * host.numel() * 4 / dt / 1e9)
    # the same chunk by one plain copy + the chain, no pipeline
    dbuf = pipe.bufs[0]
    def plain_step():
        dbuf.copy_(host, non_blocking=True); smp.run_chunk(dbuf, event_offset=0, acc=acc)
    plain_step(); barrier()
    t0 = time.perf_counter()
    for _ in range(6): plain_step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    barrier()
    report("   one copy_ of the whole chunk + chain, serial", 6 * host.numel() * 4 / dt / 1e9)
    del pipe, host, dbuf
if world > 1: dist.destroy_process_group()
