"""BASELINE configs[4] at full size: the sharded MEM-style integral over 1e10 tt-bar-H + 1 jet points (n = 4, production
flow shape T = 6, F = 10, K = 40), events sharded over the ranks, ONE all-reduce of (sum w, sum w^2, n_valid) at the end.
Synthetic per-event spline parameters from a resident buffer of one chunk (the hyper-network is out of scope), squared
matrix element = pdf = 1.   usage: [torchrun --nproc-per-node N] scripts/config5_full.py [n_total] [float32|float64]"""
import os, sys, time, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import estimator as ES
from memflow_b200.distributed import ShardedIntegrator
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "float64") else torch.float32
T, F, K = 6, 10, 40
chunk = (1 << 19) if dt == torch.float32 else (1 << 18)
g = torch.Generator(device=dev).manual_seed(7)
phi = torch.randn(T, chunk, F, 3 * K - 1, device=dev, dtype=dt, generator=g)
smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], 13000.0, bound=1.0, seed=2026, device=dev)
integ = ShardedIntegrator(smp, n_total, chunk, rank=rank, world_size=world)
provider = lambda b, e: phi if e - b == chunk else phi[:, : e - b].contiguous()
acc = ES.new_accumulator(dev)
integ_warm = ShardedIntegrator(smp, 4 * chunk * world, chunk, rank=rank, world_size=world)
integ_warm.run(provider)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
tot = integ.run(provider, acc=acc)
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt_s = time.perf_counter() - t0
I, sigma, nvalid = ES.mc_estimate(tot.cpu().tolist(), n_total)
if rank == 0:
    print(json.dumps({"config": "BASELINE configs[4]: n=4, T=6, F=10, K=40", "phi_dtype": str(dt)[6:], "n_total": n_total, "n_gpus": world,
                      "chunk_events": chunk, "seconds": dt_s, "points_per_s": n_total / dt_s, "I": I, "sigma": sigma, "n_valid": nvalid,
                      "acc": tot.cpu().tolist()}))
if world > 1: dist.destroy_process_group()
