"""Which GPUs share a host link? One process, all visible GPUs: pinned host -> device copies of 1 GiB, alone, in pairs and
all together (CUDA events per device, wall clock around the group). Explains the e2e scaling of bench.py on multi-GPU boxes:
the chain's e2e leg streams 1624 B of spline parameters per event from pinned host memory."""
import itertools, subprocess, time
import torch
n = torch.cuda.device_count()
G = 1 << 30
host = [torch.empty(G, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(G, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]
def run(ids, reps=4):
    for i in ids:
        with torch.cuda.device(i), torch.cuda.stream(streams[i]):
            dev[i].copy_(host[i], non_blocking=True)
    for i in ids: torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in ids:
            with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                dev[i].copy_(host[i], non_blocking=True)
    for i in ids: torch.cuda.synchronize(i)
    dt = time.perf_counter() - t0
    return len(ids) * reps * G / dt / 1e9
print("GPUs:", n)
for i in range(n): print("alone   %d: %6.1f GB/s" % (i, run([i])))
for i, j in itertools.combinations(range(n), 2): print("pair  %d,%d: %6.1f GB/s in total" % (i, j, run([i, j])))
print("all %d     : %6.1f GB/s in total" % (n, run(list(range(n)))))
try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
    print(subprocess.run("lspci -tv 2>/dev/null | grep -i -B2 -A2 nvidia | head -60", shell=True, capture_output=True, text=True).stdout)
except Exception as e:
    print("no topology tools:", e)
