"""BASELINE config 5 shape on one GPU: tt-bar-H + 1 jet (n=4), production flow F=10, K=40, T=6."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import estimator as ES
dev = torch.device("cuda:0")
n, T, K, F, N = 4, 6, 40, 10, 1 << 19
phi = torch.randn(T, N, F, 3 * K - 1, device=dev, dtype=torch.float32)
smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], 13000.0, seed=1, device=dev)
u = torch.empty(N, F, dtype=torch.float64, device=dev); lq = torch.empty(N, dtype=torch.float64, device=dev)
acc = ES.new_accumulator(dev)
for i in range(3): smp.sample_flow(phi, u=u, logq=lq); smp.weigh(u, lq, acc=acc)
torch.cuda.synchronize()
ta, tb = [], []
for i in range(20):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); smp.sample_flow(phi, event_offset=i * N, u=u, logq=lq); e1.record(); smp.weigh(u, lq, acc=acc); e2.record()
    torch.cuda.synchronize()
    ta.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
ta.sort(); tb.sort()
a, b = ta[len(ta) // 2], tb[len(tb) // 2]
by = T * F * (3 * K - 1) * 4
print("config5 shape: A %.3f ms (%.0f GB/s, %.1f%% of HBM)  B %.3f ms  -> %.3e points/s per GPU" % (a, by * N / a / 1e6, by * N / a / 1e6 / 64.596, b, N / (a + b) * 1e3))
