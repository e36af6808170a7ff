"""BASELINE configs[2] and configs[3] at their full sizes on one GPU (the bench's `extra` block times smaller instances):
config 3 = spline log_prob direction + inverse sampling, F = 20, K = 64, batch 4 194 304 fp32 with per-event parameters
(61 GB resident), config 4 = the importance-sampling chain over 1e8 points (chunks of 2^22, resident parameter buffer)."""
import os, sys, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import estimator as ES, transforms as TR
from memflow_b200.distributed import ShardedIntegrator
dev = torch.device("cuda:0")
out = {}
def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)
# ---- config 3 ----
B, F, K = 1 << 22, 20, 64
g = torch.Generator(device=dev).manual_seed(7)
phi = torch.empty(B, F, 3 * K - 1, device=dev, dtype=torch.float32)
for i in range(0, B, 1 << 19):
    phi[i:i + (1 << 19)].normal_(generator=g)
x = torch.rand(B, F, device=dev, generator=g) * 2.6 - 1.3
tf = ev_time(lambda: TR.rqs_forward(phi, x, bound=1.1))
ti = ev_time(lambda: TR.rqs_inverse(phi, x, bound=1.1))
y, ladj, lsum = TR.rqs_forward(phi, x, bound=1.1)
xb = TR.rqs_inverse(phi, y, bound=1.1)[0]
by = F * (3 * K - 1) * 4 + 2 * F * 4 + 4
out["config3_F20_K64_fp32_4M"] = {"forward_s": tf, "inverse_s": ti, "forward_events_per_s": B / tf, "inverse_events_per_s": B / ti,
                                  "forward_GBps": by * B / tf / 1e9, "inverse_GBps": by * B / ti / 1e9,
                                  "log_prob_plus_sample_events_per_s": B / (tf + ti),
                                  "round_trip_max_abs": float((xb - x).abs().max()), "tails_fraction": float((x.abs() > 1.1).float().mean())}
del phi, x, y, ladj, lsum, xb
torch.cuda.empty_cache()
# ---- config 4 ----
N, CH, T, Fd, Kd = 100_000_000, 1 << 22, 2, 7, 10
buf = torch.randn(T, CH, Fd, 3 * Kd - 1, device=dev, dtype=torch.float32, generator=g)
smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, bound=1.0, seed=2026, device=dev)
prov = lambda b, e: buf if e - b == CH else buf[:, : e - b].contiguous()
ShardedIntegrator(smp, 4 * CH, CH).run(prov)
torch.cuda.synchronize()
t0 = time.perf_counter()
acc = ShardedIntegrator(smp, N, CH).run(prov)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
I, s, nv = ES.mc_estimate(acc.cpu().tolist(), N)
out["config4_chain_1e8_points"] = {"seconds": dt, "points_per_s": N / dt, "I": I, "sigma": s, "n_valid": nv}
print(json.dumps(out))
