"""Training-size batch (B = 1024): the same step (phase-space map n = 4 forward + inverse, six float64 spline transforms
F = 10, K = 40 forward + backward) launched eagerly from Python and replayed from ONE captured CUDA graph."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import transforms as TR
from memflow_b200.phasespace.phasespace import PhaseSpace
dev = torch.device("cuda:0")
B, F, K, T = 1024, 10, 40, 6
ps = PhaseSpace(13000.0, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor([125.25, 172.5, 172.5, 1e-5], dtype=torch.float64), dev=dev)
u = torch.rand(B, 10, dtype=torch.float64, device=dev).clamp_(1e-6, 1 - 1e-6)
phis = [torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64, requires_grad=True) for _ in range(T)]
x0 = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.0 - 1.0)
def step():
    mom, w, x1, x2 = ps.get_momenta_from_ps(u)
    back, prob = ps.get_ps_from_momenta(mom[:, 2:], x1, x2)
    x, tot = x0, 0.0
    for p in phis:
        x, ladj, lsum = TR.rqs_forward(p, x, bound=1.1)
        tot = tot + lsum.sum()
    grads = torch.autograd.grad(tot + x.sum(), phis)
    return back, grads
def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
eager = timeit(step)
from memflow_b200.graphs import capture
g = capture(step)
graph = timeit(g)
print("B=%d: K1 + K2 (n=4) + %d x (spline forward + backward, float64 F=%d K=%d): eager %.0f us / step, one graph replay %.0f us / step (%.1fx)"
      % (B, T, F, K, eager, graph, eager / graph))
