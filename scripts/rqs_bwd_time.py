"""Time the spline backward (autograd through rqs_forward) on the GPU. Usage: rqs_bwd_time.py [K ...]
MF_RQS_BWD_LANES=1/2/4/8 forces the threads-per-row of the backward kernel (read once per process)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import transforms as TR
dev = torch.device("cuda:0")
cases = [(1 << 18, 20, 64, torch.float32), (1 << 20, 7, 10, torch.float32), (1 << 18, 10, 40, torch.float64)]
if len(sys.argv) > 1:
    cases = []
    for k in sys.argv[1:]:
        K = int(k)
        rows = (1 << 29) // (3 * K - 1) // 4  # ~0.5 GB of fp32 parameters
        cases += [(rows // 8, 8, K, torch.float32), (rows // 16, 8, K, torch.float64)]
for (B, F, K, dt) in cases:
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=dt, requires_grad=True)
    x = ((torch.rand(B, F, device=dev, dtype=dt) * 2.6 - 1.3) * 1.1)
    y, l, ls = TR.rqs_forward(phi, x, bound=1.1)
    loss = ls.sum()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(4):
        phi.grad = None
        a.record(); loss.backward(retain_graph=True); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    by = 2 * B * F * (3 * K - 1) * phi.element_size()
    print("bwd L=%s %s F=%d K=%d B=%d: %.3f ms, %.0f GB/s (params read + grads written)" % (
        os.environ.get("MF_RQS_BWD_LANES", "auto"), str(dt)[6:], F, K, B, best, by / best / 1e6))
