import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import transforms as TR
dev = torch.device("cuda:0")
for (B, F, K, dt) in ((1 << 18, 20, 64, torch.float32), (1 << 20, 7, 10, torch.float32), (1 << 18, 10, 40, torch.float64)):
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=dt, requires_grad=True)
    x = ((torch.rand(B, F, device=dev, dtype=dt) * 2.6 - 1.3) * 1.1)
    y, l, ls = TR.rqs_forward(phi, x, bound=1.1)
    loss = ls.sum()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss.backward(retain_graph=True); torch.cuda.synchronize(); phi.grad = None
    a.record(); loss.backward(retain_graph=True); b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b)
    by = 2 * B * F * (3 * K - 1) * phi.element_size()
    print("bwd %s F=%d K=%d B=%d: %.3f ms, %.0f GB/s (params read + grads written)" % (str(dt)[6:], F, K, B, t, by / t / 1e6))
