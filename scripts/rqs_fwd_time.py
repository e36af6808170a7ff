"""Time K3 / K4 (spline forward and inverse) on the GPU. Usage: rqs_fwd_time.py [K ...]
MF_RQS_COOP=0 -> generic kernel; MF_RQS_LANES=1/2/4/8 forces the threads-per-row of the cooperative kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import transforms as TR
dev = torch.device("cuda:0")
cases = [(1 << 18, 10, 40, torch.float64), (1 << 20, 7, 10, torch.float64), (1 << 18, 20, 64, torch.float64), (1 << 18, 20, 24, torch.float32), (1 << 22, 7, 10, torch.float32), (1 << 21, 10, 16, torch.float32)]
if len(sys.argv) > 1:
    cases = []
    for k in sys.argv[1:]:
        K = int(k)
        rows = (1 << 30) // (3 * K - 1) // 4  # ~1 GB of fp32 parameters
        cases += [(rows // 8, 8, K, torch.float32), (rows // 16, 8, K, torch.float64)]
tag = "coop=%s L=%s" % (os.environ.get("MF_RQS_COOP", "1"), os.environ.get("MF_RQS_LANES", "auto"))
for (B, F, K, dt) in cases:
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=dt)
    x = ((torch.rand(B, F, device=dev, dtype=dt) * 2.6 - 1.3) * 1.1)
    for name, fn in (("fwd", TR.rqs_forward), ("inv", TR.rqs_inverse)):
        fn(phi, x, bound=1.1); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(4):
            a.record(); fn(phi, x, bound=1.1); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        by = B * F * (3 * K - 1) * phi.element_size()
        print("%s %s %s F=%d K=%d B=%d: %.3f ms, %.0f GB/s" % (name, tag, str(dt)[6:], F, K, B, best, by / best / 1e6))
