import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200._lib import lib, ptr, stream_ptr
dev = torch.device("cuda:0"); L = lib()
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for (B, F, K, bound) in ((1 << 20, 20, 64, 1.1), (1 << 21, 10, 40, 1.1), (1 << 22, 7, 10, 1.0)):
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float32)
    x = (torch.rand(B, F, device=dev) * 2.6 - 1.3) * bound
    y = torch.empty_like(x); ladj = torch.empty_like(x); ls = torch.empty(B, device=dev)
    st = stream_ptr(dev)
    by = (F * (3 * K - 1) + 2 * F + 1) * 4
    for name, fn in (("fwd", L.mf_rqs_forward_f32), ("inv", L.mf_rqs_inverse_f32)):
        f = lambda: fn(ptr(phi), ptr(x), ctypes.c_int64(B), ctypes.c_int(F), ctypes.c_int(K), ctypes.c_float(bound), ctypes.c_float(1e-3),
                       ptr(y), None, ptr(ls), None, st)
        assert f() == 0
        t = timeit(f)
        print("RQS f32 F=%d K=%d B=%d %s: %.3f ms  %.3e ev/s  %.0f GB/s (%.1f%% of HBM)" % (F, K, B, name, t * 1e3, B / t, by * B / t / 1e9, by * B / t / 64.596e9))
