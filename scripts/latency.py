"""Per-call latency of the small-batch paths (training calls the kernels with B ~ 1e3): wall time per call through the
Python API and through the bare C ABI, CUDA-event time per call."""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import transforms as TR
from memflow_b200._lib import lib, ptr, stream_ptr, host_array
from memflow_b200.phasespace.phasespace import PhaseSpace
dev = torch.device("cuda:0")
def wall(f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6
B, F, K = 1024, 10, 40
for dt in (torch.float32, torch.float64):
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=dt); x = torch.rand(B, F, device=dev, dtype=dt) * 2 - 1
    y = torch.empty_like(x); l = torch.empty_like(x); ls = torch.empty(B, device=dev, dtype=dt)
    L = lib(); fn = L.mf_rqs_forward_f32 if dt == torch.float32 else L.mf_rqs_forward_f64
    ct = ctypes.c_float if dt == torch.float32 else ctypes.c_double
    st = stream_ptr(dev)
    cabi = lambda: fn(ptr(phi), ptr(x), ctypes.c_int64(B), ctypes.c_int(F), ctypes.c_int(K), ct(1.1), ct(1e-3), ptr(y), ptr(l), ptr(ls), None, st)
    print("rqs_forward %s B=%d F=%d K=%d: python API %.1f us/call, C ABI %.1f us/call" % (str(dt)[6:], B, F, K, wall(lambda: TR.rqs_forward(phi, x, bound=1.1)), wall(cabi)))
ps = PhaseSpace(13000.0, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor([125.25, 172.5, 172.5, 0.0], dtype=torch.float64), dev=dev)
u = torch.rand(B, 10, dtype=torch.float64, device=dev)
print("get_momenta_from_ps n=4 B=%d: python API %.1f us/call" % (B, wall(lambda: ps.get_momenta_from_ps(u))))
mom, w, x1, x2 = ps.get_momenta_from_ps(u)
print("get_ps_from_momenta n=4 B=%d: python API %.1f us/call" % (B, wall(lambda: ps.get_ps_from_momenta(mom[:, 2:], x1, x2))))
