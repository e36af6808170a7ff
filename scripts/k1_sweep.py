"""Times K1/K2 through the C ABI directly (pre-allocated outputs, no wrapper allocations)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200._lib import lib, ptr, host_array, stream_ptr
dev = torch.device("cuda:0")
E = 13000.0
CASES = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5], 5: [125.25, 172.5, 172.5, 1e-5, 0.0],
         6: [4.7, 4.7, 80.4, 172.5, 0.0, 1e-5], 7: [4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 172.5], 8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]}
if len(sys.argv) > 1:
    CASES = {int(a): CASES[int(a)] for a in sys.argv[1:]}
else:
    CASES = {n: CASES[n] for n in (3, 4, 8)}
N = 1 << 24
L = lib()
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for n, m in CASES.items():
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
    mom = torch.empty(N, n + 2, 4, dtype=torch.float64, device=dev)
    w, x1, x2 = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3))
    ps = torch.empty(N, 3 * n - 2, dtype=torch.float64, device=dev); prob = torch.empty(N, dtype=torch.float64, device=dev)
    mh = host_array(m, ctypes.c_double); order = host_array(list(range(n)), ctypes.c_int32)
    st = stream_ptr(dev)
    fwd = lambda: L.mf_rambo_forward_f64(ptr(u), ctypes.c_int64(N), ctypes.c_int(n), mh, ctypes.c_double(E), None, ctypes.c_uint32(0),
                                         ptr(mom), ptr(w), None, ptr(x1), ptr(x2), None, st)
    inv = lambda: L.mf_rambo_inverse_f64(ctypes.c_void_p(mom.data_ptr() + 64), ctypes.c_int64(4 * (n + 2)), ptr(x1), ptr(x2), ctypes.c_int64(N),
                                         ctypes.c_int(n), order, mh, None, ctypes.c_double(E), ctypes.c_uint32(0), ptr(ps), ptr(prob), None, None, st)
    assert fwd() == 0
    tf = timeit(fwd); assert inv() == 0; ti = timeit(inv)
    bf, bi = (3 * n - 2 + 4 * (n + 2) + 3) * 8, (4 * n + 2 + 3 * n - 2 + 1) * 8
    print("n=%d  K1 %.3f ms %.3e ev/s %.0f GB/s (%.1f%%) | K2 %.3f ms %.3e ev/s %.0f GB/s (%.1f%%) | rt max %.1e"
          % (n, tf * 1e3, N / tf, bf * N / tf / 1e9, bf * N / tf / 64.596e9, ti * 1e3, N / ti, bi * N / ti / 1e9, bi * N / ti / 64.596e9,
             float((ps - u).abs().max())))
