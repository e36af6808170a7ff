"""SM clock / power while a kernel runs back to back for ~1.5 s (NVML, 5 ms period): are the FP64-heavy maps power-capped?"""
import ctypes, os, sys, threading, time
import torch, pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200._lib import lib, ptr, host_array, stream_ptr
dev = torch.device("cuda:0"); L = lib(); E = 13000.0
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
CASES = {3: [125.25, 172.5, 172.5], 8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]}
N = 1 << 24
def probe(name, f, secs=1.5):
    smp, stop = [], threading.Event()
    def poll():
        while not stop.is_set():
            smp.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
            time.sleep(0.005)
    for _ in range(3): f()
    torch.cuda.synchronize()
    th = threading.Thread(target=poll); th.start()
    t0 = time.time(); n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < secs:
        for _ in range(20): f()
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = a.elapsed_time(b) / n
    clk = sorted(s[0] for s in smp); pw = sorted(s[1] for s in smp); rs = 0
    for s in smp: rs |= s[2]
    print("%-10s %.3f ms/launch  SM clock median %d MHz (min %d max %d)  power median %.0f W max %.0f W  reasons 0x%x  samples %d"
          % (name, ms, clk[len(clk) // 2], clk[0], clk[-1], pw[len(pw) // 2], pw[-1], rs, len(smp)))
for n, m in CASES.items():
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
    mom = torch.empty(N, n + 2, 4, dtype=torch.float64, device=dev)
    w, x1, x2 = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3))
    ps = torch.empty(N, 3 * n - 2, dtype=torch.float64, device=dev); prob = torch.empty(N, dtype=torch.float64, device=dev)
    mh = host_array(m, ctypes.c_double); order = host_array(list(range(n)), ctypes.c_int32); st = stream_ptr(dev)
    fwd = lambda: L.mf_rambo_forward_f64(ptr(u), ctypes.c_int64(N), ctypes.c_int(n), mh, ctypes.c_double(E), None, ctypes.c_uint32(0),
                                         ptr(mom), ptr(w), None, ptr(x1), ptr(x2), None, st)
    inv = lambda: L.mf_rambo_inverse_f64(ctypes.c_void_p(mom.data_ptr() + 64), ctypes.c_int64(4 * (n + 2)), ptr(x1), ptr(x2), ctypes.c_int64(N),
                                         ctypes.c_int(n), order, mh, None, ctypes.c_double(E), ctypes.c_uint32(0), ptr(ps), ptr(prob), None, None, st)
    probe("K1 n=%d" % n, fwd); probe("K2 n=%d" % n, inv)
a = torch.empty(1 << 28, dtype=torch.float32, device=dev); b = torch.empty_like(a)
probe("copy 1GiB", lambda: b.copy_(a))
