"""Times the fused chain for a few launch-bound variants (MF_CHAIN_MINB is read once per process)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from memflow_b200 import estimator as ES
    dev = torch.device("cuda:0")
    N = 1 << 22
    phi = torch.randn(2, N, 7, 29, device=dev, dtype=torch.float32)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, seed=1, device=dev)
    acc = ES.new_accumulator(dev)
    for i in range(3): smp.run_chunk(phi, acc=acc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20): smp.run_chunk(phi, event_offset=i * N, acc=acc)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20
    print("MODE=%s STAGES=%s THREADS=%s: %.3f ms  %.3e ev/s  %.0f GB/s" % (os.environ.get("MF_CHAIN_MODE"), os.environ.get("MF_CHAIN_STAGES"), os.environ.get("MF_CHAIN_MINB"), t, N / t * 1e3, 1624 * N / t / 1e6))
else:
    for mode, st, m in (("split", "1", "256"), ("split", "1", "256"), ("split", "1", "128")):
        env = dict(os.environ, MF_CHAIN_MINB=m, MF_CHAIN_THREADS=m, MF_CHAIN_MODE=mode, MF_CHAIN_STAGES=st)
        subprocess.run([sys.executable, __file__, "x"], env=env)
