"""A/B timing of the flow-sample kernel variants inside ONE process per variant, interleaved and repeated, to
beat run-to-run clock noise: prints the per-variant median of launch A (ms) over several rounds."""
import os, sys, subprocess, json
VARIANTS = [dict(MF_CHAIN_FAST="1"), dict(MF_CHAIN_FAST="0")]  # or dict(MF_CHAIN_FAST="0", MF_CHAIN_MINB="6"), dict(MEMFLOW_B200_LIB=<other build of the .so>)
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from memflow_b200 import estimator as ES
    dev = torch.device("cuda:0")
    N = 1 << 22
    phi = torch.randn(2, N, 7, 29, device=dev, dtype=torch.float32)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, seed=1, device=dev)
    u = torch.empty(N, 7, dtype=torch.float64, device=dev); lq = torch.empty(N, dtype=torch.float64, device=dev)
    acc = ES.new_accumulator(dev)
    for i in range(5): smp.sample_flow(phi, u=u, logq=lq); smp.weigh(u, lq, acc=acc)
    torch.cuda.synchronize()
    ta, tb = [], []
    for i in range(60):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); smp.sample_flow(phi, event_offset=i * N, u=u, logq=lq); e1.record(); smp.weigh(u, lq, acc=acc); e2.record()
        torch.cuda.synchronize()
        ta.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    ta.sort(); tb.sort()
    print(json.dumps({"A_ms": ta[len(ta) // 2], "B_ms": tb[len(tb) // 2]}))
else:
    res = {i: [] for i in range(len(VARIANTS))}
    for rnd in range(3):
        for i, v in enumerate(VARIANTS):
            out = subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, **v), capture_output=True, text=True).stdout
            res[i].append(json.loads(out.strip().splitlines()[-1]))
    for i, v in enumerate(VARIANTS):
        a = sorted(r["A_ms"] for r in res[i]); b = sorted(r["B_ms"] for r in res[i])
        print(v, "A median %.3f ms (%.0f GB/s)  [%s]   B median %.3f ms" % (a[1], 1624 * (1 << 22) / a[1] / 1e6, ", ".join("%.3f" % x for x in a), b[1]))
