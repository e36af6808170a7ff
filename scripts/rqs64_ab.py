"""Same-box A/B of the float64 spline forward / inverse (csrc/rqs_f64.cu) and of the float64 flow stage of the chain under
environment knobs (read once per process, so every variant runs in its own process).
usage: rqs64_ab.py [F:K ...]      variants: MF_RQS64_WARPS = 8 / 4 (threads per CTA 256 / 128); MF_RQS64_STAGES, MF_RQS64_LANES from the environment"""
import os, sys, subprocess, json
VARIANTS = [("default", {}), ("4 warps", {"MF_RQS64_WARPS": "4"}), ("8 warps", {"MF_RQS64_WARPS": "8"})]
_PREV = os.environ.get("AB_PREV_LIB")   # another build of the library to compare with
if _PREV:
    VARIANTS.append(("prev build", {"MEMFLOW_B200_LIB": _PREV}))
if os.environ.get("_AB_CHILD"):
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from memflow_b200 import transforms as TR, estimator as ES
    dev = torch.device("cuda:0")
    res = {}
    def timeit(fn, n=12):
        for _ in range(4): fn()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]
    for spec in sys.argv[1:]:
        F, K = map(int, spec.split(":"))
        B = max(1024, (1 << 30) // (F * (3 * K - 1) * 8))   # 1 GiB of parameters
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64)
        x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
        by = B * F * (3 * K - 1) * 8
        for name, fn in (("fwd", TR.rqs_forward), ("inv", TR.rqs_inverse)):
            ms = timeit(lambda: fn(phi, x, bound=1.1))
            res["%s %s" % (spec, name)] = (ms, by / ms / 1e6)
        del phi, x
    # float64 flow stage of the chain at the config-5 shape
    for K in (40, 35):
        N, T, F = 1 << 17, 6, 10
        phi = torch.randn(T, N, F, 3 * K - 1, device=dev, dtype=torch.float64)
        smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], 13000.0, seed=3, device=dev)
        acc = ES.new_accumulator(dev)
        ms = timeit(lambda: smp.run_chunk(phi, event_offset=0, acc=acc), 8)
        res["chain64 K=%d" % K] = (ms, T * N * F * (3 * K - 1) * 8 / ms / 1e6)
        del phi
    print(json.dumps(res))
else:
    specs = sys.argv[1:] or ["10:40", "10:35", "20:64", "8:20", "8:16", "7:10", "1:35"]
    out = {}
    for rnd in range(2):
        for name, env in VARIANTS:
            p = subprocess.run([sys.executable, __file__] + specs, env=dict(os.environ, _AB_CHILD="1", **env), capture_output=True, text=True)
            try:
                out.setdefault(name, []).append(json.loads(p.stdout.strip().splitlines()[-1]))
            except Exception:
                print(name, "FAILED", p.stderr[-500:])
    keys = list(next(iter(out.values()))[0].keys())
    print("%-14s" % "case" + "".join("%22s" % n for n, _ in VARIANTS if n in out))
    for k in keys:
        print("%-14s" % k + "".join("%12.3f ms %4.0f %%" % (min(r[k][0] for r in out[n]), max(r[k][1] for r in out[n]) / 64.596)
                                    for n, _ in VARIANTS if n in out))
