"""Same-box A/B of the float64 spline backward: rqs64_backward_kernel (default) against rqs_backward_coop_kernel
(MF_RQS64_BWD=0), each in its own process (the knob is read once), through torch.ops.memflow_b200.rqs_bwd.
usage: rqs_bwd_ab.py [F:K ...]"""
import os, sys, subprocess, json
if os.environ.get("_AB_CHILD"):
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import memflow_b200.ops  # noqa: F401
    dev = torch.device("cuda:0")
    res = {}
    for spec in sys.argv[1:]:
        F, K = map(int, spec.split(":"))
        B = max(1024, (1 << 28) // (F * (3 * K - 1) * 8))   # 256 MiB of parameters
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64)
        x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
        gl = torch.ones(B, F, device=dev, dtype=torch.float64)
        for inverse in (False, True):
            for _ in range(5):
                torch.ops.memflow_b200.rqs_bwd(phi, x, None, gl, 1.1, 1e-3, inverse)
            ts = []
            for _ in range(15):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); torch.ops.memflow_b200.rqs_bwd(phi, x, None, gl, 1.1, 1e-3, inverse); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            by = 2 * B * F * (3 * K - 1) * 8
            res["%s %s" % (spec, "inv" if inverse else "fwd")] = (ts[0], ts[len(ts) // 2], by / ts[len(ts) // 2] / 1e6)
    print(json.dumps(res))
else:
    specs = sys.argv[1:] or ["10:40", "10:35", "20:64", "8:20", "8:16", "7:10"]
    out = {}
    for rnd in range(2):
        for name, env in (("old", {"MF_RQS64_BWD": "0"}), ("new", {})):
            p = subprocess.run([sys.executable, __file__] + specs, env=dict(os.environ, _AB_CHILD="1", **env), capture_output=True, text=True)
            try:
                out.setdefault(name, []).append(json.loads(p.stdout.strip().splitlines()[-1]))
            except Exception:
                print(name, "FAILED", p.stderr[-500:])
    for k in out.get("new", [{}])[0]:
        o = min(r[k][1] for r in out["old"]); n = min(r[k][1] for r in out["new"])
        gb = max(r[k][2] for r in out["new"])
        print("%-12s old %.3f ms  new %.3f ms  (%.2fx)  new: %.0f GB/s = %.0f %% of 6459.6" % (k, o, n, o / n, gb, gb / 64.596))
