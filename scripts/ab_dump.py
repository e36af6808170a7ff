"""Same-box A/B of two builds of the library (MEMFLOW_B200_LIB): dumps K1 / K2 outputs on seeded inputs.
usage: python scripts/ab_dump.py <out.pt> [<momenta-from.pt>]; python scripts/ab_dump.py --compare a.pt b.pt"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5], 5: [125.25, 172.5, 172.5, 1e-5, 0.0],
         6: [4.7, 4.7, 80.4, 172.5, 0.0, 1e-5], 7: [4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 172.5], 8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]}
if sys.argv[1] == "--compare":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for k in sorted(a):
        x, y = a[k].double(), b[k].double()
        same_nan = bool((torch.isnan(x) == torch.isnan(y)).all())
        ok = ~(torch.isnan(x) | torch.isnan(y))
        d = ((x - y).abs() / (x.abs().amax() if "mom" in k else x.abs().clamp_min(1e-300)))[ok]
        print("%-12s bit-equal %-5s nan-pattern-equal %-5s nans %d max rel diff %.2e  frac>1e-13 %.2e" %
              (k, torch.equal(a[k].view(torch.int64), b[k].view(torch.int64)) if a[k].dtype == torch.float64 else torch.equal(a[k], b[k]),
               same_nan, int(torch.isnan(x).sum()), float(d.max()) if d.numel() else 0.0, float((d > 1e-13).double().mean()) if d.numel() else 0.0))
    sys.exit(0)
from memflow_b200.phasespace.phasespace import PhaseSpace
dev = torch.device("cuda:0")
src = torch.load(sys.argv[2]) if len(sys.argv) > 2 else None
out = {}
N = 1 << 20
for n, m in CASES.items():
    g = torch.Generator().manual_seed(100 + n)
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, generator=g)
    u[:64] = torch.randint(0, 5, (64, 3 * n - 2), generator=g).double() / 4          # exact 0, 1/4, 1/2, 3/4, 1 (edges)
    u[64:128, 0] = torch.tensor([0.0, 1.0, 1e-40, 1 - 1e-16] * 16, dtype=torch.float64)
    u[128, 3] = float("nan")
    ps = PhaseSpace(13000.0, [21, 21], list(range(n)), final_masses=torch.tensor(m, dtype=torch.float64), dev=dev)
    mom, w, x1, x2 = ps.get_momenta_from_ps(u.to(dev))
    out["mom%d" % n], out["w%d" % n], out["x1_%d" % n] = mom.cpu(), w.cpu(), x1.cpu()
    out["st%d" % n] = ps.generator.last_status.cpu().clone().double()
    if src is not None:
        mom, x1, x2 = src["mom%d" % n].to(dev), src["x1_%d" % n].to(dev), src["x2_%d" % n].to(dev)
    out["x2_%d" % n] = x2.cpu()
    back, prob = ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2)
    out["inv%d" % n], out["prob%d" % n] = back.cpu(), prob.cpu()
torch.save(out, sys.argv[1])
