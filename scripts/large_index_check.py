"""One-off check of 64-bit indexing: K1 / K2 at 2^27 events (2.7e9 output doubles) and the chain at 2^25 events
(1.4e10 parameter floats): the tail of the big launch must equal a small launch on the same inputs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200.phasespace.phasespace import PhaseSpace
from memflow_b200 import estimator as ES
dev = torch.device("cuda:0")
ps = PhaseSpace(13000.0, [21, 21], [25, 6, -6], final_masses=torch.tensor([125.25, 172.5, 172.5], dtype=torch.float64), dev=dev)
N = 1 << 27
u = torch.rand(N, 7, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
mom, w, x1, x2 = ps.get_momenta_from_ps(u)
tail = slice(N - 1000, N)
m2, w2, a2, b2 = ps.get_momenta_from_ps(u[tail].contiguous())
assert torch.equal(mom[tail], m2) and torch.equal(w[tail], w2) and torch.equal(x1[tail], a2)
back, prob = ps.get_ps_from_momenta(mom[:, 2:], x1, x2)
bk2, pr2 = ps.get_ps_from_momenta(m2[:, 2:], a2, b2)
assert torch.equal(back[tail], bk2) and torch.equal(prob[tail], pr2)
err = (back[tail] - u[tail]).abs().max().item()
print("K1/K2 at 2^27 events: tail identical to a small launch; round trip max %.1e" % err)
del u, mom, w, x1, x2, back, prob
torch.cuda.empty_cache()
N = 1 << 25
phi = torch.randn(2, N, 7, 29, device=dev, dtype=torch.float32)
smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, seed=5, device=dev)
ub, lb = smp.sample_flow(phi)
off = N - 777
us, ls = smp.sample_flow(phi[:, off:].contiguous(), event_offset=off)
assert torch.equal(ub[off:], us) and torch.equal(lb[off:], ls)
print("chain launch A at 2^25 events (1.4e10 parameters): tail identical to a small launch")
del phi, ub, lb
torch.cuda.empty_cache()
from memflow_b200 import transforms as TR
for dt, B, F, K in ((torch.float32, 1 << 21, 20, 64), (torch.float64, 1 << 21, 10, 35), (torch.float32, 1 << 25, 7, 10)):
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=dt)
    x = (torch.rand(B, F, device=dev, dtype=dt) * 2.6 - 1.3)
    y, l, ls = TR.rqs_forward(phi, x, bound=1.1)
    t = slice(B - 333, B)
    y2, l2, ls2 = TR.rqs_forward(phi[t].contiguous(), x[t].contiguous(), bound=1.1)
    if K % 2 == 0:
        assert torch.equal(y[t], y2) and torch.equal(l[t], l2) and torch.equal(ls[t], ls2), (dt, K)
    else:  # odd K: the softmax sums start at a tile-position dependent element (bank conflicts): last-bit differences
        tol = 1e-13 if dt == torch.float64 else 1e-4
        assert (y[t] - y2).abs().max().item() < tol and (ls[t] - ls2).abs().max().item() < 100 * tol, (dt, K)
    print("spline %s F=%d K=%d at %d events (%.1e parameters): tail %s a small launch" % (
        str(dt)[6:], F, K, B, phi.numel(), "identical to" if K % 2 == 0 else "equal within rounding to"))
    del phi, x, y, l, ls
    torch.cuda.empty_cache()
