"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import estimator as ES, transforms as TR
from memflow_b200.phasespace.phasespace import PhaseSpace
from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor
dev = torch.device("cuda:0"); E = 13000.0
for n, m in ((3, [125.25, 172.5, 172.5]), (4, [125.25, 172.5, 172.5, 1e-5]), (8, [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3])):
    ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(m, dtype=torch.float64), dev=dev)
    u = torch.rand(3001, 3 * n - 2, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
    mom, w, x1, x2 = ps.get_momenta_from_ps(u)
    ps.get_momenta_from_ps(u, pT_mincut=20.0, delR_mincut=0.4, rap_maxcut=5.0, lab_frame=True)
    ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2)
    ps.generator.bisect_vec_batch(u[:, : n - 2])
    if n == 4:
        b = torch.stack((E / 2 * (x1 + x2), 0 * x1, 0 * x1, E / 2 * (x1 - x2)), dim=1)
        Compute_ParticlesTensor.get_PS(mom[:, 2:], b, logit_eps=5e-5)
for (dt, F, K) in ((torch.float32, 20, 64), (torch.float32, 10, 40), (torch.float32, 7, 10), (torch.float64, 7, 10), (torch.float64, 5, 33), (torch.float32, 3, 9)):
    phi = torch.randn(1003, F, 3 * K - 1, device=dev, dtype=dt)
    x = (torch.rand(1003, F, device=dev, dtype=dt) * 2.6 - 1.3)
    TR.rqs_forward(phi, x, bound=1.0); TR.rqs_inverse(phi, x, bound=1.0)
for (F, K) in ((10, 40), (6, 35), (7, 10), (5, 16), (3, 64), (2, 5)):   # spline backward, float64 (rqs64_backward_kernel and the cooperative one) and float32
    for dt in (torch.float64, torch.float32):
        phi = torch.randn(1531, F, 3 * K - 1, device=dev, dtype=dt, requires_grad=True)
        x = (torch.rand(1531, F, device=dev, dtype=dt) * 2.6 - 1.3)
        for fn in (TR.rqs_forward, TR.rqs_inverse):
            y, l, ls = fn(phi, x, bound=1.0)
            (y.sum() + ls.sum()).backward()
smp4 = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], E, seed=3, device=dev)
for K in (40, 35):   # float64 flow stage of the chain
    smp4.run_chunk(torch.randn(3, 1201, 10, 3 * K - 1, device=dev, dtype=torch.float64))
smp = ES.ImportanceSampler([125.25, 172.5, 172.5], E, seed=1, device=dev)
phi = torch.randn(2, 5003, 7, 29, device=dev, dtype=torch.float32)
acc = smp.run_chunk(phi); acc, out = smp.run_chunk(phi, debug_outputs=True)
ES.accumulate_weights(None, out["weight"], out["logq"])
os.environ["MF_CHAIN_MODE"] = "split"
torch.cuda.synchronize()
print("sanitize target done", acc.tolist())
