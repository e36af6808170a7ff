"""Small single-GPU program for ncu: a few launches of each kernel at benchmark shapes.
usage: python scripts/profile_target.py [chain|k1|k2|rqs|rqs64|all]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memflow_b200 import estimator as ES, transforms as TR
from memflow_b200.phasespace.phasespace import PhaseSpace

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
E = 13000.0
if what in ("chain", "all"):
    N = 1 << 21
    phi = torch.randn(2, N, 7, 29, device=dev, dtype=torch.float32)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5], E, seed=1, device=dev)
    acc = ES.new_accumulator(dev)
    for i in range(4):
        smp.run_chunk(phi, event_offset=i * N, acc=acc)
    torch.cuda.synchronize()
    del phi
if what in ("k1", "k2", "all"):
    for n, m in ((3, [125.25, 172.5, 172.5]), (8, [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3])):
        ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(m, dtype=torch.float64), dev=dev)
        u = torch.rand(1 << 22, 3 * n - 2, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
        for i in range(3):
            mom, w, x1, x2 = ps.get_momenta_from_ps(u)
        if what in ("k2", "all"):
            for i in range(3):
                ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2)
        torch.cuda.synchronize()
if what in ("rqs", "all"):
    B, F, K = 1 << 19, 20, 64
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float32)
    x = (torch.rand(B, F, device=dev) * 2.6 - 1.3) * 1.1
    for i in range(3):
        TR.rqs_forward(phi, x, bound=1.1)
        TR.rqs_inverse(phi, x, bound=1.1)
    torch.cuda.synchronize()
if what in ("rqs64", "all"):   # the reference's configs run the flows in float64: forward + backward at F=10, K=40
    B, F, K = 1 << 18, 10, 40
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64, requires_grad=True)
    x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
    for i in range(3):
        y, l, ls = TR.rqs_forward(phi, x, bound=1.1)
        phi.grad = None
        ls.sum().backward()
    torch.cuda.synchronize()
if what == "chain5":   # the chain's flow stage at the production shape (n=4: T=6, F=10, K=40), fp32 parameters
    N, T, F, K = 1 << 18, 6, 10, 40
    phi = torch.randn(T, N, F, 3 * K - 1, device=dev, dtype=torch.float32)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], E, seed=3, device=dev)
    acc = ES.new_accumulator(dev)
    for i in range(3):
        smp.run_chunk(phi, event_offset=i * N, acc=acc)
    torch.cuda.synchronize()
if what == "k1n6":   # n = 6: padded input rows filled by 16-byte asynchronous copies
    m6 = [4.7, 4.7, 80.4, 172.5, 0.0, 1e-5]
    ps = PhaseSpace(E, [21, 21], list(range(6)), final_masses=torch.tensor(m6, dtype=torch.float64), dev=dev)
    u = torch.rand(1 << 22, 16, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
    for i in range(3):
        mom, w, x1, x2 = ps.get_momenta_from_ps(u)
    for i in range(3):
        ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2, order=list(range(6)))
    torch.cuda.synchronize()
if what == "rqs64bk35":   # float64 backward at the most common bin count of the reference's configs (odd K: padded rows)
    B, F, K = 1 << 18, 10, 35
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64, requires_grad=True)
    x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
    for i in range(3):
        y, l, ls = TR.rqs_forward(phi, x, bound=1.1)
        phi.grad = None
        ls.sum().backward()
    torch.cuda.synchronize()
if what in ("rqs64k40", "rqs64k35", "rqs64k33", "rqs64k64", "rqs64k10"):   # float64 forward / inverse alone (csrc/rqs_f64.cu)
    K = int(what[6:])
    F = {40: 10, 35: 10, 33: 10, 64: 20, 10: 7}[K]
    B = (1 << 18) if K >= 33 else (1 << 20)
    if K == 64:
        B = 1 << 17
    phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64)
    x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
    for i in range(3):
        TR.rqs_forward(phi, x, bound=1.1)
        TR.rqs_inverse(phi, x, bound=1.1)
    torch.cuda.synchronize()
if what in ("chain64", "chain64k35"):   # float64 flow stage of the chain at the config-5 shape (n=4: F=10, T=6)
    K = 35 if what.endswith("k35") else 40
    N, T, F = 1 << 17, 6, 10
    phi = torch.randn(T, N, F, 3 * K - 1, device=dev, dtype=torch.float64)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], E, seed=3, device=dev)
    acc = ES.new_accumulator(dev)
    for i in range(3):
        smp.run_chunk(phi, event_offset=i * N, acc=acc)
    torch.cuda.synchronize()
if what == "sampling":   # fused sampling branch of the training loss (csrc/sampling.cu), forward + backward, n = 4
    from memflow_b200.unfolding_flow import utils as UU
    N = 1 << 20
    ps4 = PhaseSpace(E, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor([125.25, 172.5, 172.5, 1e-5], dtype=torch.float64), dev=dev)
    u = torch.rand(N, 10, dtype=torch.float64, device=dev).clamp_(1e-6, 1 - 1e-6).requires_grad_(True)
    ang = [torch.stack((torch.randn(N, device=dev, dtype=torch.float64), (torch.rand(N, device=dev, dtype=torch.float64) * 2 - 1) * 3.1), 1).requires_grad_(True)
           for _ in range(5)]
    z3, o3, z2, o2 = [0.0] * 3, [1.0] * 3, [0.0] * 2, [1.0] * 2
    for i in range(3):
        out, bs, *_ = UU.sampled_decayPartons_from_ps(ps4, u, *ang, log_mean_parton_lab=z3, log_std_parton_lab=o3, log_mean_boost=z2,
                                                      log_std_boost=o2, log_mean_parton_Hthadtlep=z3, log_std_parton_Hthadtlep=o3)
        (out.nan_to_num().sum() + bs.sum()).backward()
    torch.cuda.synchronize()
if what == "k5":   # weight + estimator reduction alone (csrc/is_reduce.cu)
    N = 1 << 26
    jac = torch.rand(N, dtype=torch.float64, device=dev); lq = torch.randn(N, dtype=torch.float64, device=dev)
    tg = torch.rand(N, dtype=torch.float64, device=dev)
    acc = ES.new_accumulator(dev)
    for i in range(3):
        ES.accumulate_weights(tg, jac, lq, acc)
    torch.cuda.synchronize()
print("done")
