"""Exploratory GPU check: prints parity numbers of every kernel against the oracle + rough timings."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O, philox as OP
from memflow_b200.phasespace.phasespace import PhaseSpace
from memflow_b200 import transforms as TR, estimator as ES

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
E = 13000.0
CASES = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5], 5: [125.25, 172.5, 172.5, 1e-5, 1e-5],
         6: [125.25, 172.5, 172.5, 1e-5, 1e-5, 4.7], 7: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3],
         8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]}

def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for n, masses in CASES.items():
    m = torch.tensor(masses, dtype=torch.float64)
    ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=m, dev=dev)
    N = 200000
    g = torch.Generator().manual_seed(100 + n)
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, generator=g).clamp(1e-10, 1 - 1e-10)
    mom, w, x1, x2, lw = ps.get_momenta_from_ps(u.to(dev), return_logweight=True)
    torch.cuda.synchronize()
    om, ow, ox1, ox2 = O.rambo_forward(u.numpy(), masses, E)
    Ecm = np.sqrt(ox1 * ox2) * E
    dm = np.abs(mom.cpu().numpy() - om).max(axis=(1, 2)) / Ecm
    dlw = np.abs(np.log(w.cpu().numpy()) - np.log(ow))
    print("n=%d FWD: mom/Ecm max %.2e  frac>1e-12 %.2e  p99.9 %.2e | x1 %.1e | dlogw max %.2e p99.9 %.2e | logw-log(w) %.1e | status %d"
          % (n, dm.max(), (dm > 1e-12).mean(), np.quantile(dm, 0.999), np.abs(x1.cpu().numpy() - ox1).max(), np.nanmax(dlw),
             np.nanquantile(dlw, 0.999), float((lw - w.log()).abs().max()), int(ps.generator.last_status.item())))
    # inverse on the oracle's momenta
    omt = torch.from_numpy(om).to(dev)
    psb, prob, lp = ps.generator.getPSpoint_batch(omt[:, 2:], torch.from_numpy(ox1).to(dev), torch.from_numpy(ox2).to(dev), return_logprob=True)
    ops, oprob, rc = O.rambo_inverse(om[:, 2:], ox1, ox2, masses, E)
    dps = np.abs(psb.cpu().numpy() - ops)
    dlp = np.abs(np.log(prob.cpu().numpy()) - np.log(oprob))
    print("     INV: ps max %.2e frac>1e-12 %.2e | dlogprob max %.2e p99.9 %.2e | rt max|u'-u| %.2e | status %d"
          % (np.nanmax(dps), (dps > 1e-12).mean(), np.nanmax(dlp), np.nanquantile(dlp, 0.999),
             float((psb.cpu() - u).abs().max()), int(ps.generator.last_status.item())))
    # timing at 4M
    Nb = 4_000_000
    ub = torch.rand(Nb, 3 * n - 2, dtype=torch.float64, device=dev).clamp(1e-10, 1 - 1e-10)
    t = timeit(lambda: ps.get_momenta_from_ps(ub))
    bytes_ev = (3 * n - 2 + 4 * (n + 2) + 3) * 8
    momb, wb, x1b, x2b = ps.get_momenta_from_ps(ub)
    ti = timeit(lambda: ps.generator.getPSpoint_batch(momb[:, 2:], x1b, x2b))
    bytes_inv = (4 * n + 2 + 3 * n - 2 + 1) * 8
    print("     time fwd %.3f ms (%.2e ev/s, %.0f GB/s algorithmic) | inv %.3f ms (%.2e ev/s, %.0f GB/s) [wrapper incl. alloc]"
          % (t, Nb / t * 1e3, bytes_ev * Nb / t / 1e6, ti, Nb / ti * 1e3, bytes_inv * Nb / ti / 1e6))

# massvar solver vs oracle bisection (through the forward path of the oracle is implicit); direct check:
for n in (4, 8):
    ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(CASES[n], dtype=torch.float64), dev=dev)
    v = torch.rand(100000, n - 2, dtype=torch.float64)
    uu = ps.generator.bisect_vec_batch(v.to(dev)).cpu()
    e = torch.arange(n - 2, 0, -1, dtype=torch.float64)
    res = (uu ** e) * ((e + 1) - e * uu) - v
    print("massvar n=%d residual max %.2e" % (n, float(res.abs().max())))

# spline
for (dt, F, K, bound, B) in ((torch.float32, 20, 64, 1.1, 20000), (torch.float64, 7, 10, 1.0, 20000), (torch.float64, 10, 40, 1.1, 20000), (torch.float32, 7, 10, 1.0, 20000)):
    g = torch.Generator().manual_seed(7)
    phi = torch.randn(B, F, 3 * K - 1, generator=g, dtype=torch.float64).to(dt)
    x = ((torch.rand(B, F, generator=g, dtype=torch.float64) * 2.6 - 1.3) * bound).to(dt)
    y, ladj, lsum, bins = TR.rqs_forward(phi.to(dev), x.to(dev), bound=bound, return_bins=True)
    oy, oladj, obin = O.rqs_forward(phi.double().numpy(), x.double().numpy(), bound=bound)  # fp64 oracle on the same inputs
    sy, sladj, sbin = O.rqs_forward(phi.numpy(), x.numpy(), bound=bound)  # same-dtype oracle
    mism = (bins.cpu().numpy() != obin)
    ok = ~mism
    print("RQS %s F=%d K=%d FWD: bin mismatch %d/%d | y max err %.2e | ladj max err %.2e (same-dtype oracle err vs f64: y %.2e ladj %.2e) | ladj_sum err %.2e"
          % (str(dt)[6:], F, K, mism.sum(), mism.size, np.abs(y.cpu().numpy() - oy)[ok].max(), np.abs(ladj.cpu().numpy() - oladj)[ok].max(),
             np.abs(sy - oy)[sbin == obin].max(), np.abs(sladj - oladj)[sbin == obin].max(),
             np.abs(lsum.cpu().numpy() - ladj.cpu().numpy().sum(1)).max()))
    xi, ladji, _, binsi = TR.rqs_inverse(phi.to(dev), x.to(dev), bound=bound, return_bins=True)
    ox, oladji, obini = O.rqs_inverse(phi.double().numpy(), x.double().numpy(), bound=bound)
    mism = binsi.cpu().numpy() != obini; ok = ~mism
    print("                    INV: bin mismatch %d | x max err %.2e | ladj max err %.2e | roundtrip f(inv(y))-y %.2e"
          % (mism.sum(), np.abs(xi.cpu().numpy() - ox)[ok].max(), np.abs(ladji.cpu().numpy() - oladji)[ok].max(),
             float((TR.rqs_forward(phi.to(dev), xi, bound=bound)[0].cpu() - x).abs().max())))
    t = TR.MonotonicRQSTransform(phi[..., :K].to(dev), phi[..., K:2 * K].to(dev), phi[..., 2 * K:].to(dev), bound=bound)
    y2, l2 = t.call_and_ladj(x.to(dev))
    print("                    Transform API equal: %s %s ; inv: %s" % (torch.equal(y2, y), torch.equal(l2, ladj), torch.equal(t.inv(x.to(dev)), xi)))
# spline timing config 3 (1M events)
B, F, K = 1 << 20, 20, 64
phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float32)
x = (torch.rand(B, F, device=dev) * 2.6 - 1.3) * 1.1
tf = timeit(lambda: TR.rqs_forward(phi, x, bound=1.1)); ti = timeit(lambda: TR.rqs_inverse(phi, x, bound=1.1))
by = (F * (3 * K - 1) + 2 * F + 1) * 4
print("RQS f32 F20 K64 1M: fwd %.3f ms %.0f GB/s | inv %.3f ms %.0f GB/s" % (tf, by * B / tf / 1e6, ti, by * B / ti / 1e6))

# estimator
N = 3_000_001
tg = torch.rand(N, dtype=torch.float64); tg[::1000] = -1.0
jac = torch.rand(N, dtype=torch.float64) * 1e3; jac[::777] = float("nan")
lq = torch.randn(N, dtype=torch.float64)
acc = ES.accumulate_weights(tg.to(dev), jac.to(dev), lq.to(dev))
oacc = O.is_reduce(tg.numpy(), jac.numpy(), lq.numpy())
print("IS reduce: rel err", np.abs(acc.cpu().numpy() / oacc - 1), "count", acc[2].item(), oacc[2])
acc2 = ES.accumulate_weights(tg.to(dev), jac.to(dev), lq.to(dev))
print("   deterministic:", torch.equal(acc, acc2))

# chain
for n, T, K in ((3, 2, 10), (4, 3, 8)):
    masses = CASES[n]; F = 3 * n - 2; N = 5000 + 3
    g = torch.Generator().manual_seed(11)
    phi = torch.randn(T, N, F, 3 * K - 1, generator=g, dtype=torch.float32)
    smp = ES.ImportanceSampler(masses, E, bound=1.0, seed=1234, device=dev)
    acc, out = smp.run_chunk(phi.to(dev), event_offset=17, debug_outputs=True)
    z = OP.base_sample(N, F, 1.0, 1234, 17)
    xcur = z.copy(); lsum = np.zeros(N)
    for t in range(T - 1, -1, -1):
        xcur, l, _ = O.rqs_inverse(phi[t].numpy(), xcur, bound=1.0)
        lsum += l.astype(np.float64).sum(1)
    uo = (xcur.astype(np.float64) + 1.0) / 2.0
    print("chain n=%d: u err %.2e logq err %.2e" % (n, np.abs(out["u"].cpu().numpy() - uo).max(), np.abs(out["logq"].cpu().numpy() - lsum).max()))
    ug = out["u"].cpu().numpy()
    om, ow, ox1, ox2 = O.rambo_forward(ug, masses, E)
    Ecm = np.sqrt(ox1 * ox2) * E
    print("   mom/Ecm %.2e  dlogw %.2e" % ((np.abs(out["momenta"].cpu().numpy() - om).max(axis=(1, 2)) / Ecm).max(),
                                          np.nanmax(np.abs(np.log(out["weight"].cpu().numpy()) - np.log(ow)))))
    tgt = 1.0 / (2 * E * E) / (ox1 * ox2)
    oacc = O.is_reduce(tgt, ow, out["logq"].cpu().numpy())
    print("   acc rel err", np.abs(acc.cpu().numpy() / oacc - 1))
# chain timing
n, T, K, N = 3, 2, 10, 1 << 22
phi = torch.randn(T, N, 7, 29, device=dev, dtype=torch.float32)
smp = ES.ImportanceSampler(CASES[3], E, seed=1, device=dev)
acc = ES.new_accumulator(dev)
t = timeit(lambda: smp.run_chunk(phi, acc=acc))
print("chain n=3 T=2 K=10 4M: %.3f ms  %.2e ev/s  %.0f GB/s" % (t, N / t * 1e3, T * 7 * 29 * 4 * N / t / 1e6))
