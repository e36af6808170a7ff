"""Round-2 check of the float64 spline kernel (csrc/rqs_f64.cu) on the GPU: accuracy against the oracle and timings,
new kernel vs the round-1 cooperative kernel (MF_RQS64=0), incl. the float64 flow stage of the chain at the config-5 shape.
usage: python scripts/r2_rqs64.py [time|acc|all]"""
import os, subprocess, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memflow_b200 import transforms as TR, estimator as ES
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
HBM = 6459.6


def timeit(f, n=6):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


if what in ("time", "all"):
    tag = "rqs64=%s L=%s" % (os.environ.get("MF_RQS64", "1"), os.environ.get("MF_RQS64_LANES", "auto"))
    for (B, F, K) in ((1 << 18, 10, 40), (1 << 18, 10, 35), (1 << 17, 20, 64), (1 << 20, 7, 10), (1 << 18, 10, 33), (1 << 19, 8, 16)):
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64)
        x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
        y = torch.empty_like(x); l = torch.empty_like(x); ls = torch.empty(B, device=dev, dtype=torch.float64)
        by = (F * (3 * K - 1) + 2 * F + 1) * 8
        for name, fn in (("fwd", TR.rqs_forward), ("inv", TR.rqs_inverse)):
            ms = timeit(lambda: fn(phi, x, bound=1.1))
            print("%s %s f64 F=%d K=%d B=%d: %.3f ms  %.0f GB/s  frac %.3f  %.3e ev/s" % (name, tag, F, K, B, ms, by * B / ms / 1e6, by * B / ms / 1e6 / HBM, B / ms * 1e3), flush=True)
        del phi, x
    # float64 flow stage of the chain, config-5 shape (n=4: F=10, K=40, T=6) and config-4 shape (F=7, K=10, T=2)
    for (N, n, T, K) in ((1 << 17, 4, 6, 40), (1 << 17, 4, 6, 35), (1 << 20, 3, 2, 10)):
        F = 3 * n - 2
        phi = torch.randn(T, N, F, 3 * K - 1, device=dev, dtype=torch.float64)
        smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5][:n], 13000.0, seed=3, device=dev)
        u = torch.empty(N, F, dtype=torch.float64, device=dev); lq = torch.empty(N, dtype=torch.float64, device=dev)
        acc = ES.new_accumulator(dev)
        ta = timeit(lambda: smp.sample_flow(phi, u=u, logq=lq))
        tb = timeit(lambda: smp.weigh(u, lq, acc=acc))
        by = T * F * (3 * K - 1) * 8
        print("chain f64 %s n=%d T=%d F=%d K=%d N=%d: A %.3f ms (%.0f GB/s, frac %.3f)  B %.3f ms  -> %.3e ev/s" % (
            tag, n, T, F, K, N, ta, by * N / ta / 1e6, by * N / ta / 1e6 / HBM, tb, N / (ta + tb) * 1e3), flush=True)
        del phi

if what in ("bwd", "all"):
    tag = "rqs64=%s" % os.environ.get("MF_RQS64", "1")
    for (B, F, K) in ((1 << 18, 10, 40), (1 << 18, 10, 35), (1 << 17, 20, 64), (1 << 20, 7, 10)):
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64, requires_grad=True)
        x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
        y, l, ls = TR.rqs_forward(phi, x, bound=1.1)
        loss = ls.sum()
        def bwd():
            phi.grad = None
            loss.backward(retain_graph=True)
        ms = timeit(bwd)
        by = 2 * F * (3 * K - 1) * 8
        print("bwd %s f64 F=%d K=%d B=%d: %.3f ms  %.0f GB/s  frac %.3f" % (tag, F, K, B, ms, by * B / ms / 1e6, by * B / ms / 1e6 / HBM), flush=True)
        del phi, x, y, l, ls, loss

if what in ("acc", "all"):
    from oracle import oracle as O
    O.build()
    for (B, F, K, bound) in ((20000, 10, 40, 1.1), (20000, 10, 35, 1.1), (8000, 20, 64, 1.1), (30000, 7, 10, 1.0), (3000, 5, 128, 1.0), (5000, 3, 9, 5.0), (5000, 1, 1, 2.0), (5000, 6, 33, 1.0)):
        g = torch.Generator().manual_seed(7)
        phi = torch.randn(B, F, 3 * K - 1, generator=g, dtype=torch.float64)
        x = (torch.rand(B, F, generator=g, dtype=torch.float64) * 2.6 - 1.3) * bound
        for name, fn, ofn in (("fwd", TR.rqs_forward, O.rqs_forward), ("inv", TR.rqs_inverse, O.rqs_inverse)):
            y, l, ls, b = fn(phi.cuda(), x.cuda(), bound=bound, return_bins=True)
            oy, ol, ob = ofn(phi.numpy(), x.numpy(), bound=bound)
            ey, el = np.abs(y.cpu().numpy() - oy), np.abs(l.cpu().numpy() - ol)
            print("acc %s F=%d K=%d: bins equal %s (%d differ) | y p50 %.1e p99 %.1e max %.1e | ladj p50 %.1e p99 %.1e max %.1e | lsum max %.1e" % (
                name, F, K, np.array_equal(b.cpu().numpy(), ob), int((b.cpu().numpy() != ob).sum()), np.median(ey), np.quantile(ey, 0.99), ey.max(),
                np.median(el), np.quantile(el, 0.99), el.max(), np.abs(ls.cpu().numpy() - ol.sum(1)).max()), flush=True)
