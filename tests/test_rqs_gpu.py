"""GPU parity of K3/K4 (rational-quadratic spline) against the oracle / torch restatement.

fp64: 1e-12-level agreement. fp32: the problem itself is ill-conditioned in fp32 for narrow bins / steep
derivative ratios (knots are differences of a cumulative sum of O(1) values: one fp32 ulp of x moves
z = (x-x_k)/(x_{k+1}-x_k) by 6e-8/width), so the bound is  1e-5 + 8 * (change of the fp64 result when x moves by
4e-7 * bound), and the kernel must be no worse than the fp32 oracle itself (both measured against the fp64 oracle on the same fp32 inputs)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(B, F, K, dt, bound, seed=7):
    g = torch.Generator().manual_seed(seed)
    phi = torch.randn(B, F, 3 * K - 1, generator=g, dtype=torch.float64).to(dt)
    x = ((torch.rand(B, F, generator=g, dtype=torch.float64) * 2.6 - 1.3) * bound).to(dt)
    return phi, x


def fp32_conditioning(orc, p64, x64, bound):
    """Intrinsic sensitivity of the fp32 problem: how much the fp64 result moves when the input moves by a
    few fp32 ulps of the knot scale (knots are cumulative sums of O(1) fp32 values). Returns (d_out, d_ladj)."""
    delta = 4e-7 * bound
    o0, l0, b0 = orc(p64, x64, bound=bound)
    d_o, d_l = np.zeros_like(o0), np.zeros_like(l0)
    for sgn in (-1.0, 1.0):
        o1, l1, b1 = orc(p64, x64 + sgn * delta, bound=bound)
        same = b1 == b0
        d_o = np.maximum(d_o, np.where(same, np.abs(o1 - o0), np.inf))
        d_l = np.maximum(d_l, np.where(same, np.abs(l1 - l0), np.inf))
    return d_o, d_l


@pytest.mark.parametrize("F,K,bound", [(7, 10, 1.0), (10, 40, 1.1), (20, 64, 1.1), (3, 9, 5.0), (1, 1, 2.0), (5, 128, 1.0)])
def test_fp64_forward_inverse(oracle, F, K, bound):
    from memflow_b200 import transforms as TR
    B = 3001
    phi, x = make(B, F, K, torch.float64, bound)
    y, ladj, lsum, bins = TR.rqs_forward(phi.cuda(), x.cuda(), bound=bound, return_bins=True)
    oy, ol, ob = oracle.rqs_forward(phi.numpy(), x.numpy(), bound=bound)
    assert np.array_equal(bins.cpu().numpy(), ob)                      # integer indexing bit-exact
    assert np.abs(y.cpu().numpy() - oy).max() < 1e-12
    assert np.abs(ladj.cpu().numpy() - ol).max() < 5e-11               # log J, J built from knot differences
    assert np.quantile(np.abs(ladj.cpu().numpy() - ol), 0.99) < 1e-12
    np.testing.assert_allclose(lsum.cpu().numpy(), ol.sum(1), atol=1e-10)
    out = np.abs(x.numpy()) > bound * 1.0001                           # identity tails
    assert out.any() and np.array_equal(y.cpu().numpy()[out], x.numpy()[out]) and np.all(ladj.cpu().numpy()[out] == 0)
    xi, li, _, bi = TR.rqs_inverse(phi.cuda(), x.cuda(), bound=bound, return_bins=True)
    ox, oli, obi = oracle.rqs_inverse(phi.numpy(), x.numpy(), bound=bound)
    assert np.array_equal(bi.cpu().numpy(), obi)
    assert np.abs(xi.cpu().numpy() - ox).max() < 1e-12
    assert np.quantile(np.abs(li.cpu().numpy() - oli), 0.99) < 1e-12
    # inverse o forward = identity, and the inverse's ladj equals the forward's at the recovered point
    y2, l2, _ = TR.rqs_forward(phi.cuda(), xi, bound=bound)
    assert float((y2.cpu() - x).abs().max()) < 1e-12
    assert float((l2 - li).abs().max()) < 1e-9


@pytest.mark.parametrize("F,K,bound,slope", [(7, 10, 1.0, 1e-3), (4, 33, 1.1, 1e-3), (3, 8, 5.0, None)])
def test_defining_properties_without_the_oracle(F, K, bound, slope):
    """The spline is pinned here by what DEFINES a monotonic rational-quadratic spline (Durkan et al. 2019, as zuko
    parameterises it), not by the restatement under oracle/: it interpolates the knots (cumulative softmax of the
    widths / heights, scaled to [-B, B]), its derivative at an interior knot is exp(d_k) and 1 at the two ends, it is
    strictly increasing, its log-det is the log of its finite-difference slope, it is the identity outside [-B, B],
    and the inverse undoes it. The knots are rebuilt with numpy float64 independently of any package code."""
    from memflow_b200 import transforms as TR
    import math
    B = 257
    phi, _ = make(B, F, K, torch.float64, bound, seed=11)
    p = phi.numpy()
    clip = (lambda a, f: a / (1 + np.abs(f * a / math.log(slope)))) if slope is not None else (lambda a, f: a)
    def knots(a):
        e = np.exp(clip(a, 2.0) - clip(a, 2.0).max(-1, keepdims=True))
        c = np.cumsum(e / e.sum(-1, keepdims=True), -1)
        return bound * (2 * np.concatenate((np.zeros_like(c[..., :1]), c), -1) - 1)
    kx, ky = knots(p[..., :K]), knots(p[..., K:2 * K])                 # [B,F,K+1]
    dk = np.concatenate((np.ones((B, F, 1)), np.exp(clip(p[..., 2 * K:], 1.0)), np.ones((B, F, 1))), -1)
    fwd = lambda x: TR.rqs_forward(phi.cuda(), torch.from_numpy(np.ascontiguousarray(x)).cuda(), bound=bound, slope=slope)
    for k in range(1, K):                                              # interior knots: value and slope
        y, ladj, _ = fwd(kx[..., k])
        assert np.abs(y.cpu().numpy() - ky[..., k]).max() < 1e-12
        assert np.abs(ladj.cpu().numpy() - np.log(dk[..., k])).max() < 1e-9
    for k, eps in ((0, 1e-10), (K, -1e-10)):                             # end knots: joins the identity with slope 1
        y, ladj, _ = fwd(kx[..., k] + eps)
        assert np.abs(y.cpu().numpy() - (ky[..., k] + eps)).max() < 1e-12 and np.abs(ladj.cpu().numpy()).max() < 1e-6
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(-1.2 * bound, 1.2 * bound, (B, F, 64)), -1)
    ys, ls = [], []
    h = 1e-6 * bound
    for j in range(64):
        y, ladj, _ = fwd(x[..., j])
        yp, ym = fwd(x[..., j] + h)[0], fwd(x[..., j] - h)[0]
        fd = ((yp - ym) / (2 * h)).cpu().numpy()
        ys.append(y.cpu().numpy())
        # away from a knot the central difference is the derivative; within h of one it mixes two bins
        near = (np.abs(kx - x[..., j][..., None]) < 2 * h).any(-1)
        dev = np.abs(np.log(fd) - ladj.cpu().numpy())[~near]          # O(h^2 / width^2) truncation in narrow bins
        assert np.quantile(dev, 0.99) < 1e-6 and dev.max() < 1e-4
        xi = TR.rqs_inverse(phi.cuda(), y, bound=bound, slope=slope)[0].cpu().numpy()
        assert np.abs(xi - x[..., j]).max() < 2e-12 * max(bound, 1.0)   # unclipped slopes amplify the rounding of y
        outside = np.abs(x[..., j]) > bound
        assert np.array_equal(y.cpu().numpy()[outside], x[..., j][outside])
    assert (np.diff(np.stack(ys, -1), axis=-1) > 0).all()              # strictly increasing


def test_kernels_vs_high_precision_known_answers(golden_dir):
    """The kernels against 50-digit mpmath evaluations of the closed form (tests/golden/make_rqs_known_answers.py):
    independent of the oracle's arithmetic. Bins exact; fp64 to the conditioning of the case, fp32 at 1e-5."""
    import json
    from memflow_b200 import transforms as TR
    ka = json.load(open(os.path.join(golden_dir, "rqs_known_answers.json")))
    for c in ka["cases"]:
        K = len(c["widths"])
        phi = torch.tensor(c["widths"] + c["heights"] + c["derivatives"], dtype=torch.float64)
        P = len(c["points"])
        x = torch.tensor([p["x"] for p in c["points"]], dtype=torch.float64)
        phib = phi.expand(P, 1, 3 * K - 1).contiguous()
        ry, rl, rx, rli = (np.array([float(p[k]) for p in c["points"]]) for k in ("y", "ladj", "x_inv", "ladj_at_x_inv"))
        rb, rbi = (np.array([p[k] for p in c["points"]]) for k in ("bin", "bin_inv"))
        narrow = c["name"] == "K8_bound5_large_parameters"      # bins down to 1e-4 wide: 1 / width conditioning
        y, l, _, b = TR.rqs_forward(phib.cuda(), x[:, None].cuda(), bound=c["bound"], slope=c["slope"], return_bins=True)
        xi, li, _, bi = TR.rqs_inverse(phib.cuda(), x[:, None].cuda(), bound=c["bound"], slope=c["slope"], return_bins=True)
        tol = 5e-12 if narrow else 2e-14
        assert np.array_equal(b.cpu().numpy()[:, 0], rb) and np.array_equal(bi.cpu().numpy()[:, 0], rbi), c["name"]
        assert np.abs(y.cpu().numpy()[:, 0] - ry).max() < tol and np.abs(l.cpu().numpy()[:, 0] - rl).max() < tol, c["name"]
        assert np.abs(xi.cpu().numpy()[:, 0] - rx).max() < tol and np.abs(li.cpu().numpy()[:, 0] - rli).max() < tol, c["name"]
        if narrow:
            continue
        y32, l32, _ = TR.rqs_forward(phib.float().cuda(), x[:, None].float().cuda(), bound=c["bound"], slope=c["slope"])
        x64 = x.float().double().numpy()                        # the fp32 kernel sees the rounded input
        moved = np.abs(x64 - x.numpy()) > 0
        assert np.abs(y32.cpu().numpy()[:, 0] - ry)[~moved].max() < 1e-5 and np.abs(l32.cpu().numpy()[:, 0] - rl)[~moved].max() < 1e-4


@pytest.mark.parametrize("F,K,bound", [(20, 64, 1.1), (7, 10, 1.0), (10, 40, 1.1), (6, 32, 1.0), (5, 48, 1.1)])
def test_fp32_forward_inverse(oracle, F, K, bound):
    from memflow_b200 import transforms as TR
    B = 20000
    phi, x = make(B, F, K, torch.float32, bound)
    p64, x64 = phi.double().numpy(), x.double().numpy()
    for inverse in (False, True):
        run = TR.rqs_inverse if inverse else TR.rqs_forward
        orc = oracle.rqs_inverse if inverse else oracle.rqs_forward
        out, ladj, lsum, bins = run(phi.cuda(), x.cuda(), bound=bound, return_bins=True)
        o64, l64, b64 = orc(p64, x64, bound=bound)                      # fp64 truth on the same inputs
        o32, l32, b32 = orc(phi.numpy(), x.numpy(), bound=bound)        # fp32 oracle (what torch-fp32 zuko computes)
        bins = bins.cpu().numpy()
        mism = bins != b64
        assert mism.mean() < 1e-4                                       # only x within an ulp of a knot may move bin
        ok = ~mism & (b32 == b64)
        c_out, c_l = fp32_conditioning(orc, p64, x64, bound)
        e_out = np.abs(out.cpu().numpy() - o64)
        e_l = np.abs(ladj.cpu().numpy() - l64)
        print("fp32 F=%d K=%d inv=%s: out p50 %.1e max %.1e | ladj p50 %.1e p99 %.1e max %.1e | fp32-oracle ladj max %.1e"
              % (F, K, inverse, np.median(e_out), e_out[ok].max(), np.median(e_l), np.quantile(e_l, 0.99), e_l[ok].max(),
                 np.abs(l32 - l64)[ok].max()))
        assert (e_out[ok] <= 1e-5 + 8 * c_out[ok]).all()
        # log-det: no per-element bound is attempted (it also depends on the rounding of the knot differences);
        # the kernel's error distribution must match that of fp32 arithmetic itself (the fp32 oracle), below
        assert np.quantile(e_l, 0.5) < 1e-5
        e32 = np.abs(l32 - l64)
        for q in (0.5, 0.9, 0.99, 0.999, 0.9999):
            assert np.quantile(e_l[ok], q) < 3 * np.quantile(e32[ok], q) + 1e-5
        assert np.median(e_l) < 1e-5 and np.median(e_out) < 1e-6
        # never materially worse than the fp32 reference arithmetic itself
        assert e_l[ok].max() < 3 * np.abs(l32 - l64)[ok].max() + 1e-5
        np.testing.assert_allclose(lsum.cpu().numpy(), ladj.cpu().numpy().sum(1), atol=2e-4)


def test_searchsorted_bit_exact():
    """Bin search with the knots GIVEN must equal torch.searchsorted exactly."""
    from memflow_b200 import transforms as TR
    g = torch.Generator().manual_seed(0)
    for dt in (torch.float32, torch.float64):
        knots = torch.sort(torch.rand(5000, 65, generator=g, dtype=torch.float64).to(dt) * 2 - 1, dim=-1).values
        v = (torch.rand(5000, generator=g, dtype=torch.float64) * 2.4 - 1.2).to(dt)
        v[:100] = knots[:100, 7]                                        # exact hits on a knot: right=False semantics
        idx = TR.searchsorted(knots.cuda(), v.cuda()).cpu()
        ref = torch.searchsorted(knots, v[:, None]).squeeze(-1)
        assert torch.equal(idx.long(), ref)


def test_transform_api_and_golden(golden_dir):
    """zuko-shaped Transform surface; zero-copy packing of split views; golden fixture of the torch restatement."""
    from memflow_b200.transforms import MonotonicRQSTransform
    g = np.load(os.path.join(golden_dir, "rqs_f64_F7K10.npz"))
    phi = torch.from_numpy(g["phi"]).cuda()
    x = torch.from_numpy(g["x"]).cuda()
    K = 10
    w, h, d = phi.split([K, K, K - 1], dim=-1)                           # what zuko's MAF hands to the univariate
    t = MonotonicRQSTransform(w, h, d, bound=float(g["bound"]))
    assert t.phi.data_ptr() == phi.data_ptr()                            # no concatenation copy
    y, ladj = t.call_and_ladj(x)
    assert np.abs(y.cpu().numpy() - g["y"]).max() < 1e-12
    assert np.abs(ladj.cpu().numpy() - g["ladj"]).max() < 1e-11
    assert torch.equal(t(x), y)
    assert torch.equal(t.log_abs_det_jacobian(x, y), ladj)
    xi = t.inv(x)
    assert np.abs(xi.cpu().numpy() - g["x_inv"]).max() < 1e-12
    t2 = MonotonicRQSTransform(w.clone(), h.clone(), d.clone(), bound=float(g["bound"]))  # separate tensors -> packed copy
    assert torch.equal(t2(x), y)
    t3 = MonotonicRQSTransform(w, h, d, bound=float(g["bound"]), slope=None)              # zuko <= 0.1: no soft clip
    assert not torch.equal(t3(x), y)
    # parameters shared along the batch (an unconditional transform: phi [F, 3K-1], x [S, B, F]) broadcast like torch
    ts = MonotonicRQSTransform(w[3], h[3], d[3], bound=float(g["bound"]))
    xs = x[:24].reshape(2, 12, -1)
    ys, ls = ts.call_and_ladj(xs)
    tr = MonotonicRQSTransform(w[3].expand(2, 12, -1, -1), h[3].expand(2, 12, -1, -1), d[3].expand(2, 12, -1, -1), bound=float(g["bound"]))
    yr, lr = tr.call_and_ladj(xs)
    assert ys.shape == xs.shape and torch.equal(ys, yr) and torch.equal(ls, lr)
    assert torch.allclose(ts.inv(ys), xs, atol=1e-12, rtol=0)
    g32 = np.load(os.path.join(golden_dir, "rqs_f32_F20K64.npz"))
    p32, x32 = torch.from_numpy(g32["phi"]).cuda(), torch.from_numpy(g32["x"]).cuda()
    t = MonotonicRQSTransform(*p32.split([64, 64, 63], dim=-1), bound=float(g32["bound"]))
    y, ladj = t.call_and_ladj(x32)
    assert np.median(np.abs(y.cpu().numpy() - g32["y"])) < 1e-6


def test_ragged_and_unaligned(oracle):
    from memflow_b200 import transforms as TR
    for B in (1, 5, 127, 131):
        phi, x = make(B, 7, 10, torch.float32, 1.0, seed=B)
        y, l, s = TR.rqs_forward(phi.cuda(), x.cuda(), bound=1.0)
        oy, ol, _ = oracle.rqs_forward(phi.numpy(), x.numpy(), bound=1.0)
        assert np.abs(y.cpu().numpy() - oy).max() < 1e-4
    # parameter block whose base address is only 4-byte aligned -> cooperative-copy path instead of TMA
    phi, x = make(301, 7, 10, torch.float32, 1.0)
    buf = torch.empty(phi.numel() + 1, dtype=torch.float32, device="cuda")
    view = buf[1:].view(phi.shape); view.copy_(phi)
    y1 = TR._run("forward", view, x.cuda(), 1.0, 1e-3)[0] if view.is_contiguous() else None
    y0 = TR.rqs_forward(phi.cuda(), x.cuda(), bound=1.0)[0]
    assert y1 is not None and torch.equal(y1, y0)
    e = TR.rqs_forward(torch.empty(0, 7, 29, device="cuda"), torch.empty(0, 7, device="cuda"), bound=1.0)
    assert e[0].shape == (0, 7)
    # the cooperative kernels (several threads per row: every fp64 call, fp32 with other bin counts; forward, inverse and
    # backward): row counts that are not a multiple of the tile, and a base address that rules the bulk copy out
    for dt, F, K, bound in ((torch.float64, 10, 40, 1.1), (torch.float64, 3, 10, 1.0), (torch.float64, 2, 100, 1.0),
                            (torch.float32, 5, 24, 1.0)):
        tol = 1e-12 if dt == torch.float64 else 1e-4
        for B in (1, 7, 53, 200):
            phi, x = make(B, F, K, dt, bound, seed=B + K)
            buf = torch.empty(phi.numel() + 1, dtype=dt, device="cuda")
            view = buf[1:].view(phi.shape); view.copy_(phi)
            for fn, ofn in ((TR.rqs_forward, oracle.rqs_forward), (TR.rqs_inverse, oracle.rqs_inverse)):
                y, l, ls = fn(phi.cuda(), x.cuda(), bound=bound)
                y1, l1, ls1 = fn(view, x.cuda(), bound=bound)
                oy, ol, _ = ofn(phi.numpy(), x.numpy(), bound=bound)
                assert torch.equal(y, y1) and torch.equal(l, l1) and torch.equal(ls, ls1)
                assert np.abs(y.cpu().numpy() - oy).max() < tol * 100 and np.quantile(np.abs(y.cpu().numpy() - oy), 0.9) < tol
                assert torch.allclose(ls, l.sum(1), rtol=0, atol=1e-9 if dt == torch.float64 else 1e-3)
            # backward: the gradient of the aligned and of the unaligned copy must be the same bits
            grads = []
            for p in (phi.cuda(), view):
                p = p.detach().requires_grad_(True)
                yy, ll, _ = TR.rqs_forward(p, x.cuda(), bound=bound)
                (yy.sum() + 0.5 * ll.sum()).backward()
                grads.append(p.grad.clone())
            assert torch.equal(grads[0], grads[1]) and bool(torch.isfinite(grads[0]).all())


# fp64 with the soft clip, K = 6..64: rqs64_backward_kernel (2 / 4 / 8 threads per row for K <= 10 / 20 / 64, odd K in
# padded rows); otherwise K < 12: one thread per row; 12..23 / 24..47 / >= 48: 2 / 4 / 8 threads per row (rqs_backward_coop_kernel)
@pytest.mark.parametrize("F,K,bound,slope", [(7, 10, 1.0, 1e-3), (4, 33, 1.1, 1e-3), (3, 8, 5.0, None), (5, 16, 1.0, 1e-3),
                                             (3, 64, 1.1, 1e-3), (2, 50, 2.0, None), (3, 13, 1.0, None), (10, 40, 1.1, 1e-3),
                                             (6, 35, 1.1, 1e-3), (2, 20, 1.0, 1e-3), (3, 6, 1.0, 1e-3), (3, 45, 2.0, 1e-3),
                                             (2, 7, 1.0, 1e-3), (2, 19, 1.0, 1e-3), (3, 5, 1.0, 1e-3)])
def test_spline_autograd_vs_torch_restatement(F, K, bound, slope):
    """SURVEY 8f-4 for K3/K4: gradients w.r.t. the parameters and the input == torch autograd through the op-for-op
    restatement of zuko (oracle/rqs_torch.py), both directions, cotangents on the output AND on the log-det."""
    from memflow_b200 import transforms as TR
    from oracle import rqs_torch
    B = 257
    g = torch.Generator().manual_seed(21)
    phi0 = torch.randn(B, F, 3 * K - 1, generator=g, dtype=torch.float64)
    x0 = (torch.rand(B, F, generator=g, dtype=torch.float64) * 2.4 - 1.2) * bound
    gy = torch.randn(B, F, generator=g, dtype=torch.float64)
    gl = torch.randn(B, F, generator=g, dtype=torch.float64)
    for inverse in (False, True):
        # reference gradients (CPU, torch autograd)
        phi_r, x_r = phi0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        t = rqs_torch.MonotonicRQSTransform(phi_r[..., :K], phi_r[..., K:2 * K], phi_r[..., 2 * K:], bound=bound, slope=slope)
        if not inverse:
            y_r, l_r = t.call_and_ladj(x_r)
        else:
            y_r = t._inverse(x_r)
            l_r = t.call_and_ladj(y_r)[1]          # forward log-det at the recovered point
        ((y_r * gy).sum() + (l_r * gl).sum()).backward()
        # kernels
        phi_k, x_k = phi0.clone().cuda().requires_grad_(True), x0.clone().cuda().requires_grad_(True)
        run = TR.rqs_inverse if inverse else TR.rqs_forward
        y_k, l_k, ls_k = run(phi_k, x_k, bound=bound, slope=slope)
        assert y_k.requires_grad and l_k.requires_grad
        ((y_k * gy.cuda()).sum() + (l_k * gl.cuda()).sum()).backward()
        for got, ref, name in ((phi_k.grad, phi_r.grad, "phi"), (x_k.grad, x_r.grad, "x")):
            got, ref = got.cpu().numpy().reshape(B, -1), ref.numpy().reshape(B, -1)
            rel = np.abs(got - ref) / (np.abs(ref).max(axis=1, keepdims=True) + 1e-300)
            print("spline autograd K=%d inv=%s %s: max rel %.2e median %.2e" % (K, inverse, name, rel.max(), np.median(rel)))
            assert np.median(rel) < 1e-12 and np.quantile(rel, 0.999) < 1e-8 and rel.max() < 1e-6
    # the Transform surface is differentiable too, separate parameter tensors get their own gradients
    w = phi0[..., :K].clone().cuda().requires_grad_(True)
    h = phi0[..., K:2 * K].clone().cuda().requires_grad_(True)
    d = phi0[..., 2 * K:].clone().cuda().requires_grad_(True)
    tk = TR.MonotonicRQSTransform(w, h, d, bound=bound, slope=slope)
    yk, lk = tk.call_and_ladj(x0.cuda())
    (yk.sum() + lk.sum()).backward()
    assert w.grad is not None and h.grad is not None and d.grad is not None and float(w.grad.abs().sum()) > 0
    # fp32 path runs and is close to the fp64 gradient
    p32, x32 = phi0.float().cuda().requires_grad_(True), x0.float().cuda()
    y32, l32, _ = TR.rqs_forward(p32, x32, bound=bound, slope=slope)
    (y32 * gy.float().cuda()).sum().backward()
    assert torch.isfinite(p32.grad).all()


@pytest.mark.parametrize("F,K", [(10, 40), (10, 35), (7, 10), (8, 16)])
def test_fp64_backward_many_tiles_and_bad_rows(F, K):
    """rqs64_backward_kernel with more tiles than resident CTAs (every CTA loops: bulk store of tile i, then the bulk copy of
    tile i + grid into the same rows), against torch autograd of the restatement; a row with a non-finite parameter gets
    NaN gradients and leaves every other row untouched; inputs in the identity tails get zero parameter gradients."""
    from memflow_b200 import transforms as TR
    from oracle import rqs_torch
    B = 6007
    g = torch.Generator().manual_seed(5)
    phi0 = torch.randn(B, F, 3 * K - 1, generator=g, dtype=torch.float64)
    x0 = (torch.rand(B, F, generator=g, dtype=torch.float64) * 2.4 - 1.2) * 1.1
    gy = torch.randn(B, F, generator=g, dtype=torch.float64)
    gl = torch.randn(B, F, generator=g, dtype=torch.float64)
    for inverse in (False, True):
        phi_r, x_r = phi0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        t = rqs_torch.MonotonicRQSTransform(phi_r[..., :K], phi_r[..., K:2 * K], phi_r[..., 2 * K:], bound=1.1)
        if not inverse:
            y_r, l_r = t.call_and_ladj(x_r)
        else:
            y_r = t._inverse(x_r)
            l_r = t.call_and_ladj(y_r)[1]
        ((y_r * gy).sum() + (l_r * gl).sum()).backward()
        phi_k, x_k = phi0.clone().cuda().requires_grad_(True), x0.clone().cuda().requires_grad_(True)
        y_k, l_k, _ = (TR.rqs_inverse if inverse else TR.rqs_forward)(phi_k, x_k, bound=1.1)
        ((y_k * gy.cuda()).sum() + (l_k * gl.cuda()).sum()).backward()
        for got, ref, name in ((phi_k.grad, phi_r.grad, "phi"), (x_k.grad, x_r.grad, "x")):
            got, ref = got.cpu().numpy().reshape(B, -1), ref.numpy().reshape(B, -1)
            rel = np.abs(got - ref) / (np.abs(ref).max(axis=1, keepdims=True) + 1e-300)
            print("many tiles K=%d inv=%s %s: max rel %.2e median %.2e" % (K, inverse, name, rel.max(), np.median(rel)))
            assert np.median(rel) < 1e-12 and np.quantile(rel, 0.999) < 1e-8 and rel.max() < 1e-6
        tails = (x0.abs() >= 1.1)
        assert tails.any() and float(phi_k.grad.cpu()[tails].abs().max()) == 0.0
        # poisoned rows
        bad = phi0.clone()
        bad[17, 1, 3] = float("inf")
        bad[4000, F - 1, K + 2] = float("nan")
        bad[5000, 0, 0] = 1e300
        pb = bad.cuda().requires_grad_(True)
        yb, lb, _ = (TR.rqs_inverse if inverse else TR.rqs_forward)(pb, x0.cuda(), bound=1.1)
        ((yb * gy.cuda()).sum() + (lb * gl.cuda()).sum()).backward()
        gb = pb.grad.cpu()
        badrow = torch.zeros(B, F, dtype=torch.bool)
        badrow[17, 1] = badrow[4000, F - 1] = badrow[5000, 0] = True
        assert bool(torch.isnan(gb[badrow]).all())
        assert torch.equal(gb[~badrow], phi_k.grad.cpu()[~badrow])


@pytest.mark.parametrize("F,K", [(10, 40), (6, 35), (7, 10), (3, 64)])
def test_fp64_backward_unaligned_pointers_same_bits(F, K):
    """Parameters (and therefore gradients) that are 8- but not 16-byte aligned cannot take the bulk copy / bulk store:
    every tile then goes through the plain-copy path of rqs64_backward_kernel. Same arithmetic: bit-identical gradients."""
    import memflow_b200.ops  # noqa: F401
    B, PW = 2999, 3 * K - 1
    g = torch.Generator().manual_seed(3)
    phi = torch.randn(B, F, PW, generator=g, dtype=torch.float64).cuda()
    x = ((torch.rand(B, F, generator=g, dtype=torch.float64) * 2.4 - 1.2) * 1.1).cuda()
    gy = torch.randn(B, F, generator=g, dtype=torch.float64).cuda()
    gl = torch.randn(B, F, generator=g, dtype=torch.float64).cuda()
    flat = torch.empty(B * F * PW + 1, dtype=torch.float64, device="cuda")
    phi_u = flat[1:].view(B, F, PW)
    phi_u.copy_(phi)
    assert phi_u.data_ptr() % 16 == 8 and phi.data_ptr() % 16 == 0
    for inverse in (False, True):
        ga, gva = torch.ops.memflow_b200.rqs_bwd(phi, x, gy, gl, 1.1, 1e-3, inverse)
        gu, gvu = torch.ops.memflow_b200.rqs_bwd(phi_u, x, gy, gl, 1.1, 1e-3, inverse)
        assert torch.equal(ga, gu) and torch.equal(gva, gvu) and bool(torch.isfinite(ga).all())


@pytest.mark.parametrize("dt,F,K", [(torch.float64, 10, 40), (torch.float64, 6, 35), (torch.float32, 10, 40), (torch.float32, 7, 10), (torch.float64, 7, 10)])
def test_feature_subset_reads_rows_in_place(dt, F, K):
    """Autoregressive sampling: pass j of zuko's MaskedAutoregressiveTransform inverse determines feature j only, so it
    should read that feature's 3K-1 parameters, not the event's F (3K-1) (call sites memflow/unfolding_flow/unfolding_flow.py:
    97,186). The strided entry points take the column view of the parameter tensor in place: same bits as the full call."""
    from memflow_b200 import ops, transforms as TR
    B, bound = 3001, 1.1
    phi, x = make(B, F, K, dt, bound, seed=3)
    phi, x = phi.cuda(), x.cuda()
    for fn in (TR.rqs_forward, TR.rqs_inverse):
        full, lfull, _, bfull = fn(phi, x, bound=bound, return_bins=True)
        for j0, fc in ((0, 1), (F - 1, 1), (2, 3), (1, F - 1)):
            sub = phi[:, j0:j0 + fc]                      # a view: events F (3K-1) elements apart, fc rows each
            assert not sub.is_contiguous() or fc == F
            kept, stride = ops._rqs_rows(sub)
            assert kept.data_ptr() == sub.data_ptr() and stride == F * (3 * K - 1)   # read in place, no gather copy
            o, l, ls, b = fn(sub, x[:, j0:j0 + fc], bound=bound, return_bins=True)
            assert torch.equal(o, full[:, j0:j0 + fc]) and torch.equal(l, lfull[:, j0:j0 + fc]) and torch.equal(b, bfull[:, j0:j0 + fc])
            assert torch.allclose(ls, lfull[:, j0:j0 + fc].sum(1), rtol=0, atol=1e-9 if dt == torch.float64 else 1e-3)
    # F passes, one feature each == one pass over all features (parameters given), through the Transform surface
    w, h, d = phi.split([K, K, K - 1], dim=-1)
    t_full = TR.MonotonicRQSTransform(w, h, d, bound=bound)
    x_full = t_full.inv(x)
    cols = []
    for j in range(F):
        tj = TR.MonotonicRQSTransform(w[:, j], h[:, j], d[:, j], bound=bound)
        assert tj.phi.data_ptr() == phi[:, j].data_ptr() and tj.phi.stride(0) == F * (3 * K - 1)   # zero-copy column
        cols.append(tj.inv(x[:, j]))
    assert torch.equal(torch.stack(cols, dim=1), x_full)


def test_registered_ops_opcheck_and_compile():
    """The torch.library layer on the device: schema / fake kernel / autograd registration are mutually consistent
    (torch.library.opcheck), the ops trace under torch.compile, and the eager fast path gives the same bits as the op."""
    from memflow_b200 import ops, transforms as TR
    phi, x = make(33, 4, 10, torch.float64, 1.0, seed=5)
    phi, x = phi.cuda(), x.cuda()
    args = (phi.clone().requires_grad_(True), x.clone().requires_grad_(True), 1.0, 1e-3, False, ops.RQS_LADJ | ops.RQS_SUM)
    torch.library.opcheck(torch.ops.memflow_b200.rqs, args, test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    u = torch.rand(17, 7, dtype=torch.float64, device="cuda").clamp_(1e-6, 1 - 1e-6)
    torch.library.opcheck(torch.ops.memflow_b200.rambo_forward, (u.clone().requires_grad_(True), [125.25, 172.5, 172.5], 13000.0, [], 0, True),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    a = torch.ops.memflow_b200.rqs(phi, x, 1.0, 1e-3, False, ops.RQS_LADJ | ops.RQS_SUM)
    b = ops.rqs_cuda(phi, x, 1.0, 1e-3, False, ops.RQS_LADJ | ops.RQS_SUM)
    assert all(torch.equal(p, q) for p, q in zip(a, b))

    def f(p, v):
        y, ladj, lsum = TR.rqs_forward(p, v, bound=1.0)
        return y * 2.0 + lsum[:, None]

    ref = f(phi, x)
    got = torch.compile(f, backend="aot_eager", fullgraph=False)(phi, x)
    assert torch.allclose(got, ref, rtol=0, atol=0)
