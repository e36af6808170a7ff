"""GPU parity of K1 (rambo_forward) and K2 (rambo_inverse) against the oracle and the committed golden
vectors, through the reference-shaped Python API (which calls the C ABI). Tolerances: tests/parity.py."""
import os

import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu
E = 13000.0
MASSES = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5], 5: [125.25, 172.5, 172.5, 1e-5, 1e-5],
          6: [125.25, 172.5, 172.5, 4.7, 4.7, 1e-5], 7: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3],
          8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]}
TOL64 = 1e-12


def make_ps(n, **opts):
    from memflow_b200.phasespace.phasespace import PhaseSpace
    return PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(MASSES[n], dtype=torch.float64),
                      dev="cuda:0", **opts)


def rand_u(N, n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(N, 3 * n - 2, dtype=torch.float64, generator=g).clamp(1e-10, 1 - 1e-10)


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8])
def test_forward_vs_oracle(oracle, n):
    N = 100_000
    u = rand_u(N, n, 100 + n)
    ps = make_ps(n)
    mom, w, x1, x2, lw = ps.get_momenta_from_ps(u.cuda(), return_logweight=True)
    assert mom.dtype == torch.float64 and mom.shape == (N, n + 2, 4) and mom.is_cuda
    om, ow, ox1, ox2 = oracle.rambo_forward(u.numpy(), MASSES[n], E)
    kappa = parity.conditioning(u.numpy(), om, n, MASSES[n])
    err = parity.momenta_error(mom.cpu().numpy(), om, ox1, ox2, E)
    s = parity.summarize(err, TOL64, kappa)
    print("n=%d momenta" % n, s)
    assert s["p50"] < 1e-14
    assert s["p99"] < TOL64
    assert s["max_over_kappa"] < TOL64, s
    if n <= 4:  # the tt-bar-H(+j) configurations of BASELINE.json: every single event inside 1e-12
        assert s["max"] < TOL64, s
    np.testing.assert_allclose(x1.cpu().numpy(), ox1, rtol=0, atol=TOL64)
    np.testing.assert_allclose(x2.cpu().numpy(), ox2, rtol=0, atol=TOL64)
    dl = np.abs(np.log(w.cpu().numpy()) - np.log(ow))
    sl = parity.summarize(dl, TOL64, kappa)
    print("n=%d log-weight" % n, sl)
    assert sl["p99"] < TOL64 and sl["max_over_kappa"] < TOL64, sl
    assert float((lw - w.log()).abs().max()) < 1e-13
    assert int(ps.generator.last_status.item()) == 0


@pytest.mark.parametrize("n", [3, 4, 5, 8])
def test_forward_vs_reference_golden(golden_dir, n):
    """Against the UNMODIFIED reference's outputs (fp64-clean variant and literal float32-quirk variant)."""
    g = np.load(os.path.join(golden_dir, "rambo_n%d.npz" % n))
    u = torch.from_numpy(g["u"]).cuda()
    for tag, quirks in (("clean", False), ("lit", True)):
        ps = make_ps(n, ref_fp32_quirks=quirks)
        mom, w, x1, x2 = ps.get_momenta_from_ps(u)
        kappa = parity.conditioning(g["u"], g["mom_" + tag], n, MASSES[n])
        err = parity.momenta_error(mom.cpu().numpy(), g["mom_" + tag], g["x1_" + tag], g["x2_" + tag], E)
        assert np.nanmax(err / kappa) < TOL64, (tag, np.nanmax(err / kappa))
        assert np.quantile(err, 0.9) < TOL64
        rw = g["w_" + tag]
        assert np.array_equal(np.isinf(w.cpu().numpy()), np.isinf(rw))
        fin = np.isfinite(rw) & (rw > 0)
        dl = np.abs(np.log(w.cpu().numpy()[fin]) - np.log(rw[fin]))
        tol = 5e-7 if quirks else TOL64
        assert np.nanmax(dl / kappa[fin]) < tol, (tag, np.nanmax(dl / kappa[fin]))
        if quirks:  # incoming rows carry the reference's float32 rounding exactly
            assert np.array_equal(mom.cpu().numpy()[24:, 0, 0], g["mom_lit"][24:, 0, 0])
        # literal weights agree with the clean ones only to float32 precision
    lit, clean = g["w_lit"], g["w_clean"]
    f = np.isfinite(lit)
    assert 1e-9 < np.nanmax(np.abs(lit[f] / clean[f] - 1)) < 1e-6


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8])
def test_inverse_vs_oracle(oracle, n):
    N = 100_000
    u = rand_u(N, n, 200 + n)
    om, ow, ox1, ox2 = oracle.rambo_forward(u.numpy(), MASSES[n], E)
    ps = make_ps(n)
    full = torch.from_numpy(om).cuda()
    # the [:, 2:] slice is read in place (event pitch 4(n+2)), no copy
    out, prob, lp = ps.generator.getPSpoint_batch(full[:, 2:], torch.from_numpy(ox1).cuda(),
                                                  torch.from_numpy(ox2).cuda(), return_logprob=True)
    ops, oprob, rc = oracle.rambo_inverse(om[:, 2:], ox1, ox2, MASSES[n], E)
    assert rc == 0 and int(ps.generator.last_status.item()) == 0
    kappa = parity.conditioning(u.numpy(), om, n, MASSES[n])
    d = np.abs(out.cpu().numpy() - ops).max(axis=1)
    s = parity.summarize(d, TOL64, kappa)
    print("n=%d inverse ps" % n, s)
    assert s["p99"] < TOL64 and s["max_over_kappa"] < TOL64, s
    dl = np.abs(np.log(prob.cpu().numpy()) - np.log(oprob))
    assert np.nanquantile(dl, 0.99) < TOL64 and np.nanmax(dl / kappa) < TOL64
    assert float((lp - prob.log()).abs().max()) < 1e-13


@pytest.mark.parametrize("n", [4, 8])
def test_inverse_permuted_offshell_and_golden(oracle, golden_dir, n):
    g = np.load(os.path.join(golden_dir, "rambo_n%d.npz" % n))
    ps = make_ps(n)
    mom = torch.from_numpy(g["mom_clean"]).cuda()
    x1, x2 = torch.from_numpy(g["x1_clean"]).cuda(), torch.from_numpy(g["x2_clean"]).cuda()
    kappa = parity.conditioning(g["u"], g["mom_clean"], n, MASSES[n])
    out, prob = ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2)
    assert np.nanmax(np.abs(out.cpu().numpy() - g["inv_ps_clean"]).max(axis=1) / kappa) < TOL64
    perm = [int(i) for i in g["perm"]]
    out, prob = ps.generator.getPSpoint_batch(mom[:, 2:].contiguous(), x1, x2, order=perm)
    ops, oprob, _ = oracle.rambo_inverse(g["mom_clean"][:, 2:], g["x1_clean"], g["x2_clean"], MASSES[n], E, order=perm)

    def kappa_for(ps_ref, masses_perm):
        """conditioning of the PERMUTED inverse: its intermediate systems are sums over the permuted bodies, its angle /
        mass variables are the inverse's own outputs"""
        full = np.concatenate((g["mom_clean"][:, :2], g["mom_clean"][:, 2:][:, perm]), axis=1)
        k = parity.conditioning(np.nan_to_num(ps_ref, nan=0.5), full, n, masses_perm)
        return np.where(np.isfinite(k), k, np.inf)

    def check(got, ref, kap, what):
        assert np.array_equal(np.isnan(got), np.isnan(ref)), what
        d = np.nanmax(np.abs(got - ref), axis=1)
        fin = np.isfinite(d)
        s = parity.summarize(d[fin], TOL64, kap[fin])
        print("n=%d %s" % (n, what), s)
        assert s["max_over_kappa"] < TOL64, (what, s)               # every event inside 1e-12 * kappa(event)
        assert s["p50"] < 1e-14 and s["frac_gt_tol"] < 0.15, (what, s)  # (the fixture's first 24 rows are edge-of-cube points)

    mperm = [MASSES[n][i] for i in perm]
    check(out.cpu().numpy(), ops, kappa_for(ops, mperm), "permuted inverse vs oracle")
    check(out.cpu().numpy(), g["inv_ps_perm_clean"], kappa_for(g["inv_ps_perm_clean"], mperm), "permuted inverse vs reference golden")
    fin = np.isfinite(g["inv_prob_perm_clean"]) & (g["inv_prob_perm_clean"] > 0)
    dl = np.abs(np.log(prob.cpu().numpy()[fin]) - np.log(g["inv_prob_perm_clean"][fin]))
    assert np.nanmax(dl / kappa_for(g["inv_ps_perm_clean"], mperm)[fin]) < TOL64
    out, prob = ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2, order=perm, target_mass=torch.from_numpy(g["target_mass"]),
                                              ensure_onShell=False)
    ref = g["inv_ps_offshell_clean"]
    check(out.cpu().numpy(), ref, kappa_for(ref, [float(v) for v in g["target_mass"].reshape(-1)]), "off-shell inverse vs reference golden")


def test_lab_frame_direct_x1x2_and_intermediates_vs_reference_golden(golden_dir):
    """Reference-run fixtures of round 2 (tests/golden/make_golden_r2.py): the fused lab-frame epilogue against the
    reference's own boost_to_lab_frame (utils.py:180-202), uniform_x1x2=False (rambo_generator.py:229-234) and the two
    intermediates methods called on their own (:448-509)."""
    from memflow_b200.phasespace.rambo_generator import FlatInvertiblePhasespace
    gx = np.load(os.path.join(golden_dir, "rambo_extra.npz"))
    for n in (3, 4, 8):
        g = np.load(os.path.join(golden_dir, "rambo_n%d.npz" % n))
        ps = make_ps(n)
        lab = ps.get_momenta_from_ps(torch.from_numpy(g["u"]).cuda(), lab_frame=True)[0].cpu().numpy()
        ref = gx["lab_n%d" % n]
        assert np.array_equal(np.isnan(lab), np.isnan(ref))
        kappa = parity.conditioning(g["u"], g["mom_clean"], n, MASSES[n])
        scale = np.nanmax(np.abs(ref), axis=(1, 2))                 # the lab boost stretches the event by gamma_lab
        err = np.nanmax(np.abs(lab - ref), axis=(1, 2)) / scale
        fin = np.isfinite(err)
        assert np.nanmax(err[fin] / kappa[fin]) < TOL64 and np.nanquantile(err[fin], 0.9) < TOL64, n
    # uniform_x1x2 = False: the last two columns are x1, x2 themselves, no Jacobian
    n = 4
    gen = FlatInvertiblePhasespace([0.0, 0.0], torch.tensor(MASSES[n], dtype=torch.float64), E, pdf=None, pdf_active=True,
                                   uniform_x1x2=False, dev="cuda:0")
    u = gx["direct_u_n4"]
    mom, w, x1, x2 = gen.generateKinematics_batch(torch.from_numpy(u).cuda())
    assert np.array_equal(x1.cpu().numpy(), u[:, -2]) and np.array_equal(x2.cpu().numpy(), u[:, -1])   # passed through
    kappa = parity.conditioning(u, gx["direct_mom_n4"], n, MASSES[n])
    err = parity.momenta_error(mom.cpu().numpy(), gx["direct_mom_n4"], gx["direct_x1_n4"], gx["direct_x2_n4"], E)
    assert np.nanmax(err / kappa) < TOL64 and np.nanquantile(err, 0.9) < TOL64
    fin = np.isfinite(gx["direct_w_n4"]) & (gx["direct_w_n4"] > 0)
    dl = np.abs(np.log(w.cpu().numpy()[fin]) - np.log(gx["direct_w_n4"][fin]))
    assert np.nanmax(dl / kappa[fin]) < TOL64 and np.nanquantile(dl, 0.9) < TOL64
    # and the inverse undoes it: x columns come back as they went in
    back, prob = gen.getPSpoint_batch(mom[:, 2:], x1, x2)
    d = np.abs(back.cpu().numpy() - u)
    assert np.nanquantile(d, 0.9) < 1e-9 and np.array_equal(back.cpu().numpy()[:, -2:], u[:, -2:])
    # generateIntermediatesMassless_batch / generateIntermediatesMassive_batch on their own
    for n in (4, 8):
        gen = make_ps(n).generator
        u = torch.from_numpy(gx["inter_u_n%d" % n]).cuda()
        ecm = torch.from_numpy(gx["inter_ecm_n%d" % n]).cuda()
        M = torch.zeros(u.shape[0], n - 1, dtype=torch.float64, device="cuda")
        M[:, 0] = ecm
        w0, K = gen.generateIntermediatesMassless_batch(M, ecm, u)
        w1, Mi = gen.generateIntermediatesMassive_batch(M, ecm, u)
        uu = gx["inter_u_n%d" % n][:, : n - 2]
        kap = 1.0 + 1e-2 / np.minimum(uu, 1 - uu).min(axis=1) ** 2      # parity.conditioning's mass-variable factor
        for got, ref in ((K, gx["inter_massless_K_n%d" % n]), (Mi, gx["inter_massive_M_n%d" % n])):
            rel = np.abs(got.cpu().numpy() - ref).max(axis=1) / gx["inter_ecm_n%d" % n]
            assert np.max(rel / kap) < TOL64, n
        for got, ref in ((w0, gx["inter_massless_w_n%d" % n]), (w1, gx["inter_massive_w_n%d" % n])):
            dl = np.abs(np.log(got.cpu().numpy()) - np.log(ref))
            assert np.max(dl / kap) < TOL64 and np.median(dl) < 1e-14, (n, dl.max())
        w0s, _ = gen.generateIntermediatesMassless_batch(M, 13000.0, u)
        np.testing.assert_allclose(w0s.cpu().numpy(), gx["inter_massless_w_scalar_n%d" % n], rtol=1e-15)


def test_round_trip_config2_full_size():
    """BASELINE config 2: u -> p -> u' for the 8-body final state, 16M events, fp64 (size-independent property;
    the oracle is not involved). Also prob * weight == 1 and 4-momentum conservation."""
    n, N = 8, 16_000_000
    ps = make_ps(n)
    g = torch.Generator(device="cuda").manual_seed(2024)
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, device="cuda", generator=g).clamp_(1e-10, 1 - 1e-10)
    mom, w, x1, x2 = ps.get_momenta_from_ps(u)
    back, prob = ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2)
    assert int(ps.generator.last_status.item()) == 0
    res = (back - u).abs()
    del back
    med = float(res.flatten()[::97].median())
    frac = float((res.max(dim=1).values > 1e-9).double().mean())
    print("config 2 round trip: median %.2e, max %.2e, frac(>1e-9) %.2e" % (med, float(res.max()), frac))
    assert med < 1e-15 and frac < 1e-3
    pw = (prob * w - 1).abs()
    assert float(pw.median()) < 1e-14 and float((pw > 1e-9).double().mean()) < 1e-3
    tot = mom[:, 2:].sum(dim=1)
    ecm = torch.sqrt(x1 * x2) * E
    assert float(((tot[:, 0] - ecm).abs() / ecm).max()) < 1e-12
    assert float((tot[:, 1:].abs().max(dim=1).values / ecm).max()) < 1e-12


def test_edges_and_status_flags(oracle):
    n = 3
    ps = make_ps(n)
    u = torch.full((64, 7), 0.5, dtype=torch.float64)
    u[1, 0] = 0.0; u[2, 0] = 1.0; u[3, 5] = 1.0   # mass variable 0 / 1 (bisection floor 2^-180 / ~1), u_x1 = 1 -> NaN
    mom, w, x1, x2 = ps.get_momenta_from_ps(u.cuda())
    om, ow, ox1, ox2 = oracle.rambo_forward(u.numpy(), MASSES[n], E)
    a, b = mom.cpu().numpy(), om
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    ok[3] = False   # u_x1 == 1: x1 x2 = tau, K_0 = rounding noise -> only the NaN pattern is defined
    ok[2] = False   # v == 1 exactly: the bisection stalls ~1e-8 below 1 where u(2-u) rounds to 1; kernel returns 1-2^-53
    assert np.abs(a[ok] - b[ok]).max() / E < 1e-12
    wn = w.cpu().numpy()
    assert np.array_equal(np.isnan(wn), np.isnan(ow))
    assert np.isfinite(wn[1]) and np.isfinite(wn[2]) and np.isfinite(ow[1]) and np.isfinite(ow[2])   # finite
    assert int(ps.generator.last_status.item()) == 0
    u[5, 2] = float("nan")
    ps.get_momenta_from_ps(u.cuda())
    assert int(ps.generator.last_status.item()) & 1                # MF_ST_NAN_INPUT, no host sync needed
    from memflow_b200.phasespace.rambo_generator import PhaseSpaceGeneratorError
    strict = make_ps(n, sync_errors=True)
    with pytest.raises(PhaseSpaceGeneratorError):                  # reference behaviour on request (:191-199)
        strict.get_momenta_from_ps(u.cuda())
    bad = torch.zeros(8, 3, 4, dtype=torch.float64, device="cuda"); bad[:, :, 0] = 300.0; bad[:, 0, 1] = 50.0
    with pytest.raises(Exception, match="not in the CM"):          # :634-635
        strict.get_ps_from_momenta(bad, torch.full((8,), 0.1, device="cuda"), torch.full((8,), 0.2, device="cuda"))
    with pytest.raises(AssertionError):                            # :251 wrong column count
        ps.get_momenta_from_ps(torch.rand(4, 6, dtype=torch.float64, device="cuda"))
    # empty batch
    mom, w, x1, x2 = ps.get_momenta_from_ps(torch.empty(0, 7, dtype=torch.float64, device="cuda"))
    assert mom.shape == (0, 5, 4) and w.numel() == 0
    # ragged sizes around the block size
    for N in (1, 255, 257):
        uu = rand_u(N, n, N)
        m1 = ps.get_momenta_from_ps(uu.cuda())[0].cpu().numpy()
        m2 = oracle.rambo_forward(uu.numpy(), MASSES[n], E)[0]
        assert np.abs(m1 - m2).max() / E < 1e-12


def test_cuts_lab_frame_and_float_inputs(oracle):
    n, N = 4, 50_000
    u = rand_u(N, n, 9)
    ps = make_ps(n)
    mom, w, x1, x2 = ps.get_momenta_from_ps(u.cuda(), pT_mincut=20.0, delR_mincut=0.4, rap_maxcut=5.0)
    _, ow, _, _ = oracle.rambo_forward(u.numpy(), MASSES[n], E, pT_mincut=20.0, delR_mincut=0.4, rap_maxcut=5.0)
    z, oz = (w.cpu().numpy() == 0), (ow == 0)
    assert 0.05 < oz.mean() < 0.95
    assert (z != oz).mean() < 1e-4                                  # cut decisions: only knife-edge events may differ
    # lab frame fused == reference's boost_to_lab_frame applied to the CM output
    from memflow_b200.phasespace import utils as U
    mom_cm, _, x1, x2 = ps.get_momenta_from_ps(u.cuda())
    mom_lab = ps.get_momenta_from_ps(u.cuda(), lab_frame=True)[0]
    ref_lab = U.boost_to_lab_frame(mom_cm, x1, x2)
    assert float((mom_lab - ref_lab).abs().max()) / E < 1e-12
    # float32 points in -> float64 out, like the reference (rambo_generator.py:201-215)
    m32 = ps.get_momenta_from_ps(u.float().cuda())[0]
    assert m32.dtype == torch.float64
    # generate_random_phase_space_points (phasespace.py:64-80)
    r, m, ww, a, b = ps.generate_random_phase_space_points(1000)
    assert r.shape == (1000, 10) and m.shape == (1000, 6, 4) and bool((ww > 0).all())


def test_fp32_kernels(oracle):
    """fp32 compute path (north_star: 1e-5 relative in fp32 for momenta and log-Jacobians): against the fp64 oracle on the
    same (fp32-rounded) inputs, every event inside 1e-5 * kappa(event) -- the fp64 tests' bound with the fp32 tolerance and
    the same conditioning factor (tests/parity.py) -- and the bulk inside the plain 1e-5."""
    for n, N in ((3, 100_000), (4, 100_000)):
        u32 = rand_u(N, n, 77 + n).float().clamp(1e-6, 1 - 1e-6)
        ps = make_ps(n, compute_dtype=torch.float32)
        mom, w, x1, x2 = ps.get_momenta_from_ps(u32.cuda())
        u64 = u32.double().numpy()
        om, ow, ox1, ox2 = oracle.rambo_forward(u64, MASSES[n], E)
        kappa = parity.conditioning(u64, om, n, MASSES[n])
        err = parity.momenta_error(mom.cpu().numpy(), om, ox1, ox2, E)
        s = parity.summarize(err, 1e-5, kappa)
        print("fp32 forward n=%d momenta" % n, s)
        assert s["max_over_kappa"] < 1e-5 and s["p99"] < 1e-5 and s["p50"] < 2e-7, s
        dl = np.abs(np.log(w.cpu().numpy()) - np.log(ow))
        # log-weight: the float32 rounding of x1, x2 is amplified by E_cm / (E_cm - sum m) (parity.conditioning_fp32_weight)
        sl = parity.summarize(dl, 1e-5, parity.conditioning_fp32_weight(kappa, ox1, ox2, MASSES[n], E))
        print("fp32 forward n=%d ln w" % n, sl)
        assert sl["max_over_kappa"] < 1e-5 and sl["p50"] < 1e-5 and sl["frac_gt_tol"] < 0.05, sl
        for a, b in ((x1, ox1), (x2, ox2)):
            assert float(np.max(np.abs(a.cpu().numpy() - b) / b)) < 1e-5
        back, prob = ps.generator.getPSpoint_batch(torch.from_numpy(om[:, 2:]).float().cuda(),
                                                   torch.from_numpy(ox1).float().cuda(), torch.from_numpy(ox2).float().cuda())
        d = np.abs(back.cpu().numpy() - u64).max(axis=1)
        si = parity.summarize(d, 1e-5, kappa)
        print("fp32 inverse n=%d u" % n, si)
        # the inverse starts from momenta ROUNDED to fp32 (6e-8 relative on TeV-scale components against GeV-scale
        # masses): the invariant masses M_j^2 = E^2 - p^2 lose E^2 / M_j^2 digits before the kernel sees them
        assert si["p50"] < 1e-5 and si["max_over_kappa"] < 1e-5 * (1 if n == 3 else 1e3), si


def test_massvar_solver_vs_bisection(oracle):
    """bisect_vec_batch replacement: solves v = u^e((e+1) - e u) to <= 2 ulp-level residual for all exponents."""
    n = 8
    ps = make_ps(n)
    g = torch.Generator().manual_seed(3)
    v = torch.rand(200_000, n - 2, dtype=torch.float64, generator=g)
    v[:1000] = 10.0 ** (-30 * torch.rand(1000, n - 2, dtype=torch.float64, generator=g))
    v[1000:2000] = 1 - 10.0 ** (-16 * torch.rand(1000, n - 2, dtype=torch.float64, generator=g))
    v[2000:3000] = 10.0 ** (-30 - 250 * torch.rand(1000, n - 2, dtype=torch.float64, generator=g))   # below the fp32 start's range
    v[3000:4000] = 0.65 + (torch.rand(1000, n - 2, dtype=torch.float64, generator=g) - 0.5) * 1e-6        # the variable switch
    u = ps.generator.bisect_vec_batch(v.cuda()).cpu()
    e = torch.arange(n - 2, 0, -1, dtype=torch.float64)
    res = ((u ** e) * ((e + 1) - e * u) - v).abs()
    assert float((res / v.clamp_min(1e-300)).max()) < 5e-15 or float(res.max()) < 1e-15
    assert bool(((u >= 0) & (u <= 1)).all())
    # against the root itself: refined in 80-bit arithmetic in the well-conditioned variable (u for small v, d = 1 - u
    # with 1 - v = d^2 P_e(d) for large v), the returned u must be within 4 ulp (the residual's own rounding: P_e has
    # alternating coefficients)
    from math import comb
    L = np.longdouble
    un, vn = u.numpy(), v.numpy()
    for col, ee in enumerate(range(n - 2, 0, -1)):
        uc, vc = un[:, col].astype(L), vn[:, col].astype(L)
        ck = [L(-((-1) ** k) * (comb(ee, k) - ee * comb(ee, k - 1))) for k in range(2, ee + 2)]   # d^2 P_e(d) = sum ck d^k
        x_small, x_large = uc.copy(), (L(1) - uc)
        for _ in range(3):
            F = x_small ** ee * ((ee + 1) - ee * x_small) - vc
            dF = ee * (ee + 1) * x_small ** (ee - 1) * (1 - x_small)
            x_small = x_small - np.where(dF != 0, F / np.where(dF != 0, dF, 1), 0)
            G = sum(c * x_large ** k for c, k in zip(ck, range(2, ee + 2))) - (L(1) - vc)
            dG = sum(c * k * x_large ** (k - 1) for c, k in zip(ck, range(2, ee + 2)))
            x_large = x_large - np.where(dG != 0, G / np.where(dG != 0, dG, 1), 0)
        exact = np.where(vc < 0.65, x_small, L(1) - x_large)
        ok = (vc > 0) & (vc < 1) & (uc < 1 - 2e-16) & (uc > L(2.0) ** -179)   # away from the clamps at the two ends (2^-180: the bisection's floor)
        err_ulp = np.abs(uc - exact)[ok] / (np.abs(exact[ok]) * L(2.0) ** -53)
        assert float(err_ulp.max()) <= 4.0, (ee, float(err_ulp.max()))


def test_get_ps_vs_reference_golden(oracle, golden_dir):
    """get_PS semantics on K2 (x1/x2 from the boost, clamping, mask_problems, zero Jacobian)."""
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor
    g = np.load(os.path.join(golden_dir, "get_ps_n4.npz"))
    mom, boost = torch.from_numpy(g["mom"]).cuda(), torch.from_numpy(g["boost"]).cuda()
    for tag in ("id", "perm"):
        order = [int(i) for i in g["order_" + tag]]
        ps, jac, mask = Compute_ParticlesTensor.get_PS(mom, boost, order=order, target_mass=torch.from_numpy(g["target_mass"]),
                                                       E_CM=float(g["e_cm"]))
        assert mask.dtype == torch.bool and torch.equal(mask.cpu(), torch.from_numpy(g["mask_" + tag]))   # bit-exact
        ref = g["ps_" + tag]
        good = ~g["mask_" + tag]
        assert np.array_equal(np.isnan(ps.cpu().numpy()), np.isnan(ref))
        d = np.abs(ps.cpu().numpy() - ref)
        d[~good] = 0
        assert np.nanquantile(d, 0.99) < TOL64 and np.nanmax(d[24:]) < 1e-9   # first 24 rows: edge-of-cube points
        assert np.array_equal(np.isnan(jac.cpu().numpy()), np.isnan(g["jac_" + tag]))
    # larger random check against the oracle, default masses (H, t, t, g)
    n, N = 4, 50_000
    u = rand_u(N, n, 31)
    om, ow, ox1, ox2 = oracle.rambo_forward(u.numpy(), MASSES[n], E)
    b = np.stack((E / 2 * (ox1 + ox2), 0 * ox1, 0 * ox1, E / 2 * (ox1 - ox2)), axis=1)
    ps, jac, mask = Compute_ParticlesTensor.get_PS(torch.from_numpy(om[:, 2:]).cuda(), torch.from_numpy(b).cuda(), E_CM=E)
    ops, ojac, omask = oracle.get_ps(om[:, 2:], b, MASSES[n], E)
    assert (mask.cpu().numpy() != omask).mean() < 1e-4
    assert np.nanquantile(np.abs(ps.cpu().numpy() - ops), 0.999) < TOL64
    assert float(jac.abs().max()) == 0.0
    # fused logit + affine scaling epilogue == the reference's next two torch lines (unfolding_flow.py:170-171)
    g_ = torch.Generator().manual_seed(5)
    mean = torch.randn(10, dtype=torch.float64, generator=g_)
    scale = torch.rand(10, dtype=torch.float64, generator=g_) + 0.5
    ps2, _, _, lps = Compute_ParticlesTensor.get_PS(torch.from_numpy(om[:, 2:]).cuda(), torch.from_numpy(b).cuda(), E_CM=E,
                                                    logit_eps=5e-5, logit_mean=mean, logit_scale=scale)
    assert torch.equal(ps2, ps)
    ref = (torch.logit(ps.cpu(), eps=5e-5) - mean) / scale
    okm = ~torch.isnan(ref)
    assert torch.equal(torch.isnan(lps.cpu()), torch.isnan(ref))
    assert float((lps.cpu() - ref)[okm].abs().max()) < 1e-11


def test_dataset_products(oracle):
    """the tensors Dataset_Parton_Level.py:410-463 caches, from K2 (values vs the oracle's inverse)."""
    from memflow_b200.read_data import phasespace_products
    n, N = 4, 20_000
    u = rand_u(N, n, 41)
    om, ow, ox1, ox2 = oracle.rambo_forward(u.numpy(), MASSES[n], E)
    boost = np.stack((E / 2 * (ox1 + ox2), 0 * ox1, 0 * ox1, E / 2 * (ox1 - ox2)), axis=1)[:, None, :]
    out = phasespace_products(torch.from_numpy(om[:, 2:]).cuda(), torch.from_numpy(boost).cuda(), on_shell=True,
                              masses=MASSES[n])
    keys = {"phasespace_intermediateParticles_onShell", "phasespace_rambo_detjacobian_onShell",
            "phasespace_intermediateParticles_onShell_logit", "phasespace_intermediateParticles_onShell_logit_scaled",
            "mean_std_phasespace_intermediateParticles_onShell_logit"}
    assert set(out) == keys
    x1 = np.clip((boost[:, 0, 0] + boost[:, 0, 3]) / E, 0, 1)
    x2 = np.clip((boost[:, 0, 0] - boost[:, 0, 3]) / E, 0, 1)
    ops, oprob, _ = oracle.rambo_inverse(om[:, 2:], x1, x2, MASSES[n], E)
    ps = out["phasespace_intermediateParticles_onShell"].cpu().numpy()
    assert np.nanquantile(np.abs(ps - ops), 0.999) < TOL64
    dl = np.abs(np.log(out["phasespace_rambo_detjacobian_onShell"].cpu().numpy()) - np.log(oprob))
    assert np.nanquantile(dl, 0.99) < TOL64
    sc = out["phasespace_intermediateParticles_onShell_logit_scaled"]
    assert float(sc.nanmean(0).abs().max()) < 1e-9 and float(sc[~sc.isnan()].abs().max()) < 5
    plain = phasespace_products(torch.from_numpy(om[:, 2:]).cuda(), torch.from_numpy(boost).cuda(), on_shell=False,
                                masses=MASSES[n])
    assert set(plain) == {"phasespace_intermediateParticles", "phasespace_rambo_detjacobian"}


def test_dataset_products_vs_reference_golden(golden_dir):
    """The same products against what the UNMODIFIED reference methods computed (Dataset_Parton_Level.py:400-463, lifted
    with `ast` and run on CPU by tests/golden/make_golden.py dataset_products): every cached tensor by its on-disk name,
    default masses of that file (M_GLUON = 1e-3), off-shell dataset momenta, events below the minimal energy."""
    from memflow_b200.read_data import phasespace_products
    from memflow_b200.read_data.products import intermediateParticles_cartesian_onShell
    g = np.load(os.path.join(golden_dir, "dataset_products.npz"))
    data, boost = torch.from_numpy(g["data"]).cuda(), torch.from_numpy(g["boost"]).cuda()
    on = intermediateParticles_cartesian_onShell(data)
    ref_on = g["H_thad_tlep_ISR_cartesian_onShell"]
    assert np.max(np.abs(on.cpu().numpy() - ref_on) / np.abs(ref_on).max()) < 1e-15
    out = phasespace_products(torch.from_numpy(ref_on).cuda(), boost, on_shell=True)
    out.update(phasespace_products(data, boost, on_shell=False))
    quirk = phasespace_products(torch.from_numpy(ref_on).cuda(), boost, on_shell=True, ref_fp32_quirks=True)
    quirk.update(phasespace_products(data, boost, on_shell=False, ref_fp32_quirks=True))

    def close(name, tol, kappa=None):
        a, r = out[name].cpu().numpy(), g[name]
        assert np.array_equal(np.isnan(a), np.isnan(r)), name
        err = np.abs(a - r) / (1.0 + np.abs(r))
        if kappa is not None:
            err = err / kappa
        q = np.nanquantile(err, 0.99)
        print("%-60s max %.2e q99 %.2e" % (name, np.nanmax(err), q))
        assert np.nanmax(err) < tol and q < 1e-12, (name, np.nanmax(err), q)

    # conditioning of an event (tests/parity.py): gamma^2 of its intermediate systems and the mass-variable / threshold factors;
    # the ISR gluon of mass 1e-3 next to TeV-scale momenta makes the last decay ill-conditioned
    for sfx, mom in (("_onShell", ref_on), ("", g["data"])):
        ps_ref = g["phasespace_intermediateParticles" + sfx]
        full = np.concatenate((np.zeros((mom.shape[0], 2, 4)), mom), axis=1)
        kappa = parity.conditioning(np.clip(np.nan_to_num(ps_ref, nan=0.5), 1e-12, 1 - 1e-12), full, 4, list(g["masses"]))
        close("phasespace_intermediateParticles" + sfx, 1e-12, kappa[:, None])
        # det-jacobian: clean float64 by default; the literal reference's float32 weight chain with ref_fp32_quirks=True
        for variant, res in (("__clean", out), ("", quirk)):
            a, r = res["phasespace_rambo_detjacobian" + sfx].cpu().numpy(), g["phasespace_rambo_detjacobian" + sfx + variant]
            ok = np.isfinite(r) & (r > 0)
            assert np.array_equal(np.isfinite(a), np.isfinite(r))
            dl = np.abs(np.log(a[ok]) - np.log(r[ok])) / kappa[ok]
            print("ln detjac%s%s: max/kappa %.2e q99 %.2e" % (sfx, variant, dl.max(), np.quantile(dl, 0.99)))
            assert dl.max() < (1e-12 if variant else 1e-6)   # literal: a few float32 ulps (6e-8 each) of the in-place chain
    lk = 1.0 / np.minimum(np.clip(g["phasespace_intermediateParticles_onShell"], 5e-5, 1 - 5e-5),
                          1 - np.clip(g["phasespace_intermediateParticles_onShell"], 5e-5, 1 - 5e-5))   # d logit / d ps
    full = np.concatenate((np.zeros((ref_on.shape[0], 2, 4)), ref_on), axis=1)
    kappa = parity.conditioning(np.clip(np.nan_to_num(g["phasespace_intermediateParticles_onShell"], nan=0.5), 1e-12, 1 - 1e-12),
                                full, 4, list(g["masses"]))[:, None] * lk
    close("phasespace_intermediateParticles_onShell_logit", 1e-12, kappa)
    close("phasespace_intermediateParticles_onShell_logit_scaled", 1e-12, kappa)
    mean, scale = out["mean_std_phasespace_intermediateParticles_onShell_logit"]
    assert np.max(np.abs(mean.cpu().numpy() - g["mean_std_phasespace_intermediateParticles_onShell_logit__mean"])) < 1e-10
    assert np.max(np.abs(scale.cpu().numpy() / g["mean_std_phasespace_intermediateParticles_onShell_logit__scale"] - 1)) < 1e-10


@pytest.mark.parametrize("seed", range(6))
def test_random_masses_and_energies(oracle, seed):
    """Sweep over multiplicity, mass spectra (incl. exactly massless bodies) and collider energies."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 9))
    kinds = rng.integers(0, 3, n)
    masses = [float({0: 0.0, 1: 10.0 ** rng.uniform(-3, 0.5), 2: rng.uniform(1.0, 200.0)}[int(k)]) for k in kinds]
    if sum(masses) == 0.0:
        masses[0] = 1.0   # the x1x2 map needs sum(m) > 0 (log of (sum m / sqrt(s))^2)
    e_coll = float(10.0 ** rng.uniform(3.0, 5.0))
    from memflow_b200.phasespace.phasespace import PhaseSpace
    ps = PhaseSpace(e_coll, [21, 21], list(range(n)), final_masses=torch.tensor(masses, dtype=torch.float64), dev="cuda:0")
    N = 20_000
    u = rand_u(N, n, 500 + seed)
    mom, w, x1, x2 = ps.get_momenta_from_ps(u.cuda())
    om, ow, ox1, ox2 = oracle.rambo_forward(u.numpy(), masses, e_coll)
    kappa = parity.conditioning(u.numpy(), om, n, masses)
    err = parity.momenta_error(mom.cpu().numpy(), om, ox1, ox2, e_coll)
    fin = np.isfinite(kappa) & np.isfinite(err)
    assert fin.mean() > 0.99
    assert np.nanmax(err[fin] / kappa[fin]) < TOL64, (n, masses, e_coll, np.nanmax(err[fin] / kappa[fin]))
    assert np.nanquantile(err[fin], 0.9) < TOL64
    okw = fin & np.isfinite(ow) & (ow > 0)
    dl = np.abs(np.log(w.cpu().numpy()[okw]) - np.log(ow[okw]))
    assert np.nanmax(dl / kappa[okw]) < TOL64 and np.nanquantile(dl, 0.9) < TOL64
    back, prob = ps.generator.getPSpoint_batch(torch.from_numpy(om[:, 2:]).cuda(), torch.from_numpy(ox1).cuda(),
                                               torch.from_numpy(ox2).cuda())
    ops, oprob, _ = oracle.rambo_inverse(om[:, 2:], ox1, ox2, masses, e_coll)
    d = np.abs(back.cpu().numpy() - ops).max(axis=1)
    finb = fin & np.isfinite(d)
    assert np.array_equal(np.isnan(back.cpu().numpy()), np.isnan(ops))
    assert np.nanmax(d[finb] / kappa[finb]) < TOL64 and np.nanquantile(d[finb], 0.9) < TOL64


def test_decay_products_vs_reference_golden(oracle, golden_dir):
    """H -> bb and t -> bW -> bqq' epilogues on the GPU vs the reference's outputs."""
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    g = np.load(os.path.join(golden_dir, "decays_n3.npz"))
    H, t = torch.from_numpy(g["higgs"]).cuda(), torch.from_numpy(g["top"]).cuda()
    ha, tb, tw = (torch.from_numpy(g[k]).cuda() for k in ("higgs_angles", "top_b_angles", "top_W_angles"))
    for tag, flag in (("ptetaphi", True), ("cart", False)):
        h = CPT.higgs_decayProducts(H, ha, ptetaphi=flag).cpu().numpy()
        tt = CPT.top_decayProducts(t, tb, tw, ptetaphi=flag).cpu().numpy()
        assert h.shape == (H.shape[0], 3, 4) and tt.shape == (H.shape[0], 4, 4)
        assert np.nanmax(np.abs(h - g["higgs_out_" + tag]) / (1 + np.abs(g["higgs_out_" + tag]))) < TOL64
        assert np.nanmax(np.abs(tt - g["top_out_" + tag]) / (1 + np.abs(g["top_out_" + tag]))) < TOL64
    tt = CPT.top_decayProducts(t, tb, tw, b_mass=4.7, ptetaphi=False).cpu().numpy()
    assert np.nanmax(np.abs(tt - g["top_out_massive_b"]) / (1 + np.abs(g["top_out_massive_b"]))) < TOL64
    # rows of a K1 output read in place (row stride 4(n+2)), chained after the fused lab boost
    n, N = 3, 30_000
    ps = make_ps(n)
    mom = ps.get_momenta_from_ps(rand_u(N, n, 61).cuda(), lab_frame=True)[0]
    gsm = torch.Generator().manual_seed(3)
    ang = torch.stack((torch.randn(N, generator=gsm, dtype=torch.float64), (torch.rand(N, generator=gsm, dtype=torch.float64) * 2 - 1) * 3.14), 1)
    out = CPT.higgs_decayProducts(mom[:, 2], ang.cuda(), ptetaphi=False)
    ref = oracle.higgs_decay(mom[:, 2].cpu().numpy(), ang.numpy(), ptetaphi=False)
    assert np.abs(out.cpu().numpy() - ref).max() / E < TOL64
    assert float((out[:, 1] + out[:, 2] - out[:, 0]).abs().max()) / E < TOL64


def _decay_partons_call(g, name, device="cuda", dtype=torch.float64, requires_grad=False):
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    unscale_phi, pep, fs, kw = parity.DECAY_PARTONS_VARIANTS[name]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype).requires_grad_(requires_grad)
    ins = [t(g["prop_" + name])] + [t(g["angles"][i]) for i in range(5)] + [t(g["boost"])]
    c = lambda k: torch.from_numpy(g[k])
    out = CPT.get_decayPartons_fromlab_propagators_angles(
        ins[0], higgs_angles=ins[1], thad_b_angles=ins[2], thad_W_angles=ins[3], tlep_b_angles=ins[4], tlep_W_angles=ins[5],
        boost=ins[6], log_mean_parton_lab=c("mean_lab"), log_std_parton_lab=c("std_lab"), log_mean_boost=c("mean_boost"),
        log_std_boost=c("std_boost"), log_mean_parton_Hthadtlep=c("mean_prop"), log_std_parton_Hthadtlep=c("std_prop"),
        device=torch.device(device), ptetaphi=pep, unscale_phi=unscale_phi, final_scaling=fs, **kw)
    return ins, out


def test_decay_partons_vs_reference_golden(oracle, golden_dir):
    """the fused parton-level event (get_decayPartons_fromlab_propagators_angles, utils.py:826-935) on the
    GPU vs the unmodified reference, every option; then vs the oracle on 50k random events."""
    g = np.load(os.path.join(golden_dir, "decay_partons.npz"))
    for name, (unscale_phi, pep, fs, kw) in parity.DECAY_PARTONS_VARIANTS.items():
        ins, out = _decay_partons_call(g, name)
        ref = g["out_" + name]
        assert out.dtype == torch.float64 and tuple(out.shape) == ref.shape
        assert np.array_equal(ins[0].cpu().numpy(), g["prop_" + name])  # the argument is not unscaled in place
        o = out.cpu().numpy()
        assert np.array_equal(np.isfinite(o), np.isfinite(ref))
        assert np.nanmax(parity.decay_partons_scaled_err(o, ref, g["out_cart"], pep)) < TOL64, name
    # float32 in -> float32 out (the model dtype of the reference's training), computed in fp64 inside
    _, out32 = _decay_partons_call(g, "default", dtype=torch.float32)
    assert out32.dtype == torch.float32
    ok = np.abs(g["out_default"][:, 11, 0] - 0.001) > 1e-9  # fp32 inputs move events across the padding threshold
    assert np.nanquantile(np.abs(out32.cpu().numpy()[ok] - g["out_default"][ok]) / (1 + np.abs(g["out_default"][ok])), 0.9) < 1e-4
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    with pytest.raises(NameError):  # the reference's cartesian final scaling reads an undefined name (:913)
        CPT.get_decayPartons_fromlab_propagators_angles(*[x for x in _decay_partons_call(g, "cart")[0]],
                                                        g["mean_lab"], g["std_lab"], g["mean_boost"], g["std_boost"],
                                                        g["mean_prop"], g["std_prop"], ptetaphi=False, final_scaling=True)
    # random events at size: same kernel vs the oracle
    N = 50_000
    rng = np.random.default_rng(5)
    prop = np.stack((rng.normal(0, 1, (N, 3)), rng.normal(0, 1, (N, 3)), rng.uniform(-3.1, 3.1, (N, 3))), axis=2)
    ang = np.stack((rng.normal(0, 1.3, (5, N)), rng.uniform(-3.14, 3.14, (5, N))), axis=2)
    boost = rng.normal(0, 1, (N, 1, 2))
    gg = {k: g[k] for k in ("mean_prop", "std_prop", "mean_lab", "std_lab", "mean_boost", "std_boost")}
    gg.update(prop_default=prop, prop_cart=prop, angles=ang, boost=boost)
    cart = oracle.decay_partons(prop, *ang, boost, oracle.decay_partons_cfg(
        g["mean_prop"], g["std_prop"], g["mean_lab"], g["std_lab"], g["mean_boost"], g["std_boost"], ptetaphi=False,
        final_scaling=False))
    for name in ("default", "cart"):
        unscale_phi, pep, fs, kw = parity.DECAY_PARTONS_VARIANTS[name]
        ref = oracle.decay_partons(prop, *ang, boost, oracle.decay_partons_cfg(
            g["mean_prop"], g["std_prop"], g["mean_lab"], g["std_lab"], g["mean_boost"], g["std_boost"], ptetaphi=pep,
            unscale_phi=unscale_phi, final_scaling=fs))
        o = _decay_partons_call(gg, name)[1].cpu().numpy()
        assert np.array_equal(np.isfinite(o), np.isfinite(ref))
        err = parity.decay_partons_scaled_err(o, ref, cart, pep)
        assert np.nanquantile(err, 0.999) < TOL64 and np.nanmax(err) < 1e-9, name


def test_decay_partons_autograd_matches_reference_graph(golden_dir):
    """gradient of sum(out * g_out) w.r.t. propagators, angles and boost == the reference's autograd."""
    g = np.load(os.path.join(golden_dir, "decay_partons.npz"))
    ins, out = _decay_partons_call(g, "default", requires_grad=True)
    (out * torch.from_numpy(g["g_out"]).cuda()).sum().backward()
    got = [ins[0].grad.cpu().numpy(), np.stack([x.grad.cpu().numpy() for x in ins[1:6]]), ins[6].grad.cpu().numpy()]
    ref = [g["g_prop"], g["g_angles"], g["g_boost"]]
    for a, r, what in zip(got, ref, ("prop", "angles", "boost")):
        assert a.shape == r.shape, what
        fin = np.isfinite(r)
        assert np.isfinite(a[fin]).all(), what
        scale = np.nanmax(np.abs(r[fin])) if fin.any() else 1.0
        rel = np.abs(a - r)[fin] / (np.abs(r[fin]) + 1e-6 * scale)
        assert np.quantile(rel, 0.5) < 1e-12 and np.quantile(rel, 0.99) < 1e-8 and rel.max() < 1e-5, (what, rel.max())
    assert np.all(got[2][:, 0, 0] == 0.0)  # only pz of the boost enters the event (:857-862)


def test_sampling_loss_pipeline_matches_reference(golden_dir):
    """The sampling branch of the reference's training loss (unfolding_flow_withDecay_v2.py:376-417) end to end on the
    drop-in API: ps -> get_momenta_from_ps(requires_grad=True) [K1 + backward kernel] -> boost_tt / pt-eta-phi / scaling
    [torch glue] -> get_decayPartons_fromlab_propagators_angles [fused kernel + backward kernel]; the 12-parton event, the
    scaled boost and d(loss)/d(ps) must equal what the unmodified reference produced (golden, CPU float64)."""
    from memflow_b200.phasespace.phasespace import PhaseSpace
    from memflow_b200.phasespace import utils as ps_utils
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    g = np.load(os.path.join(golden_dir, "sampling_pipeline.npz"))
    c = lambda k: torch.from_numpy(g[k]).cuda()
    rambo = PhaseSpace(float(g["e_cm"]), [21, 21], [25, 6, -6, 21], final_masses=torch.from_numpy(g["masses"]), dev="cuda:0")
    ps = c("ps").requires_grad_(True)
    mom, _, x1, x2 = rambo.get_momenta_from_ps(ps, requires_grad=True)
    zeros = torch.zeros_like(x1)
    boost_Epz = torch.stack((rambo.collider_energy * (x1 + x2) / 2, zeros, zeros, rambo.collider_energy * (x1 - x2) / 2), dim=1)
    lab = ps_utils.boost_tt(mom[:, -4:-1], ps_utils.boostVector_t(boost_Epz).unsqueeze(dim=1))
    pep = CPT.get_ptetaphi_comp_batch(lab)
    prop = pep[..., 1:].clone()
    prop[..., 0] = torch.log(pep[..., 1] + 1)
    prop[..., 0:2] = (prop[..., 0:2] - c("mean_prop")[:2]) / c("std_prop")[:2]
    boost_scaled = boost_Epz[..., [0, 3]].clone()
    boost_scaled[..., 0] = torch.log(boost_Epz[..., 0] + 1)
    boost_scaled = (boost_scaled - c("mean_boost")) / c("std_boost")
    a = c("angles")
    out = CPT.get_decayPartons_fromlab_propagators_angles(
        prop, higgs_angles=a[0], thad_b_angles=a[1], thad_W_angles=a[2], tlep_b_angles=a[3], tlep_W_angles=a[4],
        boost=boost_scaled[:, None, :], log_mean_parton_lab=c("mean_lab"), log_std_parton_lab=c("std_lab"),
        log_mean_boost=c("mean_boost"), log_std_boost=c("std_boost"), log_mean_parton_Hthadtlep=c("mean_prop"),
        log_std_parton_Hthadtlep=c("std_prop"), device=torch.device("cuda"), ptetaphi=True, eps=1e-4, unscale_phi=False,
        final_scaling=True)
    ((out * c("g_out")).sum() + (boost_scaled * c("g_boost")).sum()).backward()
    o, ref = out.detach().cpu().numpy(), g["out"]
    assert np.array_equal(np.isfinite(o), np.isfinite(ref))
    # (pt, eta, phi) of a row inherit the event's rounding amplified by E_event / pt_row
    pt = np.exp(np.abs(ref[..., 1] * np.where(np.isin(np.arange(12), (0, 3, 7)), g["std_prop"][0], g["std_lab"][0])
                       + np.where(np.isin(np.arange(12), (0, 3, 7)), g["mean_prop"][0], g["mean_lab"][0]))) - 1
    kappa = np.maximum(float(g["e_cm"]) / np.maximum(pt, 1e-3), 1.0)[..., None]
    assert np.nanmax(np.abs(o - ref) / (1 + np.abs(ref)) / kappa) < TOL64
    assert np.nanmax(np.abs(boost_scaled.detach().cpu().numpy() - g["boost_scaled"])) < 1e-12
    gp, gr = ps.grad.cpu().numpy(), g["g_ps"]
    fin = np.isfinite(gr).all(axis=1)
    assert fin.sum() > 100 and np.isfinite(gp[fin]).all()
    assert np.all(gp[fin][:, :2] == 0.0) and np.all(gr[fin][:, :2] == 0.0)   # mass-variable columns: the reference's graph has none
    rel = np.abs(gp - gr)[fin] / (np.abs(gr[fin]).max(axis=1, keepdims=True) + 1e-300)
    assert np.median(rel) < 1e-12 and np.quantile(rel, 0.99) < 1e-8 and rel.max() < 1e-5, rel.max()


@pytest.mark.parametrize("n", [3, 4, 5])
def test_autograd_matches_reference_graph(golden_dir, oracle, n):
    """d(loss)/d(points) through get_momenta_from_ps == the reference's autograd (golden), incl. its
    semantics: zero for the mass-variable columns, x columns reaching x1/x2/weight/incoming rows only."""
    g = np.load(os.path.join(golden_dir, "rambo_grad_n%d.npz" % n))
    masses = [float(x) for x in g["masses"]]
    from memflow_b200.phasespace.phasespace import PhaseSpace
    ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(masses, dtype=torch.float64), dev="cuda:0")
    u = torch.from_numpy(g["u"]).cuda().requires_grad_(True)
    mom, w, x1, x2 = ps.get_momenta_from_ps(u, requires_grad=True)
    assert mom.requires_grad and w.requires_grad
    gm, gw, gx1, gx2 = (torch.from_numpy(g[k]).cuda() for k in ("g_mom", "g_w", "g_x1", "g_x2"))
    loss = (mom * gm).sum() + (w * gw).sum() + (x1 * gx1).sum() + (x2 * gx2).sum()
    loss.backward()
    got, ref = u.grad.cpu().numpy(), g["grad_u"]
    assert np.all(got[:, : n - 2] == 0.0) and np.all(ref[:, : n - 2] == 0.0)
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-300
    rel = np.abs(got - ref) / scale
    print("autograd n=%d: max rel err %.2e, median %.2e" % (n, rel.max(), np.median(rel)))
    assert np.median(rel) < 1e-12 and np.quantile(rel, 0.99) < 1e-9 and rel.max() < 1e-6
    # partial cotangents (momenta only / weight only) are accepted
    u2 = torch.from_numpy(g["u"]).cuda().requires_grad_(True)
    ps.get_momenta_from_ps(u2)[1].log().sum().backward()
    assert float(u2.grad[:, : 3 * n - 4].abs().max()) == 0.0 and float(u2.grad[:, 3 * n - 4:].abs().min()) > 0.0
    # independent check of the angle columns: central differences of the ORACLE forward
    uu = g["u"][:8].copy()
    h = 1e-7
    for col in (n - 2, n - 1, 3 * n - 5):
        up, um = uu.copy(), uu.copy()
        up[:, col] += h; um[:, col] -= h
        mp = oracle.rambo_forward(up, masses, E)[0]; mm = oracle.rambo_forward(um, masses, E)[0]
        fd = (((mp - mm) / (2 * h))[:, 2:] * g["g_mom"][:8, 2:]).sum(axis=(1, 2))
        # cotangents of w, x1, x2 do not touch angle columns, so this is the whole gradient of those columns
        assert np.abs(fd - got[:8, col]).max() / (np.abs(fd).max() + 1e-300) < 1e-5
    # no_grad / detached inputs stay on the plain path
    with torch.no_grad():
        m2 = ps.get_momenta_from_ps(u)[0]
    assert not m2.requires_grad and torch.equal(m2, mom.detach())


@pytest.mark.parametrize("n", [3, 4, 5])
def test_inverse_autograd_matches_reference_graph(golden_dir, n):
    """SURVEY 8f-4 for K2: d(ps)/d(momenta, x1, x2) == the reference's autograd on getPSpoint_batch (golden)."""
    g = np.load(os.path.join(golden_dir, "rambo_inv_grad_n%d.npz" % n))
    from memflow_b200.phasespace.phasespace import PhaseSpace
    ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.from_numpy(g["masses"]), dev="cuda:0")
    mom = torch.from_numpy(g["mom"]).cuda().requires_grad_(True)
    x1 = torch.from_numpy(g["x1"]).cuda().requires_grad_(True)
    x2 = torch.from_numpy(g["x2"]).cuda().requires_grad_(True)
    out, prob = ps.generator.getPSpoint_batch(mom, x1, x2, order=[int(i) for i in g["order"]])
    assert out.requires_grad and not prob.requires_grad
    (out * torch.from_numpy(g["g_ps"]).cuda()).sum().backward()
    for got, ref, name in ((mom.grad, g["g_mom"], "mom"), (x1.grad, g["g_x1"], "x1"), (x2.grad, g["g_x2"], "x2")):
        got = got.cpu().numpy().reshape(ref.shape[0], -1)
        ref = ref.reshape(ref.shape[0], -1)
        # the reference's own gradient is NaN where the last (near-massless) body has E^2 - p^2 < 0 by rounding
        # (sqrt of a negative number feeding an overwritten value: 0 * NaN); the kernel stays finite there
        assert np.isfinite(got).all()
        ok = np.isfinite(ref).all(axis=1)
        assert ok.mean() > 0.8
        got, ref = got[ok], ref[ok]
        rel = np.abs(got - ref) / (np.abs(ref).max(axis=1, keepdims=True) + 1e-300)
        print("inverse autograd n=%d %s: max rel %.2e median %.2e" % (n, name, rel.max(), np.median(rel)))
        assert np.median(rel) < 1e-11 and np.quantile(rel, 0.99) < 1e-7 and rel.max() < 1e-4


def test_get_ps_autograd_matches_reference_graph(golden_dir):
    """get_PS (called in every training step, unfolding_flow.py:163): gradient to the regressed momenta and boost."""
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor
    g = np.load(os.path.join(golden_dir, "get_ps_grad_n4.npz"))
    mom = torch.from_numpy(g["mom"]).cuda().requires_grad_(True)
    boost = torch.from_numpy(g["boost"]).cuda().requires_grad_(True)
    ps, jac, mask = Compute_ParticlesTensor.get_PS(mom, boost, order=[int(i) for i in g["order"]],
                                                   target_mass=torch.from_numpy(g["target_mass"]), E_CM=float(g["e_cm"]))
    assert ps.requires_grad and not jac.requires_grad
    (ps * torch.from_numpy(g["g_ps"]).cuda()).sum().backward()
    for got, ref, name in ((mom.grad, g["g_mom"], "mom"), (boost.grad, g["g_boost"], "boost")):
        got = got.cpu().numpy().reshape(ref.shape[0], -1)
        ref = ref.reshape(ref.shape[0], -1)
        assert np.isfinite(got).all()
        ok = np.isfinite(ref).all(axis=1) & np.isfinite(g["g_mom"].reshape(ref.shape[0], -1)).all(axis=1)
        assert ok.mean() > 0.8
        got, ref = got[ok], ref[ok]
        rel = np.abs(got - ref) / (np.abs(ref).max(axis=1, keepdims=True) + 1e-300)
        print("get_PS autograd %s: max rel %.2e median %.2e" % (name, rel.max(), np.median(rel)))
        assert np.median(rel) < 1e-11 and np.quantile(rel, 0.99) < 1e-7 and rel.max() < 1e-4
    # fused-epilogue variant stays differentiable (epilogue in torch on this path)
    mom2 = torch.from_numpy(g["mom"]).cuda().requires_grad_(True)
    out = Compute_ParticlesTensor.get_PS(mom2, boost.detach(), order=[1, 0, 2, 3], target_mass=torch.from_numpy(g["target_mass"]),
                                         E_CM=float(g["e_cm"]), logit_eps=5e-5)
    assert len(out) == 4 and out[3].requires_grad


def test_sampling_branch_fused_one_launch_each_way(golden_dir):
    """The same branch (unfolding_flow_withDecay_v2.py:376-417) through the fused entry: ONE forward launch
    (mf_sampling_partons_f64) and ONE backward launch (mf_sampling_partons_bwd_f64), no torch glue in between. Same
    golden outputs and gradients of the unmodified reference as the test above, and agreement with the unfused path."""
    from memflow_b200 import _lib
    from memflow_b200.phasespace.phasespace import PhaseSpace
    from memflow_b200.unfolding_flow import utils as UU
    g = np.load(os.path.join(golden_dir, "sampling_pipeline.npz"))
    c = lambda k: torch.from_numpy(g[k]).cuda()
    rambo = PhaseSpace(float(g["e_cm"]), [21, 21], [25, 6, -6, 21], final_masses=torch.from_numpy(g["masses"]), dev="cuda:0")
    ps = c("ps").requires_grad_(True)
    a = [c("angles")[i].clone().requires_grad_(True) for i in range(5)]
    L = _lib.lib()
    calls = []
    names = [n for n in _lib.SYMBOLS if n not in ("mf_error_string", "mf_abi_version", "mf_device_info", "mf_is_reduce_workspace_bytes")]
    orig = {n: getattr(L, n) for n in names}
    try:
        for n in names:
            setattr(L, n, (lambda f, n=n: (lambda *args: (calls.append(n), f(*args))[1]))(orig[n]))
        out, bs, mom, w, x1, x2 = UU.sampled_decayPartons_from_ps(
            rambo, ps, *a, log_mean_parton_lab=c("mean_lab"), log_std_parton_lab=c("std_lab"), log_mean_boost=c("mean_boost"),
            log_std_boost=c("std_boost"), log_mean_parton_Hthadtlep=c("mean_prop"), log_std_parton_Hthadtlep=c("std_prop"),
            ptetaphi=True, eps=1e-4, unscale_phi=False, final_scaling=True)
        assert calls == ["mf_sampling_partons_f64"], calls
        ((out * c("g_out")).sum() + (bs * c("g_boost")).sum()).backward()
        assert calls == ["mf_sampling_partons_f64", "mf_sampling_partons_bwd_f64"], calls
    finally:
        for n in names:
            setattr(L, n, orig[n])
    o, ref = out.detach().cpu().numpy(), g["out"]
    assert np.array_equal(np.isfinite(o), np.isfinite(ref))
    pt = np.exp(np.abs(ref[..., 1] * np.where(np.isin(np.arange(12), (0, 3, 7)), g["std_prop"][0], g["std_lab"][0])
                       + np.where(np.isin(np.arange(12), (0, 3, 7)), g["mean_prop"][0], g["mean_lab"][0]))) - 1
    kappa = np.maximum(float(g["e_cm"]) / np.maximum(pt, 1e-3), 1.0)[..., None]
    assert np.nanmax(np.abs(o - ref) / (1 + np.abs(ref)) / kappa) < TOL64
    assert np.nanmax(np.abs(bs.detach().cpu().numpy() - g["boost_scaled"])) < 1e-12
    gp, gr = ps.grad.cpu().numpy(), g["g_ps"]
    fin = np.isfinite(gr).all(axis=1)
    assert fin.sum() > 100 and np.isfinite(gp[fin]).all()
    assert np.all(gp[fin][:, :2] == 0.0)
    rel = np.abs(gp - gr)[fin] / (np.abs(gr[fin]).max(axis=1, keepdims=True) + 1e-300)
    print("fused sampling branch: d loss / d ps vs reference: median %.1e q99 %.1e max %.1e" % (np.median(rel), np.quantile(rel, 0.99), rel.max()))
    assert np.median(rel) < 1e-12 and np.quantile(rel, 0.99) < 1e-8 and rel.max() < 1e-5, rel.max()
    # the K1 outputs that come back with it are those of the plain call
    m2, w2, x12, x22 = rambo.get_momenta_from_ps(ps.detach())
    assert torch.equal(x1, x12) and torch.equal(x2, x22)
    ok = torch.isfinite(m2).all(dim=(1, 2)) & torch.isfinite(mom).all(dim=(1, 2))
    assert float((mom[ok] - m2[ok]).abs().max()) <= 1e-9 * float(g["e_cm"])
    # angle gradients against the unfused path (decay kernel backward on the torch glue)
    from memflow_b200.phasespace import utils as ps_utils
    CPT = UU.Compute_ParticlesTensor
    ps2 = c("ps").requires_grad_(True)
    a2 = [c("angles")[i].clone().requires_grad_(True) for i in range(5)]
    mom_, _, x1_, x2_ = rambo.get_momenta_from_ps(ps2, requires_grad=True)
    zeros = torch.zeros_like(x1_)
    bE = torch.stack((rambo.collider_energy * (x1_ + x2_) / 2, zeros, zeros, rambo.collider_energy * (x1_ - x2_) / 2), dim=1)
    lab = ps_utils.boost_tt(mom_[:, -4:-1], ps_utils.boostVector_t(bE).unsqueeze(dim=1))
    pep = CPT.get_ptetaphi_comp_batch(lab)
    prop = pep[..., 1:].clone()
    prop[..., 0] = torch.log(pep[..., 1] + 1)
    prop[..., 0:2] = (prop[..., 0:2] - c("mean_prop")[:2]) / c("std_prop")[:2]
    bsc = bE[..., [0, 3]].clone()
    bsc[..., 0] = torch.log(bE[..., 0] + 1)
    bsc = (bsc - c("mean_boost")) / c("std_boost")
    out2 = CPT.get_decayPartons_fromlab_propagators_angles(
        prop, higgs_angles=a2[0], thad_b_angles=a2[1], thad_W_angles=a2[2], tlep_b_angles=a2[3], tlep_W_angles=a2[4],
        boost=bsc[:, None, :], log_mean_parton_lab=c("mean_lab"), log_std_parton_lab=c("std_lab"),
        log_mean_boost=c("mean_boost"), log_std_boost=c("std_boost"), log_mean_parton_Hthadtlep=c("mean_prop"),
        log_std_parton_Hthadtlep=c("std_prop"), device=torch.device("cuda"), ptetaphi=True, eps=1e-4, unscale_phi=False,
        final_scaling=True)
    ((out2 * c("g_out")).sum() + (bsc * c("g_boost")).sum()).backward()
    for x, y in zip(a, a2):
        gx, gy = x.grad.cpu().numpy(), y.grad.cpu().numpy()
        okr = np.isfinite(gy).all(axis=1) & np.isfinite(gx).all(axis=1)
        assert okr.sum() > 100
        assert np.max(np.abs(gx - gy)[okr] / (np.abs(gy[okr]).max() + 1e-300)) < 1e-9
