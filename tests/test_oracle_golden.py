"""CPU: pin the C oracle against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known answers."""
import json
import os

import numpy as np
import pytest

NS = (3, 4, 5, 8)
EDGE = 24  # first rows of every fixture are edge-of-cube points (make_golden.make_u)


def load(golden_dir, n):
    return np.load(os.path.join(golden_dir, "rambo_n%d.npz" % n))


def test_known_answer_flat_volume(golden_dir):
    """notebooks/RamboTest.ipynb [18]: get_flatWeights(13000, 3) = 21291.052028166854."""
    from memflow_b200.phasespace.rambo_generator import FlatInvertiblePhasespace
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    assert ka["get_flatWeights(13000,3)"] == 21291.052028166854
    assert FlatInvertiblePhasespace.get_flatWeights(13000, 3) == 21291.052028166854


@pytest.mark.parametrize("n", NS)
@pytest.mark.parametrize("tag,quirks", [("clean", False), ("lit", True)])
def test_forward_matches_reference(oracle, golden_dir, n, tag, quirks):
    g = load(golden_dir, n)
    mom, w, x1, x2 = oracle.rambo_forward(g["u"], g["masses"], float(g["e_cm"]), fp32_quirks=quirks)
    ecm = np.sqrt(g["x1_" + tag] * g["x2_" + tag]) * float(g["e_cm"])
    dm = np.abs(mom - g["mom_" + tag]).max(axis=(1, 2)) / ecm
    assert dm[EDGE:].max() < 1e-12, dm.max()          # bulk: libm vs Sleef last-ulp differences only
    assert dm.max() < 1e-11
    np.testing.assert_allclose(x1, g["x1_" + tag], rtol=0, atol=5e-15)
    np.testing.assert_allclose(x2, g["x2_" + tag], rtol=0, atol=5e-15)
    ref_w = g["w_" + tag]
    assert np.array_equal(np.isinf(w), np.isinf(ref_w))        # float32 overflow pattern of the literal n=8 case
    fin = np.isfinite(ref_w) & (ref_w > 0)
    dl = np.abs(np.log(w[fin]) - np.log(ref_w[fin]))
    tol = 5e-7 if quirks else 1e-11                            # literal: float32 ulps (powf differs by 1 ulp)
    assert np.nanmax(dl[EDGE:]) < tol
    assert np.nanmax(dl) < max(tol, 1e-9)                      # edge rows: conditioning of the mass variables


@pytest.mark.parametrize("n", NS)
@pytest.mark.parametrize("tag,quirks", [("clean", False), ("lit", True)])
def test_inverse_matches_reference(oracle, golden_dir, n, tag, quirks):
    g = load(golden_dir, n)
    ps, prob, rc = oracle.rambo_inverse(g["mom_" + tag][:, 2:], g["x1_" + tag], g["x2_" + tag], g["masses"],
                                        float(g["e_cm"]), fp32_quirks=quirks)
    assert rc == 0
    ref = g["inv_ps_" + tag]
    assert np.array_equal(np.isnan(ps), np.isnan(ref))
    assert np.nanmax(np.abs(ps - ref)) < 1e-12
    refp = g["inv_prob_" + tag]
    fin = np.isfinite(refp) & (refp > 0)
    assert np.array_equal(prob == 0, refp == 0)               # 1/inf of the literal float32 volume
    dl = np.abs(np.log(prob[fin]) - np.log(refp[fin]))
    assert np.nanmax(dl[EDGE:][np.isfinite(dl[EDGE:])]) < (5e-7 if quirks else 1e-11)


@pytest.mark.parametrize("n", NS)
def test_inverse_permuted_and_offshell(oracle, golden_dir, n):
    g = load(golden_dir, n)
    E = float(g["e_cm"])
    a = (g["mom_clean"][:, 2:], g["x1_clean"], g["x2_clean"], g["masses"], E)
    ps, prob, _ = oracle.rambo_inverse(*a, order=g["perm"])
    assert np.nanmax(np.abs(ps - g["inv_ps_perm_clean"])[EDGE:]) < 1e-12
    ps, prob, _ = oracle.rambo_inverse(*a, order=g["perm"], target_mass=g["target_mass"])
    ref = g["inv_ps_offshell_clean"]
    assert np.array_equal(np.isnan(ps), np.isnan(ref))
    assert np.nanmax(np.abs(ps - ref)[EDGE:]) < 1e-12


@pytest.mark.parametrize("n", NS)
def test_cuts_match_reference(oracle, golden_dir, n):
    g = load(golden_dir, n)
    _, w, _, _ = oracle.rambo_forward(g["u"], g["masses"], float(g["e_cm"]), fp32_quirks=True, pT_mincut=20.0,
                                      delR_mincut=0.4, rap_maxcut=5.0)
    assert np.array_equal(w == 0, g["w_cuts_lit"] == 0)
    assert 0 < (w == 0).mean() < 1


def test_round_trip_property(oracle):
    """notebooks/RamboTest.ipynb [41-46]: u -> p -> u is the identity (n=3 residual ~4e-16 there)."""
    rng = np.random.default_rng(5)
    masses = [125.25, 172.5, 172.5]
    u = rng.random((2000, 7)).clip(1e-10, 1 - 1e-10)
    mom, w, x1, x2 = oracle.rambo_forward(u, masses, 13000.0)
    ps, prob, rc = oracle.rambo_inverse(mom[:, 2:], x1, x2, masses, 13000.0)
    assert rc == 0
    assert np.abs(ps - u).max() < 1e-9
    assert np.median(np.abs(ps - u)) < 1e-14
    np.testing.assert_allclose(prob * w, 1.0, rtol=1e-9)
    # momentum conservation and mass shells
    tot = mom[:, 2:].sum(axis=1)
    ecm = np.sqrt(x1 * x2) * 13000.0
    assert np.abs(tot[:, 0] - ecm).max() / 13000.0 < 1e-13 and np.abs(tot[:, 1:]).max() < 1e-9


def test_not_cm_is_reported(oracle):
    mom = np.zeros((3, 3, 4)); mom[:, :, 0] = 200.0; mom[:, 0, 1] = 5.0
    _, _, rc = oracle.rambo_inverse(mom, np.full(3, 0.1), np.full(3, 0.1), [125.25, 172.5, 172.5], 13000.0)
    assert rc == 2


def test_nan_input_raises(oracle):
    u = np.full((2, 7), 0.5); u[1, 3] = np.nan
    with pytest.raises(ValueError):
        oracle.rambo_forward(u, [125.25, 172.5, 172.5], 13000.0)


@pytest.mark.parametrize("tag", ["f64_F7K10", "f32_F20K64", "f64_F10K40"])
def test_spline_oracle_matches_torch_restatement(oracle, golden_dir, tag):
    """The spline is 'parity unpinned' (zuko absent): the C oracle is checked against the op-for-op torch
    restatement (oracle/rqs_torch.py), bins exactly and values to dtype precision."""
    g = np.load(os.path.join(golden_dir, "rqs_%s.npz" % tag))
    bound = float(g["bound"])
    y, ladj, bins = oracle.rqs_forward(g["phi"], g["x"], bound=bound)
    f32 = g["phi"].dtype == np.float32
    same = bins == g["bin_fwd"]
    assert same.mean() > (0.999 if f32 else 1.0 - 1e-12)
    tol_y, tol_l = (2e-5, 5e-3) if f32 else (1e-13, 1e-11)
    assert np.abs(y - g["y"])[same].max() < tol_y
    assert np.abs(ladj - g["ladj"])[same].max() < tol_l
    x, _, bins_i = oracle.rqs_inverse(g["phi"], g["x"], bound=bound)
    same = bins_i == g["bin_inv"]
    assert same.mean() > (0.999 if f32 else 1.0 - 1e-12)
    assert np.abs(x - g["x_inv"])[same].max() < (1e-4 if f32 else 1e-12)
    # identity tails: |x| > bound untouched, ladj exactly 0
    out = np.abs(g["x"]) > bound * 1.0001
    assert out.any() and np.array_equal(y[out], g["x"][out]) and np.all(ladj[out] == 0)


def test_spline_self_consistency(oracle):
    """inverse(forward(x)) = x and ladj = log of the numerical derivative (fp64)."""
    rng = np.random.default_rng(3)
    phi = rng.standard_normal((500, 5, 3 * 12 - 1))
    x = rng.uniform(-0.95, 0.95, (500, 5))
    y, ladj, _ = oracle.rqs_forward(phi, x, bound=1.0)
    xb, ladj_b, _ = oracle.rqs_inverse(phi, y, bound=1.0)
    assert np.abs(xb - x).max() < 1e-12
    assert np.abs(ladj_b - ladj).max() < 1e-10
    h = 1e-6
    yp, _, bp = oracle.rqs_forward(phi, x + h, bound=1.0)
    ym, _, bm = oracle.rqs_forward(phi, x - h, bound=1.0)
    ok = bp == bm
    assert np.abs(np.log((yp - ym) / (2 * h)) - ladj)[ok].max() < 1e-5


def test_estimator_oracle():
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    t = rng.random(1000); t[::10] = -1
    j = rng.random(1000) * 5; j[3] = np.nan
    lq = rng.standard_normal(1000)
    acc = O.is_reduce(t, j, lq)
    ok = (t > 0) & ~np.isnan(j)
    w = (t * j / np.exp(lq))[ok]
    np.testing.assert_allclose(acc, [w.sum(), (w * w).sum(), ok.sum()], rtol=1e-13)
    I, s = O.estimator(acc, 1000)
    assert s > 0 and np.isclose(I, w.sum() / 1000)


def test_philox_known_answer():
    """Random123 known-answer test for Philox4x32-10: counter = key = 0 and all-ones."""
    from oracle import philox as P
    r = P.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = P.philox4x32_10([0xffffffff], [0xffffffff], [0xffffffff], [0xffffffff], 0xffffffff, 0xffffffff)
    assert [int(v[0]) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_base_variate_is_strictly_inside_the_unit_interval():
    """(k + 0.5) / 2^23 from the top 23 bits: exact in float32 for every bit pattern, never 0 or 1 (a 24-bit k would
    round to 1.0 for bits = 0xFFFFFFFF and put the base point exactly on the spline's bound)."""
    from oracle import philox as P
    bits = np.array([0, 1, 0x1FF, 0x200, 0x7FFFFFFF, 0x80000000, 0xFFFFFDFF, 0xFFFFFE00, 0xFFFFFFFF], dtype=np.uint32)
    u = P.u01_from_bits(bits)
    assert u.dtype == np.float32 and (u > 0).all() and (u < 1).all()
    exact = ((bits >> 9).astype(np.float64) + 0.5) / 2.0 ** 23
    assert np.array_equal(u.astype(np.float64), exact)
    assert float(u[-1]) == 1.0 - 2.0 ** -24 and float(u[0]) == 2.0 ** -24
    z = np.float32(1.0) * (np.float32(2.0) * u - np.float32(1.0))
    assert (np.abs(z) < 1).all()


def test_get_ps_matches_reference(oracle, golden_dir):
    """Compute_ParticlesTensor.get_PS (memflow/unfolding_flow/utils.py:1056-1152), incl. the problem mask."""
    g = np.load(os.path.join(golden_dir, "get_ps_n4.npz"))
    for tag in ("id", "perm"):
        ps, jac, mask = oracle.get_ps(g["mom"], g["boost"], g["target_mass"], float(g["e_cm"]), order=g["order_" + tag])
        assert np.array_equal(mask, g["mask_" + tag]) and 10 < mask.sum() < 40
        assert np.array_equal(np.isnan(ps), np.isnan(g["ps_" + tag]))
        assert np.nanmax(np.abs(ps - g["ps_" + tag])) < 1e-13
        assert np.array_equal(np.isnan(jac), np.isnan(g["jac_" + tag])) and np.nanmax(np.abs(jac)) == 0.0


def test_decay_products_match_reference(oracle, golden_dir):
    """higgs_decayProducts / top_decayProducts (memflow/unfolding_flow/utils.py:745-815), both output conventions."""
    g = np.load(os.path.join(golden_dir, "decays_n3.npz"))
    for tag, flag in (("ptetaphi", True), ("cart", False)):
        h = oracle.higgs_decay(g["higgs"], g["higgs_angles"], ptetaphi=flag)
        t = oracle.top_decay(g["top"], g["top_b_angles"], g["top_W_angles"], ptetaphi=flag)
        assert np.nanmax(np.abs(h - g["higgs_out_" + tag]) / (1 + np.abs(g["higgs_out_" + tag]))) < 1e-13
        assert np.nanmax(np.abs(t - g["top_out_" + tag]) / (1 + np.abs(g["top_out_" + tag]))) < 1e-12
    t = oracle.top_decay(g["top"], g["top_b_angles"], g["top_W_angles"], b_mass=4.7, ptetaphi=False)
    assert np.nanmax(np.abs(t - g["top_out_massive_b"]) / (1 + np.abs(g["top_out_massive_b"]))) < 1e-12
    # physics sanity: daughters sum to the parent, b quarks massless, W on shell
    c = g["top_out_cart"]
    assert np.abs(c[:, 1:].sum(1) - c[:, 0]).max() / 13000.0 < 1e-12
    w = c[:, 2] + c[:, 3]
    assert np.abs(np.sqrt(w[:, 0] ** 2 - (w[:, 1:] ** 2).sum(1)) - 80.4).max() < 1e-8


def test_decay_partons_match_reference(oracle, golden_dir):
    """get_decayPartons_fromlab_propagators_angles (memflow/unfolding_flow/utils.py:826-935): every option the
    reference accepts, incl. distinct thad/tlep masses (pins the reference's propagator-1 = tlep pairing) and the
    ISR padding row."""
    from tests.parity import DECAY_PARTONS_VARIANTS, decay_partons_scaled_err
    g = np.load(os.path.join(golden_dir, "decay_partons.npz"))
    a = g["angles"]
    for name, (unscale_phi, pep, fs, kw) in DECAY_PARTONS_VARIANTS.items():
        cfg = oracle.decay_partons_cfg(g["mean_prop"], g["std_prop"], g["mean_lab"], g["std_lab"], g["mean_boost"],
                                       g["std_boost"], ptetaphi=pep, unscale_phi=unscale_phi, final_scaling=fs, **kw)
        out = oracle.decay_partons(g["prop_" + name], a[0], a[1], a[2], a[3], a[4], g["boost"], cfg)
        ref = g["out_" + name]
        assert out.shape == ref.shape == (a.shape[1], 12, 4)
        assert np.array_equal(np.isfinite(out), np.isfinite(ref))
        assert np.nanmax(decay_partons_scaled_err(out, ref, g["out_cart"], pep)) < 1e-12, name
    padded = np.abs(g["out_noscale"][:, 11, 0] - 0.001) < 1e-15
    assert padded.sum() >= 32 and (~padded).sum() >= 100  # both branches of :897-905 are in the fixture


def test_spline_oracles_vs_high_precision_known_answers(oracle, golden_dir):
    """Best-effort pin of the spline restatement (zuko is absent: "parity unpinned"): the C oracle and the torch restatement
    against 50-digit mpmath evaluations of the closed form (Durkan et al. 2019 eqs. 4-8 + zuko's parameterisation),
    tests/golden/make_rqs_known_answers.py. Bins exact; values to the conditioning of the case."""
    import json
    import torch
    from oracle import rqs_torch
    ka = json.load(open(os.path.join(golden_dir, "rqs_known_answers.json")))
    assert len(ka["cases"]) == 4
    for c in ka["cases"]:
        K = len(c["widths"])
        phi = np.array(c["widths"] + c["heights"] + c["derivatives"], dtype=np.float64)
        P = len(c["points"])
        x = np.array([p["x"] for p in c["points"]])
        phib = np.broadcast_to(phi, (P, 1, 3 * K - 1)).copy()
        ry, rl, rx, rli = (np.array([float(p[k]) for p in c["points"]]) for k in ("y", "ladj", "x_inv", "ladj_at_x_inv"))
        rb, rbi = (np.array([p[k] for p in c["points"]]) for k in ("bin", "bin_inv"))
        tol = 1e-14 if c["name"] != "K8_bound5_large_parameters" else 2e-12     # bins 1e-4 wide: 1 / width conditioning
        t = rqs_torch.MonotonicRQSTransform(torch.from_numpy(phib[..., :K]), torch.from_numpy(phib[..., K:2 * K]),
                                            torch.from_numpy(phib[..., 2 * K:]), bound=c["bound"], slope=c["slope"])
        y, l = t.call_and_ladj(torch.from_numpy(x[:, None]))
        xi = t._inverse(torch.from_numpy(x[:, None]))
        assert np.abs(y.numpy()[:, 0] - ry).max() < tol and np.abs(l.numpy()[:, 0] - rl).max() < tol
        assert np.abs(xi.numpy()[:, 0] - rx).max() < tol
        oy, ol, ob = oracle.rqs_forward(phib, x[:, None], bound=c["bound"], slope=c["slope"])
        ox, oli, obi = oracle.rqs_inverse(phib, x[:, None], bound=c["bound"], slope=c["slope"])
        assert np.array_equal(ob[:, 0], rb) and np.array_equal(obi[:, 0], rbi)
        assert np.abs(oy[:, 0] - ry).max() < tol and np.abs(ol[:, 0] - rl).max() < tol
        assert np.abs(ox[:, 0] - rx).max() < tol and np.abs(oli[:, 0] - rli).max() < tol
