"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/rambo_n{3,4,5,8}.npz, tests/golden/rqs_*.npz, tests/golden/known_answers.json.

Phase space: imports memflow.phasespace from /root/reference as-is (only the unused
`particle` import is stubbed). Two variants are stored:
  *_lit   : the literal reference (float32 volume / incoming rows, rambo_generator.py:132,:533)
  *_clean : same files with `.to(torch.float)` / `.float()` neutralised by patching
            torch.Tensor.float / torch.Tensor.to for the duration of the call ("fp64-clean").
Spline: zuko is not available; the fixture comes from oracle/rqs_torch.py (op-for-op torch
restatement of upstream zuko) -> "parity unpinned" for the spline.
"""
import contextlib
import io
import json
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MEMFLOW_REFERENCE", "/root/reference")


def import_reference():
    torch.set_default_dtype(torch.float64)  # BEFORE importing memflow.phasespace
    stub = types.ModuleType("particle")
    stub.Particle = type("Particle", (), {})
    sys.modules.setdefault("particle", stub)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from memflow.phasespace import phasespace, rambo_generator, utils  # noqa
    return phasespace, rambo_generator, utils


@contextlib.contextmanager
def fp64_clean():
    """Neutralise the two float32 casts of the reference without touching its files."""
    orig_float, orig_to = torch.Tensor.float, torch.Tensor.to

    def to(self, *a, **k):
        if len(a) == 1 and a[0] in (torch.float, torch.float32) and not k:
            return self
        return orig_to(self, *a, **k)

    torch.Tensor.float = lambda self: self
    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.float, torch.Tensor.to = orig_float, orig_to


CASES = {
    3: [125.25, 172.5, 172.5],
    4: [125.25, 172.5, 172.5, 1e-5],
    5: [125.25, 172.5, 172.5, 1e-5, 1e-5],
    8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3],
}
E_CM = 13000.0
N = 384


def make_u(n, seed):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, generator=g)
    # a band of edge-of-cube points: tiny / near-one mass variables and angles
    u[0:8, : n - 2] = 1e-9
    u[8:16, : n - 2] = 1 - 1e-9
    u[16:24, n - 2:3 * n - 4] = torch.tensor([1e-10, 0.5, 1 - 1e-10, 0.25, 0.75, 1e-7, 0.5 + 1e-9, 0.5 - 1e-9]
                                             ).unsqueeze(1)
    return u.clamp(1e-10, 1 - 1e-10)


def main():
    phasespace, rambo_generator, utils = import_reference()
    known = {"get_flatWeights(13000,3)": rambo_generator.FlatInvertiblePhasespace.get_flatWeights(13000, 3)}
    for n, masses in CASES.items():
        m = torch.tensor(masses, dtype=torch.float64)
        ps = phasespace.PhaseSpace(E_CM, [21, 21], list(range(n)), final_masses=m, dev=torch.device("cpu"))
        u = make_u(n, 1000 + n)
        out = {"u": u.numpy(), "masses": m.numpy(), "e_cm": np.float64(E_CM)}
        for tag, ctx in (("lit", contextlib.nullcontext()), ("clean", fp64_clean())):
            with ctx:
                mom, w, x1, x2 = ps.get_momenta_from_ps(u)
                with contextlib.redirect_stdout(io.StringIO()):
                    psb, prob = ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2, order=list(range(n)))
            out["mom_" + tag] = mom.numpy()
            out["w_" + tag] = w.numpy()
            out["x1_" + tag] = x1.numpy()
            out["x2_" + tag] = x2.numpy()
            out["inv_ps_" + tag] = psb.double().numpy()
            out["inv_prob_" + tag] = prob.double().numpy()
        # cuts (literal only): pT > 20, dR > 0.4, |eta| < 5
        mom, w, x1, x2 = ps.get_momenta_from_ps(u, pT_mincut=20.0, delR_mincut=0.4, rap_maxcut=5.0)
        out["w_cuts_lit"] = w.numpy()
        # permuted inverse with off-shell per-event masses (ensure_onShell=False)
        perm = list(reversed(range(n)))
        with fp64_clean(), contextlib.redirect_stdout(io.StringIO()):
            mom_c = torch.from_numpy(out["mom_clean"])
            x1c, x2c = torch.from_numpy(out["x1_clean"]), torch.from_numpy(out["x2_clean"])
            psb, prob = ps.generator.getPSpoint_batch(mom_c[:, 2:], x1c, x2c, order=perm)
            out["inv_ps_perm_clean"] = psb.numpy()
            out["inv_prob_perm_clean"] = prob.numpy()
            out["perm"] = np.array(perm, dtype=np.int32)
            tm = (m[perm] * 1.001).unsqueeze(0)  # [1, n]: the only shape the reference handles (it sums over ALL dims, :660)
            psb, prob = ps.generator.getPSpoint_batch(mom_c[:, 2:], x1c, x2c, order=perm, target_mass=tm,
                                                      ensure_onShell=False)
            out["inv_ps_offshell_clean"] = psb.numpy()
            out["inv_prob_offshell_clean"] = prob.numpy()
            out["target_mass"] = tm.numpy()
        np.savez_compressed(os.path.join(HERE, "rambo_n%d.npz" % n), **out)
        print("n=%d written; max|w_lit/w_clean-1| = %.3g" % (
            n, np.nanmax(np.abs(out["w_lit"] / out["w_clean"] - 1))))

    # ---- get_PS (memflow/unfolding_flow/utils.py:1056-1152), n = 4 only in the reference ----
    for name in ("vector", "awkward"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from memflow.unfolding_flow.utils import Compute_ParticlesTensor
    g4 = np.load(os.path.join(HERE, "rambo_n4.npz"))
    mom = torch.from_numpy(g4["mom_clean"][:, 2:]).clone()
    x1, x2 = torch.from_numpy(g4["x1_clean"]), torch.from_numpy(g4["x2_clean"])
    boost = torch.stack((E_CM / 2 * (x1 + x2), torch.zeros_like(x1), torch.zeros_like(x1), E_CM / 2 * (x1 - x2)), dim=1)
    mom[30:40, 0, 1] += torch.linspace(1e-4, 5.0, 10, dtype=torch.float64)   # not in the CM
    boost[40:44, 0] *= 3.0                                                    # x1 > 1
    boost[44:48, 3] = boost[44:48, 0] * 1.5                                   # x2 < 0
    mom[50:54, 3] *= 1.3                                                      # off-shell -> r outside [0,1] for some
    tmass = torch.tensor([[125.25, 172.5, 172.5, 1e-5]], dtype=torch.float64)
    out = {}
    for tag, order in (("id", [0, 1, 2, 3]), ("perm", [1, 0, 2, 3])):
        with contextlib.redirect_stdout(io.StringIO()):
            ps_, jac_, mask_ = Compute_ParticlesTensor.get_PS(mom, boost, order=order, target_mass=tmass, E_CM=E_CM)
        out["ps_" + tag], out["jac_" + tag], out["mask_" + tag] = ps_.numpy(), jac_.numpy(), mask_.numpy()
        out["order_" + tag] = np.array(order, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "get_ps_n4.npz"), mom=mom.numpy(), boost=boost.numpy(), target_mass=tmass.numpy(),
                        e_cm=np.float64(E_CM), **out)
    print("get_PS written; masked:", int(out["mask_id"].sum()))

    # ---- gradients of the forward map from the reference's own autograd graph (requires_grad=True) ----
    for n in (3, 4, 5):
        m = torch.tensor(CASES[n], dtype=torch.float64)
        ps = phasespace.PhaseSpace(E_CM, [21, 21], list(range(n)), final_masses=m, dev=torch.device("cpu"))
        u = make_u(n, 1000 + n)[24:24 + 96].clone().requires_grad_(True)      # bulk rows only
        gen = torch.Generator().manual_seed(77 + n)
        with fp64_clean():
            mom, w, x1, x2 = ps.get_momenta_from_ps(u, requires_grad=True)
        gm = torch.randn(mom.shape, generator=gen, dtype=torch.float64)
        gw = torch.randn(w.shape, generator=gen, dtype=torch.float64) / w.detach()
        gx1 = torch.randn(w.shape, generator=gen, dtype=torch.float64)
        gx2 = torch.randn(w.shape, generator=gen, dtype=torch.float64)
        loss = (mom * gm).sum() + (w * gw).sum() + (x1 * gx1).sum() + (x2 * gx2).sum()
        loss.backward()
        np.savez_compressed(os.path.join(HERE, "rambo_grad_n%d.npz" % n), u=u.detach().numpy(), masses=m.numpy(),
                            e_cm=np.float64(E_CM), g_mom=gm.numpy(), g_w=gw.numpy(), g_x1=gx1.numpy(), g_x2=gx2.numpy(),
                            grad_u=u.grad.numpy())
        print("grad n=%d written; |grad| per column:" % n, np.array2string(u.grad.abs().mean(0).numpy(), precision=2))

    # ---- gradients of the inverse map / get_PS from the reference's autograd graph ----
    for n in (3, 4, 5):
        m = torch.tensor(CASES[n], dtype=torch.float64)
        ps = phasespace.PhaseSpace(E_CM, [21, 21], list(range(n)), final_masses=m, dev=torch.device("cpu"))
        gsrc = np.load(os.path.join(HERE, "rambo_n%d.npz" % n))
        mom = torch.from_numpy(gsrc["mom_clean"][24:24 + 96, 2:]).clone().requires_grad_(True)
        x1 = torch.from_numpy(gsrc["x1_clean"][24:24 + 96]).clone().requires_grad_(True)
        x2 = torch.from_numpy(gsrc["x2_clean"][24:24 + 96]).clone().requires_grad_(True)
        gen = torch.Generator().manual_seed(177 + n)
        perm = list(range(n)) if n != 4 else [1, 0, 2, 3]
        with fp64_clean(), contextlib.redirect_stdout(io.StringIO()):
            psb, prob = ps.generator.getPSpoint_batch(mom, x1, x2, order=perm)
        gps = torch.randn(psb.shape, generator=gen, dtype=torch.float64)
        (psb * gps).sum().backward()
        np.savez_compressed(os.path.join(HERE, "rambo_inv_grad_n%d.npz" % n), mom=mom.detach().numpy(), x1=x1.detach().numpy(),
                            x2=x2.detach().numpy(), masses=m.numpy(), e_cm=np.float64(E_CM), order=np.array(perm, dtype=np.int32),
                            g_ps=gps.numpy(), g_mom=mom.grad.numpy(), g_x1=x1.grad.numpy(), g_x2=x2.grad.numpy())
        print("inverse grad n=%d written" % n)
    g4 = np.load(os.path.join(HERE, "get_ps_n4.npz"))
    mom = torch.from_numpy(g4["mom"][60:160]).clone().requires_grad_(True)
    boost = torch.from_numpy(g4["boost"][60:160]).clone().requires_grad_(True)
    tmass = torch.from_numpy(g4["target_mass"])
    gen = torch.Generator().manual_seed(31)
    ps_, jac_, mask_ = Compute_ParticlesTensor.get_PS(mom, boost, order=[1, 0, 2, 3], target_mass=tmass, E_CM=E_CM)
    gps = torch.randn(ps_.shape, generator=gen, dtype=torch.float64)
    (ps_ * gps).sum().backward()
    np.savez_compressed(os.path.join(HERE, "get_ps_grad_n4.npz"), mom=mom.detach().numpy(), boost=boost.detach().numpy(),
                        target_mass=tmass.numpy(), e_cm=np.float64(E_CM), order=np.array([1, 0, 2, 3], dtype=np.int32),
                        g_ps=gps.numpy(), g_mom=mom.grad.numpy(), g_boost=boost.grad.numpy())
    print("get_PS grad written")

    # ---- decay products (memflow/unfolding_flow/utils.py:745-815) on lab-frame propagators ----
    g3 = np.load(os.path.join(HERE, "rambo_n3.npz"))
    mom3 = torch.from_numpy(g3["mom_clean"])
    lab = utils.boost_to_lab_frame(mom3, torch.from_numpy(g3["x1_clean"]), torch.from_numpy(g3["x2_clean"]))
    gen = torch.Generator().manual_seed(99)
    def angles():
        return torch.stack((torch.randn(N, generator=gen, dtype=torch.float64) * 1.5,
                            (torch.rand(N, generator=gen, dtype=torch.float64) * 2 - 1) * math.pi), dim=1)
    ha, tb, tw = angles(), angles(), angles()
    dk = {"higgs": lab[:, 2].numpy(), "top": lab[:, 3].numpy(), "higgs_angles": ha.numpy(), "top_b_angles": tb.numpy(),
          "top_W_angles": tw.numpy()}
    for flag in (True, False):
        tag = "ptetaphi" if flag else "cart"
        dk["higgs_out_" + tag] = Compute_ParticlesTensor.higgs_decayProducts(lab[:, 2], ha, ptetaphi=flag).numpy()
        dk["top_out_" + tag] = Compute_ParticlesTensor.top_decayProducts(lab[:, 3], tb, tw, ptetaphi=flag).numpy()
    dk["top_out_massive_b"] = Compute_ParticlesTensor.top_decayProducts(lab[:, 3], tb, tw, b_mass=4.7, ptetaphi=False).numpy()
    np.savez_compressed(os.path.join(HERE, "decays_n3.npz"), **dk)
    print("decay products written")

    # ---- spline fixture from the torch restatement (parity unpinned) ----
    sys.path.insert(0, ROOT)
    from oracle import rqs_torch
    for tag, dt, F, K, bound in (("f64_F7K10", torch.float64, 7, 10, 1.0), ("f32_F20K64", torch.float32, 20, 64, 1.1),
                                 ("f64_F10K40", torch.float64, 10, 40, 1.1)):
        g = torch.Generator().manual_seed(7)
        B = 128 if K <= 16 else 40
        phi = torch.randn(B, F, 3 * K - 1, generator=g, dtype=torch.float64).to(dt)
        x = ((torch.rand(B, F, generator=g, dtype=torch.float64) * 2.6 - 1.3) * bound / 1.0).to(dt)
        t = rqs_torch.MonotonicRQSTransform(phi[..., :K], phi[..., K:2 * K], phi[..., 2 * K:], bound=bound)
        y, ladj = t.call_and_ladj(x)
        xi = t._inverse(x)
        np.savez_compressed(os.path.join(HERE, "rqs_%s.npz" % tag), phi=phi.numpy(), x=x.numpy(), y=y.numpy(),
                            ladj=ladj.numpy(), x_inv=xi.numpy(), bound=np.float64(bound),
                            bin_fwd=(t.searchsorted(t.horizontal, x) - 1).numpy().astype(np.int64),
                            bin_inv=(t.searchsorted(t.vertical, x) - 1).numpy().astype(np.int64))
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)
    print(known)


def make_decay_partons():
    """decay_partons_*.npz: Compute_ParticlesTensor.get_decayPartons_fromlab_propagators_angles (utils.py:826-935) on
    lab-frame H / thad / tlep taken from the n=4 (real ISR gluon) and n=3 (no ISR: the reference's padding row) events.
    Also stores the reference's autograd gradient of sum(out * g_out) w.r.t. every input of the first variant."""
    _, _, utils = import_reference()
    for name in ("vector", "awkward"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from memflow.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    rows = []
    for n in (4, 3):
        g = np.load(os.path.join(HERE, "rambo_n%d.npz" % n))
        lab = utils.boost_to_lab_frame(torch.from_numpy(g["mom_clean"]), torch.from_numpy(g["x1_clean"]),
                                       torch.from_numpy(g["x2_clean"]))
        keep = torch.isfinite(lab).all(dim=(1, 2))
        lab = lab[keep][:160 if n == 4 else 32]
        e_cm = float(g["e_cm"])
        x1, x2 = torch.from_numpy(g["x1_clean"])[keep][:lab.shape[0]], torch.from_numpy(g["x2_clean"])[keep][:lab.shape[0]]
        rows.append((lab[:, 2:5], torch.stack((e_cm * (x1 + x2) / 2, e_cm * (x1 - x2) / 2), dim=1)))
    Htt = torch.cat([r[0] for r in rows])          # [N,3,4] cartesian lab, order H, t, t
    boost_Epz = torch.cat([r[1] for r in rows])    # [N,2]
    N = Htt.shape[0]
    gen = torch.Generator().manual_seed(1234)
    mean_prop, std_prop = torch.tensor([4.9, 0.02, 0.01]), torch.tensor([0.8, 1.4, 1.8])
    mean_lab, std_lab = torch.tensor([3.9, -0.01, 0.02]), torch.tensor([0.9, 1.6, 1.81])
    mean_boost, std_boost = torch.tensor([6.8, 11.0]), torch.tensor([0.45, 610.0])
    pep = CPT.get_ptetaphi_comp_batch(Htt)
    def angles():
        return torch.stack((torch.randn(N, generator=gen, dtype=torch.float64) * 1.3,
                            (torch.rand(N, generator=gen, dtype=torch.float64) * 2 - 1) * math.pi), dim=1)
    ang = [angles() for _ in range(5)]
    out = {"mean_prop": mean_prop.numpy(), "std_prop": std_prop.numpy(), "mean_lab": mean_lab.numpy(),
           "std_lab": std_lab.numpy(), "mean_boost": mean_boost.numpy(), "std_boost": std_boost.numpy(),
           "angles": torch.stack(ang).numpy()}
    boost_scaled = boost_Epz.clone()
    boost_scaled[:, 0] = torch.log(boost_Epz[:, 0] + 1)
    boost_scaled = ((boost_scaled - mean_boost) / std_boost)[:, None, :]
    out["boost"] = boost_scaled.numpy()
    variants = {  # name: (unscale_phi, ptetaphi, final_scaling, masses kw)
        "default": (False, True, True, {}),
        "unscalephi": (True, True, True, {}),
        "noscale": (False, True, False, {}),
        "cart": (False, False, False, {}),
        "masses": (True, True, True, dict(thad_mass=171.0, tlep_mass=174.0, W_had_mass=79.0, W_lep_mass=81.5, b_mass=4.7,
                                          higgs_mass=124.0, eps=1e-3)),
    }
    for name, (unscale_phi, ptetaphi, final_scaling, kw) in variants.items():
        prop = pep[..., 1:].clone()
        prop[..., 0] = torch.log(prop[..., 0] + 1)
        if unscale_phi:
            prop = (prop - mean_prop) / std_prop
        else:
            prop[..., :2] = (prop[..., :2] - mean_prop[:2]) / std_prop[:2]
        out["prop_" + name] = prop.numpy().copy()
        leaves = [prop.clone().requires_grad_(name == "default")] + [a.clone().requires_grad_(name == "default") for a in ang] \
            + [boost_scaled.clone().requires_grad_(name == "default")]
        # the reference writes into its first argument: hand it a non-leaf copy
        res = CPT.get_decayPartons_fromlab_propagators_angles(
            leaves[0] * 1.0, higgs_angles=leaves[1], thad_b_angles=leaves[2], thad_W_angles=leaves[3],
            tlep_b_angles=leaves[4], tlep_W_angles=leaves[5], boost=leaves[6], log_mean_parton_lab=mean_lab,
            log_std_parton_lab=std_lab, log_mean_boost=mean_boost, log_std_boost=std_boost,
            log_mean_parton_Hthadtlep=mean_prop, log_std_parton_Hthadtlep=std_prop, device=torch.device("cpu"),
            ptetaphi=ptetaphi, unscale_phi=unscale_phi, final_scaling=final_scaling, **kw)
        out["out_" + name] = res.detach().numpy()
        if name == "default":
            g_out = torch.randn(res.shape, generator=gen, dtype=torch.float64)
            (res * g_out).sum().backward()
            out["g_out"] = g_out.numpy()
            out["g_prop"] = leaves[0].grad.numpy()
            out["g_angles"] = torch.stack([l.grad for l in leaves[1:6]]).numpy()
            out["g_boost"] = leaves[6].grad.numpy()
    out["kw_masses"] = np.array([124.0, 171.0, 174.0, 79.0, 81.5, 4.7, 1e-3])
    np.savez_compressed(os.path.join(HERE, "decay_partons.npz"), **out)
    print("decay_partons written:", N, "events;", int((np.abs(out["out_default"][:, 11, 0] - 0.001) < 1e-12).sum()), "padded ISR rows")



def make_sampling_pipeline():
    """sampling_pipeline.npz: the sampling branch of the training loss, memflow/unfolding_flow/unfolding_flow_withDecay_v2.py
    :376-417, run with the unmodified reference on CPU in float64: ps -> PhaseSpace.get_momenta_from_ps(requires_grad=True)
    -> boost to the lab -> (pt, eta, phi) -> log / scaling -> get_decayPartons_fromlab_propagators_angles; stores the
    12-parton event, the scaled boost, and d(sum(out * g))/d(ps)."""
    phasespace, _, utils = import_reference()
    for name in ("vector", "awkward"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from memflow.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    E_CM, N = 13000.0, 160
    masses = torch.tensor([125.25, 172.5, 172.5, 0.0])
    gen = torch.Generator().manual_seed(77)
    ps = (torch.rand(N, 10, generator=gen, dtype=torch.float64) * 0.9 + 0.05).requires_grad_(True)
    mean_prop, std_prop = torch.tensor([4.9, 0.02, 0.01]), torch.tensor([0.8, 1.4, 1.8])
    mean_lab, std_lab = torch.tensor([3.9, -0.01, 0.02]), torch.tensor([0.9, 1.6, 1.81])
    mean_boost, std_boost = torch.tensor([6.8, 11.0]), torch.tensor([0.45, 610.0])
    angles = torch.stack([torch.stack((torch.randn(N, generator=gen, dtype=torch.float64) * 1.3,
                                       (torch.rand(N, generator=gen, dtype=torch.float64) * 2 - 1) * math.pi), dim=1)
                          for _ in range(5)])
    with contextlib.redirect_stdout(io.StringIO()):
        rambo = phasespace.PhaseSpace(E_CM, [21, 21], [25, 6, -6, 21], final_masses=masses, dev="cpu")
        mom, _, x1, x2 = rambo.get_momenta_from_ps(ps, requires_grad=True)
    zeros = torch.zeros_like(x1)
    boost_Epz = torch.stack((rambo.collider_energy * (x1 + x2) / 2, zeros, zeros, rambo.collider_energy * (x1 - x2) / 2), dim=1)
    lab = utils.boost_tt(mom[:, -4:-1], utils.boostVector_t(boost_Epz).unsqueeze(dim=1))
    pep = CPT.get_ptetaphi_comp_batch(lab)
    prop = pep[..., 1:].clone()
    prop[..., 0] = torch.log(pep[..., 1] + 1)
    prop[..., 0:2] = (prop[..., 0:2] - mean_prop[:2]) / std_prop[:2]
    boost_scaled = boost_Epz[..., [0, 3]].clone()
    boost_scaled[..., 0] = torch.log(boost_Epz[..., 0] + 1)
    boost_scaled = (boost_scaled - mean_boost) / std_boost
    out = CPT.get_decayPartons_fromlab_propagators_angles(
        prop, higgs_angles=angles[0], thad_b_angles=angles[1], thad_W_angles=angles[2], tlep_b_angles=angles[3],
        tlep_W_angles=angles[4], boost=boost_scaled[:, None, :], log_mean_parton_lab=mean_lab, log_std_parton_lab=std_lab,
        log_mean_boost=mean_boost, log_std_boost=std_boost, log_mean_parton_Hthadtlep=mean_prop,
        log_std_parton_Hthadtlep=std_prop, device=torch.device("cpu"), ptetaphi=True, eps=1e-4, unscale_phi=False,
        final_scaling=True)
    g_out = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    g_boost = torch.randn(boost_scaled.shape, generator=gen, dtype=torch.float64)
    ((out * g_out).sum() + (boost_scaled * g_boost).sum()).backward()
    np.savez_compressed(os.path.join(HERE, "sampling_pipeline.npz"), ps=ps.detach().numpy(), masses=masses.numpy(),
                        e_cm=np.float64(E_CM), angles=angles.numpy(), mean_prop=mean_prop.numpy(), std_prop=std_prop.numpy(),
                        mean_lab=mean_lab.numpy(), std_lab=std_lab.numpy(), mean_boost=mean_boost.numpy(),
                        std_boost=std_boost.numpy(), out=out.detach().numpy(), boost_scaled=boost_scaled.detach().numpy(),
                        g_out=g_out.numpy(), g_boost=g_boost.numpy(), g_ps=ps.grad.numpy())
    print("sampling pipeline written; |g_ps| columns:", np.abs(ps.grad.numpy()).max(0))


def make_dataset_products():
    """dataset_products.npz: the tensors MEMFlow caches per dataset, memflow/read_data/Dataset_Parton_Level.py:400-463
    (get_intermediateParticles_cartesian_onShell, get_PS_intermediateParticles_onShell, get_PS_intermediateParticles).
    The module itself cannot be imported here (awkward / numba / hist are not installed), so the three methods are lifted
    UNMODIFIED from the reference's source with `ast` and run on a stub `self` that holds the three data tensors and
    collects what the methods `torch.save`. torch default dtype float64, like every other fixture of this directory."""
    import ast
    phasespace, _, _ = import_reference()
    src_path = os.path.join(REF, "memflow", "read_data", "Dataset_Parton_Level.py")
    tree = ast.parse(open(src_path).read())
    ns = {"torch": torch, "np": np, "PhaseSpace": phasespace.PhaseSpace}
    for node in tree.body:   # module-level constants M_HIGGS, M_TOP, M_GLUON
        if isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id.startswith("M_") for t in node.targets):
            exec(compile(ast.Module([node], []), src_path, "exec"), ns)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Dataset_PartonLevel")
    wanted = ("get_intermediateParticles_cartesian_onShell", "get_PS_intermediateParticles_onShell", "get_PS_intermediateParticles")
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.Module([node], []), src_path, "exec"), ns)
    saved = {}
    real_save = torch.save
    N, E_CM = 2000, 13000.0
    masses = torch.tensor([ns["M_HIGGS"], ns["M_TOP"], ns["M_TOP"], ns["M_GLUON"]], dtype=torch.float64)
    gen = torch.Generator().manual_seed(2025)
    u = torch.rand(N, 10, generator=gen, dtype=torch.float64) * 0.98 + 0.01
    with contextlib.redirect_stdout(io.StringIO()):
        rambo = phasespace.PhaseSpace(E_CM, [21, 21], [25, 6, -6, 21], final_masses=masses, dev="cpu")
        mom, _, x1, x2 = rambo.get_momenta_from_ps(u)
    data = mom[:, 2:].clone()
    # the dataset's propagators are off shell (event-by-event masses): energies moved by 1e-4 relative
    data[:, :, 0] *= 1.0 + 1e-4 * torch.randn(N, 4, generator=gen, dtype=torch.float64)
    boost = torch.stack((E_CM * (x1 + x2) / 2, torch.zeros(N), torch.zeros(N), E_CM * (x1 - x2) / 2), dim=1).unsqueeze(1)
    boost[:50, 0, 0] = boost[:50, 0, 3].abs() + 100.0   # a few events below the minimal energy (:418-420)

    class Stub:
        pass
    stub = Stub()
    stub.data_higgs_t_tbar_ISR_cartesian = data.clone()
    stub.data_boost = boost.clone()
    stub.processed_file_names = lambda name: name
    torch.save = lambda obj, name: saved.__setitem__(name, obj)
    out = {"data": data.numpy(), "boost": boost.numpy(), "masses": masses.numpy(), "e_cm": np.float64(E_CM)}
    try:
        # literal reference first, then "fp64-clean" (its two float32 casts neutralised, see the module docstring): the
        # det-jacobians differ at 1e-7 relative, everything else is identical
        for tag, ctx in (("", contextlib.nullcontext()), ("__clean", fp64_clean())):
            saved.clear()
            stub.data_boost = boost.clone()
            with contextlib.redirect_stdout(io.StringIO()), ctx:
                ns["get_intermediateParticles_cartesian_onShell"](stub)
                stub.data_higgs_t_tbar_ISR_cartesian_onShell = saved["H_thad_tlep_ISR_cartesian_onShell"]
                ns["get_PS_intermediateParticles"](stub)                 # uses data_boost as is
                ns["get_PS_intermediateParticles_onShell"](stub)         # modifies data_boost in place (:418-420)
            for k, v in saved.items():
                if tag and "detjacobian" not in k:
                    continue
                if isinstance(v, tuple):
                    out[k + "__mean"], out[k + "__scale"] = v[0].numpy(), v[1].numpy()
                else:
                    out[k + tag] = v.numpy()
    finally:
        torch.save = real_save
    np.savez_compressed(os.path.join(HERE, "dataset_products.npz"), **out)
    print("dataset products written:", sorted(saved), "; nan rows on shell:",
          int(np.isnan(out["phasespace_intermediateParticles_onShell"]).any(axis=1).sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "decay_partons":
        make_decay_partons()
    elif len(sys.argv) > 1 and sys.argv[1] == "sampling_pipeline":
        make_sampling_pipeline()
    elif len(sys.argv) > 1 and sys.argv[1] == "dataset_products":
        make_dataset_products()
    else:
        main()
        make_decay_partons()
        make_sampling_pipeline()
