"""Parity metrics, stated once and used by every test.

The reference's own formulas are ill-conditioned in three places, and any last-ulp difference between two
correct implementations (libm vs Sleef vs libdevice cos/exp/log) is amplified there:
  * sin(phi) = +-sqrt(1 - cos(phi)^2)      (rambo_generator.py:321-322): factor ~ 1/|sin phi|
  * boosts with gamma >> 1                  (utils.py:99: gamma = 1/sqrt(1 - b^2)):  factor ~ gamma^2
  * mass-variable roots with u -> 1         (bisection on a function with f'(1) = 0): factor ~ 1/(1-u)^2 on the
                                             weight and 1/(1-u) on the momenta
So the bound for an event is  tol * kappa(event)  with kappa >= 1 computed from the ORACLE's own outputs.
`kappa == 1` for a well-conditioned event, where the plain tolerance (1e-12 fp64 / 1e-5 fp32) applies.
"""
import numpy as np


def momenta_error(mom, mom_ref, x1_ref, x2_ref, e_collider):
    """max over particles/components of |p - p_ref| / E_cm(event)."""
    ecm = np.sqrt(x1_ref * x2_ref) * e_collider
    return np.abs(mom - mom_ref).reshape(mom.shape[0], -1).max(axis=1) / ecm


def conditioning(u, mom_ref, n, masses=None):
    """kappa per event from the oracle's momenta (final rows 2..) and the input point u.
    With `masses`, also the threshold cancellation of rho(M_i, M_{i+1}, m_i) (rambo_generator.py:145-149):
    M_i^2 - (M_{i+1} + m_i)^2 loses M_i^2 / (M_i^2 - (M_{i+1}+m_i)^2) digits."""
    N = u.shape[0]
    fin = mom_ref[:, 2:, :]
    # gamma of every intermediate system Q_k = sum_{j>=k} p_j, in the frame the decay is boosted from
    q = np.cumsum(fin[:, ::-1, :], axis=1)[:, ::-1, :]
    m2 = q[..., 0] ** 2 - (q[..., 1:] ** 2).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        gamma2 = np.where(m2 > 0, q[..., 0] ** 2 / m2, np.inf)
    g2 = np.nanmax(gamma2[:, : n - 1], axis=1)
    ne = n - 2
    phi = 2 * np.pi * u[:, ne + 1: 3 * n - 4: 2]
    sinphi = np.abs(np.sin(phi)).min(axis=1)
    ct = 2 * u[:, ne: 3 * n - 4: 2] - 1
    sinth = np.sqrt(np.maximum(1 - ct * ct, 0)).min(axis=1)
    k_angle = 1.0 + 1e-3 / np.maximum(np.minimum(sinphi, sinth), 1e-300)
    if ne > 0:
        uv = u[:, :ne]
        k_mass = (1.0 / np.maximum(np.minimum(uv, 1 - uv), 1e-300) ** 2).max(axis=1) * 1e-2 + 1.0
    else:
        k_mass = np.ones(N)
    k_thr = np.ones(N)
    if masses is not None:
        M = np.sqrt(np.maximum(m2, 0.0))                      # M_i, i = 0..n-1 (M_{n-1} = m_{n-1})
        for i in range(n - 1):
            nxt = M[:, i + 1] if i + 1 < n - 1 else np.full(N, float(masses[n - 1]))
            gap = m2[:, i] - (nxt + float(masses[i])) ** 2
            with np.errstate(divide="ignore", invalid="ignore"):
                k_thr = np.maximum(k_thr, np.where(gap > 0, m2[:, i] / gap, np.inf))
    return np.maximum(g2, 1.0) * k_angle * k_mass * k_thr


def conditioning_fp32_weight(kappa, x1_ref, x2_ref, masses, e_collider):
    """Extra factor for the LOG-WEIGHT in float32: x1, x2 = exp(-a) carry eps32 |a| relative error, so the available energy
    K_0 = E_cm - sum m inherits eps32 |a| E_cm / K_0 and the weight ~ K_0^(2n-4) multiplies it by its exponent. (The same
    factor exists in float64, where it stays far below 1e-12.)"""
    ecm = np.sqrt(x1_ref * x2_ref) * e_collider
    k0 = np.maximum(ecm - float(np.sum(masses)), 1e-300)
    return kappa * (1.0 + ecm / k0)


def summarize(err, tol, kappa=None):
    out = {"max": float(np.nanmax(err)), "p50": float(np.nanquantile(err, 0.5)),
           "p99": float(np.nanquantile(err, 0.99)), "frac_gt_tol": float((err > tol).mean())}
    if kappa is not None:
        out["max_over_kappa"] = float(np.nanmax(err / kappa))
    return out


# ---- get_decayPartons_fromlab_propagators_angles (tests/golden/decay_partons.npz) ----
DECAY_PARTONS_VARIANTS = {  # name: (unscale_phi, ptetaphi, final_scaling, reference keyword arguments)
    "default": (False, True, True, {}),
    "unscalephi": (True, True, True, {}),
    "noscale": (False, True, False, {}),
    "cart": (False, False, False, {}),
    "masses": (True, True, True, dict(higgs_mass=124.0, thad_mass=171.0, tlep_mass=174.0, W_had_mass=79.0,
                                      W_lep_mass=81.5, b_mass=4.7, eps=1e-3)),
}


def decay_partons_scaled_err(out, ref, ref_cart, ptetaphi):
    """|out - ref| in units of what fp64 can resolve for that row: cartesian components carry the rounding of the
    event's largest momentum S (the ISR row is a difference of TeV-scale numbers); (pt, eta, phi) of a row inherit
    it amplified by S / pt."""
    S = np.nanmax(np.abs(ref_cart), axis=(1, 2))[:, None, None]
    if not ptetaphi:
        return np.abs(out - ref) / S
    pt = np.sqrt(ref_cart[..., 1] ** 2 + ref_cart[..., 2] ** 2)[..., None]
    kappa = np.maximum(S / np.maximum(pt, 1e-300), 1.0)
    return np.abs(out - ref) / (1.0 + np.abs(ref)) / kappa
