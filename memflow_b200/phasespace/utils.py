"""Lorentz / x1x2 helpers with the names and argument meaning of memflow/phasespace/utils.py.

These are the small free functions outside callers import from the reference
(`boostVector_t, boost_t, boost_tt, rho2_t, square_t, get_uniform_from_x1x2, ...`:
memflow/unfolding_flow/unfolding_flow.py:147-148, memflow/unfolding_flow/utils.py:760-761,1080-1150).
Inside the hot path (K1 `generateKinematics_batch`, K2 `getPSpoint_batch`) none of them is called:
their math is fused into the CUDA kernels (memflow_b200/csrc/rambo.cuh). They are kept here as plain,
device-agnostic torch expressions so that glue code written against the reference keeps importing;
unlike the reference they never synchronise with the host and never clone their inputs.
"""
import math

import numpy as np
import torch


def rho2_t(p):
    """|p_vec|^2 of [N,4] vectors (reference utils.py:24-27)."""
    s = p[:, 1:]
    return (s * s).sum(-1)


def rho2_tt(p):
    """|p_vec|^2 of [N,P,4] vectors (utils.py:30-33)."""
    return (p[:, :, 1:] ** 2).sum(-1)


def dot_t(a, b):
    """Minkowski product of [N,4] vectors (utils.py:58-67)."""
    return a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2] - a[:, 3] * b[:, 3]


def dot_s(a, b):
    """Euclidean product of the first three columns (utils.py:79-85)."""
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]


def square_t(p):
    """Minkowski square for 4-column input, plain sum of squares otherwise (utils.py:45-51)."""
    if p.shape[1] == 4 or p.shape[0] == 4:
        return dot_t(p, p)
    return (p * p).sum(-1)


def mass_t(p):
    return torch.sqrt(square_t(p))


def set_square_t(p, square, negative=False):
    """Return p with its energy reset so that p^2 = square (utils.py:6-21)."""
    e = (rho2_t(p) + square) ** 0.5
    if negative:
        e = -e
    return torch.cat((e.unsqueeze(1), p[:, 1:]), dim=1)


def boostVector_t(p):
    """beta = p_vec / E (utils.py:36-42). The reference prints "Invalid boost" after two host syncs;
    here invalid boosts simply propagate NaN/inf."""
    return p[:, 1:] / p[:, 0].unsqueeze(1)


def _boost(E, space, beta, gamma):
    b2 = (beta * beta).sum(-1)
    if not torch.is_tensor(gamma) and gamma < 0.0:
        gamma = 1.0 / torch.sqrt(1.0 - b2)
    bp = (space * beta).sum(-1)
    gamma2 = torch.where(b2 > 0, (gamma - 1.0) / b2, torch.zeros_like(b2))
    factor = gamma2 * bp + gamma * E
    return gamma * (E + bp), space + factor.unsqueeze(-1) * beta


def boost_t(p, boost_vector, gamma=-1.0):
    """General Lorentz boost of [N,4] by velocity [N,3] (utils.py:88-118)."""
    e, s = _boost(p[:, 0], p[:, 1:], boost_vector, gamma)
    return torch.cat((e.unsqueeze(-1), s), dim=-1)


def boost_tt(p, boost_vector, gamma=-1.0):
    """Batched version: p [N,P,4], boost_vector [N,1,3] or [N,P,3] (utils.py:121-146)."""
    e, s = _boost(p[:, :, 0], p[:, :, 1:], boost_vector, gamma)
    return torch.cat((e.unsqueeze(-1), s), dim=-1)


def boost_to_lab_frame(momenta, xb_1, xb_2):
    """CM -> lab boost along z with beta_z = (x1-x2)/(x1+x2) (utils.py:180-202). Per event (the reference
    skips the boost for the WHOLE batch if any event has x1 == x2, :189; that quirk is not reproduced).
    For the fused version pass MF_FLAG_LAB_FRAME / `lab_frame=True` to the forward map."""
    ref = momenta[:, 0, :] * xb_1.unsqueeze(-1) + momenta[:, 1, :] * xb_2.unsqueeze(-1)
    beta = boostVector_t(ref).unsqueeze(1)
    return boost_tt(momenta, beta)


def pseudoRap(p, eps=np.finfo(float).eps ** 0.5, huge=np.finfo(float).max):
    """Pseudorapidity of [N,4] (utils.py:205-215)."""
    pt = torch.sqrt((p[..., 1:3] ** 2).sum(-1))
    th = torch.atan2(pt, p[..., 3])
    eta = -torch.log(torch.tan(th / 2.0))
    return torch.where((pt < eps) & (p[..., 3].abs() < eps), torch.full_like(eta, huge), eta)


pseudoRap_t = pseudoRap  # [N,P,4] version (utils.py:218-228): same expression on the last axis


def getdelphi(a, b, eps=np.finfo(float).eps ** 0.5, huge=np.finfo(float).max):
    """Azimuthal separation (utils.py:231-249)."""
    pt1 = torch.sqrt((a[:, 1:3] ** 2).sum(-1))
    pt2 = torch.sqrt((b[:, 1:3] ** 2).sum(-1))
    c = (a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]) / (pt1 * pt2)
    out = torch.acos(torch.where(c.abs() > 1, c / c.abs(), c))
    return torch.where((pt1 == 0.0) | (pt2 == 0.0), torch.full_like(out, huge), out)


def deltaR(a, b):
    """sqrt(d_eta^2 + d_phi^2) (utils.py:252-257)."""
    return torch.sqrt((pseudoRap(a) - pseudoRap(b)) ** 2 + getdelphi(a, b) ** 2)


def uniform_distr(r, minv, maxv):
    """Map r in [0,1] onto (minv, maxv); returns (value, jacobian) (utils.py:161-169)."""
    width = torch.ones_like(r) * maxv - torch.ones_like(r) * minv
    return torch.ones_like(r) * minv + width * r, width


def _x1x2_constants(final_state_mass, E_cm):
    q = -torch.log((final_state_mass / E_cm) ** 2) if torch.is_tensor(final_state_mass) \
        else -math.log((final_state_mass / E_cm) ** 2)
    return q, 0.5 * q * q


def get_x1x2_from_uniform(unif, final_state_mass, E_cm):
    """[N,2] uniforms -> (x1, x2, weight = A x1 x2) with x1 x2 >= (sum m / E_cm)^2 (utils.py:260-284)."""
    q, A = _x1x2_constants(final_state_mass, E_cm)
    a = q * (1.0 - torch.sqrt(1.0 - unif[:, 0]))
    b = unif[:, 1] * (q - a)
    x1, x2 = torch.exp(-a), torch.exp(-b)
    return x1, x2, A * x1 * x2


def get_uniform_from_x1x2(x1, x2, final_state_mass, E_cm):
    """Inverse of get_x1x2_from_uniform; returns ([N,2] uniforms, jac = 1/(A x1 x2)) (utils.py:287-309)."""
    q, A = _x1x2_constants(final_state_mass, E_cm)
    a, b = -torch.log(x1), -torch.log(x2)
    u1 = (q * a - 0.5 * a * a) / A
    u2 = b / (q - a)
    return torch.stack((u1, u2), dim=1), 1.0 / (A * x1 * x2)
