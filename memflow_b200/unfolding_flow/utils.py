"""The phase-space entries of `Compute_ParticlesTensor` (memflow/unfolding_flow/utils.py) on CUDA kernels:

  get_PS               :1056-1152  K2 in get_PS mode (called every training step, unfolding_flow.py:163)
  higgs_decayProducts  :745-770    H -> b b two-body decay + boost           
  top_decayProducts    :772-815    t -> b W -> b q q' decays + boosts        
  get_decayPartons_fromlab_propagators_angles :826-935  the whole 12-row parton event in one launch, differentiable

The rest of the reference's 1100-line glue class (regression scalings, dataset-specific conversions) is model
glue and out of scope.
"""
import ctypes

import torch

from .._lib import check, DecayPartonsCfg, device_guard, host_array, lib, ptr, require_cuda, stream_ptr

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 1e-5  # memflow/unfolding_flow/utils.py:10


class _GetPSFn(torch.autograd.Function):
    """get_PS with a backward for `ps` w.r.t. the regressed momenta and boost (mf_rambo_inverse_bwd_f64)."""

    @staticmethod
    def forward(ctx, data, boost_reco, order, tm, E_CM):
        ps, jac0, mask = Compute_ParticlesTensor._launch_get_ps(data.detach(), boost_reco.detach(), order, tm, E_CM, None,
                                                               None, None)[:3]
        ctx.save_for_backward(data.detach().to(torch.float64).contiguous(), boost_reco.detach().to(torch.float64).contiguous())
        ctx.order, ctx.tm, ctx.E_CM, ctx.dtypes = order, tm, float(E_CM), (data.dtype, boost_reco.dtype)
        ctx.mark_non_differentiable(jac0, mask)
        return ps, jac0, mask

    @staticmethod
    def backward(ctx, g_ps, _gj, _gm):
        from ..phasespace.rambo_generator import _inverse_vjp
        mom, boost = ctx.saved_tensors
        E = ctx.E_CM
        x1r, x2r = (boost[:, 0] + boost[:, 3]) / E, (boost[:, 0] - boost[:, 3]) / E
        x1, x2 = x1r.clamp(0.0, 1.0), x2r.clamp(0.0, 1.0)
        n = mom.shape[1]
        g_mom, g_x1, g_x2 = _inverse_vjp(mom, x1.contiguous(), x2.contiguous(), g_ps.detach().to(torch.float64).contiguous(), n,
                                         ctx.order, ctx.tm, ctx.tm, True, E, 0, True)
        g_x1 = g_x1 * ((x1r >= 0) & (x1r <= 1))   # torch.clamp passes the gradient inside [min, max] only
        g_x2 = g_x2 * ((x2r >= 0) & (x2r <= 1))
        g_boost = torch.zeros_like(boost)
        g_boost[:, 0] = (g_x1 + g_x2) / E
        g_boost[:, 3] = (g_x1 - g_x2) / E
        return g_mom.to(ctx.dtypes[0]), g_boost.to(ctx.dtypes[1]), None, None, None


class Compute_ParticlesTensor:
    @staticmethod
    def get_PS(data_HttISR, boost_reco, order=[0, 1, 2, 3], target_mass=None, E_CM=13000, logit_eps=None,
               logit_mean=None, logit_scale=None):
        """(ps [N, 3n-2], 0 * jac_x1x2 [N], mask_problems [N] bool).

        data_HttISR [N, n, 4] CM-frame momenta (E,px,py,pz); boost_reco [N, 4] whose E +- pz give x1, x2
        (clamped to [0,1]); masses = target_mass[:, order]. mask_problems flags x out of range, |sum p|^2 > 1e-5
        and RAMBO variables outside [0,1] -- no exception, no host sync, exactly like the reference.
        Generalised from the reference's hard-coded n = 4 to n = 3..8.
        With `logit_eps` the caller's next two lines (unfolding_flow.py:170-171) are fused into the same kernel and a
        fourth value is returned: (torch.logit(ps, eps=logit_eps) - logit_mean) / logit_scale."""
        require_cuda(data_HttISR, "data_HttISR")
        n = data_HttISR.shape[1]
        if target_mass is None:
            target_mass = torch.tensor([[M_HIGGS, M_TOP, M_TOP, M_GLUON]], dtype=torch.float64)
        tm = torch.as_tensor(target_mass).detach().cpu().double().reshape(-1).tolist()
        order = [int(o) for o in order]
        if len(tm) != n or len(order) != n:
            raise AssertionError("target_mass / order must have %d entries" % n)
        if torch.is_grad_enabled() and (data_HttISR.requires_grad or boost_reco.requires_grad):
            ps, jac0, mask = _GetPSFn.apply(data_HttISR, boost_reco, order, tm, E_CM)
            if logit_eps is None:
                return ps, jac0, mask
            # differentiable path: the epilogue stays in torch so that autograd sees it (unfolding_flow.py:170-171)
            lps = torch.logit(ps, eps=logit_eps)
            if logit_mean is not None:
                lps = lps - torch.as_tensor(logit_mean, dtype=ps.dtype, device=ps.device)
            if logit_scale is not None:
                lps = lps / torch.as_tensor(logit_scale, dtype=ps.dtype, device=ps.device)
            return ps, jac0, mask, lps
        out = Compute_ParticlesTensor._launch_get_ps(data_HttISR.detach(), boost_reco.detach(), order, tm, E_CM, logit_eps,
                                                     logit_mean, logit_scale)
        return out if logit_eps is not None else out[:3]

    @staticmethod
    def _launch_get_ps(data_HttISR, boost_reco, order, tm, E_CM, logit_eps, logit_mean, logit_scale):
        dev = data_HttISR.device
        n = data_HttISR.shape[1]
        mom = data_HttISR.to(torch.float64)
        if not (mom.stride(2) == 1 and mom.stride(1) == 4 and mom.stride(0) % 4 == 0 and mom.stride(0) >= 4 * n
                and mom.data_ptr() % 32 == 0):
            mom = mom.contiguous()
        boost = boost_reco.to(torch.float64).contiguous()
        N = mom.shape[0]
        ps = torch.empty((N, 3 * n - 2), dtype=torch.float64, device=dev)
        jac0 = torch.empty(N, dtype=torch.float64, device=dev)
        mask = torch.empty(N, dtype=torch.uint8, device=dev)
        W = 3 * n - 2
        lps, lm, ls = None, None, None
        if logit_eps is not None:
            lps = torch.empty((N, W), dtype=torch.float64, device=dev)
            if logit_mean is not None:
                lm = host_array(torch.as_tensor(logit_mean).detach().cpu().double().reshape(-1).tolist(), ctypes.c_double)
            if logit_scale is not None:
                ls = host_array(torch.as_tensor(logit_scale).detach().cpu().double().reshape(-1).tolist(), ctypes.c_double)
            if (lm is not None and len(lm) != W) or (ls is not None and len(ls) != W):
                raise AssertionError("logit_mean / logit_scale must have 3n-2 = %d entries" % W)
        with device_guard(dev):
            rc = lib().mf_get_ps_f64(ptr(mom), ctypes.c_int64(mom.stride(0)), ptr(boost), ctypes.c_int64(N), ctypes.c_int(n),
                                     host_array(order, ctypes.c_int32), host_array(tm, ctypes.c_double),
                                     ctypes.c_double(float(E_CM)), ptr(ps), ptr(jac0), ptr(mask),
                                     ctypes.c_double(float(logit_eps) if logit_eps is not None else 0.0), lm, ls, ptr(lps),
                                     stream_ptr(dev))
        check(rc, "mf_get_ps_f64")
        return ps, jac0, mask.bool(), lps

    @staticmethod
    def _rows(p):
        """[N,4] momenta as (tensor, row stride) readable in place by the kernels (32 B-aligned rows)."""
        p = p.detach().to(torch.float64)
        if not (p.dim() == 2 and p.shape[1] == 4 and p.stride(1) == 1 and p.stride(0) % 4 == 0 and p.stride(0) >= 4
                and p.data_ptr() % 32 == 0):
            p = p.contiguous()
        return p, p.stride(0)

    @staticmethod
    def higgs_decayProducts(higgs_momenta, higgs_angles, higgs_mass=125.25, ptetaphi=True, device=None):
        """(H, b1, b2) [N,3,4]: b1 has (eta, phi) = higgs_angles in the Higgs rest frame, E = m_H/2, b2 is back to back;
        both boosted by the Higgs velocity. ptetaphi=True returns (E, pt, eta, phi) rows like the reference."""
        require_cuda(higgs_momenta, "higgs_momenta")
        dev = higgs_momenta.device
        mom, stride = Compute_ParticlesTensor._rows(higgs_momenta)
        ang = higgs_angles.detach().to(torch.float64).contiguous()
        N = mom.shape[0]
        out = torch.empty((N, 3, 4), dtype=torch.float64, device=dev)
        with device_guard(dev):
            check(lib().mf_higgs_decay_f64(ptr(mom), ctypes.c_int64(stride), ptr(ang), ctypes.c_int64(N),
                                           ctypes.c_double(float(higgs_mass)), ctypes.c_int(int(bool(ptetaphi))), ptr(out),
                                           stream_ptr(dev)), "mf_higgs_decay_f64")
        return out

    @staticmethod
    def top_decayProducts(top_momenta, top_b_angles, top_W_angles, top_mass=172.5, W_mass=80.4, b_mass=0.0, ptetaphi=True,
                          device=None):
        """(t, b, q1, q2) [N,4,4]: t -> b W in the top rest frame, W -> q1 q2 in the W rest frame, boosted in turn."""
        require_cuda(top_momenta, "top_momenta")
        dev = top_momenta.device
        mom, stride = Compute_ParticlesTensor._rows(top_momenta)
        ab = top_b_angles.detach().to(torch.float64).contiguous()
        aw = top_W_angles.detach().to(torch.float64).contiguous()
        N = mom.shape[0]
        out = torch.empty((N, 4, 4), dtype=torch.float64, device=dev)
        with device_guard(dev):
            check(lib().mf_top_decay_f64(ptr(mom), ctypes.c_int64(stride), ptr(ab), ptr(aw), ctypes.c_int64(N),
                                         ctypes.c_double(float(top_mass)), ctypes.c_double(float(W_mass)),
                                         ctypes.c_double(float(b_mass)), ctypes.c_int(int(bool(ptetaphi))), ptr(out),
                                         stream_ptr(dev)), "mf_top_decay_f64")
        return out

    # ---- coordinate conversions used around the kernels by the sampling loss (plain differentiable torch, any device;
    #      reference memflow/unfolding_flow/utils.py:109-119, :169-187). Inside the kernels the same formulas are fused. ----
    @staticmethod
    def get_cartesian_comp(particle, mass):
        """(pt, eta, phi) [N,3] + mass -> (E, px, py, pz) [N,4]."""
        px = particle[:, 0] * torch.cos(particle[:, 2])
        py = particle[:, 0] * torch.sin(particle[:, 2])
        pz = particle[:, 0] * torch.sinh(particle[:, 1])
        return torch.stack((torch.sqrt(px ** 2 + py ** 2 + pz ** 2 + mass ** 2), px, py, pz), dim=1)

    @staticmethod
    def _ptetaphi(E, px, py, pz, dim):
        pt = (px ** 2 + py ** 2) ** 0.5
        eta = -torch.log(torch.tan(torch.atan2(pt, pz) / 2))
        cols = (pt, eta, torch.atan2(py, px))
        return torch.stack(cols if E is None else (E,) + cols, dim=dim)

    @staticmethod
    def get_ptetaphi_comp(particle):
        """(E, px, py, pz) [N,4] -> (E, pt, eta, phi) [N,4]."""
        return Compute_ParticlesTensor._ptetaphi(particle[:, 0], particle[:, 1], particle[:, 2], particle[:, 3], 1)

    @staticmethod
    def get_ptetaphi_comp_batch(particles):
        """(E, px, py, pz) [N,P,4] -> (E, pt, eta, phi) [N,P,4]."""
        return Compute_ParticlesTensor._ptetaphi(particles[..., 0], particles[..., 1], particles[..., 2], particles[..., 3], 2)

    @staticmethod
    def get_ptetaphi_comp_noEnergy(particle):
        """(px, py, pz) [N,3] -> (pt, eta, phi) [N,3]."""
        return Compute_ParticlesTensor._ptetaphi(None, particle[:, 0], particle[:, 1], particle[:, 2], 1)

    @staticmethod
    def get_decayPartons_fromlab_propagators_angles(propagators_kinematics, higgs_angles, thad_b_angles, thad_W_angles,
                                                    tlep_b_angles, tlep_W_angles, boost, log_mean_parton_lab,
                                                    log_std_parton_lab, log_mean_boost, log_std_boost,
                                                    log_mean_parton_Hthadtlep, log_std_parton_Hthadtlep, device=None,
                                                    higgs_mass=125.25, thad_mass=172.5, tlep_mass=172.5, W_had_mass=80.4,
                                                    W_lep_mass=80.4, b_mass=0.0, ptetaphi=True, eps=1e-4, pt_cut=None,
                                                    unscale_phi=False, debug=False, final_scaling=True):
        """Full parton-level event [N,12,4] (H b1 b2 | thad b q1 q2 | tlep b l nu | ISR) from the scaled lab-frame
        propagators [N,3,3], the five decay-angle pairs [N,2] and the scaled boost [N,1,2], in one launch
        (`mf_decay_partons_f64`; reference memflow/unfolding_flow/utils.py:826-935, same argument list).
        Differences: `propagators_kinematics` is not unscaled in place; the result is differentiable w.r.t. every
        tensor input (backward kernel); `final_scaling=True, ptetaphi=False` raises NameError like the reference
        (its cartesian scaling reads an undefined name, :913)."""
        if final_scaling and not ptetaphi:
            raise NameError("name 'log_mean_parton' is not defined (memflow/unfolding_flow/utils.py:913: the cartesian "
                            "final scaling of the reference never worked; use ptetaphi=True or final_scaling=False)")
        require_cuda(propagators_kinematics, "propagators_kinematics")
        cfg = _decay_cfg(log_mean_parton_lab, log_std_parton_lab, log_mean_boost, log_std_boost, log_mean_parton_Hthadtlep,
                         log_std_parton_Hthadtlep, higgs_mass, thad_mass, tlep_mass, W_had_mass, W_lep_mass, b_mass, ptetaphi,
                         eps, unscale_phi, final_scaling)
        N = propagators_kinematics.shape[0]
        if tuple(propagators_kinematics.shape) != (N, 3, 3):
            raise ValueError("propagators_kinematics must be [N,3,3] (H, thad, tlep) x (log(pt+1), eta, phi)")
        ins = [propagators_kinematics, higgs_angles, thad_b_angles, thad_W_angles, tlep_b_angles, tlep_W_angles, boost]
        for a in ins[1:6]:
            if tuple(a.shape) != (N, 2):
                raise ValueError("decay angles must be [N,2] = (eta, phi)")
        if boost.numel() != 2 * N:
            raise ValueError("boost must be [N,1,2] = scaled (log(E+1), pz)")
        out = _DecayPartonsFn.apply(cfg, *ins)
        return out.to(propagators_kinematics.dtype)


class _DecayPartonsFn(torch.autograd.Function):
    """mf_decay_partons_f64 and its backward kernel mf_decay_partons_bwd_f64."""

    @staticmethod
    def forward(ctx, cfg, *ins):
        dev = ins[0].device
        d = [t.detach().to(torch.float64).contiguous() for t in ins]
        N = d[0].shape[0]
        out = torch.empty((N, 12, 4), dtype=torch.float64, device=dev)
        with device_guard(dev):
            check(lib().mf_decay_partons_f64(*[ptr(t) for t in d], ctypes.c_int64(N), ctypes.byref(cfg), ptr(out),
                                             stream_ptr(dev)), "mf_decay_partons_f64")
        ctx.cfg = cfg
        ctx.save_for_backward(*d)
        ctx.in_meta = [(t.shape, t.dtype) for t in ins]
        return out

    @staticmethod
    def backward(ctx, g_out):
        d = ctx.saved_tensors
        dev = d[0].device
        N = d[0].shape[0]
        g = g_out.detach().to(torch.float64).contiguous()
        grads = [torch.empty_like(t) for t in d]
        with device_guard(dev):
            check(lib().mf_decay_partons_bwd_f64(*[ptr(t) for t in d], ctypes.c_int64(N), ctypes.byref(ctx.cfg), ptr(g),
                                                 *[ptr(t) for t in grads], stream_ptr(dev)), "mf_decay_partons_bwd_f64")
        outs = [gr.reshape(shape).to(dt) if need else None
                for gr, (shape, dt), need in zip(grads, ctx.in_meta, ctx.needs_input_grad[1:])]
        return (None, *outs)


def _decay_cfg(log_mean_parton_lab, log_std_parton_lab, log_mean_boost, log_std_boost, log_mean_parton_Hthadtlep,
               log_std_parton_Hthadtlep, higgs_mass, thad_mass, tlep_mass, W_had_mass, W_lep_mass, b_mass, ptetaphi, eps,
               unscale_phi, final_scaling):
    """mf_decay_partons_cfg from the reference's keyword arguments (memflow/unfolding_flow/utils.py:826-845)."""
    cfg = DecayPartonsCfg()

    def put(field, t, n):
        v = [float(x) for x in torch.as_tensor(t).detach().reshape(-1).tolist()]
        if len(v) < n:
            raise ValueError("get_decayPartons_fromlab_propagators_angles: scaling vector shorter than %d" % n)
        for i in range(n):
            field[i] = v[i]
    put(cfg.mean_prop, log_mean_parton_Hthadtlep, 3 if unscale_phi else 2)
    put(cfg.std_prop, log_std_parton_Hthadtlep, 3 if unscale_phi else 2)
    put(cfg.mean_lab, log_mean_parton_lab, 3 if unscale_phi else 2)
    put(cfg.std_lab, log_std_parton_lab, 3 if unscale_phi else 2)
    put(cfg.mean_boost, log_mean_boost, 2)
    put(cfg.std_boost, log_std_boost, 2)
    cfg.higgs_mass, cfg.thad_mass, cfg.tlep_mass = float(higgs_mass), float(thad_mass), float(tlep_mass)
    cfg.w_had_mass, cfg.w_lep_mass, cfg.b_mass, cfg.eps = float(W_had_mass), float(W_lep_mass), float(b_mass), float(eps)
    cfg.ptetaphi, cfg.unscale_phi, cfg.final_scaling = int(bool(ptetaphi)), int(bool(unscale_phi)), int(bool(final_scaling))
    return cfg


def sampled_decayPartons_from_ps(rambo, ps_samples, higgs_angles, thad_b_angles, thad_W_angles, tlep_b_angles, tlep_W_angles,
                                 log_mean_parton_lab, log_std_parton_lab, log_mean_boost, log_std_boost,
                                 log_mean_parton_Hthadtlep, log_std_parton_Hthadtlep, higgs_mass=125.25, thad_mass=172.5,
                                 tlep_mass=172.5, W_had_mass=80.4, W_lep_mass=80.4, b_mass=0.0, ptetaphi=True, eps=1e-4,
                                 unscale_phi=False, final_scaling=True):
    """The sampling branch of the reference's training loss, memflow/unfolding_flow/unfolding_flow_withDecay_v2.py:376-417, as
    ONE forward launch (`mf_sampling_partons_f64`) and ONE backward launch (`mf_sampling_partons_bwd_f64`):

        momenta, _, x1, x2 = rambo.get_momenta_from_ps(ps_samples, requires_grad=True)
        boost_Epz = stack(E (x1 + x2) / 2, 0, 0, E (x1 - x2) / 2)
        lab = ps_utils.boost_tt(momenta[:, -4:-1], ps_utils.boostVector_t(boost_Epz).unsqueeze(1))
        prop = scaled (log(pt + 1), eta, phi) of get_ptetaphi_comp_batch(lab)
        boost_scaled = ((log(boost_E + 1), boost_pz) - log_mean_boost) / log_std_boost
        partons = Compute_ParticlesTensor.get_decayPartons_fromlab_propagators_angles(prop, <angles>, boost_scaled[:, None], ...)

    `rambo` is a memflow_b200 PhaseSpace with n >= 4 final-state bodies (H, thad, tlep, ISR). Returns
    (partons [N,12,4], boost_scaled [N,2], momenta [N,n+2,4], weight [N], x1 [N], x2 [N]); partons and boost_scaled are
    differentiable w.r.t. ps_samples (the reference graph's semantics: no gradient through the mass-variable columns) and
    the five angle pairs; momenta / weight / x1 / x2 are returned detached (differentiate them through
    `rambo.get_momenta_from_ps` instead)."""
    if final_scaling and not ptetaphi:
        raise NameError("name 'log_mean_parton' is not defined (memflow/unfolding_flow/utils.py:913)")
    require_cuda(ps_samples, "ps_samples")
    gen = rambo.generator
    n = gen.n_final
    if n < 4:
        raise ValueError("sampled_decayPartons_from_ps needs n >= 4 final-state bodies (momenta[:, -4:-1])")
    N = ps_samples.shape[0]
    if tuple(ps_samples.shape) != (N, 3 * n - 2):
        raise ValueError("ps_samples must be [N, 3n-2]")
    angles = [higgs_angles, thad_b_angles, thad_W_angles, tlep_b_angles, tlep_W_angles]
    for a in angles:
        if tuple(a.shape) != (N, 2):
            raise ValueError("decay angles must be [N,2] = (eta, phi)")
    cfg = _decay_cfg(log_mean_parton_lab, log_std_parton_lab, log_mean_boost, log_std_boost, log_mean_parton_Hthadtlep,
                     log_std_parton_Hthadtlep, higgs_mass, thad_mass, tlep_mass, W_had_mass, W_lep_mass, b_mass, ptetaphi, eps,
                     unscale_phi, final_scaling)
    masses = list(gen._masses_host)
    out, bs, mom, w, x1, x2 = _SamplingPartonsFn.apply(cfg, masses, float(gen.collider_energy), ps_samples, *angles)
    return out.to(ps_samples.dtype), bs.to(ps_samples.dtype), mom, w, x1, x2


class _SamplingPartonsFn(torch.autograd.Function):
    """mf_sampling_partons_f64 / mf_sampling_partons_bwd_f64."""

    @staticmethod
    def forward(ctx, cfg, masses, e_collider, ps, *angles):
        dev = ps.device
        d = [t.detach().to(torch.float64).contiguous() for t in (ps, *angles)]
        N, n = d[0].shape[0], len(masses)
        out = torch.empty((N, 12, 4), dtype=torch.float64, device=dev)
        bs = torch.empty((N, 2), dtype=torch.float64, device=dev)
        mom = torch.empty((N, n + 2, 4), dtype=torch.float64, device=dev)
        w, x1, x2 = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3))
        prop = torch.empty((N, 3, 3), dtype=torch.float64, device=dev)
        mh = host_array(masses, ctypes.c_double)
        with device_guard(dev):
            check(lib().mf_sampling_partons_f64(ptr(d[0]), ctypes.c_int64(N), ctypes.c_int(n), mh, ctypes.c_double(e_collider),
                                                *[ptr(t) for t in d[1:]], ctypes.byref(cfg), ptr(out), ptr(bs), ptr(mom), ptr(w),
                                                ptr(x1), ptr(x2), ptr(prop), None, stream_ptr(dev)), "mf_sampling_partons_f64")
        ctx.cfg, ctx.masses, ctx.e_collider = cfg, masses, e_collider
        ctx.save_for_backward(*d, mom, w, x1, x2, prop, bs)
        ctx.in_meta = [(t.shape, t.dtype) for t in (ps, *angles)]
        ctx.mark_non_differentiable(mom, w, x1, x2)
        return out, bs, mom, w, x1, x2

    @staticmethod
    def backward(ctx, g_out, g_bs, _gm, _gw, _gx1, _gx2):
        *d, mom, w, x1, x2, prop, bs = ctx.saved_tensors
        dev = d[0].device
        N, n = d[0].shape[0], len(ctx.masses)
        g = (torch.zeros((N, 12, 4), dtype=torch.float64, device=dev) if g_out is None
             else g_out.detach().to(torch.float64).contiguous())
        gb = None if g_bs is None else g_bs.detach().to(torch.float64).contiguous()
        grads = [torch.empty_like(t) for t in d]
        mh = host_array(ctx.masses, ctypes.c_double)
        with device_guard(dev):
            check(lib().mf_sampling_partons_bwd_f64(ptr(d[0]), ctypes.c_int64(N), ctypes.c_int(n), mh,
                                                    ctypes.c_double(ctx.e_collider), *[ptr(t) for t in d[1:]],
                                                    ctypes.byref(ctx.cfg), ptr(mom), ptr(w), ptr(x1), ptr(x2), ptr(prop), ptr(bs),
                                                    ptr(g), ptr(gb), *[ptr(t) for t in grads], stream_ptr(dev)),
                  "mf_sampling_partons_bwd_f64")
        outs = [gr.reshape(shape).to(dt) if need else None
                for gr, (shape, dt), need in zip(grads, ctx.in_meta, ctx.needs_input_grad[3:])]
        return (None, None, None, *outs)
