"""The phase-space tensors MEMFlow caches on disk for every training.

Reference: `Dataset_Parton_Level.get_PS_intermediateParticles(_onShell)`,
memflow/read_data/Dataset_Parton_Level.py:410-463. Only the tensor computation is mirrored here -- parquet /
awkward reading and `torch.save` file handling stay with the caller (I/O is out of scope).
The inverse map runs on the K2 kernel (2M-event files in milliseconds); the returned dict uses the reference's
on-disk names as keys.
"""
import torch

from ..phasespace.phasespace import PhaseSpace
from ..phasespace.rambo_generator import MF_ST_NOT_CM

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 1e-3   # memflow/read_data/Dataset_Parton_Level.py:19 (the phase-space module's own default is 1e-5)


def intermediateParticles_cartesian_onShell(momenta_HttISR, masses=None):
    """`get_intermediateParticles_cartesian_onShell` (:400-408): energies recomputed from the 3-momenta and the nominal
    masses (H, t, tbar, ISR); the cached tensor `H_thad_tlep_ISR_cartesian_onShell`. Elementwise torch (offline path)."""
    if masses is None:
        masses = torch.tensor([M_HIGGS, M_TOP, M_TOP, M_GLUON], dtype=torch.float64)
    m = torch.as_tensor(masses, dtype=momenta_HttISR.dtype, device=momenta_HttISR.device)
    out = momenta_HttISR.clone()
    out[:, :, 0] = torch.sqrt(out[:, :, 1] ** 2 + out[:, :, 2] ** 2 + out[:, :, 3] ** 2 + m ** 2)
    return out


def phasespace_products(momenta_HttISR, boost, on_shell=True, E_CM=13000, masses=None, pdgs=(25, 6, -6, 21),
                        ref_fp32_quirks=False):
    """momenta_HttISR [N, 4, 4] (H, t, tbar, ISR; CM frame, E,px,py,pz), boost [N, 1, 4] or [N, 4] -> dict.

    on_shell=True  (:410-446): minimal-energy fix of the boost (:418-420), x1/x2 clamped to [0,1], plus
                   logit(ps, eps=5e-5), its nan-mean / 5*nan-std scaling and the scaled logits.
    on_shell=False (:448-463): plain x1/x2, ps and det-jacobian only.
    ref_fp32_quirks=True reproduces the float32 weight chain of the literal reference in the det-jacobian (relative 1e-7;
    DESIGN.md section 2); the default is clean float64.
    """
    dev = momenta_HttISR.device
    if masses is None:
        masses = torch.tensor([M_HIGGS, M_TOP, M_TOP, M_GLUON], dtype=torch.float64)
    masses = torch.as_tensor(masses, dtype=torch.float64)
    ps_gen = PhaseSpace(E_CM, [21, 21], list(pdgs), final_masses=masses, dev=dev, ref_fp32_quirks=ref_fp32_quirks)
    b = boost.reshape(boost.shape[0], -1, 4)[:, 0, :].to(torch.float64).clone()
    suffix = "_onShell" if on_shell else ""
    if on_shell:
        msum = float(masses.sum())
        e, pz = b[:, 0], b[:, 3]
        b[:, 0] = torch.where(torch.sqrt(e ** 2 - pz ** 2) < msum, torch.sqrt(pz ** 2 + msum ** 2 + 1e-3), e)
        x1 = torch.clamp((b[:, 0] + b[:, 3]) / E_CM, min=0.0, max=1.0)
        x2 = torch.clamp((b[:, 0] - b[:, 3]) / E_CM, min=0.0, max=1.0)
    else:
        x1 = (b[:, 0] + b[:, 3]) / E_CM
        x2 = (b[:, 0] - b[:, 3]) / E_CM
    ps, detjinv = ps_gen.get_ps_from_momenta(momenta_HttISR, x1, x2)
    # The reference raises when the batch is not in the CM frame (rambo_generator.py:630-635); the kernel only sets a status
    # bit. This is the offline caching path: one host sync is harmless, and lab-frame or mis-ordered momenta must not end up
    # silently in cached ps / detjac tensors.
    if int(ps_gen.generator.last_status.item()) & MF_ST_NOT_CM:
        raise Exception("Momenta batch not in the CM, failing to convert back to PS point")
    out = {"phasespace_intermediateParticles" + suffix: ps,
           "phasespace_rambo_detjacobian" + suffix: detjinv}
    if on_shell:
        logit_ps = torch.logit(ps, eps=5e-5)
        mean = logit_ps.nanmean(0)
        scale = torch.sqrt(torch.nanmean(logit_ps ** 2, 0) - mean ** 2) * 5
        out["phasespace_intermediateParticles_onShell_logit"] = logit_ps
        out["phasespace_intermediateParticles_onShell_logit_scaled"] = (logit_ps - mean) / scale
        out["mean_std_phasespace_intermediateParticles_onShell_logit"] = (mean, scale)
    return out
