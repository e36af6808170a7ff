"""`Compute_ParticlesTensor.get_PS` of memflow/unfolding_flow/utils.py:1056-1152 on the K2 kernel.

Only this method of the reference's 1100-line glue class is on the hot path (it is called every training
step, memflow/unfolding_flow/unfolding_flow.py:163); the rest of the class (pt/eta/phi conversions, decay
products, regression scalings) is model glue and out of scope.
"""
import ctypes

import torch

from .._lib import check, host_array, lib, ptr, require_cuda, stream_ptr

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 1e-5  # memflow/unfolding_flow/utils.py:10


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
        dev = data_HttISR.device
        n = data_HttISR.shape[1]
        if target_mass is None:
            target_mass = torch.tensor([[M_HIGGS, M_TOP, M_TOP, M_GLUON]], dtype=torch.float64)
        tm = torch.as_tensor(target_mass).detach().cpu().double().reshape(-1).tolist()
        if len(tm) != n or len(order) != n:
            raise AssertionError("target_mass / order must have %d entries" % n)
        mom = data_HttISR.detach().to(torch.float64)
        if not (mom.stride(2) == 1 and mom.stride(1) == 4 and mom.stride(0) % 4 == 0 and mom.stride(0) >= 4 * n
                and mom.data_ptr() % 32 == 0):
            mom = mom.contiguous()
        boost = boost_reco.detach().to(torch.float64).contiguous()
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
        with torch.cuda.device(dev):
            rc = lib().mf_get_ps_f64(ptr(mom), ctypes.c_int64(mom.stride(0)), ptr(boost), ctypes.c_int64(N), ctypes.c_int(n),
                                     host_array(order, ctypes.c_int32), host_array(tm, ctypes.c_double),
                                     ctypes.c_double(float(E_CM)), ptr(ps), ptr(jac0), ptr(mask),
                                     ctypes.c_double(float(logit_eps) if logit_eps is not None else 0.0), lm, ls, ptr(lps),
                                     stream_ptr(dev))
        check(rc, "mf_get_ps_f64")
        if logit_eps is not None:
            return ps, jac0, mask.bool(), lps
        return ps, jac0, mask.bool()
