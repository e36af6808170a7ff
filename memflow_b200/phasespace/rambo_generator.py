"""Host mirror of memflow/phasespace/rambo_generator.py: same class / method names, argument meaning
and return values, every batch method backed by ONE hand-written sm_100a kernel through the C ABI
(include/memflow_b200.h). The reference is adapted from NGoetz/TorchPS and follows S. Platzer,
arXiv:1308.2922 ("RAMBO on diet"), massive variant.

Deviations from the reference, all deliberate (DESIGN.md "Reference quirks"):
  * weights and incoming rows are computed in clean fp64. Set `ref_fp32_quirks=True` on the generator to
    reproduce the reference's float32 volume / weight chain / incoming rows (rambo_generator.py:132,489-507,533).
  * data-dependent raises (NaN scan :191-199, CM check :630-635) cost host syncs in the reference. Here
    they set bits in `generator.last_status` (device int32); `sync_errors=True` restores the raises.
  * CUDA tensors only: there is no CPU fallback.
  * autograd: K1 and K2 run as registered ops (torch.ops.memflow_b200.rambo_forward / rambo_inverse, memflow_b200/ops.py)
    whose autograd formulas call the backward kernels: the forward map has the reference graph's gradient semantics; the
    inverse map (`getPSpoint_batch`, `get_PS`) is differentiable in `ps` (not in `prob`).
"""
import ctypes
import math

import torch

from .. import _lib
from .. import ops  # noqa: F401  (registers torch.ops.memflow_b200.*)
from .._lib import check, device_guard, host_array, lib, MemflowB200Error, needs_dispatcher, plain_eager, ptr, require_cuda, stream_ptr

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 0.0  # rambo_generator.py:14 (phasespace.py:31 uses 1e-5)

MF_FLAG_REF_FP32_QUIRKS = 1
MF_FLAG_LAB_FRAME = 2
MF_FLAG_CUTS = 4
MF_FLAG_X1X2_DIRECT = 8
MF_FLAG_NO_CM_CHECK = 16
MF_FLAG_ZERO_STATUS = 32
MF_ST_NAN_INPUT = 1
MF_ST_NOT_CM = 2


class PhaseSpaceGeneratorError(Exception):
    pass


def _inverse_vjp(mom, x1, x2, g_ps, n, order, masses_host, target_host, permute_target, e_collider, flags, want_x):
    """mf_rambo_inverse_bwd_f64 on contiguous fp64 tensors; returns (g_momenta [N,n,4], g_x1, g_x2)."""
    dev = mom.device
    N = mom.shape[0]
    g_mom = torch.empty((N, n, 4), dtype=torch.float64, device=dev)
    g_x1 = torch.empty(N, dtype=torch.float64, device=dev) if want_x else None
    g_x2 = torch.empty(N, dtype=torch.float64, device=dev) if want_x else None
    tm = host_array(target_host, ctypes.c_double) if target_host is not None else None
    with device_guard(dev):
        rc = lib().mf_rambo_inverse_bwd_f64(
            ptr(mom), ctypes.c_int64(mom.stride(0)), ptr(x1), ptr(x2), ptr(g_ps), ctypes.c_int64(N), ctypes.c_int(n),
            host_array(order, ctypes.c_int32), host_array(masses_host, ctypes.c_double), tm,
            ctypes.c_int(int(permute_target)), ctypes.c_double(float(e_collider)), ctypes.c_uint32(flags), ptr(g_mom),
            ptr(g_x1), ptr(g_x2), stream_ptr(dev))
    check(rc, "mf_rambo_inverse_bwd_f64")
    return g_mom, g_x1, g_x2


class VirtualPhaseSpaceGenerator(object):
    """rambo_generator.py:29-83."""

    def __init__(self, initial_masses, final_masses, collider_energy, pdf=None, pdf_active=False,
                 uniform_x1x2=True, lhapdf_dir=None, dev=None):
        # `last_status` (device int32, bit-or of MF_ST_*) belongs to the most recent call made through THIS generator object:
        # use one generator per stream / thread when calls overlap
        if dev is None:
            dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
        self.dev = torch.device(dev)
        self.initial_masses = initial_masses
        self.masses_t = torch.as_tensor(final_masses).clone().to(self.dev)
        self.tot_final_state_masses = torch.sum(self.masses_t)
        self.n_initial = len(initial_masses)
        self.n_final = len(final_masses)
        self.collider_energy = collider_energy
        self.pdf = pdf
        self.pdf_active = pdf_active
        self.uniform_x1x2 = uniform_x1x2
        # host copy of the masses: travels with every launch (kernel parameter space), never re-read from HBM
        self._masses_host = [float(m) for m in self.masses_t.detach().cpu().double().tolist()]

    def generateKinematics(self, E_cm, random_variables):
        raise NotImplementedError

    @property
    def nDimPhaseSpace(self):
        if self.n_final == 1:
            return 0
        return 3 * self.n_final - 4


class FlatInvertiblePhasespace(VirtualPhaseSpaceGenerator):
    """rambo_generator.py:86-736, B200-native."""

    epsilon_border = 1e-10
    absolute_Ecm_min = 1.0

    def __init__(self, *args, **opts):
        self.ref_fp32_quirks = bool(opts.pop("ref_fp32_quirks", False))
        self.sync_errors = bool(opts.pop("sync_errors", False))
        self.compute_dtype = opts.pop("compute_dtype", torch.float64)
        super().__init__(*args, **opts)
        if self.n_initial == 1:
            raise PhaseSpaceGeneratorError("This basic generator does not support decay topologies.")
        if self.n_initial > 2:
            raise PhaseSpaceGeneratorError("This basic generator does not support more than 2 incoming particles.")
        self.last_status = None

    # ------------------------------------------------------------------ small closed-form helpers
    @staticmethod
    def get_flatWeights(E_cm, n):
        """Massless n-body volume (2pi)^(4-3n) (pi/2)^(n-1) E^(2n-4) / ((n-1)!(n-2)!)  (:111-140).
        Tensor inputs stay in their own dtype (the reference silently casts them to float32, :132)."""
        if n == 1:
            return 1.0
        c = math.pow(2 * math.pi, 4 - 3 * n) * math.pow(math.pi / 2.0, n - 1) \
            / (math.factorial(n - 1) * math.factorial(n - 2))
        if not torch.is_tensor(E_cm):
            return (math.pow(2 * math.pi, 4 - 3 * n) * math.pow((math.pi / 2.0), n - 1)
                    * (math.pow((E_cm ** 2), n - 2) / (math.factorial(n - 1) * math.factorial(n - 2))))
        return c * (E_cm ** 2) ** (n - 2)

    def massless_map(self, x, exp):
        """:142-143."""
        return (x ** exp) * ((exp + 1) - exp * x)

    @staticmethod
    def rho(M, N, m):
        """:145-149."""
        Msqr = M ** 2
        return ((Msqr - (N + m) ** 2) * (Msqr - (N - m) ** 2)) ** 0.5 / (8.0 * Msqr)

    def bisect_vec_batch(self, v_t, target=1.0e-16, maxLevel=600):
        """Solve v = u^e ((e+1) - e u), e = n-2-column, for u (:400-446). The reference bisects the whole
        batch for up to 180 levels; the kernel uses a closed form (e=1) / an fp32-seeded Newton + chord step in fp64.
        `target` / `maxLevel` are accepted for signature compatibility and ignored."""
        if v_t.size(1) == 0:
            return None
        require_cuda(v_t, "v_t")
        if v_t.shape[1] != self.n_final - 2:
            raise AssertionError("bisect_vec_batch expects [N, n_final-2]")
        v = v_t.detach().to(torch.float64).contiguous()
        u = torch.empty_like(v)
        with device_guard(v.device):
            check(lib().mf_massvar_solve_f64(ptr(v), ctypes.c_int64(v.shape[0]), ctypes.c_int(self.n_final), ptr(u),
                                             stream_ptr(v.device)), "mf_massvar_solve_f64")
        return u

    def generateIntermediatesMassless_batch(self, M_t, E_cm, random_variables):
        """Intermediate masses of a MASSLESS final state, on their own (:448-473): K_i = K_0 prod_{j<i} u_j with
        u = bisect_vec_batch(random_variables[:, :n-2]) (the mass-variable kernel, mf_massvar_solve_f64), and the massless
        volume get_flatWeights(E_cm, n). Returns (weight [N], K [N, n-1]). K1 has this fused in; the method exists for
        callers of the reference's API. The volume stays float64 (the reference's tensor branch is float32, :132)."""
        require_cuda(random_variables, "random_variables")
        n = self.n_final
        u = self.bisect_vec_batch(random_variables[:, : n - 2])
        N = random_variables.shape[0]
        ones = torch.ones(N, 1, dtype=torch.float64, device=M_t.device)
        prod_u = torch.cumprod(torch.cat((ones, u), dim=1) if u is not None else ones, dim=1)
        K = M_t[:, 0].unsqueeze(1).to(torch.float64) * prod_u
        if not torch.is_tensor(E_cm) or E_cm.dim() == 0:
            w = torch.full((N,), float(self.get_flatWeights(float(E_cm), n)), dtype=torch.float64, device=random_variables.device)
        else:
            w = self.get_flatWeights(E_cm.to(torch.float64), n)
        return w, K

    def generateIntermediatesMassive_batch(self, M, E_cm, random_variables):
        """Intermediate masses of the MASSIVE final state and the massive / massless volume ratio (:475-509):
        M_i = K_i + sum_{j>=i} m_j, weight = V_n(E) 8 rho(M_{n-2}, m_{n-1}, m_{n-2})
        prod_{i<=n-3} [rho(M_i, M_{i+1}, m_i) / rho(K_i, K_{i+1}, 0)] (M_{i+1} / K_{i+1}) (K_0 / M_0)^(2n-4).
        Returns (weight [N], M [N, n-1]); float64 throughout (the reference's in-place chain runs in float32)."""
        n = self.n_final
        masses = self.masses_t.to(device=M.device, dtype=torch.float64)
        Mc = M.clone().to(torch.float64)
        Mc[:, 0] -= masses.sum()
        weight, K = self.generateIntermediatesMassless_batch(Mc, E_cm, random_variables)
        weight = weight.clone()
        msum = torch.flip(torch.cumsum(torch.flip(masses, (-1,)), -1), (-1,))
        Mi = K + msum[:-1]
        weight = weight * 8.0 * self.rho(Mi[:, n - 2], masses[n - 1], masses[n - 2])
        weight = weight * torch.prod((self.rho(Mi[:, : n - 2], Mi[:, 1:], masses[: n - 2]) / self.rho(K[:, : n - 2], K[:, 1:], 0.0))
                                     * (Mi[:, 1: n - 1] / K[:, 1: n - 1]), -1)
        weight = weight * torch.pow(K[:, 0] / Mi[:, 0], 2 * n - 4)
        return weight, Mi

    # ------------------------------------------------------------------ K1
    def generateKinematics_batch(self, random_variables_full, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1,
                                 pdgs=[0, 0], requires_grad=False, lab_frame=False, return_logweight=False):
        """u [N, 3n-2] -> (momenta [N, n+2, 4] in the partonic CM frame, weight [N], x1 [N], x2 [N]), all
        float64 on the input's device (:168-398). Extras: `lab_frame=True` returns lab-frame momenta
        (utils.py:180-202 fused in), `return_logweight=True` appends ln(weight)."""
        if not self.pdf_active:
            # reference: E_cm is never assigned on this branch (:264-266) -> NameError
            raise PhaseSpaceGeneratorError("pdf_active=False is not supported (it is broken in the reference too)")
        require_cuda(random_variables_full, "random_variables_full")
        assert random_variables_full.shape[1] - 2 == self.nDimPhaseSpace
        cuts_on = (pT_mincut > 0) or (delR_mincut > 0) or (rap_maxcut > 0)
        dt = self.compute_dtype
        if dt not in (torch.float64, torch.float32):
            raise MemflowB200Error("compute_dtype must be torch.float64 or torch.float32")
        needs_grad = torch.is_grad_enabled() and random_variables_full.requires_grad
        if needs_grad and (cuts_on or lab_frame or dt != torch.float64 or self.ref_fp32_quirks):
            # differentiable path = the reference's requires_grad=True (scripts/run_model_labframe_sampling.py:432)
            raise MemflowB200Error("autograd through the phase-space map supports the plain fp64 CM-frame output only "
                                   "(no cuts, no lab_frame, no fp32 compute / quirks)")
        u = (random_variables_full if needs_grad else random_variables_full.detach()).to(dt)
        flags = 0
        if self.ref_fp32_quirks:
            flags |= MF_FLAG_REF_FP32_QUIRKS
        if not self.uniform_x1x2:
            flags |= MF_FLAG_X1X2_DIRECT
        if lab_frame:
            flags |= MF_FLAG_LAB_FRAME
        if cuts_on:   # the reference evaluates the cuts always; with the default -1 thresholds they cannot fire
            flags |= MF_FLAG_CUTS
        # ONE launch through the registered op (memflow_b200/ops.py): K1, its status word, autograd via the backward kernel
        call = ops.rambo_forward_cuda if not needs_dispatcher(u) else (ops.RamboForwardEager.apply if plain_eager(u) else torch.ops.memflow_b200.rambo_forward)
        momenta, weight, x1, x2, logw, status = call(
            u, self._masses_host, float(self.collider_energy),
            [float(pT_mincut), float(delR_mincut), float(rap_maxcut)] if cuts_on else [], flags, bool(return_logweight))
        self.last_status = status
        if self.sync_errors and int(status.item()) & MF_ST_NAN_INPUT:
            raise PhaseSpaceGeneratorError("Some of the random variables passed to the phase-space generator are NaN")
        if dt != torch.float64:  # reference contract: outputs are float64
            momenta, weight, x1, x2 = momenta.double(), weight.double(), x1.double(), x2.double()
            logw = logw.double()
        if return_logweight:
            return momenta, weight, x1, x2, logw
        return momenta, weight, x1, x2

    # ------------------------------------------------------------------ K2
    def getPSpoint_batch(self, momenta_batch, x1, x2, order=None, target_mass=None, ensure_CM=True,
                         ensure_onShell=True, return_logprob=False):
        """Final-state momenta [N, n, 4] (CM frame) + x1, x2 -> (ps [N, 3n-2], prob [N]) (:615-736).
        `order=None` means the identity permutation for any n (the reference's default [0,1,2,3] only
        fits n=4). `target_mass` ([n] or [1,n]) is used when ensure_onShell=False (:649-652)."""
        require_cuda(momenta_batch, "momenta_batch")
        n = self.n_final
        if momenta_batch.dim() != 3 or momenta_batch.shape[1] != n or momenta_batch.shape[2] != 4:
            raise AssertionError("momenta_batch must be [N, %d, 4]" % n)
        order = list(range(n)) if order is None else [int(o) for o in order]
        if len(order) != n:
            raise AssertionError("order must have n_final=%d entries" % n)
        target_host = None
        if not ensure_onShell:
            if target_mass is None:
                raise AssertionError("ensure_onShell=False needs target_mass")
            target_host = torch.as_tensor(target_mass).detach().cpu().double().reshape(-1).tolist()
            if len(target_host) != n:
                raise AssertionError("target_mass must hold n_final=%d masses ([n] or [1,n])" % n)
        dt = self.compute_dtype
        x1 = torch.as_tensor(x1, device=momenta_batch.device)
        x2 = torch.as_tensor(x2, device=momenta_batch.device)
        needs_grad = torch.is_grad_enabled() and (momenta_batch.requires_grad or x1.requires_grad or x2.requires_grad)
        if needs_grad and (dt != torch.float64 or self.ref_fp32_quirks):
            raise MemflowB200Error("autograd through the inverse map supports fp64 compute without quirks only")
        mom = momenta_batch if needs_grad else momenta_batch.detach()
        if mom.dtype != dt:
            mom = mom.to(dt)
        flags = 0
        if self.ref_fp32_quirks:
            flags |= MF_FLAG_REF_FP32_QUIRKS
        if not self.uniform_x1x2:
            flags |= MF_FLAG_X1X2_DIRECT
        if not ensure_CM:
            flags |= MF_FLAG_NO_CM_CHECK
        # ONE launch through the registered op: a strided [N, n, 4] view (e.g. forward_output[:, 2:]) is read in place;
        # `ps` is differentiable w.r.t. the momenta and x1, x2 (backward kernel), `prob` carries no gradient
        call = ops.rambo_inverse_cuda if not needs_dispatcher(mom, x1, x2) else (ops.RamboInverseEager.apply if plain_eager(mom, x1, x2) else torch.ops.memflow_b200.rambo_inverse)
        ps, prob, logp, status = call(
            mom, x1 if needs_grad else x1.detach(), x2 if needs_grad else x2.detach(), order, self._masses_host, target_host,
            float(self.collider_energy), flags, bool(return_logprob))
        if needs_grad:
            prob, logp = prob.detach(), logp.detach()
        self.last_status = status
        if ensure_CM and self.sync_errors and int(status.item()) & MF_ST_NOT_CM:
            raise Exception("Momenta batch not in the CM, failing to convert back to PS point")
        if dt != torch.float64:
            ps, prob, logp = ps.double(), prob.double(), logp.double()
        if return_logprob:
            return ps, prob, logp
        return ps, prob

    def setInitialStateMomenta_batch(self, output_momenta, E_cm):
        """Rows 0,1 <- (E/2, 0, 0, +-E/2) for massless beams (:511-551); in place, like the reference."""
        if self.n_initial != 2:
            raise PhaseSpaceGeneratorError("This PS generator only supports 2 initial states")
        if not (self.initial_masses[0] == 0.0 or self.initial_masses[1] == 0.0):
            raise PhaseSpaceGeneratorError("massive beams are unreachable from PhaseSpace and not supported")
        h = (E_cm if torch.is_tensor(E_cm) else torch.as_tensor(E_cm, dtype=output_momenta.dtype,
                                                                device=output_momenta.device)) / 2
        output_momenta[:, 0, :] = 0
        output_momenta[:, 1, :] = 0
        output_momenta[:, 0, 0] = h
        output_momenta[:, 0, 3] = h
        output_momenta[:, 1, 0] = h
        output_momenta[:, 1, 3] = -h
