"""Importance-sampling weight, Monte-Carlo estimator and the fused sampling chain (K5 + chain kernel).

Reference: notebooks/TestFlows_ImportanceSampling.ipynb cells [50-54] (weights, I; sigma as in cell [54]), mask of cells
[35,43]; notebooks/TestMadgraphChain.ipynb [19-24]. The reference has no module for this -- only notebook
cells -- so the function names here are ours; the arithmetic is the notebook's.
"""
import ctypes
import math

import torch

from . import ops  # noqa: F401  (registers torch.ops.memflow_b200.*)
from ._lib import check, device_guard, host_array, lib, ptr, reduce_workspace, require_cuda, stream_ptr


def new_accumulator(device):
    """Device double[3] = (sum w, sum w^2, n_valid)."""
    return torch.zeros(3, dtype=torch.float64, device=device)


def accumulate_weights(target, rambo_jac, logq, acc=None):
    """acc += (sum w, sum w^2, n_valid) with w = target * rambo_jac / exp(logq); valid = target > 0, rambo_jac not
    NaN (notebook cells [35,43]) and w finite (one bad log q must not turn a whole run's sums into NaN).
    `target=None` means 1. Returns acc (device tensor; no host sync)."""
    require_cuda(rambo_jac, "rambo_jac")
    if acc is None:
        acc = new_accumulator(rambo_jac.device)
    torch.ops.memflow_b200.is_reduce_(None if target is None else target.detach(), rambo_jac.detach(), logq.detach(), acc)
    return acc


def mc_estimate(acc, n_total, variant="N+1"):
    """I = sum w / N, sigma = sqrt((sum w^2 / N - I^2) / (N + 1))  (cell [54]; variant "N-1" = cell [35]).
    `acc` may be a device tensor (one host sync here) or a sequence of 3 floats."""
    a = acc.detach().cpu().tolist() if torch.is_tensor(acc) else list(acc)
    n = float(n_total)
    I = a[0] / n
    den = n + 1.0 if variant == "N+1" else n - 1.0
    return I, math.sqrt(max(a[1] / n - I * I, 0.0) / den), int(a[2])


class ImportanceSampler:
    """Sampling chain: Philox base sample -> T spline inverses (launch A) -> PhaseSpace map -> weight ->
    accumulators (launch B), one C-ABI call per chunk (TestFlows_ImportanceSampling.ipynb [45,50-54]; ME = pdf = 1).

    phi: float32 [T, N, F, 3K-1] spline parameters (the hyper-network output, given), F = 3 n_final - 2.
    """

    def __init__(self, final_masses, collider_energy, bound=1.0, slope=1e-3, seed=0, device=None):
        self.masses = [float(m) for m in torch.as_tensor(final_masses).double().cpu().tolist()]
        self.n_final = len(self.masses)
        self.collider_energy = float(collider_energy)
        self.bound = float(bound)
        self.slope = slope
        self.seed = int(seed)
        self.device = torch.device(device if device is not None else "cuda:0")

    def _check_phi(self, phi):
        require_cuda(phi, "phi")
        if phi.dtype not in (torch.float32, torch.float64) or phi.dim() != 4:
            raise AssertionError("phi must be float32 or float64 [T, N, F, 3K-1]")
        T, N, F, PW = phi.shape
        if F != 3 * self.n_final - 2 or (PW + 1) % 3 != 0:
            raise AssertionError("phi shape does not match n_final (F must be 3n-2, last dim 3K-1)")
        return T, N, F, (PW + 1) // 3

    def sample_flow(self, phi, event_offset=0, u=None, logq=None):
        """Stage A alone (mf_flow_sample_f32 / _f64 by the dtype of phi): u [N, F] float64 in [0,1] and log q(u) [N] of the
        flow's samples."""
        T, N, F, K = self._check_phi(phi)
        if u is None and logq is None:   # the registered op (memflow_b200/ops.py) allocates the two outputs
            return torch.ops.memflow_b200.flow_sample(phi.detach(), self.seed, int(event_offset), self.bound,
                                                      float(self.slope) if self.slope is not None else 0.0)
        phi = phi.contiguous()
        dev = phi.device
        if u is None:
            u = torch.empty((N, F), dtype=torch.float64, device=dev)
        if logq is None:
            logq = torch.empty(N, dtype=torch.float64, device=dev)
        f64 = phi.dtype == torch.float64   # MEMFlow's configs run the flows in float64: same chain on the fp64 spline kernel
        fn = lib().mf_flow_sample_f64 if f64 else lib().mf_flow_sample_f32
        ct = ctypes.c_double if f64 else ctypes.c_float
        with device_guard(dev):
            rc = fn(ptr(phi), ctypes.c_int64(N), ctypes.c_int(F), ctypes.c_int(T), ctypes.c_int(K), ct(self.bound),
                    ct(self.slope if self.slope is not None else 0.0), ctypes.c_uint64(self.seed),
                    ctypes.c_uint64(int(event_offset)), ptr(u), ptr(logq), stream_ptr(dev))
        check(rc, "mf_flow_sample_f64" if f64 else "mf_flow_sample_f32")
        return u, logq

    def weigh(self, u, logq, acc=None, out=None):
        """Stage B alone (mf_rambo_is_f64): PhaseSpace map of u + weight + accumulation into acc. `out` (optional dict
        with momenta / weight / x1 / x2 tensors) receives the per-event values."""
        require_cuda(u, "u")
        dev = u.device
        if acc is None:
            acc = new_accumulator(dev)
        if not out:
            torch.ops.memflow_b200.rambo_is_(u, logq, self.masses, self.collider_energy, acc)
            return acc
        ws = reduce_workspace(dev)
        with device_guard(dev):
            rc = lib().mf_rambo_is_f64(ptr(u), ptr(logq), ctypes.c_int64(u.shape[0]), ctypes.c_int(self.n_final),
                                       host_array(self.masses, ctypes.c_double), ctypes.c_double(self.collider_energy),
                                       ctypes.c_uint32(0), ptr(acc), ptr(ws), ptr(out.get("momenta")), ptr(out.get("weight")),
                                       ptr(out.get("x1")), ptr(out.get("x2")), stream_ptr(dev))
        check(rc, "mf_rambo_is_f64")
        return acc

    def run_chunk(self, phi, event_offset=0, acc=None, debug_outputs=False):
        T, N, F, K = self._check_phi(phi)
        n = self.n_final
        phi = phi.contiguous()
        dev = phi.device
        if acc is None:
            acc = new_accumulator(dev)
        ws = reduce_workspace(dev)
        out = {}
        if debug_outputs:
            out = dict(u=torch.empty((N, F), dtype=torch.float64, device=dev),
                       logq=torch.empty(N, dtype=torch.float64, device=dev),
                       momenta=torch.empty((N, n + 2, 4), dtype=torch.float64, device=dev),
                       weight=torch.empty(N, dtype=torch.float64, device=dev),
                       x1=torch.empty(N, dtype=torch.float64, device=dev),
                       x2=torch.empty(N, dtype=torch.float64, device=dev))
        if phi.dtype == torch.float64:   # float64 flow parameters: stage A on the fp64 spline kernel, then stage B
            u, logq = self.sample_flow(phi, event_offset=event_offset, u=out.get("u"), logq=out.get("logq"))
            self.weigh(u, logq, acc=acc, out=out)
            return (acc, out) if debug_outputs else acc
        with device_guard(dev):
            rc = lib().mf_is_chain_f64(
                ptr(phi), ctypes.c_int64(N), ctypes.c_int(n), ctypes.c_int(T), ctypes.c_int(K),
                ctypes.c_float(self.bound), ctypes.c_float(self.slope if self.slope is not None else 0.0),
                ctypes.c_uint64(self.seed), ctypes.c_uint64(int(event_offset)),
                host_array(self.masses, ctypes.c_double), ctypes.c_double(self.collider_energy), ctypes.c_uint32(0),
                ptr(acc), ptr(ws), ptr(out.get("u")), ptr(out.get("logq")), ptr(out.get("momenta")),
                ptr(out.get("weight")), ptr(out.get("x1")), ptr(out.get("x2")), stream_ptr(dev))
        check(rc, "mf_is_chain_f64")
        return (acc, out) if debug_outputs else acc
