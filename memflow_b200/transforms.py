"""zuko-style monotonic rational-quadratic spline transform backed by the K3/K4 CUDA kernels.

`MonotonicRQSTransform(widths, heights, derivatives, bound=5.0, slope=1e-3)` has the constructor and the
`torch.distributions.Transform` surface of `zuko.transforms.MonotonicRQSTransform` (third-party dependency
of the reference, constructed at memflow/unfolding_flow/unfolding_flow.py:88-97 and
memflow/unfolding_flow/custom_spline_flow_ps.py:16-33): `t(x)`, `t.inv(y)`, `t.call_and_ladj(x)`,
`t.log_abs_det_jacobian(x, y)`. The functional forms `rqs_forward / rqs_inverse` take the hyper-network
output row `phi[B, F, 3K-1]` directly and also return the per-event sum of log-dets.

Every call goes through the registered op `torch.ops.memflow_b200.rqs` (memflow_b200/ops.py): differentiable in phi and in
the input through the backward kernels (mf_rqs_{forward,inverse}_bwd), traceable with fake tensors. CUDA tensors
only.
"""
import ctypes

import torch
from torch.distributions import Transform, constraints

from . import ops
from ._lib import check, device_guard, lib, MemflowB200Error, needs_dispatcher, plain_eager, ptr, require_cuda, stream_ptr


def _run(direction, phi, v, bound, slope, want_ladj=True, want_sum=False, want_bin=False):
    """One call of the registered op memflow_b200::rqs (memflow_b200/ops.py: schema, fake implementation, autograd through
    the backward kernels). CUDA tensors only."""
    require_cuda(phi, "phi")
    require_cuda(v, "x")
    if phi.dtype not in (torch.float32, torch.float64):
        raise MemflowB200Error("spline kernels support float32 / float64, got %s" % phi.dtype)
    outputs = (ops.RQS_LADJ if want_ladj else 0) | (ops.RQS_SUM if want_sum else 0) | (ops.RQS_BINS if want_bin else 0)
    # plain eager inference goes straight to the C ABI; plain eager autograd through a directly-applied Function;
    # torch.compile / fake tensors / functorch transforms through the dispatcher (the registered op)
    if not needs_dispatcher(phi, v):
        call = ops.rqs_cuda
    elif plain_eager(phi, v):
        call = ops.RqsEager.apply
    else:
        call = torch.ops.memflow_b200.rqs
    out, ladj, lsum, bins = call(phi, v, float(bound), float(slope) if slope is not None else 0.0, direction == "inverse", outputs)
    return out, (ladj if want_ladj else None), (lsum if want_sum else None), (bins if want_bin else None)


def rqs_forward(phi, x, bound=5.0, slope=1e-3, return_bins=False):
    """y, ladj [B,F], ladj_sum [B] = spline(x) with parameters phi[B, F, 3K-1] = [widths|heights|derivatives]."""
    y, ladj, lsum, bins = _run("forward", phi, x, bound, slope, True, True, return_bins)
    return (y, ladj, lsum, bins) if return_bins else (y, ladj, lsum)


def rqs_inverse(phi, y, bound=5.0, slope=1e-3, return_bins=False):
    """x, ladj [B,F], ladj_sum [B]: x = spline^-1(y); ladj is the FORWARD log|dy/dx| at x (sample + log q in one pass)."""
    x, ladj, lsum, bins = _run("inverse", phi, y, bound, slope, True, True, return_bins)
    return (x, ladj, lsum, bins) if return_bins else (x, ladj, lsum)


def searchsorted(knots, v):
    """torch.searchsorted(knots[..., :], v[..., None], right=False) on the GPU kernel's bin search (bit-exact)."""
    require_cuda(knots, "knots")
    knots = knots.detach().contiguous()
    v = v.detach().to(knots.dtype).contiguous()
    length = knots.shape[-1]
    BF = v.numel()
    idx = torch.empty(v.shape, dtype=torch.int32, device=knots.device)
    fn = lib().mf_searchsorted_f32 if knots.dtype == torch.float32 else lib().mf_searchsorted_f64
    with device_guard(knots.device):
        check(fn(ptr(knots), ptr(v), ctypes.c_int64(BF), ctypes.c_int(length), ptr(idx), stream_ptr(knots.device)),
              "mf_searchsorted")
    return idx


def _pack(widths, heights, derivatives):
    """Return phi[rows, 1, 3K-1]. Zero-copy when the three tensors are the split views of one hyper-network
    output row (what zuko's MAF passes); otherwise one torch.cat."""
    K = widths.shape[-1]
    if heights.shape[-1] != K or derivatives.shape[-1] != K - 1:
        raise AssertionError("shapes must be widths[*,K], heights[*,K], derivatives[*,K-1]")
    lead = torch.broadcast_shapes(widths.shape[:-1], heights.shape[:-1], derivatives.shape[:-1])
    PW = 3 * K - 1
    es = widths.element_size()
    needs_grad = torch.is_grad_enabled() and (widths.requires_grad or heights.requires_grad or derivatives.requires_grad)
    same = (widths.dtype == heights.dtype == derivatives.dtype and widths.shape[:-1] == lead
            and heights.shape[:-1] == lead and derivatives.shape[:-1] == lead
            and widths.stride() == heights.stride() == derivatives.stride()
            and widths.stride(-1) == 1
            and heights.data_ptr() == widths.data_ptr() + K * es
            and derivatives.data_ptr() == widths.data_ptr() + 2 * K * es)
    if same and not needs_grad:
        # rows of a CONSTANT pitch P >= PW (one feature's rows out of a [.., F, 3K-1] parameter tensor, as an autoregressive
        # sampling pass selects them): read in place through the strided entry points, no gather copy
        sizes = [(size, stride) for size, stride in zip(lead, widths.stride()[:-1]) if size != 1]
        if sizes:
            P = sizes[-1][1]
            exp, okp = P, P >= PW
            for size, stride in reversed(sizes):
                if stride != exp:
                    okp = False
                    break
                exp *= size
            if okp and P != PW:
                rows = 1
                for s_ in lead:
                    rows *= s_
                return torch.as_strided(widths, (rows, 1, PW), (P, PW, 1)), lead
    if same:
        # contiguous rows of pitch PW?
        exp, ok = PW, True
        for size, stride in zip(reversed(lead), reversed(widths.stride()[:-1])):
            if size != 1 and stride != exp:
                ok = False
                break
            exp *= size
        if ok:
            rows = 1
            for s in lead:
                rows *= s
            if not needs_grad:
                return torch.as_strided(widths, (rows, 1, PW), (PW, PW, 1)), lead
            # under autograd the three tensors are views of ONE hyper-network output (zuko: unflatten + split): a view
            # of that common base is the same memory and differentiable, so no concatenation copy on the way in and no
            # split of the gradient on the way back
            base = widths._base
            if (base is not None and heights._base is base and derivatives._base is base and base.is_contiguous()
                    and base.data_ptr() == widths.data_ptr() and base.numel() == rows * PW):
                return base.view(rows, 1, PW), lead
    w = widths.expand(*lead, K)
    h = heights.expand(*lead, K)
    d = derivatives.expand(*lead, K - 1)
    return torch.cat((w, h, d), dim=-1).reshape(-1, 1, PW), lead


class MonotonicRQSTransform(Transform):
    r"""Monotonic rational-quadratic spline (Durkan et al. 2019) on [-bound, bound], identity outside."""

    domain = constraints.real
    codomain = constraints.real
    bijective = True
    sign = +1

    def __init__(self, widths, heights, derivatives, bound=5.0, slope=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.bound = float(bound)
        self.slope = slope
        self.phi, self.lead = _pack(widths, heights, derivatives)

    @property
    def bins(self):
        return (self.phi.shape[-1] + 1) // 3

    def _rows(self, x):
        """x as rows [R, 1] plus the parameter rows [R, 1, 3K-1] they go with. Per-element parameters (what a
        hyper-network emits) are used in place; parameters shared along some dimension of x (an unconditional
        transform) are expanded to one row per element, as torch broadcasting would."""
        shape = torch.broadcast_shapes(self.lead, x.shape)
        phi = self.phi
        if tuple(shape) != tuple(self.lead):
            PW = phi.shape[-1]
            phi = phi.reshape(*self.lead, PW).expand(*shape, PW).reshape(-1, 1, PW)
        return x.expand(shape).reshape(-1, 1), shape, phi

    def _call(self, x):
        v, shape, phi = self._rows(x)
        y, _, _, _ = _run("forward", phi, v, self.bound, self.slope, want_ladj=False)
        return y.reshape(shape).to(x.dtype)

    def _inverse(self, y):
        v, shape, phi = self._rows(y)
        x, _, _, _ = _run("inverse", phi, v, self.bound, self.slope, want_ladj=False)
        return x.reshape(shape).to(y.dtype)

    def call_and_ladj(self, x):
        v, shape, phi = self._rows(x)
        y, ladj, _, _ = _run("forward", phi, v, self.bound, self.slope)
        return y.reshape(shape).to(x.dtype), ladj.reshape(shape).to(x.dtype)

    def inverse_and_ladj(self, y):
        """(x, forward ladj at x): lets a sampler get log q without a second pass."""
        v, shape, phi = self._rows(y)
        x, ladj, _, _ = _run("inverse", phi, v, self.bound, self.slope)
        return x.reshape(shape).to(y.dtype), ladj.reshape(shape).to(y.dtype)

    def log_abs_det_jacobian(self, x, y=None):
        return self.call_and_ladj(x)[1]
