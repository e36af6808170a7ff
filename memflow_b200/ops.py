"""torch.library registration of the hot-path kernels: the "thin torch custom-op layer" over the C ABI.

Namespace `memflow_b200`, CUDA dispatch key ONLY: a call with CPU tensors fails in the dispatcher ("could not run ... with
arguments from the 'CPU' backend") -- there is no CPU / PyTorch fallback to fall into. Every op has a schema, a fake (meta)
implementation for FakeTensor / torch.compile shape propagation, and -- where the reference's graph is differentiable -- an
autograd formula whose backward is itself a registered op backed by a backward kernel.

    torch.ops.memflow_b200.rqs(phi, v, bound, slope, inverse, outputs)        -> out, ladj, ladj_sum, bins        (K3 / K4)
    torch.ops.memflow_b200.rqs_bwd(phi, v, g_out, g_ladj, bound, slope, inverse) -> g_phi, g_v
    torch.ops.memflow_b200.rambo_forward(u, masses, e, cuts, flags, want_logw) -> momenta, weight, x1, x2, logw, status   (K1)
    torch.ops.memflow_b200.rambo_forward_bwd(...)                              -> grad_u
    torch.ops.memflow_b200.rambo_inverse(momenta, x1, x2, order, masses, target_mass, e, flags, want_logp)
                                                                               -> ps, prob, logp, status           (K2)
    torch.ops.memflow_b200.rambo_inverse_bwd(...)                              -> g_momenta, g_x1, g_x2
    torch.ops.memflow_b200.is_reduce_(target, jac, logq, acc)                  acc += (sum w, sum w^2, n)          (K5)
    torch.ops.memflow_b200.flow_sample(phi, seed, event_offset, bound, slope)  -> u, logq                          (chain A)
    torch.ops.memflow_b200.rambo_is_(u, logq, masses, e, acc)                  chain B

The host classes (transforms.MonotonicRQSTransform, phasespace.rambo_generator.FlatInvertiblePhasespace,
estimator.ImportanceSampler) call these ops; the ops call the C ABI (include/memflow_b200.h) through ctypes with the
argument types taken from that header (_lib.header_signatures).
"""
import torch
from torch.library import Library

from ._lib import check, device_guard, host_array, lib, MemflowB200Error, ptr, reduce_workspace, stream_ptr
import ctypes

_NS = "memflow_b200"
_DEF = Library(_NS, "DEF")


def _register(name, schema, cuda_impl, fake_impl):
    _DEF.define(name + schema)
    _DEF.impl(name, cuda_impl, "CUDA")
    torch.library.register_fake(_NS + "::" + name, fake_impl)


_FN = {}


def _abi(name):
    """cached attribute lookup on the loaded library"""
    f = _FN.get(name)
    if f is None:
        f = _FN[name] = getattr(lib(), name)
    return f


def _suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise MemflowB200Error("memflow_b200 kernels support float32 / float64, got %s" % dtype)


# ---------------------------------------------------------------------------------------------------------------------
# K3 / K4: monotonic rational-quadratic spline
# ---------------------------------------------------------------------------------------------------------------------
def _rqs_shapes(phi, v):
    if phi.dim() != 3:
        raise AssertionError("phi must be [B, F, 3K-1]")
    B, F, PW = phi.shape
    if (PW + 1) % 3 != 0:
        raise AssertionError("last dim of phi must be 3K-1")
    if tuple(v.shape) != (B, F):
        raise AssertionError("x must be [B, F] = %s, got %s" % ((B, F), tuple(v.shape)))
    return B, F, (PW + 1) // 3


RQS_LADJ, RQS_SUM, RQS_BINS = 1, 2, 4   # `outputs` bit mask: which of ladj [B,F], ladj_sum [B], bins [B,F] are produced


def _rqs_rows(phi):
    """(phi usable in place, event stride in elements): a [B, F, 3K-1] view whose rows are contiguous and whose events are
    a constant number of elements apart -- e.g. phi_all[:, j:j+1] in an autoregressive sampling pass -- is read in place by
    the strided entry points; anything else is packed first."""
    B, F, PW = phi.shape
    es = phi.element_size()
    if phi.stride(2) == 1 and (F == 1 or phi.stride(1) == PW) and phi.stride(0) >= F * PW and phi.data_ptr() % es == 0 and B > 0:
        return phi, int(phi.stride(0)) if B > 1 else F * PW
    phi = phi.contiguous()
    return phi, F * PW


def _rqs_cuda(phi, v, bound, slope, inverse, outputs):
    B, F, K = _rqs_shapes(phi, v)
    phi, ev_stride = _rqs_rows(phi)
    v = v.to(phi.dtype).contiguous()
    dev = phi.device
    out = torch.empty_like(v)
    ladj = torch.empty((B, F) if outputs & RQS_LADJ else (0,), dtype=phi.dtype, device=dev)
    lsum = torch.empty((B,) if outputs & RQS_SUM else (0,), dtype=phi.dtype, device=dev)
    bins = torch.empty((B, F) if outputs & RQS_BINS else (0,), dtype=torch.int32, device=dev)
    fn = _abi(("mf_rqs_inverse_strided_" if inverse else "mf_rqs_forward_strided_") + _suffix(phi.dtype))
    with device_guard(dev):
        rc = fn(ptr(phi), ev_stride, ptr(v), B, F, K, bound, slope, ptr(out), ptr(ladj) if outputs & RQS_LADJ else None,
                ptr(lsum) if outputs & RQS_SUM else None, ptr(bins) if outputs & RQS_BINS else None, stream_ptr(dev))
    check(rc, "mf_rqs_%s" % ("inverse" if inverse else "forward"))
    return out, ladj, lsum, bins


def _rqs_fake(phi, v, bound, slope, inverse, outputs):
    B, F, K = _rqs_shapes(phi, v)
    return (phi.new_empty((B, F)), phi.new_empty((B, F) if outputs & RQS_LADJ else (0,)),
            phi.new_empty((B,) if outputs & RQS_SUM else (0,)),
            phi.new_empty((B, F) if outputs & RQS_BINS else (0,), dtype=torch.int32))


def _rqs_bwd_cuda(phi, v, g_out, g_ladj, bound, slope, inverse):
    B, F, K = _rqs_shapes(phi, v)
    phi = phi.contiguous()
    v = v.to(phi.dtype).contiguous()
    go = None if g_out is None else g_out.to(phi.dtype).contiguous()
    gl = None if g_ladj is None else g_ladj.to(phi.dtype).contiguous()
    g_phi = torch.empty_like(phi)
    g_v = torch.empty_like(v)
    fn = getattr(lib(), "mf_rqs_%s_bwd_%s" % ("inverse" if inverse else "forward", _suffix(phi.dtype)))
    with device_guard(phi.device):
        rc = fn(ptr(phi), ptr(v), ptr(go), ptr(gl), B, F, K, bound, slope, ptr(g_phi), ptr(g_v), stream_ptr(phi.device))
    check(rc, "mf_rqs_bwd")
    return g_phi, g_v


def _rqs_bwd_fake(phi, v, g_out, g_ladj, bound, slope, inverse):
    return torch.empty_like(phi), phi.new_empty(v.shape)


_register("rqs", "(Tensor phi, Tensor v, float bound, float slope, bool inverse, int outputs) -> (Tensor, Tensor, Tensor, Tensor)",
          _rqs_cuda, _rqs_fake)
_register("rqs_bwd", "(Tensor phi, Tensor v, Tensor? g_out, Tensor? g_ladj, float bound, float slope, bool inverse) -> (Tensor, Tensor)",
          _rqs_bwd_cuda, _rqs_bwd_fake)


def _rqs_setup(ctx, inputs, output):
    phi, v, bound, slope, inverse, _ = inputs
    ctx.save_for_backward(phi, v)
    ctx.cfg = (bound, slope, inverse, v.dtype)
    ctx.set_materialize_grads(False)


def _rqs_backward(ctx, g_out, g_ladj, g_lsum, _g_bins):
    phi, v = ctx.saved_tensors
    bound, slope, inverse, vdt = ctx.cfg
    if g_lsum is not None and g_lsum.numel():   # the per-event sum's cotangent is the per-element one broadcast over the features
        g_ladj = g_lsum.unsqueeze(1).expand(v.shape) if (g_ladj is None or not g_ladj.numel()) else g_ladj + g_lsum.unsqueeze(1)
    if g_ladj is not None and not g_ladj.numel():
        g_ladj = None
    if g_out is None and g_ladj is None:
        return None, None, None, None, None, None
    g_phi, g_v = torch.ops.memflow_b200.rqs_bwd(phi, v, g_out, g_ladj, bound, slope, inverse)
    return g_phi, g_v.to(vdt), None, None, None, None


torch.library.register_autograd(_NS + "::rqs", _rqs_backward, setup_context=_rqs_setup)


class RqsEager(torch.autograd.Function):
    """The same forward / backward pair applied directly in plain eager mode. The registered op above costs ~55 us of
    dispatcher, default-filling and redispatch per call under autograd (scripts/graph_latency.py: 79 us per recorded
    forward against 23 us without grad at B = 1024); this path keeps the recording and drops that. torch.compile, fake
    tensors and functorch transforms still go through torch.ops.memflow_b200.rqs (memflow_b200/transforms.py::_run)."""

    @staticmethod
    def forward(ctx, phi, v, bound, slope, inverse, outputs):
        out, ladj, lsum, bins = _rqs_cuda(phi.detach(), v.detach(), bound, slope, inverse, outputs)
        ctx.save_for_backward(phi, v)
        ctx.cfg = (bound, slope, inverse, v.dtype)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(bins)
        return out, ladj, lsum, bins

    @staticmethod
    def backward(ctx, g_out, g_ladj, g_lsum, _g_bins):
        phi, v = ctx.saved_tensors
        bound, slope, inverse, vdt = ctx.cfg
        if g_lsum is not None and g_lsum.numel():
            g_ladj = g_lsum.unsqueeze(1).expand(v.shape) if (g_ladj is None or not g_ladj.numel()) else g_ladj + g_lsum.unsqueeze(1)
        if g_ladj is not None and not g_ladj.numel():
            g_ladj = None
        if g_out is None and g_ladj is None:
            return None, None, None, None, None, None
        g_phi, g_v = _rqs_bwd_cuda(phi.detach(), v.detach(), g_out, g_ladj, bound, slope, inverse)
        return g_phi, g_v.to(vdt), None, None, None, None


def _eager_function(name, impl, setup, backward, n_inputs):
    """A torch.autograd.Function that applies `impl` / `backward` of a registered op directly (plain eager autograd: the
    dispatcher round trip of the registered op costs ~55 us per recorded call, see RqsEager). `setup` and `backward` are
    the functions given to torch.library.register_autograd: one formula, two ways in."""
    def forward(ctx, *inputs):
        out = impl(*(t.detach() if isinstance(t, torch.Tensor) else t for t in inputs))
        setup(ctx, inputs, out)
        ctx.mark_non_differentiable(*(o for o in out if isinstance(o, torch.Tensor) and not o.is_floating_point()))
        return out

    def bwd(ctx, *grads):
        res = backward(ctx, *grads)
        return tuple(res) + (None,) * (n_inputs - len(res))

    return type(name, (torch.autograd.Function,), {"forward": staticmethod(forward), "backward": staticmethod(bwd)})


# ---------------------------------------------------------------------------------------------------------------------
# K1: RAMBO forward
# ---------------------------------------------------------------------------------------------------------------------
MF_FLAG_ZERO_STATUS = 32
MF_FLAG_CUTS = 4


def _k1_cuda(u, masses, e_collider, cuts, flags, want_logw):
    n = len(masses)
    if u.dim() != 2 or u.shape[1] != 3 * n - 2:
        raise AssertionError("points must be [N, 3 n_final - 2]")
    u = u.contiguous()
    dt, dev, N = u.dtype, u.device, u.shape[0]
    ct = ctypes.c_double if dt == torch.float64 else ctypes.c_float
    momenta = torch.empty((N, n + 2, 4), dtype=dt, device=dev)
    weight = torch.empty(N, dtype=dt, device=dev)
    x1 = torch.empty(N, dtype=dt, device=dev)
    x2 = torch.empty(N, dtype=dt, device=dev)
    logw = torch.empty(N if want_logw else 0, dtype=dt, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)   # cleared inside the call (MF_FLAG_ZERO_STATUS): no fill launch
    fn = _abi("mf_rambo_forward_" + _suffix(dt))
    with device_guard(dev):
        rc = fn(ptr(u), N, n, host_array(masses, ct), e_collider, host_array(cuts, ct) if (flags & MF_FLAG_CUTS) else None,
                flags | MF_FLAG_ZERO_STATUS, ptr(momenta), ptr(weight), ptr(logw) if want_logw else None, ptr(x1), ptr(x2),
                ptr(status), stream_ptr(dev))
    check(rc, "mf_rambo_forward")
    return momenta, weight, x1, x2, logw, status


def _k1_fake(u, masses, e_collider, cuts, flags, want_logw):
    n, N = len(masses), u.shape[0]
    return (u.new_empty((N, n + 2, 4)), u.new_empty((N,)), u.new_empty((N,)), u.new_empty((N,)),
            u.new_empty((N if want_logw else 0,)), u.new_empty((1,), dtype=torch.int32))


def _k1_bwd_cuda(u, momenta, weight, x1, x2, g_mom, g_w, g_x1, g_x2, masses, e_collider, flags):
    def c(t):
        return None if t is None else t.to(torch.float64).contiguous()

    g_mom, g_w, g_x1, g_x2 = c(g_mom), c(g_w), c(g_x1), c(g_x2)
    grad_u = torch.empty_like(u)
    with device_guard(u.device):
        rc = lib().mf_rambo_forward_bwd_f64(ptr(u), ptr(momenta), ptr(weight), ptr(x1), ptr(x2), ptr(g_mom), ptr(g_w), ptr(g_x1),
                                            ptr(g_x2), u.shape[0], len(masses), host_array(masses, ctypes.c_double), e_collider,
                                            flags, ptr(grad_u), stream_ptr(u.device))
    check(rc, "mf_rambo_forward_bwd_f64")
    return grad_u


_register("rambo_forward", "(Tensor u, float[] masses, float e_collider, float[] cuts, int flags, bool want_logw) -> "
          "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", _k1_cuda, _k1_fake)
_register("rambo_forward_bwd", "(Tensor u, Tensor momenta, Tensor weight, Tensor x1, Tensor x2, Tensor? g_momenta, Tensor? g_weight, "
          "Tensor? g_x1, Tensor? g_x2, float[] masses, float e_collider, int flags) -> Tensor", _k1_bwd_cuda,
          lambda u, *a: torch.empty_like(u))


def _k1_setup(ctx, inputs, output):
    u, masses, e_collider, cuts, flags, _ = inputs
    momenta, weight, x1, x2, _, _ = output
    ctx.save_for_backward(u, momenta, weight, x1, x2)
    ctx.cfg = (masses, e_collider, flags)
    ctx.set_materialize_grads(False)


def _k1_backward(ctx, g_mom, g_w, g_x1, g_x2, g_logw, _g_status):
    u, momenta, weight, x1, x2 = ctx.saved_tensors
    masses, e_collider, flags = ctx.cfg
    if flags & (1 | 2 | 4) or u.dtype != torch.float64:
        raise MemflowB200Error("autograd through the phase-space map supports the plain fp64 CM-frame output only "
                               "(no cuts, no lab_frame, no fp32 compute / quirks)")
    if g_logw is not None:   # d ln w = dw / w
        g_w = g_logw / weight if g_w is None else g_w + g_logw / weight
    grad_u = torch.ops.memflow_b200.rambo_forward_bwd(u, momenta, weight, x1, x2, g_mom, g_w, g_x1, g_x2, masses, e_collider, flags & 8)
    return grad_u, None, None, None, None, None


torch.library.register_autograd(_NS + "::rambo_forward", _k1_backward, setup_context=_k1_setup)
RamboForwardEager = _eager_function("RamboForwardEager", lambda *a: _k1_cuda(*a), _k1_setup, _k1_backward, 6)


# ---------------------------------------------------------------------------------------------------------------------
# K2: RAMBO inverse
# ---------------------------------------------------------------------------------------------------------------------
def _k2_view(mom, n):
    """a [N, n, 4] view the kernel can read in place (e.g. forward_output[:, 2:]); a packed copy otherwise"""
    esz = mom.element_size()
    if not (mom.stride(2) == 1 and mom.stride(1) == 4 and mom.stride(0) % 4 == 0 and mom.stride(0) >= 4 * n
            and mom.data_ptr() % (4 * esz) == 0):
        mom = mom.contiguous()
    return mom


def _k2_cuda(momenta, x1, x2, order, masses, target_mass, e_collider, flags, want_logp):
    n = len(masses)
    if momenta.dim() != 3 or momenta.shape[1] != n or momenta.shape[2] != 4:
        raise AssertionError("momenta_batch must be [N, %d, 4]" % n)
    dt, dev, N = momenta.dtype, momenta.device, momenta.shape[0]
    mom = _k2_view(momenta, n)
    x1c, x2c = x1.to(dt).contiguous(), x2.to(dt).contiguous()
    ct = ctypes.c_double if dt == torch.float64 else ctypes.c_float
    ps = torch.empty((N, 3 * n - 2), dtype=dt, device=dev)
    prob = torch.empty(N, dtype=dt, device=dev)
    logp = torch.empty(N if want_logp else 0, dtype=dt, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    fn = _abi("mf_rambo_inverse_" + _suffix(dt))
    with device_guard(dev):
        rc = fn(ptr(mom), mom.stride(0), ptr(x1c), ptr(x2c), N, n, host_array(order, ctypes.c_int32), host_array(masses, ct),
                host_array(target_mass, ct) if target_mass is not None else None, e_collider, flags | MF_FLAG_ZERO_STATUS,
                ptr(ps), ptr(prob), ptr(logp) if want_logp else None, ptr(status), stream_ptr(dev))
    check(rc, "mf_rambo_inverse")
    return ps, prob, logp, status


def _k2_fake(momenta, x1, x2, order, masses, target_mass, e_collider, flags, want_logp):
    n, N = len(masses), momenta.shape[0]
    return (momenta.new_empty((N, 3 * n - 2)), momenta.new_empty((N,)), momenta.new_empty((N if want_logp else 0,)),
            momenta.new_empty((1,), dtype=torch.int32))


def _k2_bwd_cuda(momenta, x1, x2, g_ps, order, masses, target_mass, permute_target, e_collider, flags):
    n = len(masses)
    mom = _k2_view(momenta.to(torch.float64), n)
    dev, N = mom.device, mom.shape[0]
    g_mom = torch.empty((N, n, 4), dtype=torch.float64, device=dev)
    g_x1 = torch.empty(N, dtype=torch.float64, device=dev)
    g_x2 = torch.empty(N, dtype=torch.float64, device=dev)
    with device_guard(dev):
        rc = lib().mf_rambo_inverse_bwd_f64(ptr(mom), mom.stride(0), ptr(x1.to(torch.float64).contiguous()),
                                            ptr(x2.to(torch.float64).contiguous()), ptr(g_ps.to(torch.float64).contiguous()), N, n,
                                            host_array(order, ctypes.c_int32), host_array(masses, ctypes.c_double),
                                            host_array(target_mass, ctypes.c_double) if target_mass is not None else None,
                                            int(permute_target), e_collider, flags, ptr(g_mom), ptr(g_x1), ptr(g_x2), stream_ptr(dev))
    check(rc, "mf_rambo_inverse_bwd_f64")
    return g_mom, g_x1, g_x2


_register("rambo_inverse", "(Tensor momenta, Tensor x1, Tensor x2, int[] order, float[] masses, float[]? target_mass, float e_collider, "
          "int flags, bool want_logp) -> (Tensor, Tensor, Tensor, Tensor)", _k2_cuda, _k2_fake)
_register("rambo_inverse_bwd", "(Tensor momenta, Tensor x1, Tensor x2, Tensor g_ps, int[] order, float[] masses, float[]? target_mass, "
          "bool permute_target, float e_collider, int flags) -> (Tensor, Tensor, Tensor)", _k2_bwd_cuda,
          lambda momenta, x1, x2, *a: (momenta.new_empty(momenta.shape), x1.new_empty(x1.shape), x2.new_empty(x2.shape)))


def _k2_setup(ctx, inputs, output):
    momenta, x1, x2, order, masses, target_mass, e_collider, flags, _ = inputs
    ctx.save_for_backward(momenta, x1, x2)
    ctx.cfg = (order, masses, target_mass, e_collider, flags)
    ctx.set_materialize_grads(False)


def _k2_backward(ctx, g_ps, _g_prob, _g_logp, _g_status):
    momenta, x1, x2 = ctx.saved_tensors
    order, masses, target_mass, e_collider, flags = ctx.cfg
    if g_ps is None:   # `prob` is not differentiable (the reference's prob is a float32 side product)
        return (None,) * 9
    if momenta.dtype != torch.float64 or flags & 1:
        raise MemflowB200Error("autograd through the inverse map supports fp64 compute without quirks only")
    g_mom, g_x1, g_x2 = torch.ops.memflow_b200.rambo_inverse_bwd(momenta, x1, x2, g_ps, order, masses, target_mass, False, e_collider,
                                                                 flags & 8)
    return (g_mom.to(momenta.dtype), g_x1.to(x1.dtype), g_x2.to(x2.dtype), None, None, None, None, None, None)


torch.library.register_autograd(_NS + "::rambo_inverse", _k2_backward, setup_context=_k2_setup)
RamboInverseEager = _eager_function("RamboInverseEager", lambda *a: _k2_cuda(*a), _k2_setup, _k2_backward, 9)


# ---------------------------------------------------------------------------------------------------------------------
# K5 and the sampling chain
# ---------------------------------------------------------------------------------------------------------------------
def _is_reduce_cuda(target, jac, logq, acc):
    dev = jac.device
    jac = jac.to(torch.float64).contiguous()
    lq = logq.to(torch.float64).contiguous()
    tg = None if target is None else target.to(torch.float64).contiguous()
    ws = reduce_workspace(dev)
    with device_guard(dev):
        check(lib().mf_is_reduce_f64(ptr(tg), ptr(jac), ptr(lq), jac.shape[0], ptr(acc), ptr(ws), stream_ptr(dev)), "mf_is_reduce_f64")


def _flow_sample_cuda(phi, seed, event_offset, bound, slope):
    T, N, F, PW = phi.shape
    K = (PW + 1) // 3
    phi = phi.contiguous()
    dev = phi.device
    u = torch.empty((N, F), dtype=torch.float64, device=dev)
    logq = torch.empty(N, dtype=torch.float64, device=dev)
    fn = getattr(lib(), "mf_flow_sample_" + _suffix(phi.dtype))
    with device_guard(dev):
        rc = fn(ptr(phi), N, F, T, K, bound, slope, seed, event_offset, ptr(u), ptr(logq), stream_ptr(dev))
    check(rc, "mf_flow_sample")
    return u, logq


def _rambo_is_cuda(u, logq, masses, e_collider, acc):
    dev = u.device
    u = u.to(torch.float64).contiguous()
    logq = logq.to(torch.float64).contiguous()
    ws = reduce_workspace(dev)
    with device_guard(dev):
        rc = lib().mf_rambo_is_f64(ptr(u), ptr(logq), u.shape[0], len(masses), host_array(masses, ctypes.c_double), e_collider, 0,
                                   ptr(acc), ptr(ws), None, None, None, None, stream_ptr(dev))
    check(rc, "mf_rambo_is_f64")


_register("is_reduce_", "(Tensor? target, Tensor jac, Tensor logq, Tensor(a!) acc) -> ()", _is_reduce_cuda, lambda *a: None)
_register("flow_sample", "(Tensor phi, int seed, int event_offset, float bound, float slope) -> (Tensor, Tensor)", _flow_sample_cuda,
          lambda phi, *a: (phi.new_empty((phi.shape[1], phi.shape[2]), dtype=torch.float64), phi.new_empty((phi.shape[1],), dtype=torch.float64)))
_register("rambo_is_", "(Tensor u, Tensor logq, float[] masses, float e_collider, Tensor(a!) acc) -> ()", _rambo_is_cuda, lambda *a: None)

# the CUDA implementations, callable directly by the host classes when neither autograd nor a tracer is involved
rqs_cuda, rambo_forward_cuda, rambo_inverse_cuda = _rqs_cuda, _k1_cuda, _k2_cuda

OPS = ("rqs", "rqs_bwd", "rambo_forward", "rambo_forward_bwd", "rambo_inverse", "rambo_inverse_bwd", "is_reduce_", "flow_sample",
       "rambo_is_")
