"""ctypes front-end of oracle/memflow_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- see the header of memflow_oracle.c. The library is built by
`make -C oracle` (or __graft_entry__.build()); nothing here touches CUDA.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmemflow_oracle.so")
_lib = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_i32_p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "memflow_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmemflow_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.mfo_rambo_forward.restype = ctypes.c_int
        _lib.mfo_rambo_inverse.restype = ctypes.c_int
        _lib.mfo_get_ps.restype = ctypes.c_int
        for n in ("mfo_rqs_forward_f32", "mfo_rqs_forward_f64", "mfo_rqs_inverse_f32",
                  "mfo_rqs_inverse_f64", "mfo_is_reduce", "mfo_max_threads", "mfo_get_threads"):
            getattr(_lib, n).restype = ctypes.c_int
        _lib.mfo_set_threads.restype = None
    return _lib


def set_threads(t):
    lib().mfo_set_threads(ctypes.c_int(int(t)))


def max_threads():
    return int(lib().mfo_max_threads())


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def rambo_forward(u, masses, e_collider, fp32_quirks=False, pT_mincut=-1.0, delR_mincut=-1.0,
                  rap_maxcut=-1.0):
    """rambo_generator.py:168-398. u [N, 3n-2] -> momenta [N, n+2, 4], weight, x1, x2 (float64)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    masses = np.ascontiguousarray(masses, dtype=np.float64)
    n = masses.shape[0]
    N = u.shape[0]
    assert u.shape[1] == 3 * n - 2
    mom = np.empty((N, n + 2, 4), dtype=np.float64)
    w = np.empty(N, dtype=np.float64)
    x1 = np.empty(N, dtype=np.float64)
    x2 = np.empty(N, dtype=np.float64)
    rc = lib().mfo_rambo_forward(_dp(u), ctypes.c_int64(N), ctypes.c_int(n), _dp(masses),
                                 ctypes.c_double(e_collider), ctypes.c_int(int(fp32_quirks)),
                                 ctypes.c_double(pT_mincut), ctypes.c_double(delR_mincut),
                                 ctypes.c_double(rap_maxcut), _dp(mom), _dp(w), _dp(x1), _dp(x2))
    if rc == 1:
        raise ValueError("NaN in random variables (reference raises PhaseSpaceGeneratorError)")
    if rc != 0:
        raise RuntimeError("mfo_rambo_forward rc=%d" % rc)
    return mom, w, x1, x2


def rambo_inverse(momenta, x1, x2, masses, e_collider, order=None, target_mass=None, ensure_cm=True,
                  fp32_quirks=False):
    """rambo_generator.py:615-736. momenta [N, n, 4] -> ps [N, 3n-2], prob [N]; also returns rc
    (2 = the reference would have raised 'Momenta batch not in the CM')."""
    momenta = np.ascontiguousarray(momenta, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    masses = np.ascontiguousarray(masses, dtype=np.float64)
    N, n, _ = momenta.shape
    order = np.ascontiguousarray(list(range(n)) if order is None else order, dtype=np.intc)
    tm = None
    if target_mass is not None:
        tm = np.ascontiguousarray(np.broadcast_to(np.asarray(target_mass, dtype=np.float64), (N, n)))
    ps = np.empty((N, 3 * n - 2), dtype=np.float64)
    prob = np.empty(N, dtype=np.float64)
    rc = lib().mfo_rambo_inverse(_dp(momenta), _dp(x1), _dp(x2), ctypes.c_int64(N), ctypes.c_int(n),
                                 order.ctypes.data_as(c_int_p), _dp(masses),
                                 _dp(tm) if tm is not None else None, ctypes.c_double(e_collider),
                                 ctypes.c_int(int(ensure_cm)), ctypes.c_int(int(fp32_quirks)), _dp(ps),
                                 _dp(prob))
    if rc not in (0, 2):
        raise RuntimeError("mfo_rambo_inverse rc=%d" % rc)
    return ps, prob, rc


def get_ps(momenta, boost, target_mass, e_collider, order=None):
    """memflow/unfolding_flow/utils.py:1056-1152: returns ps [N,3n-2], 0*jac [N], mask [N] bool."""
    momenta = np.ascontiguousarray(momenta, dtype=np.float64)
    boost = np.ascontiguousarray(boost, dtype=np.float64)
    N, n, _ = momenta.shape
    order = np.ascontiguousarray(list(range(n)) if order is None else order, dtype=np.intc)
    tm = np.ascontiguousarray(np.asarray(target_mass, dtype=np.float64).reshape(-1))
    ps = np.empty((N, 3 * n - 2), dtype=np.float64)
    jac0 = np.empty(N, dtype=np.float64)
    mask = np.empty(N, dtype=np.uint8)
    rc = lib().mfo_get_ps(_dp(momenta), _dp(boost), ctypes.c_int64(N), ctypes.c_int(n), order.ctypes.data_as(c_int_p),
                          _dp(tm), ctypes.c_double(e_collider), _dp(ps), _dp(jac0),
                          mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    if rc != 0:
        raise RuntimeError("mfo_get_ps rc=%d" % rc)
    return ps, jac0, mask.astype(bool)


def higgs_decay(parent, angles, mass=125.25, ptetaphi=True):
    """memflow/unfolding_flow/utils.py:745-770: parent [N,4], angles [N,2] -> [N,3,4] = (H, b1, b2)."""
    parent = np.ascontiguousarray(parent, dtype=np.float64)
    angles = np.ascontiguousarray(angles, dtype=np.float64)
    N = parent.shape[0]
    out = np.empty((N, 3, 4), dtype=np.float64)
    lib().mfo_higgs_decay(_dp(parent), _dp(angles), ctypes.c_int64(N), ctypes.c_double(mass), ctypes.c_int(int(ptetaphi)),
                          _dp(out))
    return out


def top_decay(parent, b_angles, w_angles, top_mass=172.5, w_mass=80.4, b_mass=0.0, ptetaphi=True):
    """memflow/unfolding_flow/utils.py:772-815: -> [N,4,4] = (t, b, q1, q2)."""
    parent = np.ascontiguousarray(parent, dtype=np.float64)
    b_angles = np.ascontiguousarray(b_angles, dtype=np.float64)
    w_angles = np.ascontiguousarray(w_angles, dtype=np.float64)
    N = parent.shape[0]
    out = np.empty((N, 4, 4), dtype=np.float64)
    lib().mfo_top_decay(_dp(parent), _dp(b_angles), _dp(w_angles), ctypes.c_int64(N), ctypes.c_double(top_mass),
                        ctypes.c_double(w_mass), ctypes.c_double(b_mass), ctypes.c_int(int(ptetaphi)), _dp(out))
    return out


def decay_partons_cfg(mean_prop, std_prop, mean_lab, std_lab, mean_boost, std_boost, higgs_mass=125.25, thad_mass=172.5,
                      tlep_mass=172.5, W_had_mass=80.4, W_lep_mass=80.4, b_mass=0.0, eps=1e-4, ptetaphi=True,
                      unscale_phi=False, final_scaling=True):
    return np.array(list(mean_prop) + list(std_prop) + list(mean_lab) + list(std_lab) + list(mean_boost) + list(std_boost)
                    + [higgs_mass, thad_mass, tlep_mass, W_had_mass, W_lep_mass, b_mass, eps, float(ptetaphi),
                       float(unscale_phi), float(final_scaling)], dtype=np.float64)


def decay_partons(prop, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, boost, cfg):
    """memflow/unfolding_flow/utils.py:826-935: prop [N,3,3], angles [N,2], boost [N,2] (or [N,1,2]) -> [N,12,4]."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
            (prop, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, boost)]
    N = arrs[0].shape[0]
    out = np.empty((N, 12, 4), dtype=np.float64)
    cfg = np.ascontiguousarray(cfg, dtype=np.float64)
    assert cfg.size == 26
    rc = lib().mfo_decay_partons(*[_dp(a) for a in arrs], ctypes.c_int64(N), _dp(cfg), _dp(out))
    if rc != 0:
        raise RuntimeError("mfo_decay_partons rc=%d" % rc)
    return out


def _rqs(direction, phi, v, bound, slope):
    phi = np.ascontiguousarray(phi)
    dt = phi.dtype
    assert dt in (np.float32, np.float64)
    v = np.ascontiguousarray(v, dtype=dt)
    B, F, PW = phi.shape
    K = (PW + 1) // 3
    assert 3 * K - 1 == PW and v.shape == (B, F)
    out = np.empty((B, F), dtype=dt)
    ladj = np.empty((B, F), dtype=dt)
    bins = np.empty((B, F), dtype=np.int32)
    suf = "f32" if dt == np.float32 else "f64"
    fn = getattr(lib(), "mfo_rqs_%s_%s" % (direction, suf))
    ct = ctypes.c_float if dt == np.float32 else ctypes.c_double
    pt = c_float_p if dt == np.float32 else c_double_p
    rc = fn(phi.ctypes.data_as(pt), v.ctypes.data_as(pt), ctypes.c_int64(B), ctypes.c_int(F),
            ctypes.c_int(K), ct(bound), ct(slope if slope is not None else 0.0), out.ctypes.data_as(pt),
            ladj.ctypes.data_as(pt), bins.ctypes.data_as(c_i32_p))
    if rc != 0:
        raise RuntimeError("mfo_rqs rc=%d" % rc)
    return out, ladj, bins


def rqs_forward(phi, x, bound=5.0, slope=1e-3):
    """zuko MonotonicRQSTransform.call_and_ladj: returns y, ladj (per feature), bin index k."""
    return _rqs("forward", phi, x, bound, slope)


def rqs_inverse(phi, y, bound=5.0, slope=1e-3):
    """zuko MonotonicRQSTransform._inverse (+ forward ladj at the recovered x), bin index k."""
    return _rqs("inverse", phi, y, bound, slope)


def is_reduce(target, jac, logq):
    """TestFlows_ImportanceSampling.ipynb [50-54]: returns (sum w, sum w^2, n_valid)."""
    jac = np.ascontiguousarray(jac, dtype=np.float64)
    logq = np.ascontiguousarray(logq, dtype=np.float64)
    N = jac.shape[0]
    t = None if target is None else np.ascontiguousarray(target, dtype=np.float64)
    acc = np.zeros(3, dtype=np.float64)
    lib().mfo_is_reduce(_dp(t) if t is not None else None, _dp(jac), _dp(logq), ctypes.c_int64(N),
                        _dp(acc))
    return acc


def estimator(acc, n_total, variant="N+1"):
    """I = sum w / N ; sigma = sqrt((sum w^2 / N - I^2) / (N +- 1)) (cells [53-54] / [35])."""
    I = acc[0] / n_total
    den = n_total + 1 if variant == "N+1" else n_total - 1
    return I, float(np.sqrt(max(acc[1] / n_total - I * I, 0.0) / den))
