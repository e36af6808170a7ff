"""Loader and timing legs for the staged reference (oracle/_ref, see oracle/make_ref.py) -- TEST INFRASTRUCTURE ONLY.

Runs the reference's own torch CPU path: `PhaseSpace.get_momenta_from_ps` (memflow/phasespace/phasespace.py:82-103)
and `generator.getPSpoint_batch` (rambo_generator.py:615-736), unmodified, float64 default dtype set BEFORE the import
and masses passed explicitly (SURVEY 8c usage rules). zuko is absent from the reference tree and from the image, so
the spline step of the chain uses oracle/rqs_torch.py (torch restatement of zuko's transform, same torch ops).
"""
import contextlib
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "memflow", "phasespace", "rambo_generator.py"))


_mods = None


def load():
    """(phasespace, rambo_generator, utils) modules of the staged reference."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
        import torch
        torch.set_default_dtype(torch.float64)   # before the import: default arguments are evaluated at import time
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from memflow.phasespace import phasespace, rambo_generator, utils
        _mods = (phasespace, rambo_generator, utils)
    return _mods


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _best(f, reps):
    f()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        best = min(best, time.perf_counter() - t0)
    return best


def time_phase_space(n_events=1_000_000, reps=2, threads=None):
    """BASELINE.md section 3's own plan: config 1 forward (n = 3, 1 M points, seed 12345) and config 2's inverse
    (n = 8, same count), all host threads and 1 thread. Returns a dict of points/s."""
    import torch
    ps_mod, _, _ = load()
    ncpu = os.cpu_count() or 1
    out = {"cpu": cpu_model(), "cores": ncpu, "points": int(n_events), "torch": torch.__version__}
    E = 13000.0
    m3 = torch.tensor([125.25, 172.5, 172.5], dtype=torch.float64)
    m8 = torch.tensor([4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3], dtype=torch.float64)
    g = torch.Generator().manual_seed(12345)
    u3 = torch.rand(n_events, 7, dtype=torch.float64, generator=g).clamp_(1e-10, 1 - 1e-10)
    g = torch.Generator().manual_seed(2024)
    u8 = torch.rand(n_events, 22, dtype=torch.float64, generator=g).clamp_(1e-10, 1 - 1e-10)
    ps3 = ps_mod.PhaseSpace(E, [21, 21], [25, 6, -6], final_masses=m3, dev="cpu")
    ps8 = ps_mod.PhaseSpace(E, [21, 21], list(range(8)), final_masses=m8, dev="cpu")
    with contextlib.redirect_stdout(io.StringIO()):
        mom8, _, x1, x2 = ps8.get_momenta_from_ps(u8)
    fin8 = mom8[:, 2:].contiguous()
    prev = torch.get_num_threads()
    for label, nt in (("all_threads", threads or ncpu), ("one_thread", 1)):
        torch.set_num_threads(nt)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = _best(lambda: ps3.get_momenta_from_ps(u3), reps)
            ti = _best(lambda: ps8.generator.getPSpoint_batch(fin8, x1, x2, order=list(range(8)), target_mass=m8), reps)
        out[label] = {"threads": nt, "config1_forward_n3_points_per_s": n_events / tf,
                      "config2_inverse_n8_points_per_s": n_events / ti}
    torch.set_num_threads(prev)
    return out


def chain_step(phi, z, masses, e_cm, ps=None):
    """One step of the benchmark chain on the host with the reference's torch path: T spline inverses (torch restatement
    of zuko, oracle/rqs_torch.py) -> u -> the reference PhaseSpace.get_momenta_from_ps -> weight -> (sum w, sum w^2, n).
    phi: float tensor [T, N, F, 3K-1]; z: base points [N, F]."""
    import torch
    from . import rqs_torch
    ps_mod, _, _ = load()
    T, N, F, PW = phi.shape
    K = (PW + 1) // 3
    x = z
    lsum = torch.zeros(N, dtype=torch.float64)
    for t in range(T - 1, -1, -1):
        tr = rqs_torch.MonotonicRQSTransform(phi[t, ..., :K], phi[t, ..., K:2 * K], phi[t, ..., 2 * K:], bound=1.0)
        x = tr._inverse(x)
        lsum += tr.call_and_ladj(x)[1].double().sum(-1)   # what flow.log_prob adds: the forward log-det at the sample
    u = (x.double() + 1.0) / 2.0
    if ps is None:
        ps = ps_mod.PhaseSpace(e_cm, [21, 21], list(range(len(masses))), final_masses=torch.as_tensor(masses, dtype=torch.float64), dev="cpu")
    with contextlib.redirect_stdout(io.StringIO()):
        mom, w, x1, x2 = ps.get_momenta_from_ps(u)
    tgt = 1.0 / (2 * e_cm * e_cm) / (x1 * x2)
    wt = tgt * w / torch.exp(lsum)
    ok = (tgt > 0) & ~torch.isnan(w) & torch.isfinite(wt)
    wt = wt[ok]
    return float(wt.sum()), float((wt * wt).sum()), int(ok.sum())
