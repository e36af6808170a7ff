"""Out-of-bounds WRITE check of every kernel behind the Python API (compute-sanitizer is closed on this pool).

Every CUDA buffer the host layer allocates while the kernels run (torch.empty / empty_like / zeros: all outputs, scratch
and gradient buffers) is carved out of the middle of a larger arena whose two margins hold a canary pattern; after the
calls every margin must still hold it. Sizes are ragged on purpose (batch sizes that are no multiple of any tile, odd
bin counts): the last tile of a kernel is where a full-tile bulk store or a vector store would run past the end.
Reads past the end of an input cannot be seen this way; the parity tests at the same ragged sizes cover their effect."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

E = 13000.0
PAD = 512          # bytes on either side (keeps the 256-byte alignment torch gives, so the bulk-copy paths are the ones that run)
CANARY = 0xA5


class Arenas:
    def __init__(self):
        self.items = []
        self._empty, self._empty_like, self._zeros = torch.empty, torch.empty_like, torch.zeros

    def _carve(self, shape, dtype, device, zero=False):
        numel = int(np.prod(shape)) if len(shape) else 1
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        body = (nbytes + 255) // 256 * 256
        raw = self._empty(PAD + body + PAD, dtype=torch.uint8, device=device)
        raw.fill_(CANARY)
        out = raw[PAD:PAD + nbytes].view(dtype).view(shape)
        if zero:
            out.zero_()
        self.items.append((raw, nbytes))
        return out

    def empty(self, *size, dtype=None, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dev = torch.device(device) if device is not None else torch.device("cpu")
        if dev.type != "cuda" or kw:
            return self._empty(*size, dtype=dtype, device=device, **kw)
        return self._carve(shape, dtype or torch.get_default_dtype(), dev)

    def zeros(self, *size, dtype=None, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dev = torch.device(device) if device is not None else torch.device("cpu")
        if dev.type != "cuda" or kw:
            return self._zeros(*size, dtype=dtype, device=device, **kw)
        return self._carve(shape, dtype or torch.get_default_dtype(), dev, zero=True)

    def empty_like(self, t, **kw):
        if not t.is_cuda or kw or not t.is_contiguous():
            return self._empty_like(t, **kw)
        return self._carve(tuple(t.shape), t.dtype, t.device)

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for i, (raw, nbytes) in enumerate(self.items):
            body_end = PAD + nbytes
            lo, hi = raw[:PAD], raw[body_end:]        # (the rounding slack after the body counts as margin too)
            if not bool((lo == CANARY).all()) or not bool((hi == CANARY).all()):
                bad.append((i, nbytes, int((lo != CANARY).sum()), int((hi != CANARY).sum())))
        return bad


@pytest.fixture
def arenas(monkeypatch):
    a = Arenas()
    monkeypatch.setattr(torch, "empty", a.empty)
    monkeypatch.setattr(torch, "empty_like", a.empty_like)
    monkeypatch.setattr(torch, "zeros", a.zeros)
    yield a
    monkeypatch.undo()


MASSES = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5], 5: [4.7, 4.7, 0.1, 0.0, 0.3], 6: [4.7, 4.7, 4.7, 0.1, 0.0, 0.3],
          7: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3], 8: [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]}


def test_the_canaries_see_an_overrun(arenas):
    """Negative control: one element written past either end of a carved buffer is reported."""
    t = torch.empty((5, 3), dtype=torch.float64, device="cuda")
    assert arenas.check() == []
    flat = t.view(-1)
    torch.as_strided(flat, (1,), (1,), storage_offset=flat.storage_offset() + flat.numel()).fill_(1.0)   # one past the end
    assert len(arenas.check()) == 1
    z = torch.zeros(7, dtype=torch.int32, device="cuda")
    torch.as_strided(z, (1,), (1,), storage_offset=z.storage_offset() - 1).fill_(3)                        # one before the start
    assert len(arenas.check()) == 2


def test_phase_space_kernels_write_inside_their_buffers(arenas):
    from memflow_b200.phasespace.phasespace import PhaseSpace
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor
    for n in range(3, 9):
        ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(MASSES[n], dtype=torch.float64), dev="cuda:0")
        for N in (1, 255, 257, 3001):
            u = torch.rand(N, 3 * n - 2, dtype=torch.float64, device="cuda").clamp_(1e-9, 1 - 1e-9)
            mom, w, x1, x2 = ps.get_momenta_from_ps(u)
            ps.get_momenta_from_ps(u, pT_mincut=20.0, delR_mincut=0.4, rap_maxcut=5.0, lab_frame=True)
            ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2, order=list(range(n)))
            ur = u.clone().requires_grad_(True)
            m2, w2, _, _ = ps.get_momenta_from_ps(ur)
            (m2.sum() + w2.log().sum()).backward()
            if n == 4:
                b = torch.stack((E / 2 * (x1 + x2), 0 * x1, 0 * x1, E / 2 * (x1 - x2)), dim=1)
                Compute_ParticlesTensor.get_PS(mom[:, 2:], b, logit_eps=5e-5)
        ps32 = ps.get_momenta_from_ps(torch.rand(1001, 3 * n - 2, dtype=torch.float32, device="cuda").clamp_(1e-6, 1 - 1e-6))
        assert ps32[0].shape[0] == 1001
    assert len(arenas.items) > 100
    assert arenas.check() == []


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_spline_kernels_write_inside_their_buffers(arenas, dt):
    from memflow_b200 import transforms as TR
    for (F, K) in ((10, 40), (6, 35), (7, 10), (5, 16), (20, 64), (3, 9), (2, 5), (1, 33), (4, 24), (3, 128)):
        for B in (1, 37, 1531):
            phi = torch.randn(B, F, 3 * K - 1, device="cuda", dtype=dt, requires_grad=True)
            x = (torch.rand(B, F, device="cuda", dtype=dt) * 2.6 - 1.3)
            for fn in (TR.rqs_forward, TR.rqs_inverse):
                y, l, ls, b = fn(phi, x, bound=1.0, return_bins=True)
                phi.grad = None
                (y.sum() + ls.sum()).backward()
            if F > 2:   # a feature subset read in place (autoregressive sampling passes)
                TR.rqs_inverse(phi.detach()[:, 1:2], x[:, 1:2], bound=1.0)
    assert len(arenas.items) > 200
    assert arenas.check() == []


def test_chain_kernels_write_inside_their_buffers(arenas):
    from memflow_b200 import estimator as ES
    for (n, T, K, dt) in ((3, 2, 10, torch.float32), (4, 3, 40, torch.float32), (3, 2, 16, torch.float32), (4, 3, 40, torch.float64),
                          (3, 2, 35, torch.float64), (3, 2, 10, torch.float64)):
        smp = ES.ImportanceSampler(MASSES[n], E, seed=3, device="cuda:0")
        for N in (1, 35, 37, 5003):
            phi = torch.randn(T, N, 3 * n - 2, 3 * K - 1, device="cuda", dtype=dt)
            acc = smp.run_chunk(phi)
            acc, out = smp.run_chunk(phi, event_offset=11, debug_outputs=True)
            ES.accumulate_weights(None, out["weight"], out["logq"])
    assert len(arenas.items) > 100
    assert arenas.check() == []
