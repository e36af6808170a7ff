"""The hot-path calls are CUDA-graph capturable (no host synchronisation, no pageable copies, no stream-ordered
allocation outside torch's allocator on the default paths): training-size batches (B ~ 1e3) are launch-bound, a captured
step replays the phase-space map, the spline (forward and backward) and the chain without any Python between launches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
E = 13000.0


def _capture(fn):
    from memflow_b200.graphs import capture
    step = capture(fn)
    return step.graph, step.outputs


def test_phase_space_maps_in_a_graph():
    from memflow_b200.phasespace.phasespace import PhaseSpace
    ps = PhaseSpace(E, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor([125.25, 172.5, 172.5, 1e-5], dtype=torch.float64), dev="cuda:0")
    u = torch.rand(1024, 10, dtype=torch.float64, device="cuda").clamp_(1e-6, 1 - 1e-6)
    def step():
        mom, w, x1, x2 = ps.get_momenta_from_ps(u)
        back, prob = ps.get_ps_from_momenta(mom[:, 2:], x1, x2)
        return mom, w, back, prob
    ref = [t.clone() for t in step()]
    g, out = _capture(step)
    u2 = torch.rand(1024, 10, dtype=torch.float64, device="cuda").clamp_(1e-6, 1 - 1e-6)
    u.copy_(u2)                      # new inputs in the captured buffer
    g.replay()
    torch.cuda.synchronize()
    mom2, w2, x12, x22 = ps.get_momenta_from_ps(u2)
    assert torch.equal(out[0], mom2) and torch.equal(out[1], w2)
    assert not torch.equal(out[0], ref[0])
    assert float((out[2] - u2).abs().max()) < 1e-9


@pytest.mark.parametrize("dt,F,K", [(torch.float64, 10, 40), (torch.float64, 6, 35), (torch.float32, 7, 10)])
def test_spline_forward_and_backward_in_a_graph(dt, F, K):
    from memflow_b200 import transforms as TR
    B = 1024
    phi = torch.randn(B, F, 3 * K - 1, device="cuda", dtype=dt, requires_grad=True)
    x = (torch.rand(B, F, device="cuda", dtype=dt) * 2.4 - 1.2)
    def step():
        y, ladj, lsum = TR.rqs_forward(phi, x, bound=1.0)
        (gphi,) = torch.autograd.grad((y.sum() + lsum.sum()), (phi,))
        return y, lsum, gphi
    g, out = _capture(step)
    with torch.no_grad():
        phi.copy_(torch.randn_like(phi))
        x.copy_(torch.rand_like(x) * 2.4 - 1.2)
    g.replay()
    torch.cuda.synchronize()
    y, lsum, gphi = step()
    assert torch.equal(out[0], y) and torch.equal(out[1], lsum) and torch.equal(out[2], gphi)
    assert bool(torch.isfinite(out[2]).all()) and float(out[2].abs().sum()) > 0


def test_chain_in_a_graph():
    from memflow_b200 import estimator as ES
    N, T, Fd, K = 4096, 2, 7, 10
    phi = torch.randn(T, N, Fd, 3 * K - 1, device="cuda", dtype=torch.float32)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5], E, seed=9, device="cuda:0")
    acc = ES.new_accumulator("cuda:0")
    def step():
        smp.run_chunk(phi, event_offset=0, acc=acc)
        return acc
    g, _ = _capture(step)
    acc.zero_()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    two = acc.clone()
    ref = ES.new_accumulator("cuda:0")
    smp.run_chunk(phi, event_offset=0, acc=ref); smp.run_chunk(phi, event_offset=0, acc=ref)
    torch.cuda.synchronize()
    np.testing.assert_allclose(two.cpu().numpy(), ref.cpu().numpy(), rtol=1e-13)
    assert float(two[2]) > 0
