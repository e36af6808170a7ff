"""GPU parity of K5 (IS weight + estimator reduction) and of the fused chain kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
E = 13000.0
MASSES = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5]}


def test_is_reduce_vs_oracle(oracle):
    from memflow_b200 import estimator as ES
    for N in (1, 1000, 3_000_001):
        rng = np.random.default_rng(N)
        tg = rng.random(N); tg[::1000] = -1.0
        jac = rng.random(N) * 1e3; jac[::777] = np.nan
        lq = rng.standard_normal(N)
        acc = ES.accumulate_weights(torch.from_numpy(tg).cuda(), torch.from_numpy(jac).cuda(), torch.from_numpy(lq).cuda())
        ref = oracle.is_reduce(tg, jac, lq)
        a = acc.cpu().numpy()
        assert a[2] == ref[2]                                          # counts bit-exact
        if ref[2] > 0:
            np.testing.assert_allclose(a[:2], ref[:2], rtol=1e-12)
        acc2 = ES.accumulate_weights(torch.from_numpy(tg).cuda(), torch.from_numpy(jac).cuda(), torch.from_numpy(lq).cuda())
        assert torch.equal(acc, acc2)                                  # deterministic reduction order
        ES.accumulate_weights(torch.from_numpy(tg).cuda(), torch.from_numpy(jac).cuda(), torch.from_numpy(lq).cuda(), acc=acc)
        np.testing.assert_allclose(acc.cpu().numpy(), 2 * a, rtol=1e-15)  # accumulates into acc
    I, s, nv = ES.mc_estimate(acc, 3_000_001)
    Io, so = oracle.estimator(acc.cpu().numpy(), 3_000_001)
    assert I == pytest.approx(Io, rel=1e-15) and s == pytest.approx(so, rel=1e-12)
    acc = ES.accumulate_weights(None, torch.ones(10, device="cuda", dtype=torch.float64), torch.zeros(10, device="cuda", dtype=torch.float64))
    assert acc.tolist() == [10.0, 10.0, 10.0]


@pytest.mark.parametrize("n,T,K", [(3, 2, 10), (4, 3, 8), (3, 1, 16), (4, 2, 40), (3, 2, 64), (3, 2, 9), (3, 2, 32), (4, 2, 48)])
def test_chain_vs_oracle_composition(oracle, n, T, K):
    """Fused chain == Philox -> K4 x T -> affine -> K1 -> weight -> K5 done step by step with the oracle."""
    from memflow_b200 import estimator as ES
    from oracle import philox as OP
    from tests import parity
    F = 3 * n - 2
    N = 20_000 + 3
    g = torch.Generator().manual_seed(11)
    phi = torch.randn(T, N, F, 3 * K - 1, generator=g, dtype=torch.float32)
    smp = ES.ImportanceSampler(MASSES[n], E, bound=1.0, seed=1234, device="cuda:0")
    acc, out = smp.run_chunk(phi.cuda(), event_offset=17, debug_outputs=True)
    z = OP.base_sample(N, F, 1.0, 1234, 17)
    xcur, lsum = z.copy(), np.zeros(N)
    for t in range(T - 1, -1, -1):
        xcur, l, _ = oracle.rqs_inverse(phi[t].numpy(), xcur, bound=1.0)
        lsum += l.astype(np.float64).sum(1)
    uo = (xcur.astype(np.float64) + 1.0) / 2.0
    ug = out["u"].cpu().numpy()
    # spline part is fp32: medians at fp32 precision, tails limited by narrow bins (tests/test_rqs_gpu.py)
    assert np.median(np.abs(ug - uo)) < 2e-7 and np.quantile(np.abs(ug - uo), 0.999) < 1e-4
    dq = np.abs(out["logq"].cpu().numpy() - lsum)
    ksc = max(1.0, K / 10.0)   # narrower bins -> larger fp32 rounding of every log-det (both sides are fp32 here)
    assert np.median(dq) < 2e-5 * ksc and np.quantile(dq, 0.99) < 2e-3 * ksc
    # phase-space part is fp64: feed the oracle the kernel's own u
    om, ow, ox1, ox2 = oracle.rambo_forward(ug, MASSES[n], E)
    kappa = parity.conditioning(ug, om, n, MASSES[n])
    err = parity.momenta_error(out["momenta"].cpu().numpy(), om, ox1, ox2, E)
    assert np.nanmax(err / kappa) < 1e-12 and np.nanquantile(err, 0.99) < 1e-12
    dl = np.abs(np.log(out["weight"].cpu().numpy()) - np.log(ow))
    assert np.nanmax(dl / kappa) < 1e-12 and np.nanquantile(dl, 0.5) < 1e-12
    tgt = 1.0 / (2 * E * E) / (ox1 * ox2)
    ref = oracle.is_reduce(tgt, out["weight"].cpu().numpy(), out["logq"].cpu().numpy())
    a = acc.cpu().numpy()
    assert a[2] == ref[2]
    np.testing.assert_allclose(a[:2], ref[:2], rtol=1e-11)


# K = 35: the most common bin count of the reference's configs; odd K takes the rotated passes of the spline kernel,
# whose start follows the GLOBAL row index so that chunking stays bit-invariant
@pytest.mark.parametrize("n,T,K", [(3, 2, 10), (4, 3, 40), (3, 2, 35)])
def test_chain_with_float64_flow_parameters(oracle, n, T, K):
    """MEMFlow's configs run the flows in float64: the same chain on the fp64 spline kernel (mf_flow_sample_f64) against
    the step-by-step oracle composition, at fp64 tolerances; chunk / offset invariant like the fp32 chain."""
    from memflow_b200 import estimator as ES
    from oracle import philox as OP
    F = 3 * n - 2
    N = 10_000 + 7
    g = torch.Generator().manual_seed(5)
    phi = torch.randn(T, N, F, 3 * K - 1, generator=g, dtype=torch.float64)
    smp = ES.ImportanceSampler(MASSES[n], E, bound=1.0, seed=99, device="cuda:0")
    acc, out = smp.run_chunk(phi.cuda(), event_offset=5, debug_outputs=True)
    xcur = OP.base_sample(N, F, 1.0, 99, 5).astype(np.float64)
    lsum = np.zeros(N)
    for t in range(T - 1, -1, -1):
        xcur, l, _ = oracle.rqs_inverse(phi[t].numpy(), xcur, bound=1.0)
        lsum += l.sum(1)
    uo = (xcur + 1.0) / 2.0
    ug = out["u"].cpu().numpy()
    # T chained inverses: each amplifies the previous one's rounding by its slope (soft-clipped to <= 1e3)
    assert np.median(np.abs(ug - uo)) < 1e-15 and np.quantile(np.abs(ug - uo), 0.999) < 1e-11
    dq = np.abs(out["logq"].cpu().numpy() - lsum)
    assert np.median(dq) < 1e-12 and np.quantile(dq, 0.999) < 1e-9   # a sum of T F log-dets of O(1..10)
    om, ow, ox1, ox2 = oracle.rambo_forward(ug, MASSES[n], E)
    tgt = 1.0 / (2 * E * E) / (ox1 * ox2)
    ref = oracle.is_reduce(tgt, out["weight"].cpu().numpy(), out["logq"].cpu().numpy())
    a = acc.cpu().numpy()
    assert a[2] == ref[2]
    np.testing.assert_allclose(a[:2], ref[:2], rtol=1e-11)
    # two half-chunks == one chunk, per event
    h = N // 2
    u1, q1 = smp.sample_flow(phi[:, :h].contiguous().cuda(), event_offset=5)
    u2, q2 = smp.sample_flow(phi[:, h:].contiguous().cuda(), event_offset=5 + h)
    assert torch.equal(torch.cat((u1, u2)), out["u"]) and torch.equal(torch.cat((q1, q2)), out["logq"])


def test_chain_is_chunk_and_offset_invariant():
    """Philox counter = global event index: splitting a range into chunks gives the same per-event values,
    and the accumulators add up (what makes the multi-GPU shards reproducible)."""
    from memflow_b200 import estimator as ES
    n, T, K, N = 3, 2, 10, 8192
    phi = torch.randn(T, N, 7, 3 * K - 1, device="cuda", dtype=torch.float32)
    smp = ES.ImportanceSampler(MASSES[n], E, seed=5, device="cuda:0")
    acc, out = smp.run_chunk(phi, event_offset=100, debug_outputs=True)
    h = N // 2
    acc1, o1 = smp.run_chunk(phi[:, :h].contiguous(), event_offset=100, debug_outputs=True)
    acc2, o2 = smp.run_chunk(phi[:, h:].contiguous(), event_offset=100 + h, debug_outputs=True)
    assert torch.equal(torch.cat((o1["u"], o2["u"])), out["u"])
    assert torch.equal(torch.cat((o1["weight"], o2["weight"])), out["weight"])
    np.testing.assert_allclose((acc1 + acc2).cpu().numpy(), acc.cpu().numpy(), rtol=1e-12)
    assert not torch.equal(smp.run_chunk(phi, event_offset=101, debug_outputs=True)[1]["u"], out["u"])
    # estimator sanity: flat flow (phi = 0 -> identity-like spline), integral of 1/(2 x1 x2 s) dPS is finite and positive
    I, s, nv = ES.mc_estimate(acc, N)
    assert nv == N and I > 0 and s > 0


def test_sharded_integrator_single_rank():
    from memflow_b200 import estimator as ES
    from memflow_b200.distributed import ShardedIntegrator
    n, T, K = 3, 2, 10
    buf = torch.randn(T, 4096, 7, 3 * K - 1, device="cuda", dtype=torch.float32)
    smp = ES.ImportanceSampler(MASSES[n], E, seed=9, device="cuda:0")
    it = ShardedIntegrator(smp, n_total=10_000, chunk=4096)
    acc = it.run(lambda b, e: buf[:, : e - b])
    assert acc[2].item() == 10_000


def test_two_stage_calls_equal_chain_call():
    """mf_flow_sample_f32 + mf_rambo_is_f64 called separately == mf_is_chain_f64 (bit-identical accumulators)."""
    from memflow_b200 import estimator as ES
    n, T, K, N = 4, 2, 10, 10_000
    phi = torch.randn(T, N, 10, 3 * K - 1, device="cuda", dtype=torch.float32)
    smp = ES.ImportanceSampler(MASSES[n], E, seed=77, device="cuda:0")
    acc, out = smp.run_chunk(phi, event_offset=5, debug_outputs=True)
    u, lq = smp.sample_flow(phi, event_offset=5)
    assert torch.equal(u, out["u"]) and torch.equal(lq, out["logq"])
    assert bool(((u >= 0) & (u <= 1)).all())
    acc2 = smp.weigh(u, lq)
    assert torch.equal(acc, acc2)


def test_fused_single_kernel_mode_agrees():
    """MF_CHAIN_MODE=fused (one launch, nothing but the accumulators leaves the SM; kept as the measured alternative of
    DESIGN.md section 4) gives the same estimator inputs as the default two-launch chain."""
    import json, os, subprocess, sys
    code = (
        "import sys, json, torch; sys.path.insert(0, %r);"
        "from memflow_b200 import estimator as ES;"
        "g = torch.Generator().manual_seed(5);"
        "phi = torch.randn(2, 6000, 7, 29, generator=g, dtype=torch.float32).cuda();"
        "smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, seed=3, device='cuda:0');"
        "acc, out = smp.run_chunk(phi, event_offset=9, debug_outputs=True);"
        "print(json.dumps({'acc': acc.tolist(), 'u': out['u'].sum().item(), 'w': out['weight'].log().sum().item()}))"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("split", "fused"):
        env = dict(os.environ, MF_CHAIN_MODE=mode)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        res[mode] = json.loads(out.stdout.strip().splitlines()[-1])
    a, b = res["split"], res["fused"]
    assert a["acc"][2] == b["acc"][2] == 6000
    np.testing.assert_allclose(a["acc"][:2], b["acc"][:2], rtol=1e-11)
    assert abs(a["u"] - b["u"]) < 1e-9 and abs(a["w"] - b["w"]) < 1e-7
