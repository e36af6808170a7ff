"""CPU: host-side mirror of the reference API (names, attributes, error behaviour) and the sharding logic."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def make_ps(n=3):
    from memflow_b200.phasespace.phasespace import PhaseSpace
    masses = {3: [125.25, 172.5, 172.5], 4: [125.25, 172.5, 172.5, 1e-5]}[n]
    return PhaseSpace(13000.0, [21, 21], [25, 6, -6, 21][:n], final_masses=torch.tensor(masses, dtype=torch.float64),
                      dev="cpu")


def test_phasespace_surface():
    """phasespace.py:37-62 attributes and rambo_generator.py:76-83 nDimPhaseSpace."""
    from memflow_b200.phasespace import phasespace, rambo_generator
    assert (phasespace.M_HIGGS, phasespace.M_TOP, phasespace.M_GLUON) == (125.25, 172.5, 1e-5)
    assert rambo_generator.M_GLUON == 0.0
    ps = make_ps(4)
    for a in ("dev", "collider_energy", "initial_pdgs", "final_pdgs", "final_masses", "final_state_mass", "pdf",
              "generator"):
        assert hasattr(ps, a)
    g = ps.generator
    assert g.nDimPhaseSpace == 8 and g.n_final == 4 and g.n_initial == 2
    assert float(g.tot_final_state_masses) == pytest.approx(125.25 + 345 + 1e-5)
    assert g.epsilon_border == 1e-10 and g.pdf_active and g.uniform_x1x2
    for m in ("generateKinematics_batch", "getPSpoint_batch", "get_flatWeights", "rho", "massless_map",
              "bisect_vec_batch", "setInitialStateMomenta_batch"):
        assert callable(getattr(g, m))
    # default masses = (H, t, t, g) like the reference, but float64
    from memflow_b200.phasespace.phasespace import PhaseSpace
    d = PhaseSpace(13000.0, [21, 21], [25, 6, -6, 21], dev="cpu")
    assert d.final_masses.dtype == torch.float64 and d.final_masses.tolist() == [125.25, 172.5, 172.5, 1e-5]


def test_generator_errors_match_reference():
    from memflow_b200.phasespace.rambo_generator import FlatInvertiblePhasespace, PhaseSpaceGeneratorError
    m = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    with pytest.raises(PhaseSpaceGeneratorError):  # rambo_generator.py:102-105
        FlatInvertiblePhasespace([0.0], m, 13000.0, dev="cpu")
    with pytest.raises(PhaseSpaceGeneratorError):  # :106-109
        FlatInvertiblePhasespace([0.0, 0.0, 0.0], m, 13000.0, dev="cpu")
    assert issubclass(PhaseSpaceGeneratorError, Exception)


def test_no_cpu_fallback():
    from memflow_b200 import MemflowB200Error
    ps = make_ps(3)
    with pytest.raises(MemflowB200Error, match="no CPU fallback"):
        ps.get_momenta_from_ps(torch.rand(4, 7, dtype=torch.float64))
    with pytest.raises(MemflowB200Error, match="no CPU fallback"):
        ps.get_ps_from_momenta(torch.rand(4, 3, 4, dtype=torch.float64), torch.rand(4), torch.rand(4))
    from memflow_b200.transforms import rqs_forward
    with pytest.raises(MemflowB200Error):
        rqs_forward(torch.randn(2, 3, 29), torch.rand(2, 3))


def test_parton_event_surface_on_cpu():
    """get_decayPartons_fromlab_propagators_angles: the reference's argument list, its NameError for the cartesian final
    scaling (utils.py:913), no CPU fallback; the coordinate helpers are plain torch and invert each other."""
    import inspect
    import numpy as np
    from memflow_b200._lib import MemflowB200Error
    from memflow_b200.unfolding_flow.utils import Compute_ParticlesTensor as CPT
    names = list(inspect.signature(CPT.get_decayPartons_fromlab_propagators_angles).parameters)
    assert names[:13] == ["propagators_kinematics", "higgs_angles", "thad_b_angles", "thad_W_angles", "tlep_b_angles",
                          "tlep_W_angles", "boost", "log_mean_parton_lab", "log_std_parton_lab", "log_mean_boost",
                          "log_std_boost", "log_mean_parton_Hthadtlep", "log_std_parton_Hthadtlep"]
    N = 4
    z3, z2 = torch.zeros(3), torch.ones(2)
    args = [torch.zeros(N, 3, 3)] + [torch.zeros(N, 2)] * 5 + [torch.zeros(N, 1, 2), z3, z3 + 1, z2, z2, z3, z3 + 1]
    with pytest.raises(NameError):
        CPT.get_decayPartons_fromlab_propagators_angles(*args, ptetaphi=False, final_scaling=True)
    with pytest.raises(MemflowB200Error):
        CPT.get_decayPartons_fromlab_propagators_angles(*args)
    g = torch.Generator().manual_seed(2)
    pep = torch.stack((torch.rand(50, generator=g, dtype=torch.float64) * 100 + 1, torch.randn(50, generator=g, dtype=torch.float64),
                       (torch.rand(50, generator=g, dtype=torch.float64) * 2 - 1) * 3.1), dim=1)
    cart = CPT.get_cartesian_comp(pep, 4.7)
    back = CPT.get_ptetaphi_comp(cart)
    assert torch.allclose(back[:, 1:], pep, rtol=1e-12, atol=1e-12)
    assert torch.allclose(back[:, 0] ** 2 - (cart[:, 1:] ** 2).sum(1), torch.full((50,), 4.7 ** 2, dtype=torch.float64), rtol=1e-9)
    assert torch.equal(CPT.get_ptetaphi_comp_batch(cart[:, None, :])[:, 0], back)
    assert torch.equal(CPT.get_ptetaphi_comp_noEnergy(cart[:, 1:]), back[:, 1:])


def test_small_closed_forms_against_oracle(oracle):
    """get_flatWeights / rho / massless_map keep the reference's formulas (:111-149)."""
    ps = make_ps(3)
    g = ps.generator
    assert g.get_flatWeights(13000, 3) == 21291.052028166854
    e = torch.tensor([500.0, 13000.0], dtype=torch.float64)
    np.testing.assert_allclose(g.get_flatWeights(e, 4).numpy(),
                               [g.get_flatWeights(500.0, 4), g.get_flatWeights(13000.0, 4)], rtol=1e-14)
    M, N, m = torch.tensor([500.0]), torch.tensor([172.5]), torch.tensor([125.25])
    r = g.rho(M.double(), N.double(), m.double())
    ref = np.sqrt((500.0 ** 2 - (172.5 + 125.25) ** 2) * (500.0 ** 2 - (172.5 - 125.25) ** 2)) / (8 * 500.0 ** 2)
    assert float(r) == pytest.approx(ref, rel=1e-15)
    x = torch.tensor([0.3], dtype=torch.float64)
    assert float(g.massless_map(x, 2)) == pytest.approx(0.3 ** 2 * (3 - 2 * 0.3))


def test_utils_match_oracle_math(oracle):
    """utils.py helpers: x1x2 map round trip, boost to rest frame, lab boost conserves mass."""
    from memflow_b200.phasespace import utils as U
    rng = np.random.default_rng(1)
    unif = torch.from_numpy(rng.random((100, 2)))
    fsm = torch.tensor(470.25, dtype=torch.float64)
    x1, x2, w = U.get_x1x2_from_uniform(unif, fsm, 13000.0)
    assert bool(((x1 * x2) >= (470.25 / 13000.0) ** 2 * (1 - 1e-12)).all())
    back, jac = U.get_uniform_from_x1x2(x1, x2, fsm, 13000.0)
    np.testing.assert_allclose(back.numpy(), unif.numpy(), atol=1e-12)
    np.testing.assert_allclose((jac * w).numpy(), 1.0, rtol=1e-13)
    # the oracle's forward uses the same map
    u = np.concatenate([rng.random((100, 5)), unif.numpy()], axis=1)
    _, _, ox1, ox2 = oracle.rambo_forward(u, [125.25, 172.5, 172.5], 13000.0)
    np.testing.assert_allclose(x1.numpy(), ox1, rtol=1e-14)
    np.testing.assert_allclose(x2.numpy(), ox2, rtol=1e-13)
    p = torch.tensor([[200.0, 30.0, -40.0, 100.0]], dtype=torch.float64)
    rest = U.boost_t(p, -U.boostVector_t(p))
    assert float(rest[0, 1:].abs().max()) < 1e-12
    assert float(rest[0, 0]) == pytest.approx(float(U.mass_t(p)), rel=1e-13)
    mom = torch.tensor([[[50.0, 0, 0, 50.0], [50.0, 0, 0, -50.0], [60.0, 10.0, 20.0, 30.0]]], dtype=torch.float64)
    lab = U.boost_to_lab_frame(mom, torch.tensor([0.2]), torch.tensor([0.05]))
    assert float(U.mass_t(lab[:, 2])) == pytest.approx(float(U.mass_t(mom[:, 2])), rel=1e-12)


def test_registered_ops_have_schemas_fake_kernels_and_no_cpu_backend():
    """torch.library layer (memflow_b200/ops.py): every op carries a schema and a fake implementation (shape propagation on
    meta tensors, no GPU needed), the differentiable ones an autograd formula, and there is NO CPU kernel to fall back to."""
    from memflow_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.memflow_b200, name)
    phi = torch.empty(4, 3, 29, device="meta")
    x = torch.empty(4, 3, device="meta")
    out, ladj, lsum, bins = torch.ops.memflow_b200.rqs(phi, x, 1.0, 1e-3, False, ops.RQS_LADJ | ops.RQS_SUM | ops.RQS_BINS)
    assert out.shape == ladj.shape == bins.shape == (4, 3) and lsum.shape == (4,) and bins.dtype == torch.int32
    out, ladj, lsum, bins = torch.ops.memflow_b200.rqs(phi, x, 1.0, 1e-3, True, 0)
    assert out.shape == (4, 3) and ladj.numel() == lsum.numel() == bins.numel() == 0
    u = torch.empty(5, 7, device="meta", dtype=torch.float64)
    mom, w, x1, x2, lw, st = torch.ops.memflow_b200.rambo_forward(u, [125.25, 172.5, 172.5], 13000.0, [], 0, True)
    assert mom.shape == (5, 5, 4) and w.shape == x1.shape == x2.shape == lw.shape == (5,) and st.dtype == torch.int32
    ps, prob, lp, st = torch.ops.memflow_b200.rambo_inverse(mom[:, 2:], x1, x2, [0, 1, 2], [125.25, 172.5, 172.5], None, 13000.0, 0, False)
    assert ps.shape == (5, 7) and prob.shape == (5,) and lp.numel() == 0
    uu, lq = torch.ops.memflow_b200.flow_sample(torch.empty(2, 8, 7, 29, device="meta"), 1, 0, 1.0, 1e-3)
    assert uu.shape == (8, 7) and lq.shape == (8,) and uu.dtype == torch.float64
    # autograd formulas are registered: gradients propagate through the fake kernels
    phi_g = torch.empty(4, 3, 29, device="meta", requires_grad=True)
    out, ladj, lsum, _ = torch.ops.memflow_b200.rqs(phi_g, x, 1.0, 1e-3, False, ops.RQS_LADJ | ops.RQS_SUM)
    (out.sum() + lsum.sum()).backward()
    assert phi_g.grad.shape == phi_g.shape
    ug = torch.empty(5, 7, device="meta", dtype=torch.float64, requires_grad=True)
    mom = torch.ops.memflow_b200.rambo_forward(ug, [125.25, 172.5, 172.5], 13000.0, [], 0, False)[0]
    mom.sum().backward()
    assert ug.grad.shape == ug.shape
    # no CPU backend: the dispatcher refuses
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.memflow_b200.rqs(torch.zeros(4, 3, 29), torch.zeros(4, 3), 1.0, 1e-3, False, 0)
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.memflow_b200.rambo_forward(torch.zeros(5, 7, dtype=torch.float64), [1.0, 2.0, 3.0], 13000.0, [], 0, False)


def test_transform_packing_is_zero_copy_also_under_autograd():
    """zuko hands the univariate transform three split views of ONE hyper-network output: the parameter block is used in
    place, with and without autograd (a view of the common base: no torch.cat on the way in, no gradient split back)."""
    from memflow_b200 import transforms as TR
    B, F, K = 6, 3, 5
    hyper = torch.randn(B, F * (3 * K - 1), requires_grad=True) * 1.0        # non-leaf, like a network output
    w, h, d = hyper.unflatten(-1, (F, 3 * K - 1)).split([K, K, K - 1], dim=-1)
    p, lead = TR._pack(w, h, d)
    assert p.shape == (B * F, 1, 3 * K - 1) and tuple(lead) == (B, F) and p.data_ptr() == hyper.data_ptr() and p.requires_grad
    p2, _ = TR._pack(w.detach(), h.detach(), d.detach())
    assert p2.data_ptr() == hyper.data_ptr() and not p2.requires_grad
    p3, _ = TR._pack(w.clone(), h.clone(), d.clone())                          # separate tensors: one concatenation
    assert p3.data_ptr() != hyper.data_ptr() and p3.requires_grad and torch.equal(p3.detach().reshape(B, F, -1), hyper.detach().reshape(B, F, -1))


def test_shard_and_chunk_ranges():
    from memflow_b200.distributed import chunk_ranges, shard_range
    for n_total in (0, 1, 7, 10 ** 10, 12345678901):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n_total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n_total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1
    assert chunk_ranges(5, 17, 5) == [(5, 10), (10, 15), (15, 17)]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class _OracleSampler:
    """CPU stand-in for ImportanceSampler.run_chunk (the oracle's estimator as the per-chunk engine): lets the host logic of
    ShardedIntegrator (chunking, checkpoint / resume) run without a GPU."""

    def __init__(self, jac, lq, seed=7):
        self.jac, self.lq, self.seed, self.device, self.calls = jac, lq, seed, torch.device("cpu"), []

    def run_chunk(self, phi, event_offset=0, acc=None):
        from oracle import oracle as O
        b, e = event_offset, event_offset + phi
        self.calls.append((b, e))
        acc += torch.from_numpy(O.is_reduce(None, self.jac[b:e], self.lq[b:e]))
        return acc


def test_restartable_accumulator(tmp_path, oracle):
    """SURVEY section 5, checkpoint row: a run killed after some chunks and resumed from its state file processes every
    chunk exactly once and ends on bit-identical accumulators; a state file of another run is refused."""
    from memflow_b200.distributed import ShardedIntegrator
    rng = np.random.default_rng(3)
    n = 1000
    jac, lq = rng.random(n) * 3, rng.standard_normal(n)
    full = ShardedIntegrator(_OracleSampler(jac, lq), n, 128).run(lambda b, e: e - b, acc=torch.zeros(3, dtype=torch.float64))
    path = str(tmp_path / "state.json")

    class Killed(Exception):
        pass

    smp = _OracleSampler(jac, lq)

    def provider(b, e):
        if len(smp.calls) == 5:
            raise Killed()
        return e - b

    it = ShardedIntegrator(smp, n, 128)
    with pytest.raises(Killed):
        it.run(provider, acc=torch.zeros(3, dtype=torch.float64), state_path=path, save_every=2)
    assert os.path.exists(path) and len(smp.calls) == 5          # chunks 0..4 done, the last save was after chunk 3
    smp2 = _OracleSampler(jac, lq)
    res = ShardedIntegrator(smp2, n, 128).run(lambda b, e: e - b, acc=torch.zeros(3, dtype=torch.float64), state_path=path, save_every=2)
    assert smp2.calls[0] == (512, 640) and smp2.calls[-1] == (896, 1000)   # resumed after the last SAVED chunk
    assert torch.equal(res, full)                                   # bit-identical (hex-encoded doubles in the state file)
    done = ShardedIntegrator(_OracleSampler(jac, lq), n, 128)
    s3 = done.sampler
    assert torch.equal(done.run(lambda b, e: e - b, acc=torch.zeros(3, dtype=torch.float64), state_path=path), full) and s3.calls == []
    with pytest.raises(ValueError):                                 # another chunking: not this run's checkpoint
        ShardedIntegrator(_OracleSampler(jac, lq), n, 100).run(lambda b, e: e - b, acc=torch.zeros(3, dtype=torch.float64), state_path=path)
    with pytest.raises(ValueError):                                 # another seed
        ShardedIntegrator(_OracleSampler(jac, lq, seed=8), n, 128).run(lambda b, e: e - b, acc=torch.zeros(3, dtype=torch.float64), state_path=path)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from memflow_b200.distributed import allreduce_accumulators, chunk_ranges, shard_range
    from oracle import oracle as O
    # every rank integrates its shard with the ORACLE as the per-chunk engine (CPU stand-in for the kernel);
    # the test checks the sharding + single all-reduce logic, which is what runs on N GPUs.
    n_total = 1000
    rng = np.random.default_rng(0)
    jac = rng.random(n_total) * 3
    lq = rng.standard_normal(n_total)
    b, e = shard_range(n_total, rank, world)
    acc = torch.zeros(3, dtype=torch.float64)
    for cb, ce in chunk_ranges(b, e, 128):
        acc += torch.from_numpy(O.is_reduce(None, jac[cb:ce], lq[cb:ce]))
    allreduce_accumulators(acc)
    q.put((rank, acc.tolist()))
    dist.destroy_process_group()


def test_world_size_2_gloo_allreduce(oracle):
    """N>1 path on CPU: 2 gloo ranks, disjoint shards, one all-reduce == single-process result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    jac = rng.random(1000) * 3
    lq = rng.standard_normal(1000)
    ref = oracle.is_reduce(None, jac, lq)
    for _, acc in res:
        np.testing.assert_allclose(acc, ref, rtol=1e-13)
    assert res[0][1] == res[1][1]


def test_call_path_selection():
    """Which way a call takes into the kernels (memflow_b200/_lib.py): plain eager inference -> C ABI directly; plain eager
    autograd -> the directly-applied autograd.Function; functorch transforms / tensor subclasses -> the registered op."""
    import torch
    from memflow_b200._lib import needs_dispatcher, plain_eager
    t = torch.zeros(3)
    assert plain_eager(t) and not needs_dispatcher(t)
    g = torch.zeros(3, requires_grad=True)
    assert needs_dispatcher(g) and plain_eager(g)
    with torch.no_grad():
        assert not needs_dispatcher(g)
    seen = []

    def f(x):
        seen.append((needs_dispatcher(x), plain_eager(x)))
        return (x * x).sum()

    torch.func.grad(f)(torch.ones(3))
    assert seen == [(True, False)]
