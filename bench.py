#!/usr/bin/env python
"""bench.py -- headline benchmark of the phase-space / importance-sampling hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], "Full importance-sampling chain ... on 1 GPU", sharded as configs[4] for
N > 1): one STEP = one chunk of 2^22 ttH points through the fused chain
    Philox base sample -> 2 RQ-spline inverses (F=7, K=10, per-event parameters streamed from HBM)
    -> PhaseSpace map u -> momenta (n=3: H, t, tbar at sqrt(s) = 13 TeV, fp64) + Jacobian
    -> weight w = jac / (2 x1 x2 s) / q(u) -> (sum w, sum w^2, n_valid),
two kernel launches per step (memflow_b200/csrc/is_chain.cu: A = flow sample, B = phase space + weight + reduce),
one all-reduce of the 3 accumulators at the end.
Metric: phase-space points/sec (map + Jacobian + weights), whole job.

Prints ONE JSON line (rank 0). Keys follow the driver's contract; see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout. Libraries (NCCL prints its version banner there) must not add
# lines, so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


METRIC = "phase-space points/sec (map+Jacobian+weights)"
UNIT = "points/s"
E_CM = 13000.0
MASSES = [125.25, 172.5, 172.5]
N_FINAL, T_FLOW, K_BINS, F_DIM = 3, 2, 10, 7
PW = 3 * K_BINS - 1
CHUNK = 1 << 22
BYTES_PER_EVENT = T_FLOW * F_DIM * PW * 4  # algorithmic: only the spline parameters touch HBM
CONFIG = {"workload": "is_chain_ttH: philox -> 2x RQS inverse (F=7,K=10,fp32) -> RAMBO n=3 fp64 -> weight -> sum w, sum w^2",
          "chunk_events": CHUNK, "n_final": N_FINAL, "flow": {"T": T_FLOW, "F": F_DIM, "K": K_BINS, "bound": 1.0},
          "sqrt_s": E_CM, "masses": MASSES, "me2": 1, "pdf": 1,
          "l2_policy": "inputs larger than L2 (6.8 GB of parameters per step)", "parallelism": "events sharded, 1 allreduce"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (2 ms period; the timed
    region of the default run lasts tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        import threading
        self.samples, self.reasons, self.max_mhz, self.p, self.f = [], set(), None, None, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = int(get_reasons(h))
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            self.source = "nvml"
        except Exception:
            self.source = "nvidia-smi"
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                           "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                          stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": self.source}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            sm = sorted(self.samples)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
            out["reasons"] = sorted(self.reasons)
            out["samples"] = len(sm)
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def cpu_chain(O, OP, phi, seed, event_offset):
    """The chain on the host with the oracle port (the reference's algorithm restated in C, oracle/)."""
    import numpy as np
    n = phi.shape[1]
    z = OP.base_sample(n, F_DIM, 1.0, seed, event_offset)
    lsum = np.zeros(n)
    for t in range(T_FLOW - 1, -1, -1):
        z, l, _ = O.rqs_inverse(phi[t], z, bound=1.0)
        lsum += l.astype(np.float64).sum(1)
    u = (z.astype(np.float64) + 1.0) / 2.0
    mom, w, x1, x2 = O.rambo_forward(u, MASSES, E_CM)
    tgt = 1.0 / (2 * E_CM * E_CM) / (x1 * x2)
    return O.is_reduce(tgt, w, lsum)


def cpu_leg(steps, warmup, budget_s):
    """Times the CPU port on a bounded sample; returns (points/s, cores, sample description, ms/step)."""
    import numpy as np
    from oracle import oracle as O, philox as OP
    O.build()
    cores = O.max_threads()
    O.set_threads(cores)
    rng = np.random.default_rng(0)
    probe = 1 << 13
    phi = rng.standard_normal((T_FLOW, probe, F_DIM, PW)).astype(np.float32)
    cpu_chain(O, OP, phi, 1, 0)
    t0 = time.perf_counter()
    cpu_chain(O, OP, phi, 1, 0)
    per_ev = (time.perf_counter() - t0) / probe
    n = int(budget_s / max(steps + warmup, 1) / per_ev)
    n = max(1 << 12, min(n, 1 << 20))
    phi = rng.standard_normal((T_FLOW, n, F_DIM, PW)).astype(np.float32)
    for i in range(warmup):
        cpu_chain(O, OP, phi, 1, i * n)
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_chain(O, OP, phi, 1, (warmup + i) * n)
    dt = time.perf_counter() - t0
    return steps * n / dt, cores, "%d steps x %d points of the same chain (oracle port: C, pthreads)" % (steps, n), dt / steps * 1e3


def reference_torch_leg(steps, warmup, budget_s):
    """The reference's OWN torch CPU path on the host cores (oracle/_ref = the unmodified memflow/phasespace files staged by
    oracle/make_ref.py; zuko is absent from the reference tree, so the spline step is the torch restatement
    oracle/rqs_torch.py): the same chain on a bounded sample. Returns None when oracle/_ref is not staged."""
    from oracle import ref_torch as R
    if not R.available():
        return None
    import numpy as np
    import torch
    from oracle import philox as OP
    ps_mod, _, _ = R.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ps = ps_mod.PhaseSpace(E_CM, [21, 21], [25, 6, -6], final_masses=torch.tensor(MASSES, dtype=torch.float64), dev="cpu")
    g = torch.Generator().manual_seed(0)

    def run(n, off):
        phi = torch.randn(T_FLOW, n, F_DIM, PW, generator=g, dtype=torch.float32)
        z = torch.from_numpy(OP.base_sample(n, F_DIM, 1.0, 1, off))
        t0 = time.perf_counter()
        R.chain_step(phi, z, MASSES, E_CM, ps=ps)
        return time.perf_counter() - t0

    probe = 1 << 14
    run(probe, 0)
    per_ev = run(probe, 0) / probe
    n = int(budget_s / max(steps + warmup, 1) / per_ev)
    n = max(1 << 12, min(n, 1 << 20))
    for i in range(warmup):
        run(n, i * n)
    dt = sum(run(n, (warmup + i) * n) for i in range(steps))
    return {"value": steps * n / dt, "cores": cores, "ms_per_step": dt / steps * 1e3, "cpu": R.cpu_model(),
            "sample": "%d steps x %d points of the same chain: torch CPU, %d threads; phase space = unmodified reference "
                      "PhaseSpace.get_momenta_from_ps (oracle/_ref), spline = torch restatement of zuko (zuko is absent)"
                      % (steps, n, cores)}


def cpu_baseline_block(steps, warmup, budget_s, with_phase_space=True):
    """cpu_baseline of the JSON line: the reference torch path when it is staged (kind "reference"), with the C port's number
    next to it; the port alone (kind "port") otherwise."""
    v, cores, sample, ms = cpu_leg(steps, warmup, budget_s=min(budget_s, 20.0))
    port = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    ref = None
    try:
        ref = reference_torch_leg(steps, warmup, budget_s)
    except Exception as e:  # the reference leg must never break the line
        port["reference_error"] = repr(e)
    if ref is None:
        return port, ms
    block = {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": "reference", "sample": ref["sample"],
             "cpu": ref["cpu"], "port": port}
    if with_phase_space:
        try:
            from oracle import ref_torch as R
            block["reference_torch"] = R.time_phase_space(1_000_000, reps=2)
        except Exception as e:
            block["reference_torch"] = {"error": repr(e)}
    return block, ref["ms_per_step"]


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (oracle/_ref: its torch
    phase-space files, unmodified; the spline step is the torch restatement of the absent zuko). Falls back to the C port
    when oracle/_ref is not staged. Rank 0 only."""
    if rank != 0:
        return
    block, ms = cpu_baseline_block(args.steps, args.warmup, budget_s=100.0, with_phase_space=False)
    v = block["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": CONFIG, "cpu_baseline": block,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def extras(torch, dev, hbm_peak):
    """Per-kernel rates for the other BASELINE configs (device-resident, CUDA events): K1, K2, K3, K4."""
    from memflow_b200.phasespace.phasespace import PhaseSpace
    from memflow_b200 import transforms as TR

    def timeit(f, n=5, warm=2):
        for _ in range(warm):
            f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            f()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e-3

    out = {}
    try:
        N = 1 << 24
        ps3 = PhaseSpace(E_CM, [21, 21], [25, 6, -6], final_masses=torch.tensor(MASSES, dtype=torch.float64), dev=dev)
        u = torch.rand(N, 7, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
        t = timeit(lambda: ps3.get_momenta_from_ps(u))
        out["K1_forward_n3_f64_16M"] = {"points_per_s": N / t, "GBps": 240 * N / t / 1e9, "frac_hbm": 240 * N / t / 1e9 / hbm_peak}
        del u
        m8 = [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3]
        ps8 = PhaseSpace(E_CM, [21, 21], list(range(8)), final_masses=torch.tensor(m8, dtype=torch.float64), dev=dev)
        u = torch.rand(N, 22, dtype=torch.float64, device=dev).clamp_(1e-10, 1 - 1e-10)
        tf = timeit(lambda: ps8.get_momenta_from_ps(u))
        mom, w, x1, x2 = ps8.get_momenta_from_ps(u)
        ti = timeit(lambda: ps8.generator.getPSpoint_batch(mom[:, 2:], x1, x2))
        out["K1_forward_n8_f64_16M"] = {"points_per_s": N / tf, "GBps": 520 * N / tf / 1e9, "frac_hbm": 520 * N / tf / 1e9 / hbm_peak}
        out["K2_inverse_n8_f64_16M"] = {"points_per_s": N / ti, "GBps": 456 * N / ti / 1e9, "frac_hbm": 456 * N / ti / 1e9 / hbm_peak}
        out["config2_round_trip_n8_16M"] = {"points_per_s": N / (tf + ti), "GBps": 976 * N / (tf + ti) / 1e9,
                                            "frac_hbm": 976 * N / (tf + ti) / 1e9 / hbm_peak}
        del u, mom, w, x1, x2
        B, F, K = 1 << 20, 20, 64
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float32)
        x = (torch.rand(B, F, device=dev) * 2.6 - 1.3) * 1.1
        by = (F * (3 * K - 1) + 2 * F + 1) * 4
        tf = timeit(lambda: TR.rqs_forward(phi, x, bound=1.1))
        ti = timeit(lambda: TR.rqs_inverse(phi, x, bound=1.1))
        out["K3_rqs_forward_F20_K64_f32_1M"] = {"points_per_s": B / tf, "GBps": by * B / tf / 1e9, "frac_hbm": by * B / tf / 1e9 / hbm_peak}
        out["K4_rqs_inverse_F20_K64_f32_1M"] = {"points_per_s": B / ti, "GBps": by * B / ti / 1e9, "frac_hbm": by * B / ti / 1e9 / hbm_peak}
        del phi, x
        torch.cuda.empty_cache()
        # the flows of the reference's configs run in float64 (production shape F=10, K=40): forward, and the
        # backward of the training loss (parameters read + gradients written)
        B, F, K = 1 << 18, 10, 40
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64, requires_grad=True)
        x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
        by = (F * (3 * K - 1) + 2 * F + 1) * 8
        tf = timeit(lambda: TR.rqs_forward(phi.detach(), x, bound=1.1))
        out["K3_rqs_forward_F10_K40_f64_256k"] = {"points_per_s": B / tf, "GBps": by * B / tf / 1e9, "frac_hbm": by * B / tf / 1e9 / hbm_peak}
        loss = TR.rqs_forward(phi, x, bound=1.1)[2].sum()
        def bwd():
            phi.grad = None
            loss.backward(retain_graph=True)
        tb = timeit(bwd)
        byb = 2 * F * (3 * K - 1) * 8
        out["K3_rqs_backward_F10_K40_f64_256k"] = {"points_per_s": B / tb, "GBps": byb * B / tb / 1e9, "frac_hbm": byb * B / tb / 1e9 / hbm_peak}
        del phi, x, loss
        torch.cuda.empty_cache()
        # the most common flow shape of the reference's configs: float64, 10 features, 35 bins (odd K: even row pitch)
        B, F, K = 1 << 18, 10, 35
        phi = torch.randn(B, F, 3 * K - 1, device=dev, dtype=torch.float64)
        x = (torch.rand(B, F, device=dev, dtype=torch.float64) * 2.6 - 1.3) * 1.1
        by = (F * (3 * K - 1) + 2 * F + 1) * 8
        tf = timeit(lambda: TR.rqs_forward(phi, x, bound=1.1))
        out["K3_rqs_forward_F10_K35_f64_256k"] = {"points_per_s": B / tf, "GBps": by * B / tf / 1e9, "frac_hbm": by * B / tf / 1e9 / hbm_peak}
        # the same parameters through the zuko-shaped Transform API: MonotonicRQSTransform(widths, heights, derivatives) on the
        # three split views of one hyper-network output (rows [B*F, 1, 3K-1], packed without a copy), y and log|dy/dx|
        wv, hv, dv = phi.split((K, K, K - 1), dim=-1)
        tt = timeit(lambda: TR.MonotonicRQSTransform(wv, hv, dv, bound=1.1).call_and_ladj(x))
        out["K3_transform_api_rows_F1_K35_f64_2.6M"] = {"rows_per_s": B * F / tt, "points_per_s": B / tt, "GBps": by * B / tt / 1e9,
                                                         "frac_hbm": by * B / tt / 1e9 / hbm_peak}
        del phi, x, wv, hv, dv
        torch.cuda.empty_cache()
        # BASELINE config 5 shape on this GPU: tt-bar-H + 1 jet (n=4), production flow F=10, K=40, T=6
        from memflow_b200 import estimator as ES
        T5, N5, F5, K5 = 6, 1 << 19, 10, 40
        phi5 = torch.randn(T5, N5, F5, 3 * K5 - 1, device=dev, dtype=torch.float32)
        smp5 = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], E_CM, seed=3, device=dev)
        u5 = torch.empty(N5, F5, dtype=torch.float64, device=dev)
        lq5 = torch.empty(N5, dtype=torch.float64, device=dev)
        acc5 = ES.new_accumulator(dev)
        ta = timeit(lambda: smp5.sample_flow(phi5, u=u5, logq=lq5))
        tb = timeit(lambda: smp5.weigh(u5, lq5, acc=acc5))
        by5 = T5 * F5 * (3 * K5 - 1) * 4
        out["chain_config5_shape_n4_F10_K40_T6_512k"] = {
            "points_per_s": N5 / (ta + tb), "flow_sample_GBps": by5 * N5 / ta / 1e9, "frac_hbm": by5 * N5 / ta / 1e9 / hbm_peak,
            "algorithmic_bytes_per_event": by5, "ms_flow_sample": ta * 1e3, "ms_phase_space_and_reduce": tb * 1e3}
        del phi5, u5, lq5
        torch.cuda.empty_cache()
        # K5 alone: weight + estimator accumulation over resident (target, jac, log q), 24 B/event
        N6 = 1 << 26
        jac6 = torch.rand(N6, dtype=torch.float64, device=dev)
        lq6 = torch.randn(N6, dtype=torch.float64, device=dev)
        tg6 = torch.rand(N6, dtype=torch.float64, device=dev)
        acc6 = ES.new_accumulator(dev)
        t6 = timeit(lambda: ES.accumulate_weights(tg6, jac6, lq6, acc6))
        out["K5_is_reduce_64M"] = {"points_per_s": N6 / t6, "GBps": 24 * N6 / t6 / 1e9, "frac_hbm": 24 * N6 / t6 / 1e9 / hbm_peak}
        del jac6, lq6, tg6
        torch.cuda.empty_cache()
        # the sampling branch of the training loss (unfolding_flow_withDecay_v2.py:376-417) through the fused entry: one
        # forward and one backward launch; bytes = ps + 5 angle pairs in, partons + boost + saved K1 outputs + propagators out
        from memflow_b200.unfolding_flow import utils as UU
        N7 = 1 << 20
        ps4 = PhaseSpace(E_CM, [21, 21], [25, 6, -6, 21], final_masses=torch.tensor([125.25, 172.5, 172.5, 1e-5], dtype=torch.float64), dev=dev)
        u7 = torch.rand(N7, 10, dtype=torch.float64, device=dev).clamp_(1e-6, 1 - 1e-6).requires_grad_(True)
        ang7 = [torch.stack((torch.randn(N7, device=dev, dtype=torch.float64),
                             (torch.rand(N7, device=dev, dtype=torch.float64) * 2 - 1) * 3.1), 1) for _ in range(5)]
        kw7 = dict(log_mean_parton_lab=[0.0] * 3, log_std_parton_lab=[1.0] * 3, log_mean_boost=[0.0] * 2, log_std_boost=[1.0] * 2,
                   log_mean_parton_Hthadtlep=[0.0] * 3, log_std_parton_Hthadtlep=[1.0] * 3)
        tf7 = timeit(lambda: UU.sampled_decayPartons_from_ps(ps4, u7.detach(), *ang7, **kw7))
        o7, b7, *_ = UU.sampled_decayPartons_from_ps(ps4, u7, *ang7, **kw7)
        go7 = torch.ones_like(o7)
        def bwd7():
            u7.grad = None
            torch.autograd.backward([o7, b7], [go7, torch.ones_like(b7)], retain_graph=True)
        tb7 = timeit(bwd7)
        out["sampling_branch_n4_1M"] = {"forward_events_per_s": N7 / tf7, "backward_events_per_s": N7 / tb7,
                                        "forward_GBps": 848 * N7 / tf7 / 1e9, "launches_forward": 1, "launches_backward": 1}
        out["note"] = "wrapper-level timings (include the output allocations of the Python API)"
    except Exception as e:  # extras must never break the headline line
        out["error"] = repr(e)
    return out


def config5_block(torch, dist, dev, rank, world, hbm_peak, barrier):
    """BASELINE configs[4] ("Sharded MEM integral ... ttH+1j ... across 2/4/8 B200") at its own shape: n = 4 with
    masses (125.25, 172.5, 172.5, 1e-5), the production flow F = 10, K = 40, T = 6
    (configs/flow_pretrain_labframe_sampling/flow_config_forEnric.yaml), float32 parameters (the notebook's dtype) and
    float64 parameters (`.double()`, transfer_flow_paper.py:116-127). Device-resident parameters, whole job = all ranks,
    time = max over ranks of the CUDA-event time. Roofline: algorithmic bytes = T F (3K-1) s per event."""
    from memflow_b200 import estimator as ES
    from memflow_b200.distributed import allreduce_accumulators
    out = {"workload": "is_chain_ttH+1j: philox -> 6x RQS inverse (F=10,K=40) -> RAMBO n=4 fp64 -> weight -> sum w, sum w^2",
           "n_final": 4, "masses": [125.25, 172.5, 172.5, 1e-5], "flow": {"T": 6, "F": 10, "K": 40, "bound": 1.0}, "n_gpus": world}
    try:
        T5, F5, K5 = 6, 10, 40
        smp = ES.ImportanceSampler([125.25, 172.5, 172.5, 1e-5], E_CM, seed=3, device=dev)
        for name, dt, N5, steps in (("phi_f32", torch.float32, 1 << 19, 6), ("phi_f64", torch.float64, 1 << 18, 6)):
            g = torch.Generator(device=dev).manual_seed(77 + rank)
            phi5 = torch.randn(T5, N5, F5, 3 * K5 - 1, device=dev, dtype=dt, generator=g)
            u5 = torch.empty(N5, F5, dtype=torch.float64, device=dev)
            lq5 = torch.empty(N5, dtype=torch.float64, device=dev)
            acc5 = ES.new_accumulator(dev)
            base = rank * (steps + 2) * N5
            for i in range(2):
                smp.sample_flow(phi5, event_offset=base + i * N5, u=u5, logq=lq5)
                smp.weigh(u5, lq5, acc=acc5)
            acc5.zero_()
            a, m, b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ta = 0.0
            barrier()
            a.record()
            for i in range(steps):
                if i == steps - 1:
                    m.record()
                smp.sample_flow(phi5, event_offset=base + (2 + i) * N5, u=u5, logq=lq5)
                if i == steps - 1:
                    lastA = torch.cuda.Event(enable_timing=True)
                    lastA.record()
                smp.weigh(u5, lq5, acc=acc5)
            allreduce_accumulators(acc5)
            b.record()
            barrier()
            ta = m.elapsed_time(lastA) * 1e-3      # launch A of the last step alone
            t = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            by = T5 * F5 * (3 * K5 - 1) * phi5.element_size()
            I, sigma, nv = ES.mc_estimate(acc5, world * steps * N5)
            out[name] = {"value": world * steps * N5 / float(t.item()), "unit": UNIT, "events_per_step_per_gpu": N5, "steps": steps,
                         "ms_per_step": float(t.item()) / steps * 1e3,
                         "roofline": {"bound": "hbm", "achieved": by * N5 / ta / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": by * N5 / ta / 1e9 / hbm_peak, "algorithmic_bytes_per_event": by,
                                      "kernel": "flow_sample_reg2_kernel<40>" if dt == torch.float32 else "rqs64_kernel<inverse, 4, chain>",
                                      "kernel_ms": ta * 1e3, "rank": 0,
                                      "note": "peak = the measured read + write COPY rate (MEASURED_PEAKS.json); a read-only "
                                              "parameter stream can run a few per cent above it"},
                         "estimate": {"I": I, "sigma": sigma, "n_valid": nv}}
            del phi5, u5, lq5
            torch.cuda.empty_cache()
    except Exception as e:  # must never break the headline line
        out["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.ensure_built()   # fresh checkout: compile the C-ABI library (and the CPU port of the reference arm) first
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: memflow_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    from memflow_b200 import estimator as ES
    from memflow_b200.distributed import allreduce_accumulators

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = load_peaks()
    K, W = args.steps, max(args.warmup, 0)

    # resident synthetic parameters for one chunk: [T, CHUNK, F, 3K-1] fp32 = 6.8 GB (> L2)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    phi = torch.randn(T_FLOW, CHUNK, F_DIM, PW, device=dev, dtype=torch.float32, generator=g)
    smp = ES.ImportanceSampler(MASSES, E_CM, bound=1.0, seed=2026, device=dev)
    acc = ES.new_accumulator(dev)
    per_rank_events = (W + K) * CHUNK
    base = rank * per_rank_events  # disjoint global event ranges per rank (Philox counter = global index)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # scratch of the two launches of one step (A: flow sample -> u, log q;  B: phase space + weight + reduce)
    u_buf = torch.empty((CHUNK, F_DIM), dtype=torch.float64, device=dev)
    lq_buf = torch.empty(CHUNK, dtype=torch.float64, device=dev)

    def step(i):
        smp.sample_flow(phi, event_offset=base + i * CHUNK, u=u_buf, logq=lq_buf)
        smp.weigh(u_buf, lq_buf, acc=acc)

    for i in range(W):
        step(i)
    barrier()
    acc.zero_()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    t_wall = time.perf_counter()
    evs[0].record()
    for i in range(K):
        smp.sample_flow(phi, event_offset=base + (W + i) * CHUNK, u=u_buf, logq=lq_buf)
        mids[i].record()              # between the two launches: times the dominant kernel alone
        smp.weigh(u_buf, lq_buf, acc=acc)
        evs[i + 1].record()
    allreduce_accumulators(acc)       # the path's only collective: 3 doubles, once
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    t_dev = evs[0].elapsed_time(end) * 1e-3
    kern_ms = sorted(evs[i].elapsed_time(mids[i]) for i in range(K))          # launch A (flow_sample_fast_kernel)
    kern_b_ms = sorted(mids[i].elapsed_time(evs[i + 1]) for i in range(K))    # launch B (rambo_is_kernel)
    t = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    total_events = world * K * CHUNK
    value = total_events / t_max
    acc_host = acc.cpu().tolist()
    I, sigma, nvalid = ES.mc_estimate(acc_host, total_events)

    # ---- e2e: host parameters in pinned memory -> H2D -> chain -> D2H accumulators, every step, through the package's
    #      host-buffer call (memflow_b200.distributed.HostPipeline: double-buffered, 2 copy streams) ----
    from memflow_b200.distributed import HostPipeline, bind_to_gpu_cpus
    numa = bind_to_gpu_cpus(local_rank)          # before the pinned allocation: first touch on the GPU's NUMA node
    e2e_chunk = 1 << 20
    e2e_steps = max(3, min(K, 12))
    pipe = HostPipeline(smp, (T_FLOW, e2e_chunk, F_DIM, PW), dtype=torch.float32, n_copy_streams=2, device=dev)
    host = pipe.pinned_chunk()
    host.copy_(phi[:, :e2e_chunk].cpu())
    res_host = torch.empty(3, dtype=torch.float64).pin_memory()
    acc2 = ES.new_accumulator(dev)

    def e2e_pass(nsteps, off0):
        for i in range(nsteps):
            pipe.submit(host, off0 + i * e2e_chunk, acc2)
            res_host.copy_(acc2, non_blocking=True)   # the step's result goes back to the host
        torch.cuda.synchronize()

    # Five passes of e2e_steps chunks each, the MEDIAN pass is reported and every pass is listed: on multi-GPU boxes of this
    # pool single passes are hit by host-side stalls that are not bandwidth contention (scripts/e2e_probe.py: two ranks copy
    # at 55.5 GB/s each concurrently, yet one pass in four of the SAME loop runs at 19-37 GB/s on both ranks at once).
    e2e_pass(3, 0)
    pass_s = []
    for rep in range(5):
        barrier()
        t0 = time.perf_counter()
        e2e_pass(e2e_steps, (3 + rep * e2e_steps) * e2e_chunk)
        barrier()
        tr = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)      # a pass lasts as long as its slowest rank
        pass_s.append(float(tr.item()))
    te_med = sorted(pass_s)[len(pass_s) // 2]
    h2d_gbps = torch.tensor([host.numel() * 4 * e2e_steps / te_med / 1e9], dtype=torch.float64, device=dev)
    e2e_value = world * e2e_steps * e2e_chunk / te_med
    e2e_passes = [world * e2e_steps * e2e_chunk / t for t in pass_s]
    # The ceiling of that leg on THIS box: the same pinned chunk copied host -> device by every rank at once, nothing else
    # running. (On the pool's multi-GPU boxes GPUs can share a host link: 55 GB/s per GPU alone, 29 GB/s each when the
    # ranks' GPUs sit in pairs behind one PCIe switch uplink -- scripts/h2d_ranks.py, scripts/e2e_probe.py.)
    dplain = pipe.bufs[0]
    dplain.copy_(host, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        dplain.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    tpl = torch.tensor([4 * host.numel() * 4 / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tpl, op=dist.ReduceOp.MIN)
    h2d_plain = float(tpl.item())
    del pipe, dplain

    # ---- BASELINE configs[4] at its own shape on every N: n = 4 (H, t, tbar, j), production flow F = 10, K = 40, T = 6,
    #      float32 and float64 parameters, events sharded, max over ranks ----
    del phi
    torch.cuda.empty_cache()
    config5 = config5_block(torch, dist, dev, rank, world, hbm_peak, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    med_kernel_s = kern_ms[len(kern_ms) // 2] * 1e-3
    avg_kernel_s = sum(kern_ms) / len(kern_ms) * 1e-3
    achieved = BYTES_PER_EVENT * CHUNK / avg_kernel_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("flow_sample_kernel_bytes_per_launch_4M")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": CONFIG,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": 24,
                "steps": e2e_steps, "chunk_events": e2e_chunk, "api": "memflow_b200.distributed.HostPipeline.submit",
                "h2d_GBps_per_rank": float(h2d_gbps.item()), "copy_streams": 2,
                "h2d_plain_concurrent_GBps_per_rank_min": h2d_plain,
                "passes": 5, "statistic": "median pass (each pass = max over ranks)", "passes_points_per_s": e2e_passes,
                "cpu_binding": {"numa_node": numa[0], "cpus": numa[1]},
                "note": "spline parameters start in pinned host memory (1624 B/event): PCIe-bound by construction; "
                        "H2D double-buffered against the kernels on 2 copy streams; h2d_plain_concurrent = the bare pinned "
                        "copy of the same chunk by all ranks at once (the leg's ceiling on this box: GPUs may share a host link)"},
        "config5": config5,
        "gpu_launches": 2 * K * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "kernel": "flow_sample_fast_kernel<10> (launch A of each step)", "peak_source": peak_src,
                     "algorithmic_bytes_per_event": BYTES_PER_EVENT, "events_per_launch": CHUNK,
                     "kernel_ms_avg": avg_kernel_s * 1e3, "kernel_ms_median": med_kernel_s * 1e3,
                     "second_kernel": {"name": "rambo_is_kernel<3> (launch B)", "ms_avg": sum(kern_b_ms) / len(kern_b_ms),
                                       "algorithmic_bytes_per_event": 64, "note": "FP64-pipe bound, see profiles/"},
                     "share_of_step": avg_kernel_s / (t_max / K),
                     "note": "achieved counts only the spline parameters (SURVEY 8d); the kernel also writes 64 B/event (u, log q)"},
        "clocks": clocks,
        "estimate": {"I": I, "sigma": sigma, "n_valid": nvalid, "n_total": total_events},
        "wall_s_timed_region": t_wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_baseline_block(3, 1, budget_s=20.0)
    if world == 1 and not args.no_extras:
        torch.cuda.empty_cache()
        line["extra"] = extras(torch, dev, hbm_peak)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
