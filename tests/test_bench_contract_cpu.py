"""CPU: the reference arm of bench.py (the reference's own torch CPU path staged under oracle/_ref, else the oracle port)
prints exactly ONE JSON line that carries the driver's contract keys; the GPU arm refuses to run without a device instead
of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("phase-space points/sec") and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "memflow", "phasespace", "rambo_generator.py"))
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    if staged:   # the C port's number travels next to the reference's
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 0 and "oracle/_ref" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("is_chain_ttH") and d["dtype"] == "f64" and d["data"] == "synthetic"
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling"):
        assert k in d


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert out.stdout.strip() == ""


def test_staged_reference_matches_the_port():
    """oracle/make_ref.py stages the UNMODIFIED reference files (sha256 in the manifest == the files under /root/reference
    when that tree is here) and the reference torch chain agrees with the C port on the same inputs."""
    import hashlib
    import numpy as np
    import torch
    from oracle import make_ref, ref_torch as R, oracle as O, philox as OP
    if not make_ref.make_ref(quiet=True):
        import pytest
        pytest.skip("no reference tree and no staged copy")
    man = json.load(open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")))
    for rel, sha in man["files"].items():
        assert hashlib.sha256(open(os.path.join(ROOT, "oracle", "_ref", rel), "rb").read()).hexdigest() == sha
        src = os.path.join(make_ref.REF, rel)
        if os.path.exists(src):
            assert hashlib.sha256(open(src, "rb").read()).hexdigest() == sha
    T, N, F, K = 2, 4000, 7, 10
    rng = np.random.default_rng(0)
    phi = rng.standard_normal((T, N, F, 3 * K - 1)).astype(np.float32)
    z = OP.base_sample(N, F, 1.0, 1, 0)
    a = R.chain_step(torch.from_numpy(phi), torch.from_numpy(z), [125.25, 172.5, 172.5], 13000.0)
    zz, ls = z.copy(), np.zeros(N)
    for t in range(T - 1, -1, -1):
        zz, l, _ = O.rqs_inverse(phi[t], zz, bound=1.0)
        ls += l.astype(np.float64).sum(1)
    mom, w, x1, x2 = O.rambo_forward((zz.astype(np.float64) + 1) / 2, [125.25, 172.5, 172.5], 13000.0)
    b = O.is_reduce(1 / (2 * 13000.0 ** 2) / (x1 * x2), w, ls)
    assert a[2] == b[2] == N
    np.testing.assert_allclose(a[:2], b[:2], rtol=2e-4)      # the spline step is float32 on both sides
