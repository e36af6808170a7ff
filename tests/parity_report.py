"""Writes the parity evidence table of the phase-space kernels against the oracle at BASELINE-like sizes
(run on the B200 box:  python tests/parity_report.py > gpurun_out/parity_report.txt). Test infrastructure:
uses the oracle as the checker."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from tests import parity  # noqa: E402
from memflow_b200.phasespace.phasespace import PhaseSpace  # noqa: E402

E = 13000.0
CASES = [("config 1: ttH forward, n=3, 1M, seed 12345", 3, [125.25, 172.5, 172.5], 1_000_000, 12345),
         ("config 5 shape: ttH+j, n=4", 4, [125.25, 172.5, 172.5, 1e-5], 500_000, 7),
         ("config 2 shape: 8-body, n=8", 8, [4.7, 4.7, 4.7, 4.7, 0.1, 0.0, 0.3, 0.3], 500_000, 2024)]
O.set_threads(O.max_threads())
print("| case | quantity | median | p99 | p99.99 | max | frac > 1e-12 | max / kappa |")
print("|---|---|---|---|---|---|---|---|")
for name, n, masses, N, seed in CASES:
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(N, 3 * n - 2, dtype=torch.float64, generator=g).clamp(1e-10, 1 - 1e-10)
    ps = PhaseSpace(E, [21, 21], list(range(n)), final_masses=torch.tensor(masses, dtype=torch.float64), dev="cuda:0")
    mom, w, x1, x2 = ps.get_momenta_from_ps(u.cuda())
    om, ow, ox1, ox2 = O.rambo_forward(u.numpy(), masses, E)
    kappa = parity.conditioning(u.numpy(), om, n, masses)
    rows = [("momenta |dp|/E_cm", parity.momenta_error(mom.cpu().numpy(), om, ox1, ox2, E)),
            ("ln weight", np.abs(np.log(w.cpu().numpy()) - np.log(ow))),
            ("x1, x2", np.maximum(np.abs(x1.cpu().numpy() - ox1), np.abs(x2.cpu().numpy() - ox2)))]
    back, prob = ps.generator.getPSpoint_batch(torch.from_numpy(om[:, 2:]).cuda(), torch.from_numpy(ox1).cuda(),
                                               torch.from_numpy(ox2).cuda())
    ops, oprob, _ = O.rambo_inverse(om[:, 2:], ox1, ox2, masses, E)
    rows += [("inverse u (all columns)", np.abs(back.cpu().numpy() - ops).max(axis=1)),
             ("inverse ln prob", np.abs(np.log(prob.cpu().numpy()) - np.log(oprob))),
             ("round trip |u' - u| (kernel fwd -> kernel inv)",
              (ps.generator.getPSpoint_batch(mom[:, 2:], x1, x2)[0].cpu() - u).abs().max(dim=1).values.numpy())]
    for q, err in rows:
        err = np.where(np.isfinite(err), err, np.nan)
        print("| %s | %s | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (
            name, q, np.nanmedian(err), np.nanquantile(err, 0.99), np.nanquantile(err, 0.9999), np.nanmax(err),
            np.nanmean(err > 1e-12), np.nanmax(err / kappa)))
