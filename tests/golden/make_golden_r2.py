"""Round-2 additions to the committed golden fixtures, again by running the UNMODIFIED reference (see make_golden.py
for the import recipe; the round-1 files are not rewritten):

  rambo_extra.npz
    lab_n{3,4,8}          utils.boost_to_lab_frame (memflow/phasespace/utils.py:180-202) applied to the fp64-clean CM momenta of
                          rambo_n*.npz -- what K1's fused MF_FLAG_LAB_FRAME epilogue must reproduce
    direct_*_n4           generateKinematics_batch with uniform_x1x2=False (rambo_generator.py:229-234): the last two columns
                          ARE x1, x2 and carry no Jacobian
    inter_*_n{4,8}        generateIntermediatesMassless_batch / generateIntermediatesMassive_batch called on their own
                          (rambo_generator.py:448-509): weights and intermediate masses
Run in the build container only:  python tests/golden/make_golden_r2.py
"""
import contextlib
import io
import os

import numpy as np
import torch

from make_golden import CASES, E_CM, HERE, fp64_clean, import_reference, make_u


def main():
    phasespace, rambo_generator, utils = import_reference()
    out = {}
    for n in (3, 4, 8):
        g = np.load(os.path.join(HERE, "rambo_n%d.npz" % n))
        lab = utils.boost_to_lab_frame(torch.from_numpy(g["mom_clean"]), torch.from_numpy(g["x1_clean"]), torch.from_numpy(g["x2_clean"]))
        out["lab_n%d" % n] = lab.numpy()
    # uniform_x1x2 = False: last two columns are the momentum fractions themselves
    n = 4
    m = torch.tensor(CASES[n], dtype=torch.float64)
    gen = rambo_generator.FlatInvertiblePhasespace([0.0, 0.0], m, E_CM, pdf=None, pdf_active=True, uniform_x1x2=False,
                                                   dev=torch.device("cpu"))
    u = make_u(n, 2000 + n)
    g_ = torch.Generator().manual_seed(5)
    tau = (float(m.sum()) / E_CM) ** 2
    x1 = tau ** (torch.rand(u.shape[0], dtype=torch.float64, generator=g_) * 0.9)          # x1 in (tau^0.9, 1)
    x2 = (tau / x1) ** (torch.rand(u.shape[0], dtype=torch.float64, generator=g_) * 0.95)  # x1 x2 > tau
    u[:, -2], u[:, -1] = x1, x2
    with fp64_clean(), contextlib.redirect_stdout(io.StringIO()):
        mom, w, xa, xb = gen.generateKinematics_batch(u)
    out.update(direct_u_n4=u.numpy(), direct_mom_n4=mom.numpy(), direct_w_n4=w.numpy(), direct_x1_n4=xa.numpy(),
               direct_x2_n4=xb.numpy())
    # the two intermediates methods on their own
    for n in (4, 8):
        m = torch.tensor(CASES[n], dtype=torch.float64)
        gen = rambo_generator.FlatInvertiblePhasespace([0.0, 0.0], m, E_CM, pdf=None, pdf_active=True, uniform_x1x2=True,
                                                       dev=torch.device("cpu"))
        u = make_u(n, 3000 + n)[24:24 + 128]
        g_ = torch.Generator().manual_seed(9)
        ecm = 1500.0 + 9000.0 * torch.rand(u.shape[0], dtype=torch.float64, generator=g_)
        M = torch.zeros(u.shape[0], n - 1, dtype=torch.float64)
        M[:, 0] = ecm
        with fp64_clean():
            w0, K = gen.generateIntermediatesMassless_batch(M, ecm, u)
            w1, Mi = gen.generateIntermediatesMassive_batch(M, ecm, u)
        out.update({"inter_u_n%d" % n: u.numpy(), "inter_ecm_n%d" % n: ecm.numpy(), "inter_massless_w_n%d" % n: np.asarray(w0, dtype=np.float64),
                    "inter_massless_K_n%d" % n: K.numpy(), "inter_massive_w_n%d" % n: np.asarray(w1, dtype=np.float64),
                    "inter_massive_M_n%d" % n: Mi.numpy()})
        # scalar E_cm variant of the massless method (python float volume, rambo_generator.py:465-470)
        w0s, _ = gen.generateIntermediatesMassless_batch(M, 13000.0, u)
        out["inter_massless_w_scalar_n%d" % n] = w0s.numpy()
    np.savez_compressed(os.path.join(HERE, "rambo_extra.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
