"""Host mirror of memflow/phasespace/phasespace.py (PhaseSpace facade, :36-109)."""
import torch

from . import rambo_generator
from . import utils  # noqa: F401  (the reference exposes it as phasespace.utils)

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 1e-5

_DEFAULT_MASSES = (M_HIGGS, M_TOP, M_TOP, M_GLUON)


class PhaseSpace:
    """Same constructor, attributes and three methods as the reference's PhaseSpace.

    `final_masses` defaults to (H, t, t, g) in float64 (the reference's default literal is evaluated at
    import time in the then-active default dtype, normally float32: SURVEY App. C.1)."""

    def __init__(self, collider_energy, initial_pdgs, final_pdgs, final_masses=None, pdf=None, dev=None, **gen_opts):
        if dev is None:
            self.dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
        else:
            self.dev = torch.device(dev)
        if final_masses is None:
            final_masses = torch.tensor(_DEFAULT_MASSES, dtype=torch.float64)
        self.collider_energy = collider_energy
        self.initial_pdgs = initial_pdgs
        self.final_pdgs = final_pdgs
        self.final_masses = torch.as_tensor(final_masses).to(self.dev)
        self.final_state_mass = torch.sum(self.final_masses)
        self.pdf = pdf
        self.generator = rambo_generator.FlatInvertiblePhasespace(
            [0.0, 0.0], self.final_masses, self.collider_energy, pdf=pdf, pdf_active=True, uniform_x1x2=True,
            dev=self.dev, **gen_opts)

    def generate_random_phase_space_points(self, N, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1):
        """(ps_rand [N, 3n-2], momenta, weight, x1, x2)  (:64-80). ps_rand is float64 here (the reference
        draws it in the default dtype)."""
        ps_rand = torch.rand(N, self.generator.nDimPhaseSpace + 2, device=self.dev, dtype=torch.float64)
        momenta, weight, x1, x2 = self.generator.generateKinematics_batch(
            ps_rand, pdgs=self.initial_pdgs, pT_mincut=pT_mincut, delR_mincut=delR_mincut, rap_maxcut=rap_maxcut)
        return ps_rand, momenta, weight, x1, x2

    def get_momenta_from_ps(self, points, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1, requires_grad=False,
                            **extra):
        """points [N, (3n-4)+2] -> momenta [N, n+2, 4] (gluon1, gluon2, finals; partonic CM frame),
        weight, x1, x2  (:82-103)."""
        return self.generator.generateKinematics_batch(
            points, pdgs=self.initial_pdgs, pT_mincut=pT_mincut, delR_mincut=delR_mincut, rap_maxcut=rap_maxcut,
            requires_grad=requires_grad, **extra)

    def get_ps_from_momenta(self, momenta, x1, x2, ensure_CM=True, **extra):
        """Final-state momenta [N, n, 4] -> (ps, det-jac-inverse)  (:105-109). The reference ignores its
        own ensure_CM argument and hard-codes order=[0,1,2,3]; here ensure_CM is honoured and the order
        is the identity for any n."""
        return self.generator.getPSpoint_batch(momenta, x1, x2, ensure_CM=ensure_CM, **extra)
