"""Host mirror of memflow/phasespace/phasespace.py (the `PhaseSpace` facade, reference lines 36-109).

Same constructor signature, attribute names and three public methods; every batch call lands in one CUDA
kernel through `FlatInvertiblePhasespace` (K1 / K2).
"""
import torch

from . import rambo_generator
from . import utils  # noqa: F401  (reference users reach it as `phasespace.utils`)

M_HIGGS, M_TOP, M_GLUON = 125.25, 172.5, 1e-5
_HTTG = (M_HIGGS, M_TOP, M_TOP, M_GLUON)


def _pick_device(dev):
    if dev is not None:
        return torch.device(dev)
    return torch.device("cuda:0" if torch.cuda.is_available() else "cpu")


class PhaseSpace:
    """Drop-in for `memflow.phasespace.phasespace.PhaseSpace`.

    Differences: `final_masses=None` means (H, t, tbar, g) in float64 -- the reference evaluates a float32 literal
    at import time -- and extra keyword arguments (`ref_fp32_quirks`, `sync_errors`,
    `compute_dtype`) are forwarded to the generator."""

    def __init__(self, collider_energy, initial_pdgs, final_pdgs, final_masses=None, pdf=None, dev=None, **gen_opts):
        self.dev = _pick_device(dev)
        masses = torch.tensor(_HTTG, dtype=torch.float64) if final_masses is None else torch.as_tensor(final_masses)
        self.collider_energy, self.pdf = collider_energy, pdf
        self.initial_pdgs, self.final_pdgs = initial_pdgs, final_pdgs
        self.final_masses = masses.to(self.dev)
        self.final_state_mass = self.final_masses.sum()
        # two massless beams, x1/x2 sampled from two extra uniforms: the only mode the reference facade uses (:54-62)
        self.generator = rambo_generator.FlatInvertiblePhasespace(
            [0.0, 0.0], self.final_masses, collider_energy, pdf=pdf, pdf_active=True, uniform_x1x2=True, dev=self.dev,
            **gen_opts)

    def _forward(self, points, cuts, **kw):
        pt, dr, eta = cuts
        return self.generator.generateKinematics_batch(points, pT_mincut=pt, delR_mincut=dr, rap_maxcut=eta,
                                                       pdgs=self.initial_pdgs, **kw)

    def generate_random_phase_space_points(self, N, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1):
        """N uniform points of the (3n-4)+2 cube and their image: (ps_rand, momenta, weight, x1, x2) (ref :64-80).
        The points are drawn in float64 (the reference uses the default dtype)."""
        width = self.generator.nDimPhaseSpace + 2
        ps_rand = torch.rand(N, width, dtype=torch.float64, device=self.dev)
        return (ps_rand,) + tuple(self._forward(ps_rand, (pT_mincut, delR_mincut, rap_maxcut)))

    def get_momenta_from_ps(self, points, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1, requires_grad=False, **extra):
        """points [N, (3n-4)+2] -> (momenta [N, n+2, 4], weight, x1, x2) (ref :82-103). Rows of `momenta`: the two
        incoming gluons, then the final state in the requested order, all in the partonic CM frame (what a matrix
        element wants); x1, x2 give the boost to the lab."""
        return self._forward(points, (pT_mincut, delR_mincut, rap_maxcut), requires_grad=requires_grad, **extra)

    def get_ps_from_momenta(self, momenta, x1, x2, ensure_CM=True, **extra):
        """Final-state momenta [N, n, 4] -> (ps [N, 3n-2], det-jac-inverse [N]) (ref :105-109). Unlike the reference,
        `ensure_CM` is honoured and the particle order defaults to the identity for any n."""
        return self.generator.getPSpoint_batch(momenta, x1, x2, ensure_CM=ensure_CM, **extra)
