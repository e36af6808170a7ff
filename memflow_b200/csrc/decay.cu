// decay.cu -- two-body decay epilogues of the sampled propagators:
//   H -> b b              Compute_ParticlesTensor.higgs_decayProducts, memflow/unfolding_flow/utils.py:745-770
//   t -> b W, W -> q q'   Compute_ParticlesTensor.top_decayProducts,   memflow/unfolding_flow/utils.py:772-815
// Daughters are built from (eta, phi) in the parent's rest frame with E = m/2 (or the massive two-body energies), then
// boosted by the parent's velocity (utils.boost_t) -- the W daughters by the boosted W. One event per thread; the
// parent 4-momentum is one 256-bit load (row stride given, so a row of a K1 output is read in place), every
// output row one 256-bit store. Optionally converted to (E, pt, eta, phi) like the reference's ptetaphi=True.
#include "rambo.cuh"

namespace mf {

// (pt, eta, phi) + mass -> (E, px, py, pz): get_cartesian_comp, utils.py:109-119
MF_D void cartesian(double pt, double eta, double phi, double mass, double (&p)[4]) {
    p[1] = pt * cos(phi);
    p[2] = pt * sin(phi);
    p[3] = pt * sinh(eta);
    p[0] = sqrt(((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]) + mass * mass);
}
// (E, px, py, pz) -> (E, pt, eta, phi): get_ptetaphi_comp, utils.py:169-173
MF_D void ptetaphi(const double (&p)[4], double (&o)[4]) {
    const double pt = sqrt(p[1] * p[1] + p[2] * p[2]);
    o[0] = p[0];
    o[1] = pt;
    o[2] = -log(tan(atan2(pt, p[3]) / 2.0));
    o[3] = atan2(p[2], p[1]);
}
MF_D void boost_by_parent(const double (&p)[4], const double (&parent)[4], double (&o)[4]) {
    // utils.boost_t(p, utils.boostVector_t(parent)): library divisions, this kernel is store-bound
    const double bx = parent[1] / parent[0], by = parent[2] / parent[0], bz = parent[3] / parent[0];
    const double b2 = (bx * bx + by * by) + bz * bz;
    const double gamma = 1.0 / sqrt(1.0 - b2);
    const double bp = (p[1] * bx + p[2] * by) + p[3] * bz;
    const double gamma2 = (b2 > 0.0) ? (gamma - 1.0) / b2 : 0.0;
    const double factor = gamma2 * bp + gamma * p[0];
    o[1] = p[1] + factor * bx;
    o[2] = p[2] + factor * by;
    o[3] = p[3] + factor * bz;
    o[0] = gamma * (p[0] + bp);
}
MF_D void emit_row(double* dst, const double (&p)[4], int as_ptetaphi) {
    if (as_ptetaphi) {
        double o[4];
        ptetaphi(p, o);
        st4(dst, o[0], o[1], o[2], o[3]);
    } else {
        st4(dst, p[0], p[1], p[2], p[3]);
    }
}

__global__ void __launch_bounds__(256)
higgs_decay_kernel(const double* __restrict__ parent, int64_t parent_stride, const double* __restrict__ angles, int64_t N,
                   double mass, int as_ptetaphi, double* __restrict__ out) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double H[4], eta, phi;
    ld4(parent + ev * parent_stride, H[0], H[1], H[2], H[3]);
    ld2(angles + 2 * ev, eta, phi);
    const double E = mass / 2.0;
    const double pt = E / cosh(eta);
    double b1[4], b2[4], l1[4], l2[4];
    cartesian(pt, eta, phi, 0.0, b1);
    cartesian(pt, -eta, 3.141592653589793 + phi, 0.0, b2);  // b2 = "-b1" in the rest frame (:754)
    boost_by_parent(b1, H, l1);
    boost_by_parent(b2, H, l2);
    double* o = out + ev * 12;
    emit_row(o, H, as_ptetaphi);
    emit_row(o + 4, l1, as_ptetaphi);
    emit_row(o + 8, l2, as_ptetaphi);
}

__global__ void __launch_bounds__(256)
top_decay_kernel(const double* __restrict__ parent, int64_t parent_stride, const double* __restrict__ b_angles,
                 const double* __restrict__ w_angles, int64_t N, double E_b, double E_W, double w_mass, double b_mass,
                 int as_ptetaphi, double* __restrict__ out) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    const double two_pi = 2.0 * 3.141592653589793, pi = 3.141592653589793;
    double t[4], eb, pb, ew, pw;
    ld4(parent + ev * parent_stride, t[0], t[1], t[2], t[3]);
    ld2(b_angles + 2 * ev, eb, pb);
    ld2(w_angles + 2 * ev, ew, pw);
    const double pt_b = sqrt(E_b * E_b - b_mass * b_mass) / cosh(eb);  // pt of W = pt of b (:781-785)
    double phiW = pi + pb;
    if (pb > 0.0) phiW = phiW - two_pi;
    const double Eq = w_mass / 2.0;
    const double pt_q = Eq / cosh(ew);
    double phi2 = pi + pw;
    if (pw > 0.0) phi2 = phi2 - two_pi;
    double b[4], W[4], q1[4], q2[4], bl[4], Wl[4], q1l[4], q2l[4];
    cartesian(pt_b, eb, pb, b_mass, b);
    cartesian(pt_q, ew, pw, 0.0, q1);
    cartesian(pt_q, -ew, phi2, 0.0, q2);
    cartesian(pt_b, -eb, phiW, w_mass, W);
    (void)E_W;  // the reference computes E_W but rebuilds the W energy from its momentum and mass (:800)
    boost_by_parent(b, t, bl);
    boost_by_parent(W, t, Wl);
    boost_by_parent(q1, Wl, q1l);
    boost_by_parent(q2, Wl, q2l);
    double* o = out + ev * 16;
    emit_row(o, t, as_ptetaphi);
    emit_row(o + 4, bl, as_ptetaphi);
    emit_row(o + 8, q1l, as_ptetaphi);
    emit_row(o + 12, q2l, as_ptetaphi);
}

}  // namespace mf

extern "C" int mf_higgs_decay_f64(const double* parent, int64_t parent_stride, const double* angles, int64_t N, double mass,
                                  int as_ptetaphi, double* out, mf_stream_t stream) {
    if (N < 0 || (N > 0 && (!parent || !angles || !out))) return MF_E_BADARG;
    if (parent_stride < 4 || parent_stride % 4 != 0) return MF_E_SHAPE;
    if (!mf::aligned(parent, 32) || !mf::aligned(angles, 16) || !mf::aligned(out, 32)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    mf::higgs_decay_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parent, parent_stride, angles, N,
                                                                                          mass, as_ptetaphi, out);
    return mf::launch_status();
}

extern "C" int mf_top_decay_f64(const double* parent, int64_t parent_stride, const double* b_angles, const double* w_angles,
                                int64_t N, double top_mass, double w_mass, double b_mass, int as_ptetaphi, double* out,
                                mf_stream_t stream) {
    if (N < 0 || (N > 0 && (!parent || !b_angles || !w_angles || !out))) return MF_E_BADARG;
    if (parent_stride < 4 || parent_stride % 4 != 0) return MF_E_SHAPE;
    if (!mf::aligned(parent, 32) || !mf::aligned(b_angles, 16) || !mf::aligned(w_angles, 16) || !mf::aligned(out, 32))
        return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    const double E_b = (top_mass * top_mass - w_mass * w_mass + b_mass * b_mass) / 2. / top_mass;
    const double E_W = (top_mass * top_mass + w_mass * w_mass - b_mass * b_mass) / 2. / top_mass;
    mf::top_decay_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        parent, parent_stride, b_angles, w_angles, N, E_b, E_W, w_mass, b_mass, as_ptetaphi, out);
    return mf::launch_status();
}
