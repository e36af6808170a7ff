// decay.cuh -- device functions of the decay epilogues, generic over the scalar S: double for the
// forward kernels (decay.cu), a dual number carrying input tangents for the backward kernel (decay_backward.cu).
// Reference: memflow/unfolding_flow/utils.py:109-119 (get_cartesian_comp), :169-173 (get_ptetaphi_comp),
// :745-770 (higgs_decayProducts), :772-815 (top_decayProducts), :826-935 (get_decayPartons_fromlab_propagators_angles),
// memflow/phasespace/utils.py (boost_t, boostVector_t).
#pragma once
#include "rambo.cuh"

namespace mf {

constexpr double DK_PI = 3.141592653589793;

// Forward-mode dual number with NT tangent directions: the decay blocks have 5 (H: 3 propagator components + 2 angles)
// or 7 (top: 3 + 2 + 2) inputs and 12-16 outputs, so the vector-Jacobian product is the contraction of each output
// row's tangents with its incoming gradient, accumulated as the rows are produced.
template <int NT>
struct Dual {
    double v;
    double d[NT];
    MF_D Dual() {}
    MF_D Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < NT; ++i) d[i] = 0.0;
    }
};
// f(a) with value fv and derivative df
template <int NT> MF_D Dual<NT> dual_chain(const Dual<NT>& a, double fv, double df) {
    Dual<NT> r;
    r.v = fv;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = df * a.d[i];
    return r;
}
// f(a, b) with partials fa, fb
template <int NT> MF_D Dual<NT> dual_chain2(const Dual<NT>& a, const Dual<NT>& b, double fv, double fa, double fb) {
    Dual<NT> r;
    r.v = fv;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = fa * a.d[i] + fb * b.d[i];
    return r;
}
template <int NT> MF_D Dual<NT> operator+(const Dual<NT>& a, const Dual<NT>& b) { return dual_chain2(a, b, a.v + b.v, 1.0, 1.0); }
template <int NT> MF_D Dual<NT> operator-(const Dual<NT>& a, const Dual<NT>& b) { return dual_chain2(a, b, a.v - b.v, 1.0, -1.0); }
template <int NT> MF_D Dual<NT> operator*(const Dual<NT>& a, const Dual<NT>& b) { return dual_chain2(a, b, a.v * b.v, b.v, a.v); }
template <int NT> MF_D Dual<NT> operator/(const Dual<NT>& a, const Dual<NT>& b) {
    const double q = a.v / b.v;
    return dual_chain2(a, b, q, 1.0 / b.v, -q / b.v);
}
template <int NT> MF_D Dual<NT> operator-(const Dual<NT>& a) { return dual_chain(a, -a.v, -1.0); }
template <int NT> MF_D Dual<NT> operator+(const Dual<NT>& a, double b) { return dual_chain(a, a.v + b, 1.0); }
template <int NT> MF_D Dual<NT> operator+(double b, const Dual<NT>& a) { return dual_chain(a, b + a.v, 1.0); }
template <int NT> MF_D Dual<NT> operator-(const Dual<NT>& a, double b) { return dual_chain(a, a.v - b, 1.0); }
template <int NT> MF_D Dual<NT> operator-(double b, const Dual<NT>& a) { return dual_chain(a, b - a.v, -1.0); }
template <int NT> MF_D Dual<NT> operator*(const Dual<NT>& a, double b) { return dual_chain(a, a.v * b, b); }
template <int NT> MF_D Dual<NT> operator*(double b, const Dual<NT>& a) { return dual_chain(a, b * a.v, b); }
template <int NT> MF_D Dual<NT> operator/(const Dual<NT>& a, double b) { return dual_chain(a, a.v / b, 1.0 / b); }
template <int NT> MF_D Dual<NT> operator/(double b, const Dual<NT>& a) {
    const double q = b / a.v;
    return dual_chain(a, q, -q / a.v);
}
template <int NT> MF_D bool operator>(const Dual<NT>& a, double b) { return a.v > b; }
template <int NT> MF_D bool operator<(const Dual<NT>& a, double b) { return a.v < b; }

// math on both scalar kinds (a function named like a libm one inside namespace mf would hide the global overloads)
namespace dk {
MF_D double value(double x) { return x; }
MF_D double sqrt(double x) { return ::sqrt(x); }
MF_D double exp(double x) { return ::exp(x); }
MF_D double log(double x) { return ::log(x); }
MF_D double fabs(double x) { return ::fabs(x); }
MF_D double sin(double x) { return ::sin(x); }
MF_D double cos(double x) { return ::cos(x); }
MF_D double tan(double x) { return ::tan(x); }
MF_D double sinh(double x) { return ::sinh(x); }
MF_D double cosh(double x) { return ::cosh(x); }
MF_D double atan2(double y, double x) { return ::atan2(y, x); }
template <int NT> MF_D double value(const Dual<NT>& a) { return a.v; }
template <int NT> MF_D Dual<NT> sqrt(const Dual<NT>& a) { const double s = ::sqrt(a.v); return dual_chain(a, s, 0.5 / s); }
template <int NT> MF_D Dual<NT> exp(const Dual<NT>& a) { const double e = ::exp(a.v); return dual_chain(a, e, e); }
template <int NT> MF_D Dual<NT> log(const Dual<NT>& a) { return dual_chain(a, ::log(a.v), 1.0 / a.v); }
template <int NT> MF_D Dual<NT> fabs(const Dual<NT>& a) {  // torch: d|x|/dx = sign(x), 0 at 0
    return dual_chain(a, ::fabs(a.v), a.v > 0.0 ? 1.0 : (a.v < 0.0 ? -1.0 : 0.0));
}
template <int NT> MF_D Dual<NT> sin(const Dual<NT>& a) { return dual_chain(a, ::sin(a.v), ::cos(a.v)); }
template <int NT> MF_D Dual<NT> cos(const Dual<NT>& a) { return dual_chain(a, ::cos(a.v), -::sin(a.v)); }
template <int NT> MF_D Dual<NT> tan(const Dual<NT>& a) { const double t = ::tan(a.v); return dual_chain(a, t, 1.0 + t * t); }
template <int NT> MF_D Dual<NT> sinh(const Dual<NT>& a) { return dual_chain(a, ::sinh(a.v), ::cosh(a.v)); }
template <int NT> MF_D Dual<NT> cosh(const Dual<NT>& a) { return dual_chain(a, ::cosh(a.v), ::sinh(a.v)); }
template <int NT> MF_D Dual<NT> atan2(const Dual<NT>& y, const Dual<NT>& x) {
    const double r2 = x.v * x.v + y.v * y.v;
    return dual_chain2(y, x, ::atan2(y.v, x.v), x.v / r2, -y.v / r2);
}
}  // namespace dk

// (pt, eta, phi) + mass -> (E, px, py, pz): get_cartesian_comp, utils.py:109-119
template <typename S>
MF_D void cartesian(const S& pt, const S& eta, const S& phi, double mass, S (&p)[4]) {
    p[1] = pt * dk::cos(phi);
    p[2] = pt * dk::sin(phi);
    p[3] = pt * dk::sinh(eta);
    p[0] = dk::sqrt(((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]) + mass * mass);
}
// (E, px, py, pz) -> (E, pt, eta, phi): get_ptetaphi_comp, utils.py:169-173
template <typename S>
MF_D void ptetaphi(const S (&p)[4], S (&o)[4]) {
    const S pt = dk::sqrt(p[1] * p[1] + p[2] * p[2]);
    o[0] = p[0];
    o[1] = pt;
    o[2] = -dk::log(dk::tan(dk::atan2(pt, p[3]) / 2.0));
    o[3] = dk::atan2(p[2], p[1]);
}
// utils.boost_t(p, utils.boostVector_t(parent)): library divisions, these kernels are store- or transcendental-bound
template <typename S>
MF_D void boost_by_parent(const S (&p)[4], const S (&parent)[4], S (&o)[4]) {
    const S bx = parent[1] / parent[0], by = parent[2] / parent[0], bz = parent[3] / parent[0];
    const S b2 = (bx * bx + by * by) + bz * bz;
    const S gamma = 1.0 / dk::sqrt(1.0 - b2);
    const S bp = (p[1] * bx + p[2] * by) + p[3] * bz;
    const S gamma2 = (b2 > 0.0) ? (gamma - 1.0) / b2 : S(0.0);
    const S factor = gamma2 * bp + gamma * p[0];
    o[1] = p[1] + factor * bx;
    o[2] = p[2] + factor * by;
    o[3] = p[3] + factor * bz;
    o[0] = gamma * (p[0] + bp);
}

// H -> b1 b2 (utils.py:745-770): rows = (b1, b2) in the frame of H
template <typename S>
MF_D void higgs_decay_rows(const S (&H)[4], const S& eta, const S& phi, double mass, S (&l1)[4], S (&l2)[4]) {
    const double E = mass / 2.0;
    const S pt = E / dk::cosh(eta);
    S b1[4], b2[4];
    cartesian(pt, eta, phi, 0.0, b1);
    cartesian(pt, S(-eta), S(DK_PI + phi), 0.0, b2);  // b2 = "-b1" in the rest frame (:754)
    boost_by_parent(b1, H, l1);
    boost_by_parent(b2, H, l2);
}

// t -> b W, W -> q1 q2 (utils.py:772-815): rows = (b, q1, q2) in the frame of t
template <typename S>
MF_D void top_decay_rows(const S (&t)[4], const S& eb, const S& pb, const S& ew, const S& pw, double E_b, double w_mass,
                         double b_mass, S (&bl)[4], S (&q1l)[4], S (&q2l)[4]) {
    const S pt_b = ::sqrt(E_b * E_b - b_mass * b_mass) / dk::cosh(eb);  // pt of W = pt of b (:781-785)
    S phiW = DK_PI + pb;
    if (pb > 0.0) phiW = phiW - 2.0 * DK_PI;
    const double Eq = w_mass / 2.0;
    const S pt_q = Eq / dk::cosh(ew);
    S phi2 = DK_PI + pw;
    if (pw > 0.0) phi2 = phi2 - 2.0 * DK_PI;
    S b[4], W[4], q1[4], q2[4], Wl[4];
    cartesian(pt_b, eb, pb, b_mass, b);
    cartesian(pt_q, ew, pw, 0.0, q1);
    cartesian(pt_q, S(-ew), phi2, 0.0, q2);
    cartesian(pt_b, S(-eb), phiW, w_mass, W);  // the reference computes E_W but rebuilds the W energy from p and m (:800)
    boost_by_parent(b, t, bl);
    boost_by_parent(W, t, Wl);
    boost_by_parent(q1, Wl, q1l);
    boost_by_parent(q2, Wl, q2l);
}
MF_HD double top_Eb(double top_mass, double w_mass, double b_mass) {
    return (top_mass * top_mass - w_mass * w_mass + b_mass * b_mass) / 2. / top_mass;
}

// ---- pieces of get_decayPartons_fromlab_propagators_angles (utils.py:826-935) shared by forward and backward ----
// scaled (log(pt+1), eta, phi) -> cartesian propagator (:848-855, :869)
template <typename S>
MF_D void propagator_cartesian(const S (&k_scaled)[3], double mass, const mf_decay_partons_cfg& C, S (&P)[4]) {
    S k[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) k[c] = (c < 2 || C.unscale_phi) ? S(k_scaled[c] * C.std_prop[c] + C.mean_prop[c]) : k_scaled[c];
    cartesian(S(dk::exp(dk::fabs(k[0])) - 1.0), k[1], k[2], mass, P);
}
// output convention of one row (:892-925): (E,pt,eta,phi), the ISR padding row, log(pt+1) and the affine scaling;
// rows 0,3,7 (the propagators) use the propagator scaling, the others the lab scaling
template <typename S>
MF_D void finish_row(const S (&p)[4], bool is_prop, bool is_gluon_pad, const mf_decay_partons_cfg& C, S (&r)[4]) {
    if (!C.ptetaphi) {
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = p[c];
        return;
    }
    ptetaphi(p, r);
    if (is_gluon_pad) { r[0] = S(0.0010); r[1] = S(0.0010); r[2] = S(0.0); r[3] = S(0.0); }
    if (C.final_scaling) {
        const double* mean = is_prop ? C.mean_prop : C.mean_lab;
        const double* sd = is_prop ? C.std_prop : C.std_lab;
        r[1] = dk::log(r[1] + 1.0);
        r[1] = (r[1] - mean[0]) / sd[0];
        r[2] = (r[2] - mean[1]) / sd[1];
        if (C.unscale_phi) r[3] = (r[3] - mean[2]) / sd[2];
    }
}
// ISR gluon = boost - partons (:881-890); pad = the reference's all-soft replacement (:897-905, ptetaphi only)
template <typename S>
MF_D bool gluon_pad(const S (&g)[4], const mf_decay_partons_cfg& C) {
    return C.ptetaphi && ::fabs(dk::value(g[1])) < 0.1 && ::fabs(dk::value(g[2])) < 0.1 && ::fabs(dk::value(g[3])) < 0.1;
}

// ---- the whole parton-level event from the three scaled lab-frame propagators k[3][3] = (log(pt+1), eta, phi), the five
//      decay angle pairs (global memory, this event's) and the scaled boost pz:
//      Compute_ParticlesTensor.get_decayPartons_fromlab_propagators_angles, memflow/unfolding_flow/utils.py:826-935
//      (call site: unfolding_flow_withDecay_v2.py:399-417). unscale -> pt = exp|.|-1 -> cartesian -> H->bb, t->bqq',
//      t->blv -> ISR gluon from the momentum balance -> (E,pt,eta,phi) -> log(pt+1) + scaling.
//      Output rows: H b1 b2 | thad b q1 q2 | tlep b l nu | ISR. The reference's naming quirk is kept: propagator 1
//      takes (tlep_mass, tlep angles), propagator 2 takes (thad_mass, thad angles) (:866-873). ----
MF_D void decay_partons_event(const double (&k)[9], const double* __restrict__ higgs_angles,
                              const double* __restrict__ thad_b_angles, const double* __restrict__ thad_w_angles,
                              const double* __restrict__ tlep_b_angles, const double* __restrict__ tlep_w_angles,
                              double boost_pz_scaled, const mf_decay_partons_cfg& C, double* __restrict__ o) {
    const double mass[3] = {C.higgs_mass, C.tlep_mass, C.thad_mass};
    double P[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double ki[3] = {k[3 * i], k[3 * i + 1], k[3 * i + 2]};
        propagator_cartesian(ki, mass[i], C, P[i]);
    }
    const double boost_pz = boost_pz_scaled * C.std_boost[1] + C.mean_boost[1];
    auto emit = [&](int row, const double (&p)[4], bool is_prop, bool is_gluon_pad) {
        double r[4];
        finish_row(p, is_prop, is_gluon_pad, C, r);
        st4(o + 4 * row, r[0], r[1], r[2], r[3]);
    };
    {
        double eta, phi, l1[4], l2[4];
        ld2(higgs_angles, eta, phi);
        higgs_decay_rows(P[0], eta, phi, C.higgs_mass, l1, l2);
        emit(0, P[0], true, false);
        emit(1, l1, false, false);
        emit(2, l2, false, false);
    }
    {
        double eb, pb, ew, pw, bl[4], q1l[4], q2l[4];
        ld2(thad_b_angles, eb, pb);
        ld2(thad_w_angles, ew, pw);
        top_decay_rows(P[2], eb, pb, ew, pw, top_Eb(C.thad_mass, C.w_had_mass, C.b_mass), C.w_had_mass, C.b_mass, bl, q1l, q2l);
        emit(3, P[2], true, false);
        emit(4, bl, false, false);
        emit(5, q1l, false, false);
        emit(6, q2l, false, false);
    }
    {
        double eb, pb, ew, pw, bl[4], q1l[4], q2l[4];
        ld2(tlep_b_angles, eb, pb);
        ld2(tlep_w_angles, ew, pw);
        top_decay_rows(P[1], eb, pb, ew, pw, top_Eb(C.tlep_mass, C.w_lep_mass, C.b_mass), C.w_lep_mass, C.b_mass, bl, q1l, q2l);
        emit(7, P[1], true, false);
        emit(8, bl, false, false);
        emit(9, q1l, false, false);
        emit(10, q2l, false, false);
    }
    // ISR gluon = boost - partons (:881-905); the all-soft case is replaced by the reference's padding row
    double g[4];
    g[1] = -((P[0][1] + P[2][1]) + P[1][1]);
    g[2] = -((P[0][2] + P[2][2]) + P[1][2]);
    g[3] = boost_pz - ((P[0][3] + P[2][3]) + P[1][3]);
    g[0] = ::sqrt((g[1] * g[1] + g[2] * g[2]) + g[3] * g[3]) + C.eps;
    emit(11, g, false, gluon_pad(g, C));
}

// ---- the sampling branch's glue between K1 and the decay chain (unfolding_flow_withDecay_v2.py:381-397): one CM-frame
//      propagator -> lab frame by the event's boost (E_b, 0, 0, pz_b) -> (pt, eta, phi) -> (scaled log(pt+1), scaled eta, phi).
//      Generic over the scalar like the decay blocks: double in the forward kernel, a dual number in the backward one. ----
template <typename S>
MF_D void lab_scaled_propagator(const S (&p_cm)[4], const S& boost_E, const S& boost_pz, const mf_decay_partons_cfg& C, S (&k)[3]) {
    const S parent[4] = {boost_E, S(0.0), S(0.0), boost_pz};
    S lab[4], o[4];
    boost_by_parent(p_cm, parent, lab);   // ps_utils.boost_tt(p, ps_utils.boostVector_t(boost_Epz))  (:384-387)
    ptetaphi(lab, o);                     // get_ptetaphi_comp_batch (:390)
    k[0] = (dk::log(o[1] + 1.0) - C.mean_prop[0]) / C.std_prop[0];   // :392-393
    k[1] = (o[2] - C.mean_prop[1]) / C.std_prop[1];
    k[2] = o[3];
}

}  // namespace mf
