// rambo.cuh -- per-event device math of the RAMBO-on-diet map (Platzer, arXiv:1308.2922, massive
// variant) as MEMFlow implements it: memflow/phasespace/rambo_generator.py, memflow/phasespace/utils.py.
// One event per thread; all per-particle state lives in registers (loops over particles are fully
// unrolled on the compile-time multiplicity NF).
#pragma once
#include <cmath>
#include <type_traits>
#include <utility>

#include "common.cuh"

namespace mf {

// Launch-constant parameters, passed by value (kernel parameter space / constant bank).
template <typename T, int NF> struct RamboParams {
    T m[NF];      // final-state masses (already permuted / target masses for the inverse)
    T msum_rev[NF];  // forward: flip(cumsum(flip(m))) (:484-486).  inverse: sum(m[j:]) forward order (:660)
    T msum;       // torch.sum(masses_t)
    T q, A;       // x1x2 map constants, utils.py:266-269
    T t1, t1sq, neg_inv_A, inv_A, neg_A;
    T e_collider;
    T flat_c, flat_ff;  // get_flatWeights: c = (2pi)^(4-3n) (pi/2)^(n-1), ff = (n-1)!(n-2)!
    float flat_c32, flat_ff32;
    T cut_pt, cut_dr, cut_eta;
    int order[NF];
    uint32_t flags;
    // get_PS epilogue (memflow/unfolding_flow/unfolding_flow.py:170-171): logit(ps, eps) then (x - mean) / scale
    T logit_eps;
    T lmean[3 * NF - 2], lscale[3 * NF - 2];
};

// rambo_generator.py:145-149 rho(M,N,m) = sqrt((M^2-(N+m)^2)(M^2-(N-m)^2)) / (8 M^2)
template <typename T> MF_D T rho_f(T M, T N, T m) {
    T Msqr = mul_rn(M, M);
    T a = N + m, b = N - m;
    T t = mul_rn(sub_rn(Msqr, mul_rn(a, a)), sub_rn(Msqr, mul_rn(b, b)));
    return sqrt_(t) / ((T)8 * Msqr);
}

// sqrt((M^2-(N+m)^2)(M^2-(N-m)^2)): numerator of 8 M^2 rho(M,N,m); the products are non-contractible
template <typename Pr = PrimIEEE, typename T> MF_D T rho_num(T M, T N, T m) {
    T Msqr = mul_rn(M, M);
    T a = N + m, b = N - m;
    return Pr::sqrt(mul_rn(sub_rn(Msqr, mul_rn(a, a)), sub_rn(Msqr, mul_rn(b, b))));
}

template <typename T> MF_D T rho2_3(T x, T y, T z) { return (x * x + y * y) + z * z; }

// utils.py:88-118 boost_t: general Lorentz boost by velocity bv
template <typename Pr = PrimIEEE, typename T> MF_D void boost(const T (&p)[4], T bx, T by, T bz, T (&o)[4]) {
    T b2 = add_rn(add_rn(mul_rn(bx, bx), mul_rn(by, by)), mul_rn(bz, bz));
    T gamma = Pr::rsqrt(sub_rn((T)1, b2));
    T bp = (p[1] * bx + p[2] * by) + p[3] * bz;
    T gamma2 = (b2 > (T)0) ? Pr::div(gamma - (T)1, b2) : (T)0;
    T factor = gamma2 * bp + gamma * p[0];
    o[1] = p[1] + factor * bx;
    o[2] = p[2] + factor * by;
    o[3] = p[3] + factor * bz;
    o[0] = gamma * (p[0] + bp);
}

// ---------------------------------------------------------------------------------------------
// Mass-variable root solve: v = u^E ((E+1) - E u), u in [0,1]  (rambo_generator.py:142-143).
// The reference finds u by up to 180 levels of dyadic bisection over the whole batch
// (bisect_vec_batch, :400-446: 2 pow per level -> ~360 pow per root). Here:
//   E == 1 : closed form u = v / (1 + sqrt(1 - v)).
//   E >= 2 : v is the CDF of Beta(E,2). Work in the well-conditioned variable: x = u and
//            target t = v when v < 0.65, else x = 1-u and t = 1-v with
//            1 - v = x^2 P_E(x)  (polynomial, no cancellation). fp32 asymptotic start + 2 fp32
//            Newton steps (MUFU log2/exp2), then one fp64 Newton and one chord step on a per-lane selected
//            polynomial (MassVarPoly): within 4 ulp of the exact root for all v in (0, 1),
//            tests/test_rambo_gpu.py::test_massvar_solver_vs_bisection.
// Edge semantics follow the bisection: v <= 0 -> 2^-180 (its floor after 180 levels); v >= 1 -> the largest
// double below 1 (the bisection stalls just below 1, so the reference's weight stays finite there).
// ---------------------------------------------------------------------------------------------
MF_HD constexpr double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return r;
}
// coefficient of x^k (k = 2..E+1) in 1 - (1-x)^E (1 + E x)
MF_HD constexpr double tail_coef(int E, int k) {
    double s = (k % 2 == 0) ? 1.0 : -1.0;
    return -s * (binom(E, k) - (double)E * binom(E, k - 1));
}
template <int E, int K, typename T> MF_D T tail_horner(T x, T acc) {
    if constexpr (K < 2) return acc;
    else return tail_horner<E, K - 1>(x, acc * x + (T)tail_coef(E, K));
}
// P_E(x) with 1 - g(1-x) = x^2 P_E(x)
template <int E, typename T> MF_D T tail_poly(T x) { return tail_horner<E, E>(x, (T)tail_coef(E, E + 1)); }

template <int E, typename T> struct MassVarFn {
    // F(x) - t, F'(x), F''(x) in the chosen variable
    static MF_D void eval(bool small, T x, T t, T& f, T& d1, T& d2) {
        const T ee1 = (T)(E * (E + 1));
        if (small) {
            T xm2 = ipow<E - 2>(x);
            T xm1 = xm2 * x;
            f = xm1 * x * ((T)(E + 1) - (T)E * x) - t;
            d1 = ee1 * xm1 * ((T)1 - x);
            d2 = ee1 * xm2 * ((T)(E - 1) - (T)E * x);
        } else {
            T om = (T)1 - x;
            T omm2 = ipow<E - 2>(om);
            T omm1 = omm2 * om;
            f = x * x * tail_poly<E>(x) - t;
            d1 = ee1 * x * omm1;
            d2 = ee1 * omm2 * ((T)1 - (T)E * x);
        }
    }
};

// F and F' as Horner forms over per-lane selected coefficients (fp64 polish of solve_massvar)
MF_D double rcp_seed_refined(double b) {  // 2^-40 reciprocal: a Newton slope does not need more
    const double r0 = rcp_seed(b);
    return fma(r0, fma(-b, r0, 1.0), r0);
}
// 0 < x < 1 on the bit pattern (integer pipe; false for 0, negatives, NaN, inf)
MF_D bool in_open_unit(double x) {
    const unsigned hi = (unsigned)__double2hiint(x);
    return hi < 0x3ff00000u && (hi | (unsigned)__double2loint(x)) != 0u;
}
template <int E> struct MassVarPoly {
    bool small;
    MF_D explicit MassVarPoly(bool s) : small(s) {}
    // coefficient of x^(j+2), j = 0..E-1: x^E((E+1) - E x) for the small variable, x^2 P_E(x) for the large one.
    // Selected at every use from two compile-time constants (integer pipe); an array of selected values ended up in
    // local memory.
    template <int J> MF_D double coef(double scale) const {
        constexpr double a = (J + 2 == E) ? (double)(E + 1) : ((J + 2 == E + 1) ? -(double)E : 0.0);
        constexpr double b = tail_coef(E, J + 2);
        return small ? scale * a : scale * b;
    }
    template <int J> MF_D double horner_f(double x, double h) const {
        if constexpr (J < 0) return h;
        else return horner_f<J - 1>(x, fma(h, x, coef<J>(1.0)));
    }
    template <int J> MF_D double horner_d(double x, double g) const {
        if constexpr (J < 0) return g;
        else return horner_d<J - 1>(x, fma(g, x, coef<J>((double)(J + 2))));
    }
    MF_D double residual(double x, double t) const {   // F(x) - t, F(x) = x^2 (c_0 + c_1 x + ...)
        return fma(x * x, horner_f<E - 2>(x, coef<E - 1>(1.0)), -t);
    }
    MF_D double slope(double x) const {   // F'(x) = x (2 c_0 + 3 c_1 x + ...)
        return x * horner_d<E - 2>(x, coef<E - 1>((double)(E + 1)));
    }
};

// log2 of the asymptote prefactors (compile-time tables, E = 1..7)
template <int E> MF_HD constexpr float kLog2InvEp1() {   // log2(1 / (E + 1))
    constexpr float t[8] = {0.0f, -1.0f, -1.5849625007211562f, -2.0f, -2.321928094887362f, -2.584962500721156f,
                            -2.807354922057604f, -3.0f};
    return t[E];
}
template <int E> MF_HD constexpr float kLog2TwoOverEEp1() {   // log2(2 / (E (E + 1)))
    constexpr float t[8] = {0.0f, 0.0f, -1.5849625007211562f, -2.584962500721156f, -3.321928094887362f, -3.9068905956085187f,
                            -4.392317422778761f, -4.807354922057604f};
    return t[E];
}
// MUFU approximations without the denormal / special-case wrappers of exp2f, __log2f, sqrtf (each wrapper is a compare
// and a branch per call: they cut the event's instruction stream into basic blocks the scheduler cannot interleave)
MF_D float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MF_D float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MF_D float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// fp32 start of the root, BRANCH-FREE: both variables share one instruction stream.
//   asymptote   x0 = (a t)^p          small variable: a = 1/(E+1), p = 1/E;   large variable: a = 2/(E(E+1)), p = 1/2
//   refinement  small: (t / ((E+1) - E x0))^(1/E)      large: x0 (1 + (E-1)/3 x0)
//   2 Newton steps on F(x) - t = x^2 H(x) - t, F'(x) = x G(x), H and G Horner forms over per-lane selected coefficients
//   (the same polynomial the fp64 polish uses). Valid for 1e-30 < t <= 0.65; callers route everything else elsewhere.
template <int E> struct MassVarPolyF {
    bool small;
    MF_D explicit MassVarPolyF(bool s) : small(s) {}
    template <int J> MF_D float coef(float scale) const {
        constexpr float a = (J + 2 == E) ? (float)(E + 1) : ((J + 2 == E + 1) ? -(float)E : 0.0f);
        constexpr float b = (float)tail_coef(E, J + 2);
        return small ? scale * a : scale * b;
    }
    template <int J> MF_D float horner_f(float x, float h) const {
        if constexpr (J < 0) return h;
        else return horner_f<J - 1>(x, fmaf(h, x, coef<J>(1.0f)));
    }
    template <int J> MF_D float horner_d(float x, float g) const {
        if constexpr (J < 0) return g;
        else return horner_d<J - 1>(x, fmaf(g, x, coef<J>((float)(J + 2))));
    }
    MF_D float residual(float x, float t) const { return fmaf(x * x, horner_f<E - 2>(x, coef<E - 1>(1.0f)), -t); }
    MF_D float slope(float x) const { return x * horner_d<E - 2>(x, coef<E - 1>((float)(E + 1))); }
};
template <int E> MF_D float massvar_start_f32(bool small, float tf) {
    constexpr float ie = 1.0f / (float)E;
    const float lt = lg2_approx(tf);
    // log2 of the asymptote's prefactor a, times the exponent p
    const float p = small ? ie : 0.5f;
    // constants folded at compile time: log2(1/(E+1)) and log2(2/(E(E+1)))
    const float lga = small ? kLog2InvEp1<E>() : kLog2TwoOverEEp1<E>();
    const float x0 = ex2_approx((lt + lga) * p);
    const float xs = ex2_approx((lt - lg2_approx(fmaf(-(float)E, x0, (float)(E + 1)))) * ie);
    const float xl = x0 * fmaf((float)(E - 1) / 3.0f, x0, 1.0f);
    float x = small ? xs : xl;
    const MassVarPolyF<E> pl(small);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float xn = fmaf(-pl.residual(x, tf), rcp_approx(pl.slope(x)), x);
        x = (xn > 0.0f && xn < 1.0f) ? xn : x;  // also rejects NaN/inf steps
    }
    return x;
}

// t <= 1e-30: fp64 asymptotic start (library pow) and two full Newton steps. Kept OUT OF LINE: inlined into the five
// solver instances of an n = 8 event it was a quarter of the kernel's instructions, all of them dead weight in the
// instruction cache (ncu: no_instruction stalls).
template <int E> __device__ __noinline__ double massvar_tiny_target(bool small, double t) {
    double x = small ? pow(t / (double)(E + 1), 1.0 / (double)E) : sqrt(2.0 * t / (double)(E * (E + 1)));
    const MassVarPoly<E> pl(small);
    for (int it = 0; it < 2; ++it) {
        const double xn = x - div_fast(pl.residual(x, t), pl.slope(x));
        x = in_open_unit(xn) ? xn : x;
    }
    return x;
}

// `rare` (PrimRaw only): set when the target leaves the fp32 start's domain (t <= 1e-30, t <= 0, NaN): the caller redoes the
// event on PrimIEEE, whose result is bit-identical for every other t (same start, same polish).
template <int E, typename Pr = PrimIEEE> MF_D double solve_massvar(double v, bool& rare) {
    if constexpr (E == 1) {
        double u = Pr::div(v, 1.0 + Pr::sqrt(1.0 - v));
        return (v <= 0.0) ? Limits<double>::tiny_u() : ((v >= 1.0 || u > Limits<double>::max_u()) ? Limits<double>::max_u() : u);
    } else {
        // v < 0.65 on the high word (integer pipe; either variable is fine within 2^-20 of the boundary; negative v -> small,
        // NaN -> large, both end in the t <= 0 / NaN handling)
        const bool small = __double2hiint(v) < 0x3fe4cccc;
        const double t = small ? v : 1.0 - v;
        const float tf = (float)t;
        double x;
        if constexpr (std::is_same_v<Pr, PrimRaw>) {
            rare |= !(tf > 1e-30f);
            x = (double)massvar_start_f32<E>(small, tf);
        } else {
            if (tf > 1e-30f) {
                x = (double)massvar_start_f32<E>(small, tf);
            } else if (t > 0.0) {  // outside fp32 range (never taken for u in [1e-10, 1-1e-10]): out of line, see below
                x = massvar_tiny_target<E>(small, t);
            } else {
                x = (t == 0.0 || t < 0.0) ? 0.0 : t;  // t <= 0 -> 0, NaN -> NaN
            }
        }
        // fp64 polish. In either variable F is a polynomial sum_{k=2}^{E+1} c_k x^k (x^E((E+1) - E x), or x^2 P_E(x)):
        // the coefficients are selected per lane, so a warp with both kinds of lanes runs ONE instruction stream.
        // One Newton step from the fp32 start (relative error ~1e-6 -> ~1e-12) and one chord step with the same
        // reciprocal slope (-> ~1e-18): 3E + 7 fp64 operations per root, the previous two Halley steps cost ~85
        // because both variables' formulas were executed by every diverged warp.
        const MassVarPoly<E> pl(small);
        const double inv = rcp_seed_refined(pl.slope(x));
        double xn = fma(-pl.residual(x, t), inv, x);
        x = in_open_unit(xn) ? xn : x;  // also rejects NaN / inf steps
        xn = fma(-pl.residual(x, t), inv, x);
        x = in_open_unit(xn) ? xn : x;
        double u = small ? x : 1.0 - x;
        u = (u > Limits<double>::max_u()) ? Limits<double>::max_u() : u;
        return (u < Limits<double>::tiny_u()) ? Limits<double>::tiny_u() : u;  // NaN stays NaN
    }
}
template <int E, typename Pr = PrimIEEE> MF_D double solve_massvar(double v) {
    bool rare = false;
    return solve_massvar<E, Pr>(v, rare);
}
template <int E, typename Pr = PrimIEEE> MF_D float solve_massvar(float v) {
    if constexpr (E == 1) {
        float u = v / (1.0f + sqrtf(1.0f - v));
        return (v <= 0.0f) ? Limits<float>::tiny_u() : ((v >= 1.0f || u > Limits<float>::max_u()) ? Limits<float>::max_u() : u);
    } else {
        const bool small = v < 0.65f;
        const float t = small ? v : 1.0f - v;
        float x = (t > 0.0f) ? massvar_start_f32<E>(small, t) : ((t <= 0.0f) ? 0.0f : t);
        float f, d1, d2;
        MassVarFn<E, float>::eval(small, x, t, f, d1, d2);
        float xn = x - f / d1;
        x = (xn > 0.0f && xn < 1.0f) ? xn : x;
        float u = small ? x : 1.0f - x;
        u = (u > Limits<float>::max_u()) ? Limits<float>::max_u() : u;
        return (u < Limits<float>::tiny_u()) ? Limits<float>::tiny_u() : u;
    }
}

template <int E, typename Pr = PrimIEEE> MF_D float solve_massvar(float v, bool&) { return solve_massvar<E, Pr>(v); }

// get_flatWeights (:111-140) * massive correction (:489-507), from K_i, M_i and the shared pieces
// sM_i = sqrt((M_i^2-(M_{i+1}+m_i)^2)(M_i^2-(M_{i+1}-m_i)^2)), invM_i = 1/M_i:
//   8 rho(M_{n-2}, m_{n-1}, m_{n-2})                         = sM_{n-2} invM_{n-2}^2
//   prod_i rho(M_i,M_{i+1},m_i)/rho(K_i,K_{i+1},0) M_{i+1}/K_{i+1} = prod sM_i invM_i^2 K_i^2 M_{i+1} / prod (K_i^2-K_{i+1}^2) K_{i+1}
// (rho(K,K',0) = |K^2-K'^2|/(8K^2) exactly: sqrt(fl(t*t)) == |t|). One division instead of 4(n-2)+2.
// quirks reproduces the reference's float32 in-place chain.
template <typename T, int NF>
MF_D T rambo_weight(const RamboParams<T, NF>& P, T E_cm, const T (&K)[NF], const T (&M)[NF], const T (&sM)[NF],
                    const T (&invM)[NF], bool quirks) {
    const T f8 = sM[NF - 2] * (invM[NF - 2] * invM[NF - 2]);
    T num = (T)1, den = (T)1;
#pragma unroll
    for (int i = 0; i <= NF - 3; ++i) {
        num *= (sM[i] * (invM[i] * invM[i])) * ((K[i] * K[i]) * M[i + 1]);
        den *= fabs_(sub_rn(mul_rn(K[i], K[i]), mul_rn(K[i + 1], K[i + 1]))) * K[i + 1];
    }
    const T pr = (NF >= 3) ? div_fast(num, den) : (T)1;
    const T pw = ipow<2 * NF - 4>(K[0] * invM[0]);
    if (!quirks) {
        T w = P.flat_c * (ipow<NF - 2>(E_cm * E_cm) / P.flat_ff);
        w *= f8;
        w *= pr;
        w *= pw;
        return w;
    } else {
        float e = (float)E_cm;
        float e2 = __fmul_rn(e, e);
        float wf = __fmul_rn(P.flat_c32, __fdiv_rn(ipow<NF - 2>(e2), P.flat_ff32));
        wf = (float)((double)wf * (double)f8);
        wf = (float)((double)wf * (double)pr);
        wf = (float)((double)wf * (double)pw);
        return (T)wf;
    }
}

template <typename T> MF_D T pseudo_rap(const T (&p)[4]) {  // utils.py:205-215
    T pt = sqrt_(p[1] * p[1] + p[2] * p[2]);
    T th = atan2_(pt, p[3]);
    if (pt < Limits<T>::sqrt_eps() && fabs_(p[3]) < Limits<T>::sqrt_eps()) return Limits<T>::huge();
    return -log_(tan_(th / (T)2));
}
template <typename T> MF_D T delphi(const T (&a)[4], const T (&b)[4]) {  // utils.py:231-249
    T pt1 = sqrt_(a[1] * a[1] + a[2] * a[2]);
    T pt2 = sqrt_(b[1] * b[1] + b[2] * b[2]);
    T tmp = (a[1] * b[1] + a[2] * b[2]) / (pt1 * pt2);
    T r = (fabs_(tmp) > (T)1) ? acos_(tmp / fabs_(tmp)) : acos_(tmp);
    if (pt1 == (T)0 || pt2 == (T)0) r = Limits<T>::huge();
    return r;
}

// cos(2 pi u) in fp64 without the library's coefficient-table loads (cospi fetches 6 coefficients with 3 LDG.128 per call:
// ncu showed the dependent DFMAs waiting on them, `long_scoreboard`). Exact reduction u -> quarter turns: k = rint(4u),
// r = 2u - k/2, |r| <= 1/4; cos(pi r) or sin(pi r) by one degree-7 Horner form in r^2 over per-lane selected minimax
// coefficients (Remez fit in 60-digit arithmetic: relative error 4e-20 / 2e-21 before rounding); measured <= 1 ulp
// against 50-digit cos(2 pi u) for u in [0, 1]. Valid for |u| < 2^49 (callers route anything larger to the library).
MF_D double cos2pi_fast(double u) {
    const double tk = fma(4.0, u, 6755399441055744.0);   // 1.5 * 2^52: rint(4u) in the low word
    const double kd = tk - 6755399441055744.0;
    const int k = __double2loint(tk);
    const double r = fma(kd, -0.5, u + u);               // exact
    const bool odd = (k & 1) != 0;
    const double s = r * r;
    double p = odd ? -2.1716765487702665e-05 : -0.00010355601112394462;
    p = fma(p, s, odd ? 0.0004662825901993505 : 0.0019294632268243347);
    p = fma(p, s, odd ? -0.007370429872719428 : -0.025806885435598757);
    p = fma(p, s, odd ? 0.08214588657954407 : 0.23533063018157302);
    p = fma(p, s, odd ? -0.5992645293202866 : -1.3352627688517122);
    p = fma(p, s, odd ? 2.5501640398773415 : 4.058712126416745);
    p = fma(p, s, odd ? -5.16771278004997 : -4.934802200544679);
    p = fma(p, s, odd ? 3.141592653589793 : 1.0);
    const double res = (odd ? r : 1.0) * p;
    // cos(pi (k/2 + r)): k mod 4 = 0: cos, 1: -sin, 2: -cos, 3: sin
    return __hiloint2double(__double2hiint(res) ^ (((k + 1) & 2) << 30), __double2loint(res));
}
// |u| >= 2^49, inf, NaN (high word, integer pipe)
MF_D bool beyond_cos2pi_fast(double u) { return ((unsigned)__double2hiint(u) & 0x7fffffffu) >= 0x43000000u; }
// u > 1/2 on the bit pattern (false for negative u and -0)
MF_D bool above_half(double u) {
    const int hi = __double2hiint(u);
    return hi > 0x3fe00000 || (hi == 0x3fe00000 && __double2loint(u) != 0);
}

// ---------------------------------------------------------------------------------------------
// Streaming core of generateKinematics_batch (rambo_generator.py:168-398). The event is produced decay by decay:
// each iteration solves one mass variable, updates the weight accumulators, builds one two-body decay, boosts it and
// hands the finished 4-momentum to `emit(row, p)` immediately, so the live state is O(1) in the multiplicity
// (Q, two masses, four accumulators) instead of the K/M/sM/invM arrays plus all 4(n+2) output components.
//   load(i)    -> the event's i-th uniform (i is a compile-time constant after unrolling)
//   emit(j, p) -> row j of the output (0,1 incoming; 2.. final), p = (E,px,py,pz)
// ---------------------------------------------------------------------------------------------
template <typename T, int NF, int K_IDX, typename Pr, typename Load, typename Emit>
MF_D void rambo_decay_step(const RamboParams<T, NF>& P, Load& load, Emit& emit, const T K0, T& prod, T& Kk, T& Mk,
                           T (&Q)[4], T& num, T& den, T& f8, T& pw, bool& rare) {
    constexpr int NE = NF - 2;
    constexpr int k = K_IDX;
    // next intermediate mass (:448-509): K_{k+1} = K_0 prod_{j<=k} t_j, M_{k+1} = K_{k+1} + sum_{j>k} m_j
    T Kn, Mn;
    if constexpr (k < NF - 2) {
        prod *= solve_massvar<NF - 2 - k, Pr>(load(k), rare);  // exponent of column k is NF-2-k (:405-407)
        Kn = K0 * prod;
        Mn = Kn + P.msum_rev[k + 1];
    } else {
        Kn = (T)0;
        Mn = P.m[NF - 1];
    }
    const T sM = rho_num<Pr>(Mk, Mn, P.m[k]);
    const T invM = Pr::rcp(Mk);
    if constexpr (k == 0) pw = ipow<2 * NF - 4>(K0 * invM);
    if constexpr (k <= NF - 3) {  // weight accumulators, see rambo_weight() for the algebra
        num *= (sM * (invM * invM)) * ((Kk * Kk) * Mn);
        den *= fabs_(sub_rn(mul_rn(Kk, Kk), mul_rn(Kn, Kn))) * Kn;
    } else {
        f8 = sM * (invM * invM);
    }
    // two-body decay M_k -> m_k + M_{k+1} in the rest frame of Q, then boost (:310-343)
    const T qk = sM * ((T)0.5 * invM);  // 4 M_k rho(M_k, M_{k+1}, m_k)
    const T ct = (T)2 * load(NE + 2 * k) - (T)1;
    const T st = Pr::sqrt(sub_rn((T)1, mul_rn(ct, ct)));
    // cos(2 pi u) as cospi(2u): exact range reduction and no large-argument slow path (the library cos carries a
    // Payne-Hanek branch that is never taken for phi in [0, 2 pi] but was inlined once per decay: instruction-cache
    // weight). Differs from cos(fl(2 pi u)) by <= 1e-15 |sin phi|: the rounding of phi itself.
    const T uphi = load(NE + 2 * k + 1);
    T cp, sp;
    if constexpr (sizeof(T) == 8) {
        if constexpr (std::is_same_v<Pr, PrimRaw>) {
            rare |= beyond_cos2pi_fast(uphi);
            cp = cos2pi_fast(uphi);
        } else {
            cp = beyond_cos2pi_fast(uphi) ? cospi_((T)2 * uphi) : cos2pi_fast(uphi);
        }
        const T sq = Pr::sqrt(sub_rn((T)1, mul_rn(cp, cp)));
        // sin(phi) = -sqrt(..) for phi > pi (:321-322); fl(2 pi u) > fl(pi) <=> u > 1/2
        sp = above_half(uphi) ? -sq : sq;
    } else {
        const T phi = (T)6.283185307179586 * uphi;
        cp = cospi_((T)2 * uphi);
        const T sq = Pr::sqrt(sub_rn((T)1, mul_rn(cp, cp)));
        sp = (phi > (T)3.141592653589793) ? -sq : sq;
    }
    T p2[4], pb[4];
    p2[1] = (qk * st) * cp;
    p2[2] = (qk * st) * sp;
    p2[3] = qk * ct;
    const T m2 = P.m[k] * P.m[k];
    p2[0] = Pr::sqrt(rho2_3(p2[1], p2[2], p2[3]) + m2);
    if constexpr (k == 0) {  // Q = (M_0, 0, 0, 0): the boost is the identity (beta = 0 exactly)
#pragma unroll
        for (int c = 0; c < 4; ++c) pb[c] = p2[c];
    } else {
        const T iq = Pr::rcp(Q[0]);
        boost<Pr>(p2, Q[1] * iq, Q[2] * iq, Q[3] * iq, pb);
        pb[0] = Pr::sqrt(rho2_3(pb[1], pb[2], pb[3]) + m2);
    }
    emit(2 + k, pb);
    Q[1] -= pb[1]; Q[2] -= pb[2]; Q[3] -= pb[3];
    Q[0] = Pr::sqrt(rho2_3(Q[1], Q[2], Q[3]) + Mn * Mn);
    Kk = Kn;
    Mk = Mn;
}

template <typename T, int NF, typename Pr = PrimIEEE, typename Load, typename Emit>
MF_D void rambo_forward_stream(const RamboParams<T, NF>& P, Load&& load, Emit&& emit, T& wgt, T& x1, T& x2, bool& rare) {
    constexpr int D = 3 * NF - 4, NP = NF + 2;
    // ---- x1, x2 from the last two uniforms: utils.py:260-284 ----
    T wx;
    if (P.flags & MF_FLAG_X1X2_DIRECT) {  // uniform_x1x2=False, rambo_generator.py:229-234
        x1 = load(D); x2 = load(D + 1); wx = (T)1;
    } else {
        T ml1;
        if constexpr (sizeof(T) == 8) {
            // the reference's inverse-CDF expression as written (utils.py:271-272): same operations, same rounding
            ml1 = (P.t1 + Pr::sqrt(P.t1sq + ((T)-2 * load(D)) * P.inv_A)) * P.neg_A;
        } else {
            // float32: the algebraically identical q (1 - sqrt(1 - u)). The literal form subtracts 2u/A from (q/A)^2, which
            // for u -> 1 loses eps32 / (1 - u) of the radicand: 2.5e-4 relative error on x1 at 1 - u = 1e-6 (measured);
            // 1 - u is exact here and the error stays at eps32 q
            ml1 = P.q * ((T)1 - Pr::sqrt((T)1 - load(D)));
        }
        T ml2 = (P.q - ml1) * load(D + 1);
        x1 = exp_(-ml1);
        x2 = exp_(-ml2);
        wx = (P.A * x1) * x2;
    }
    const T E_cm = Pr::sqrt(x1 * x2) * P.e_collider;  // :226
    const bool quirks = (P.flags & MF_FLAG_REF_FP32_QUIRKS) != 0;
    {   // incoming partons :511-551
        T h = quirks ? (T)((float)E_cm / 2.0f) : E_cm / (T)2;
        T in1[4] = {h, (T)0, (T)0, h}, in2[4] = {h, (T)0, (T)0, -h};
        emit(0, in1);
        emit(1, in2);
    }
    const T K0 = E_cm - P.msum;
    T prod = (T)1, Kk = K0, Mk = K0 + P.msum_rev[0];
    T Q[4] = {Mk, (T)0, (T)0, (T)0};
    T num = (T)1, den = (T)1, f8 = (T)1, pw = (T)1;
    [&]<int... I>(std::integer_sequence<int, I...>) {
        (rambo_decay_step<T, NF, I, Pr>(P, load, emit, K0, prod, Kk, Mk, Q, num, den, f8, pw, rare), ...);
    }(std::make_integer_sequence<int, NF - 1>{});
    emit(NP - 1, Q);
    // weight: get_flatWeights (:111-140) * massive correction (:489-507), same product order as the reference
    const T pr = Pr::div(num, den);
    T w;
    if (!quirks) {
        w = P.flat_c * (ipow<NF - 2>(E_cm * E_cm) / P.flat_ff);
        w *= f8;
        w *= pr;
        w *= pw;
    } else {  // the reference's float32 in-place chain (:132, :489-507)
        float e = (float)E_cm;
        float e2 = __fmul_rn(e, e);
        float wf = __fmul_rn(P.flat_c32, __fdiv_rn(ipow<NF - 2>(e2), P.flat_ff32));
        wf = (float)((double)wf * (double)f8);
        wf = (float)((double)wf * (double)pr);
        wf = (float)((double)wf * (double)pw);
        w = (T)wf;
    }
    wgt = wx * w;
}

template <typename T, int NF, typename Pr = PrimIEEE, typename Load, typename Emit>
MF_D void rambo_forward_stream(const RamboParams<T, NF>& P, Load&& load, Emit&& emit, T& wgt, T& x1, T& x2) {
    bool rare = false;
    rambo_forward_stream<T, NF, Pr>(P, load, emit, wgt, x1, x2, rare);
}

// Array interface (chain kernels, lab-frame / cut epilogue): r = the event's 3NF-2 uniforms, out = all rows.
template <typename T, int NF>
MF_D void rambo_forward_event(const T (&r)[3 * NF - 2], const RamboParams<T, NF>& P, T (&out)[NF + 2][4], T& wgt,
                              T& x1, T& x2) {
    constexpr int NP = NF + 2;
    rambo_forward_stream<T, NF>(
        P, [&](int i) { return r[i]; },
        [&](int j, const T (&p)[4]) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out[j][c] = p[c];
        },
        wgt, x1, x2);

    // ---- lab frame (utils.py:180-202) for the cuts (:350-393) and/or as the requested output ----
    if (P.flags & (MF_FLAG_CUTS | MF_FLAG_LAB_FRAME)) {
        T lab[NP][4];
        if (x1 != (T)1 || x2 != (T)1) {
            T ref0 = out[0][0] * x1 + out[1][0] * x2;
            T ref3 = out[0][3] * x1 + out[1][3] * x2;
            T bz = div_fast(ref3, ref0);
#pragma unroll
            for (int k = 0; k < NP; ++k) boost(out[k], (T)0, (T)0, bz, lab[k]);
        } else {
#pragma unroll
            for (int k = 0; k < NP; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) lab[k][c] = out[k][c];
        }
        if (P.flags & MF_FLAG_CUTS) {
            T ptmin = Limits<T>::huge();
            bool ptnan = false;
#pragma unroll
            for (int k = 2; k < NP; ++k) {
                T pt = sqrt_(lab[k][1] * lab[k][1] + lab[k][2] * lab[k][2]);
                ptnan |= (pt != pt);
                ptmin = pt < ptmin ? pt : ptmin;
            }
            T f2 = (!ptnan && ptmin < P.cut_pt) ? (T)0 : (T)1;
            T eta[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) eta[k] = pseudo_rap(lab[k + 2]);
#pragma unroll
            for (int i = 0; i < NF; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    T de = eta[i] - eta[j];
                    T dp = delphi(lab[i + 2], lab[j + 2]);
                    T dr = sqrt_(de * de + dp * dp);
                    f2 *= (fabs_(dr) < P.cut_dr) ? (T)0 : (T)1;
                }
            if (P.cut_eta > (T)0) {
                T emax = -Limits<T>::huge();
                bool en = false;
#pragma unroll
                for (int k = 0; k < NF; ++k) { en |= (eta[k] != eta[k]); emax = eta[k] > emax ? eta[k] : emax; }
                f2 *= (!en && P.cut_eta < fabs_(emax)) ? (T)0 : (T)1;
            }
            wgt *= f2;
        }
        if (P.flags & MF_FLAG_LAB_FRAME) {
#pragma unroll
            for (int k = 0; k < NP; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) out[k][c] = lab[k][c];
        }
    }

}

// Host: fill the launch constants exactly as the reference's 0-dim tensors are computed.
template <typename T, int NF>
static RamboParams<T, NF> make_forward_params(const T* masses, T e_collider, const T* cuts, uint32_t flags) {
    RamboParams<T, NF> P{};
    for (int i = 0; i < NF; ++i) { P.m[i] = masses[i]; P.order[i] = i; }
    {
        T acc = (T)0;
        for (int i = NF - 1; i >= 0; --i) { acc += masses[i]; P.msum_rev[i] = acc; }
        T s = (T)0;
        for (int i = 0; i < NF; ++i) s += masses[i];
        P.msum = s;
    }
    {   // utils.py:266-269, evaluated in T like the reference's 0-dim tensors
        T ratio = P.msum / e_collider;
        T logtau = (T)std::log((double)(ratio * ratio));
        if (sizeof(T) == 4) logtau = (T)logf((float)(ratio * ratio));
        P.q = -logtau;
        P.A = (T)(-0.5) * (P.q * P.q) + (P.q * P.q);
        P.t1 = (-P.q) / P.A;
        P.t1sq = P.t1 * P.t1;
        P.neg_inv_A = (T)-1 * ((T)1 / P.A);
        P.inv_A = (T)1 / P.A;
        P.neg_A = -P.A;
    }
    P.e_collider = e_collider;
    {
        const double pi = 3.141592653589793;
        double c = std::pow(2.0 * pi, 4 - 3 * NF) * std::pow(pi / 2.0, NF - 1);
        double ff = 1.0;
        for (int i = 2; i <= NF - 1; ++i) ff *= i;
        for (int i = 2; i <= NF - 2; ++i) ff *= i;
        P.flat_c = (T)c; P.flat_ff = (T)ff;
        P.flat_c32 = (float)c; P.flat_ff32 = (float)ff;
    }
    if (cuts != nullptr && (flags & MF_FLAG_CUTS)) { P.cut_pt = cuts[0]; P.cut_dr = cuts[1]; P.cut_eta = cuts[2]; }
    else { P.cut_pt = (T)-1; P.cut_dr = (T)-1; P.cut_eta = (T)-1; flags &= ~MF_FLAG_CUTS; }
    P.flags = flags;
    return P;
}

}  // namespace mf
