// rqs.cuh -- device math of zuko's MonotonicRQSTransform (monotonic rational-quadratic spline) on one
// parameter row [w(K) | h(K) | d(K-1)] staged in shared memory. Used by rqs.cu (K3/K4) and by the fused
// importance-sampling chain (is_chain.cu).
//
// zuko (third party, un-vendored): soft clip -> softmax -> pad -> cumsum -> knots
// X_j = B(2 c_j - 1) -> searchsorted(right=False) -> gather -> rational quadratic / its inverse.
// Here the row is processed IN PLACE in shared memory by LANES (1 or 2) threads:
//   pass 1  e_k = exp(clip(w_k)) (lane 0: widths, lane 1: heights), sequential sum S
//   pass 2  knots in place: arr[k] <- B(2 * cumsum_{i<=k}(e_i / S) - 1), counting knots < v on the fly
//   epilogue (one lane)  gather the 4 knots + 2 derivatives of the selected bin, evaluate.
// Only the two derivatives of the selected bin are exponentiated (the reference exponentiates all K-1).
#pragma once
#include <cmath>

#include "common.cuh"

namespace mf {

template <typename T> struct FastMath;
template <> struct FastMath<float> {
    // softmax inputs are bounded by |log slope|/2 after the soft clip, so ex2.approx on x*log2(e)
    // (rel. error <= ~3e-7 there) and rcp.approx are within the fp32 parity budget (1e-5).
    static MF_D float exp(float x) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
        return r;
    }
    static MF_D float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static MF_D float log(float x) { return logf(x); }
};
template <> struct FastMath<double> {
    static MF_D double exp(double x) { return ::exp(x); }
    static MF_D double rcp(double x) { return 1.0 / x; }
    static MF_D double log(double x) { return ::log(x); }
};

// CTAs per SM asked of the two-sided register-row kernels (rqs_reg2_kernel, flow_sample_reg2_kernel; 128 threads). Their
// single-stage TMA tiles want as many tiles in flight per SM as fit: the compiler keeps K <= 48 within 80 registers (six
// CTAs) and K <= 32 within 72 (seven) without spills; K = 64 is limited to four by its 46 KB tiles. Measured on B200
// (profiles/r2/reg2_minb_ab.log): the chain's flow stage at the production shape (K = 40) 2.57 ms with four CTAs, 2.38 ms
// with five, 2.26 ms with six (6.6 TB/s)
#ifndef MF_REG2_MINB
#define MF_REG2_MINB(K) ((K) <= 32 ? 7 : ((K) <= 48 ? 6 : 4))
#endif
template <typename T> struct RqsConst {
    T bound;
    T two_over_L;  // 2 / log(slope)   (soft clip of widths/heights), 0 when clip disabled
    T one_over_L;  // 1 / log(slope)   (soft clip of derivatives)
    // exp(a / (1 + |a| c)) = exp2(a / (ln2 + |a| c ln2)): the log2(e) factor folded into the reciprocal's argument
    T ln2, k1w, k1d;  // ln 2, |2/L| ln 2, |1/L| ln 2
    int clip;      // slope > 0
};

template <typename T> inline RqsConst<T> make_rqs_const(T bound, T slope) {
    RqsConst<T> C;
    C.bound = bound;
    C.clip = slope > (T)0 ? 1 : 0;
    const double L = C.clip ? std::log((double)slope) : 1.0;
    const double ln2 = 0.6931471805599453;
    C.two_over_L = C.clip ? (T)(2.0 / L) : (T)0;
    C.one_over_L = C.clip ? (T)(1.0 / L) : (T)0;
    C.ln2 = (T)ln2;
    C.k1w = C.clip ? (T)(std::fabs(2.0 / L) * ln2) : (T)0;
    C.k1d = C.clip ? (T)(std::fabs(1.0 / L) * ln2) : (T)0;
    return C;
}

MF_D double soft_clip(double a, double c) { return a / (1.0 + fabs(a * c)); }
MF_D float soft_clip(float a, float c) { return a * FastMath<float>::rcp(fmaf(fabsf(a), fabsf(c), 1.0f)); }  // 1 FFMA (|.| modifiers)

// One thread transforms one of the two softmax arrays of a row into knots, in place.
// Returns the number of knots (including X_0 = -B) strictly below v  == torch.searchsorted(knots, v).
template <typename T>
MF_D int rqs_knots_inplace(T* __restrict__ arr, int K, const RqsConst<T>& C, T v) {
    T S = (T)0;
    if (C.clip) {
        for (int k = 0; k < K; ++k) {
            T a = arr[k];
            a = soft_clip(a, C.two_over_L);
            T e = FastMath<T>::exp(a);
            S += e;
            arr[k] = e;
        }
    } else {
        T mx = arr[0];
        for (int k = 1; k < K; ++k) mx = fmax_(mx, arr[k]);
        for (int k = 0; k < K; ++k) {
            T e = FastMath<T>::exp(arr[k] - mx);
            S += e;
            arr[k] = e;
        }
    }
    const T inv = FastMath<T>::rcp(S);
    T c = (T)0;
    int cnt = ((T)-C.bound < v) ? 1 : 0;  // X_0 = B(2*0 - 1)
    for (int k = 0; k < K; ++k) {
        c += arr[k] * inv;
        T x = C.bound * ((T)2 * c - (T)1);
        arr[k] = x;
        cnt += (x < v) ? 1 : 0;
    }
    return cnt;
}

// Epilogue, knots already in place: Xk[j-1] = X_j (j = 1..K), X_0 = -B. bin = searchsorted - 1.
// INV == false: v = x -> y, ladj = log|dy/dx|.  INV == true: v = y -> x, ladj = log|dy/dx| at that x.
template <typename T, bool INV>
MF_D void rqs_eval(const T* __restrict__ Xk, const T* __restrict__ Yk, const T* __restrict__ d, int K,
                   const RqsConst<T>& C, int bin, T v, T& out, T& ladj) {
    if (bin < 0 || bin >= K) { out = v; ladj = (T)0; return; }  // identity tails: mask == False
    const T x0 = bin == 0 ? -C.bound : Xk[bin - 1], x1 = Xk[bin];
    const T y0 = bin == 0 ? -C.bound : Yk[bin - 1], y1 = Yk[bin];
    T d0 = (T)1, d1 = (T)1;  // D_0 = D_K = exp(0)
    if (bin > 0) { T t = d[bin - 1]; if (C.clip) t = soft_clip(t, C.one_over_L); d0 = exp_(t); }
    if (bin < K - 1) { T t = d[bin]; if (C.clip) t = soft_clip(t, C.one_over_L); d1 = exp_(t); }
    const T dx = x1 - x0, dy = y1 - y0;
    const T s = dy / dx;
    const T dd = d0 + d1 - (T)2 * s;
    T z;
    if constexpr (!INV) {
        z = (v - x0) / dx;
        const T zz = z * ((T)1 - z);
        const T den = s + dd * zz;
        out = y0 + dy * (s * (z * z) + d0 * zz) / den;
    } else {
        const T y_ = v - y0;
        const T a = dy * (s - d0) + y_ * dd;
        const T b = dy * d0 - y_ * dd;
        const T c = -s * y_;
        z = (T)2 * c / (-b - sqrt_(b * b - (T)4 * a * c));
        out = x0 + z * dx;
    }
    const T zz = z * ((T)1 - z);
    const T den = s + dd * zz;
    const T omz = (T)1 - z;
    const T jac = (s * s) * ((T)2 * s * zz + d0 * (omz * omz) + d1 * (z * z)) / (den * den);
    ladj = log_(jac);
}

// Whole row by ONE thread (used when rows are small: the fused chain, K <= ~16).
template <typename T, bool INV>
MF_D void rqs_row_1lane(T* __restrict__ row, int K, const RqsConst<T>& C, T v, T& out, T& ladj, int& bin) {
    const T big = Limits<T>::huge();
    int cx = rqs_knots_inplace(row, K, C, INV ? big : v);
    int cy = rqs_knots_inplace(row + K, K, C, INV ? v : big);
    bin = (INV ? cy : cx) - 1;
    rqs_eval<T, INV>(row, row + K, row + 2 * K, K, C, bin, v, out, ladj);
}

// exp(soft_clip(a)) in 4 instructions: FFMA (|a| modifier), MUFU.RCP, FMUL, MUFU.EX2
MF_D float exp_clip_w(float a, const RqsConst<float>& C) {
    float r, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(fabsf(a), C.k1w, C.ln2)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * r));
    return e;
}
MF_D float exp_clip_d(float a, const RqsConst<float>& C) {
    float r, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(fabsf(a), C.k1d, C.ln2)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * r));
    return e;
}

// Rational quadratic of one bin from its 4 knot coordinates (fp32, reciprocal-multiply instead of divisions).
template <bool INV>
MF_D void rqs_eval_fast(const float* __restrict__ d, int K, const RqsConst<float>& C, int bin, float x0, float x1, float y0,
                        float y1, float v, float& out, float& ladj) {
    float d0 = 1.0f, d1 = 1.0f;
    if (bin > 0) d0 = C.clip ? exp_clip_d(d[bin - 1], C) : FastMath<float>::exp(d[bin - 1]);
    if (bin < K - 1) d1 = C.clip ? exp_clip_d(d[bin], C) : FastMath<float>::exp(d[bin]);
    const float dx = x1 - x0, dy = y1 - y0;
    const float rdx = FastMath<float>::rcp(dx);
    const float s = dy * rdx;
    const float dd = d0 + d1 - 2.0f * s;
    float z;
    if constexpr (!INV) {
        z = (v - x0) * rdx;
        const float zz = z * (1.0f - z);
        out = y0 + dy * (s * (z * z) + d0 * zz) * FastMath<float>::rcp(s + dd * zz);
    } else {
        const float y_ = v - y0;
        const float a = dy * (s - d0) + y_ * dd;
        const float b = dy * d0 - y_ * dd;
        const float cc = -s * y_;
        z = 2.0f * cc * FastMath<float>::rcp(-b - sqrtf(b * b - 4.0f * a * cc));
        out = x0 + z * dx;
    }
    const float zz = z * (1.0f - z);
    const float rden = FastMath<float>::rcp(s + dd * zz);
    const float omz = 1.0f - z;
    const float jac = (s * s) * (2.0f * s * zz + d0 * (omz * omz) + d1 * (z * z)) * (rden * rden);
    ladj = logf(jac);
}

// One softmax array of a row into registers: e[k] = exp(clip(a_k)), returns the sequential sum.
template <int K>
MF_D float rqs_exp_row(const float* __restrict__ arr, const RqsConst<float>& C, float (&e)[K]) {
    float S = 0.0f;
    if (C.clip) {
#pragma unroll
        for (int k = 0; k < K; ++k) { e[k] = exp_clip_w(arr[k], C); S += e[k]; }
    } else {
        float m = arr[0];
#pragma unroll
        for (int k = 1; k < K; ++k) m = fmaxf(m, arr[k]);
#pragma unroll
        for (int k = 0; k < K; ++k) { e[k] = FastMath<float>::exp(arr[k] - m); S += e[k]; }
    }
    return S;
}
// Bin search on the unnormalised running sum; returns bin = searchsorted - 1 and the two knots of that bin.
template <int K>
MF_D int rqs_search_row(const float (&e)[K], float S, float bound, float v, float& k0, float& k1) {
    const float target = fmaf(v, 0.5f / bound, 0.5f) * S;
    int cnt = (-bound < v) ? 1 : 0;
    float c = 0.0f, c_lo = 0.0f, c_hi = 3.0e38f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        c += e[k];
        const bool below = c < target;
        cnt += below ? 1 : 0;
        c_lo = below ? c : c_lo;
        c_hi = below ? c_hi : fminf(c_hi, c);
    }
    const float is2 = 2.0f * FastMath<float>::rcp(S);
    k0 = bound * fmaf(c_lo, is2, -1.0f);
    k1 = bound * fmaf(c_hi, is2, -1.0f);
    return cnt - 1;
}
// Knots of bin `bin` of the other coordinate: prefix sums at bin and bin + 1.
template <int K>
MF_D void rqs_prefix_row(const float (&e)[K], float S, float bound, int bin, float& k0, float& k1) {
    float o = 0.0f, o_lo = 0.0f, o_hi = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        o += e[k];
        o_lo = (k < bin) ? o : o_lo;
        o_hi = (k <= bin) ? o : o_hi;
    }
    const float io2 = 2.0f * FastMath<float>::rcp(S);
    k0 = bound * fmaf(o_lo, io2, -1.0f);
    k1 = bound * fmaf(o_hi, io2, -1.0f);
}

// ---------------------------------------------------------------------------------------------
// Register-resident row for small compile-time K (the sampling chain: K = 8..16). One thread, fully unrolled:
// each softmax array is exponentiated once into registers; the bin is found by comparing the UNNORMALISED
// running sum with v mapped into that space ((v/B+1)/2 * S), so there is no per-element normalisation; the
// running sums go back to the row in place and the 4 knots of the selected bin are 4 indexed loads. fp32 only.
// ---------------------------------------------------------------------------------------------
template <int K, bool INV>
MF_D void rqs_row_reg(float* __restrict__ row, const RqsConst<float>& C, float v, float& out, float& ladj, int& bin) {
    // x = searched array (heights for the inverse, widths for the forward), y = the other one, handled as float2
    // pairs so that the multiply and the two running sums issue as packed FMUL2 / FADD2 (sm_100 f32x2). One pass:
    // exponentiate, accumulate, write the running sums back in place; then a binary search for the bin and 4 indexed
    // loads for its knots.
    float* rs = row + (INV ? K : 0);
    float* ro = row + (INV ? 0 : K);
    float2 c = make_float2(0.0f, 0.0f);  // running sums (searched, other); live state is O(1) in K
    if (C.clip) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 a = make_float2(rs[k], ro[k]);
            float2 r, e;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(fmaf(fabsf(a.x), C.k1w, C.ln2)));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(fmaf(fabsf(a.y), C.k1w, C.ln2)));
            const float2 y = __fmul2_rn(a, r);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(y.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(y.y));
            c = __fadd2_rn(c, e);
            rs[k] = c.x;
            ro[k] = c.y;
        }
    } else {
        float ms = rs[0], mo = ro[0];
#pragma unroll
        for (int k = 1; k < K; ++k) { ms = fmaxf(ms, rs[k]); mo = fmaxf(mo, ro[k]); }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            c = __fadd2_rn(c, make_float2(FastMath<float>::exp(rs[k] - ms), FastMath<float>::exp(ro[k] - mo)));
            rs[k] = c.x;
            ro[k] = c.y;
        }
    }
    const float Ss = c.x, So = c.y;
    const float target = fmaf(v, 0.5f / C.bound, 0.5f) * Ss;  // v in units of the unnormalised cumulative sum
    // number of running sums below the target: binary search on the (ascending) sums just written
    int lo = 0, hi = K;
#pragma unroll
    for (int step = 0; (1 << step) <= K; ++step) {
        const int mid = (lo + hi) >> 1;
        const bool go_right = (lo < hi) && (rs[mid < K ? mid : K - 1] < target);
        const bool go_left = (lo < hi) && !go_right;
        lo = go_right ? mid + 1 : lo;
        hi = go_left ? mid : hi;
    }
    const int cnt = ((-C.bound < v) ? 1 : 0) + lo;
    bin = cnt - 1;
    if (bin < 0 || bin >= K) { out = v; ladj = 0.0f; return; }
    const float is2 = 2.0f * FastMath<float>::rcp(Ss), io2 = 2.0f * FastMath<float>::rcp(So);
    const float c_lo = bin > 0 ? rs[bin - 1] : 0.0f, c_hi = rs[bin];
    const float o_lo = bin > 0 ? ro[bin - 1] : 0.0f, o_hi = ro[bin];
    const float s0 = C.bound * fmaf(c_lo, is2, -1.0f), s1 = C.bound * fmaf(c_hi, is2, -1.0f);
    const float o0 = C.bound * fmaf(o_lo, io2, -1.0f), o1 = C.bound * fmaf(o_hi, io2, -1.0f);
    const float x0 = INV ? o0 : s0, x1 = INV ? o1 : s1, y0 = INV ? s0 : o0, y1 = INV ? s1 : o1;
    rqs_eval_fast<INV>(row + 2 * K, K, C, bin, x0, x1, y0, y1, v, out, ladj);
}

// ---------------------------------------------------------------------------------------------
// float64 exp(soft_clip(a)) = 2^(a / (ln2 + |a| k1)), k1 = |f / log slope| ln2 (f = 2 widths / heights, 1 derivatives), as ONE
// branch-free sequence of 18 FP64 instructions (the library's exp + a range-checked division cost ~95 issued instructions
// per parameter in round 1): the log2(e) factor is folded into the divisor, the reciprocal is MUFU.RCP64H (2^-20) with a
// cubic correction r0 (1 + e + e^2) (2^-60 before rounding), 2^q = 2^k P11(q - k) with a degree-11 minimax polynomial
// (|error| <= 3.2e-18, Chebyshev fit in 60-digit arithmetic) and an integer add into the exponent field. The soft clip
// bounds |q| <= |log2 slope| / 2 (5 for slope = 1e-3), so there is no range check; callers track the largest |a| (one
// integer max per element) and take an exact out-of-line path for non-finite / astronomically large parameters.
// Relative error <= ~6e-16 (tests compare against exp() of the IEEE quotient).
// ---------------------------------------------------------------------------------------------
static __constant__ double kExp2[12] = {1.0, 6.9314718055994530925e-1, 2.4022650695910159567e-1, 5.5504108664821627039e-2,
                                        9.6181291075872566681e-3, 1.3333558146406470697e-3, 1.5403530463724354209e-4,
                                        1.5252733841556772589e-5, 1.3215432535912376166e-6, 1.0178057087733941105e-7,
                                        7.0741942972885210056e-9, 4.455817908336064493e-10};
// parameters at or above 2^993 in magnitude, inf and NaN leave the fast sequence's domain
constexpr int kHiAbsLimit = 0x7e000000;
// 2^q for |q| < 1000: q = k + r, |r| <= 1/2
MF_D double exp2_of_q(double q) {
    const double tk = q + 6755399441055744.0;       // 1.5 * 2^52: rint(q) in the low word
    const double kd = tk - 6755399441055744.0;
    const double r = q - kd;                        // exact
    double p = kExp2[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, r, kExp2[i]);
    return __hiloint2double(__double2hiint(p) + (__double2loint(tk) << 20), __double2loint(p));
}
// q = a / (ln2 + |a| k1) = log2 of exp(soft_clip(a)); hi_abs_max tracks the largest exponent field seen
MF_D double clip_log2(double a, double k1, double ln2, int& hi_abs_max) {
    hi_abs_max = max(hi_abs_max, __double2hiint(a) & 0x7fffffff);
    const double bq = fma(fabs(a), k1, ln2);
    const double r0 = rcp_seed(bq);                 // MUFU.RCP64H: 2^-20
    const double e = fma(-bq, r0, 1.0);
    const double t = fma(e, e, e);
    const double s = fma(r0, t, r0);                // 1 / bq
    return a * s;
}
MF_D double exp2_clip(double a, double k1, double ln2, int& hi_abs_max) { return exp2_of_q(clip_log2(a, k1, ln2, hi_abs_max)); }

// Bulk async copy (TMA, 1-D) global -> shared with mbarrier completion.
MF_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MF_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MF_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MF_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MF_DONE;\n"
        "bra MF_WAIT;\n"
        "MF_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
MF_D void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk store (TMA, 1-D) with bulk-group completion
MF_D void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
MF_D void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
MF_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MF_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// rqs_coop.cu: several threads per row; MF_E_SHAPE when no variant fits the shape
template <typename T, bool INV>
int rqs_coop_launch(const T* phi, const T* vin, int64_t B, int F, int K, T bound, T slope, T* vout, T* ladj, T* ladj_sum,
                    int32_t* bin, cudaStream_t stream, int64_t row_offset = 0);


// rqs_f64.cu: float64 with the soft clip (every MEMFlow production flow): K3 / K4 (T = 1) and the chain's whole flow stage
// (chain = true: Philox base point -> T inverses -> u in [0,1], log q) in one launch; MF_E_SHAPE when the shape is outside it
int rqs64_launch(bool inverse, bool chain, const double* phi, const double* vin, int64_t B, int F, int K, int T, double bound,
                 double slope, uint64_t seed, uint64_t event_offset, double* vout, double* ladj, double* ladj_sum,
                 int32_t* bin, cudaStream_t stream, int64_t event_stride = 0);

}  // namespace mf
