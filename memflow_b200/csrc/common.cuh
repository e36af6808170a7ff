// common.cuh -- shared device helpers for the sm_100a kernels of memflow_b200.
#pragma once
// Build-time checks with a clear message: the kernels use 256-bit ld / st.global.v4.f64, 1-D
// cp.async.bulk with mbarrier completion and __grid_constant__ parameters, and are tuned for 148 SMs / 227 KB of shared memory.
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ < 1209)
#error "memflow_b200 needs nvcc >= 12.9 (sm_100a, 256-bit global loads / stores); found an older CUDA toolkit"
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "memflow_b200 is written for sm_100a (B200) only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/memflow_b200.h"

namespace mf {

#define MF_HD __host__ __device__ __forceinline__
#define MF_D __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Non-contractible arithmetic. nvcc fuses a*b+c into FMA by default; the reference (torch CPU/GPU
// elementwise kernels) rounds every product. These wrappers are used ONLY in the ill-conditioned
// expressions (1 - c*c, M^2 - (N+m)^2, 1 - beta^2, E^2 - p^2) where one extra/missing rounding is
// amplified by cancellation. Everything else may contract.
// ---------------------------------------------------------------------------------------------
MF_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
MF_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
MF_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
MF_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
MF_D double sub_rn(double a, double b) { return __dsub_rn(a, b); }
MF_D float sub_rn(float a, float b) { return __fsub_rn(a, b); }

MF_D double sqrt_(double x) { return sqrt(x); }
MF_D float sqrt_(float x) { return sqrtf(x); }
MF_D double exp_(double x) { return exp(x); }
MF_D float exp_(float x) { return expf(x); }
MF_D double log_(double x) { return log(x); }
MF_D float log_(float x) { return logf(x); }
MF_D double cospi_(double x) { return cospi(x); }
MF_D float cospi_(float x) { return cospif(x); }
MF_D double cos_(double x) { return cos(x); }
MF_D float cos_(float x) { return cosf(x); }
MF_D double atan_(double x) { return atan(x); }
MF_D float atan_(float x) { return atanf(x); }
MF_D double atan2_(double y, double x) { return atan2(y, x); }
MF_D float atan2_(float y, float x) { return atan2f(y, x); }
MF_D double acos_(double x) { return acos(x); }
MF_D float acos_(float x) { return acosf(x); }
MF_D double tan_(double x) { return tan(x); }
MF_D float tan_(float x) { return tanf(x); }
MF_D double fabs_(double x) { return fabs(x); }
MF_D float fabs_(float x) { return fabsf(x); }
MF_D double fmax_(double a, double b) { return fmax(a, b); }
MF_D float fmax_(float a, float b) { return fmaxf(a, b); }
MF_D double pow_(double a, double b) { return pow(a, b); }
MF_D float pow_(float a, float b) { return powf(a, b); }

// ---------------------------------------------------------------------------------------------
// fp64 reciprocal / division / square root without the special-case slow paths of the CUDA math library
// (each `a / b` or `sqrt(x)` there costs ~25 issue slots: MUFU seed + Newton + range checks + a call stub).
// The operands on this path are masses, energies and their ratios, far from the denormal / overflow range.
// MUFU.RCP64H / RSQ64H seed (~2^-20) + 2 Newton steps + one residual correction: faithfully rounded, equal to
// the IEEE result except for rare last-bit ties. 0, inf, NaN and negative operands keep their IEEE results through a
// final select (the MUFU seeds are IEEE-exact there); magnitudes outside [1e-280, 1e280] are not supported.
// ---------------------------------------------------------------------------------------------
// Range tests on the exponent field with integer instructions (a DSETP would occupy the half-rate FP64 pipe):
// biased exponent in [93, 1953] <=> |x| in [2^-930, 2^931) ~ [1e-280, 1e280]; false for 0, denormals, inf, NaN.
MF_D bool in_fast_range(double x) { return ((((unsigned)__double2hiint(x) >> 20) & 0x7ffu) - 93u) < 1861u; }
MF_D bool in_fast_range_pos(double x) { return (((unsigned)__double2hiint(x) >> 20) - 93u) < 1861u; }  // and x > 0
MF_D bool below_fast_max(double x) { return (((unsigned)__double2hiint(x) >> 20) & 0x7ffu) < 1954u; }   // |x| < 2^931, not NaN/inf
MF_D double rcp_seed(double b) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); return r; }
MF_D double rsqrt_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
MF_D double rcp_fast(double b) {
    const double r0 = rcp_seed(b);  // already IEEE for 0 / inf / NaN
    double e = fma(-b, r0, 1.0);
    double r = fma(r0, e, r0);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    return in_fast_range(b) ? r : r0;
}
MF_D float rcp_fast(float b) { return 1.0f / b; }
MF_D double div_fast(double a, double b) {
    const double r0 = rcp_seed(b);   // 2^-20 (scripts/fastmath_check.cu)
    const double e = fma(-b, r0, 1.0);
    const double r = fma(r0, e, r0);  // 2^-40: enough, the residual step below squares it again
    double q = a * r;
    const double rem = fma(-b, q, a);
    q = fma(rem, r, q);
    return (in_fast_range(b) && below_fast_max(a)) ? q : a * r0;  // a/0, a/inf, inf/b, NaN keep their IEEE results
}
MF_D float div_fast(float a, float b) { return a / b; }
MF_D double rsqrt_fast(double x) {
    const double y0 = rsqrt_seed(x);  // IEEE for 0 / inf / negative / NaN; 2^-20 otherwise
    const double y1 = y0 * fma(-0.5 * x * y0, y0, 1.5);
    const double e = fma(-(x * y1), y1, 1.0);  // second step in residual form: <= 2 ulp of 1 / sqrt(x)
    const double y = fma(0.5 * y1, e, y1);
    return in_fast_range_pos(x) ? y : y0;
}
MF_D float rsqrt_fast(float x) { return rsqrtf(x); }
MF_D double sqrt_fast(double x) {
    // coupled iteration on g ~ sqrt(x), h ~ 1 / (2 sqrt(x)) from the 2^-20 seed, then one residual step: 7 operations,
    // bit-identical to sqrt() on 6e8 random inputs (scripts/fastmath_check.cu)
    const double y0 = rsqrt_seed(x);
    const double g = x * y0, h = 0.5 * y0;
    const double r = fma(-g, h, 0.5);
    const double g1 = fma(g, r, g), h1 = fma(h, r, h);
    const double s = fma(fma(-g1, g1, x), h1, g1);
    // outside the fast range: sqrt(+-0) = +-0, sqrt(inf) = inf, sqrt(negative or NaN) = NaN
    // (sign clear, or -0: integer tests, a DSETP would take an fp64 issue slot in every call)
    const unsigned hi = (unsigned)__double2hiint(x);
    const bool keep = hi < 0x80000000u || (hi == 0x80000000u && __double2loint(x) == 0);
    return in_fast_range_pos(x) ? s : (keep ? x : __longlong_as_double(0x7ff8000000000000LL));
}
MF_D float sqrt_fast(float x) { return sqrtf(x); }

// ---------------------------------------------------------------------------------------------
// Primitive policies of the phase-space cores. PrimIEEE = the guarded sequences above (IEEE results for 0 / inf / NaN /
// denormal operands through a final select). PrimRaw = the same sequences WITHOUT the select: for an in-range operand
// the result is bit-identical; for 0, a denormal (flushed by the MUFU seed), inf or NaN the Newton residual is
// 0 * inf or NaN, so the result is NaN -- never a silently wrong finite number. Kernels run the event once on PrimRaw,
// look at the exponent field of every value they store, and redo the (astronomically rare) event whose outputs are not
// finite on PrimIEEE, out of line. The guards cost ~7 issued instructions per primitive (exponent extract, compare,
// BSSY / BRA / BSYNC or two FSEL plus moves): ~560 of the 3250 instructions of an n = 8 event (ncu, profiles/r1).
// ---------------------------------------------------------------------------------------------
struct PrimIEEE {
    template <typename T> static MF_D T rcp(T b) { return rcp_fast(b); }
    template <typename T> static MF_D T div(T a, T b) { return div_fast(a, b); }
    template <typename T> static MF_D T sqrt(T x) { return sqrt_fast(x); }
    template <typename T> static MF_D T rsqrt(T x) { return rsqrt_fast(x); }
};
struct PrimRaw {
    static MF_D double rcp(double b) {
        const double r0 = rcp_seed(b);
        double e = fma(-b, r0, 1.0);
        double r = fma(r0, e, r0);
        e = fma(-b, r, 1.0);
        return fma(r, e, r);
    }
    static MF_D double div(double a, double b) {
        const double r0 = rcp_seed(b);
        const double e = fma(-b, r0, 1.0);
        const double r = fma(r0, e, r0);
        const double q = a * r;
        return fma(fma(-b, q, a), r, q);
    }
    static MF_D double rsqrt(double x) {
        const double y0 = rsqrt_seed(x);
        const double y1 = y0 * fma(-0.5 * x * y0, y0, 1.5);
        const double e = fma(-(x * y1), y1, 1.0);
        return fma(0.5 * y1, e, y1);
    }
    static MF_D double sqrt(double x) {
        const double y0 = rsqrt_seed(x);
        const double g = x * y0, h = 0.5 * y0;
        const double r = fma(-g, h, 0.5);
        const double g1 = fma(g, r, g), h1 = fma(h, r, h);
        return fma(fma(-g1, g1, x), h1, g1);
    }
    static MF_D float rcp(float b) { return 1.0f / b; }
    static MF_D float div(float a, float b) { return a / b; }
    static MF_D float rsqrt(float x) { return rsqrtf(x); }
    static MF_D float sqrt(float x) { return sqrtf(x); }
};
// exponent field all ones: NaN or inf (integer pipe)
MF_D bool not_finite_bits(double x) { return ((unsigned)__double2hiint(x) & 0x7ff00000u) == 0x7ff00000u; }
MF_D bool not_finite_bits(float x) { return ((unsigned)__float_as_int(x) & 0x7f800000u) == 0x7f800000u; }

template <typename T> struct Limits;
template <> struct Limits<double> {
    static MF_HD double huge() { return 1.7976931348623157e308; }
    static MF_HD double sqrt_eps() { return 1.4901161193847656e-08; }  // np.finfo(float).eps ** 0.5
    static MF_HD double tiny_u() { return 6.525304467998525e-55; }     // 2^-180, bisection floor (see rambo.cuh)
    static MF_HD double max_u() { return 0.9999999999999999; }         // 1 - 2^-53
};
template <> struct Limits<float> {
    static MF_HD float huge() { return 3.4028234663852886e38f; }
    static MF_HD float sqrt_eps() { return 1.4901161193847656e-08f; }
    static MF_HD float tiny_u() { return 1e-30f; }
    static MF_HD float max_u() { return 0.99999994f; }
};

// integer power by repeated multiplication, compile-time exponent
template <int E, typename T> MF_D T ipow(T x) {
    if constexpr (E == 0) return (T)1;
    else if constexpr (E == 1) return x;
    else if constexpr (E % 2 == 0) { T h = ipow<E / 2>(x); return h * h; }
    else return ipow<E - 1>(x) * x;
}

// ---------------------------------------------------------------------------------------------
// Vector global memory access. A 4-momentum row is 32 B in fp64 = exactly one DRAM sector, moved
// with one 256-bit LDG/STG (sm_100 only); 16 B in fp32.
// ---------------------------------------------------------------------------------------------
MF_D void st4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
MF_D void st4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
MF_D void ld4(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
MF_D void ld4(const float* p, float& a, float& b, float& c, float& d) {
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
}
MF_D void ld2(const double* p, double& a, double& b) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
MF_D void ld2(const float* p, float& a, float& b) {
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "l"(p));
}
MF_D double ld1(const double* p) { return __ldg(p); }
MF_D float ld1(const float* p) { return __ldg(p); }

// Load one row of W elements (AoS row of the reference's `points[N, W]`). The row start is only
// guaranteed sizeof(T)*gcd-aligned, so the widest vector that divides the row pitch is used.
template <int W, typename T> MF_D void load_row(const T* __restrict__ p, T (&r)[W]) {
    if constexpr ((W % 4 == 0) && sizeof(T) * 4 <= 32) {
#pragma unroll
        for (int i = 0; i < W; i += 4) ld4(p + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
    } else if constexpr (W % 2 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 2) ld2(p + i, r[i], r[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) r[i] = ld1(p + i);
    }
}
MF_D void st2(double* p, double a, double b) {
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
MF_D void st2(float* p, float a, float b) {
    asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
template <int W, typename T> MF_D void store_row(T* __restrict__ p, const T (&r)[W]) {
    if constexpr ((W % 4 == 0) && sizeof(T) * 4 <= 32) {
#pragma unroll
        for (int i = 0; i < W; i += 4) st4(p + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
    } else if constexpr (W % 2 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 2) st2(p + i, r[i], r[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) p[i] = r[i];
    }
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MF_OK : -(int)e;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace mf
