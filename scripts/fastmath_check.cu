// Accuracy of the fp64 MUFU-seeded sqrt / rsqrt / rcp variants of common.cuh against the IEEE library results:
// max error in ulps over random inputs, for 1, 2 and 3 Newton iterations, plus the raw seed accuracy.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/fastmath_check scripts/fastmath_check.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsqrt_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rcp_seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

template <int IT> __device__ double sqrt_v(double x) {
    const double y0 = rsqrt_seed(x);
    const double h = 0.5 * x;
    double y = y0;
#pragma unroll
    for (int i = 0; i < IT; ++i) y = y * fma(-h * y, y, 1.5);
    double s = x * y;
    s = fma(0.5 * y, fma(-s, s, x), s);
    return s;
}
template <int IT> __device__ double rsqrt_v(double x) {
    const double y0 = rsqrt_seed(x);
    const double h = 0.5 * x;
    double y = y0;
#pragma unroll
    for (int i = 0; i < IT; ++i) y = y * fma(-h * y, y, 1.5);
    return y;
}
template <int IT> __device__ double rcp_v(double b) {
    double r = rcp_seed(b);
#pragma unroll
    for (int i = 0; i < IT; ++i) { double e = fma(-b, r, 1.0); r = fma(r, e, r); }
    return r;
}
template <int IT> __device__ double div_v(double a, double b) {
    double r = rcp_seed(b);
#pragma unroll
    for (int i = 0; i < IT; ++i) { double e = fma(-b, r, 1.0); r = fma(r, e, r); }
    double q = a * r;
    const double rem = fma(-b, q, a);
    return fma(rem, r, q);
}
__device__ double rsqrt_r(double x) {  // one standard step, then one step in residual form
    const double y0 = rsqrt_seed(x);
    const double y1 = y0 * fma(-0.5 * x * y0, y0, 1.5);
    const double e = fma(-(x * y1), y1, 1.0);
    return fma(0.5 * y1, e, y1);
}
__device__ double sqrt_g(double x) {  // coupled (Goldschmidt) form: 7 fp64 operations
    const double y0 = rsqrt_seed(x);
    const double g = x * y0, h = 0.5 * y0;
    const double r = fma(-g, h, 0.5);
    const double g1 = fma(g, r, g), h1 = fma(h, r, h);
    return fma(fma(-g1, g1, x), h1, g1);
}
__device__ double exp_bounded(double x) {  // rqs.cuh: Cody-Waite reduction, degree-13 Taylor in Estrin form, exponent add
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = 1.0 + r, p23 = fma(r, 1.0 / 6, 0.5), p45 = fma(r, 1.0 / 120, 1.0 / 24), p67 = fma(r, 1.0 / 5040, 1.0 / 720);
    const double p89 = fma(r, 1.0 / 362880, 1.0 / 40320), pab = fma(r, 1.0 / 39916800, 1.0 / 3628800);
    const double pcd = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600);
    const double q0 = fma(r2, p23, p01), q1 = fma(r2, p67, p45), q2 = fma(r2, pab, p89);
    const double o0 = fma(r4, q1, q0), o1 = fma(r4, pcd, q2);
    const double p = fma(r8, o1, o0);
    const unsigned hi = (unsigned)__double2hiint(x);
    const bool under = hi > 0xC0862000u && (hi < 0xFFF00000u || (hi == 0xFFF00000u && __double2loint(x) == 0));
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return under ? 0.0 : res;
}
__device__ double ulps(double a, double ref) {
    if (a == ref) return 0.0;
    int ex; frexp(ref, &ex);
    return fabs(a - ref) / ldexp(1.0, ex - 53);
}
__device__ uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }

__global__ void check(double* out /* [gridDim*blockDim][16] */, int per_thread) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = 0x1234567ULL * (t + 1);
    double m[16] = {0};
    double mr = 0;
    for (int i = 0; i < per_thread; ++i) {
        const uint64_t bits = splitmix(s);
        // mantissa random, exponent in [-60, 60]
        const double mant = 1.0 + (double)(bits >> 12) * (1.0 / 4503599627370496.0);
        const int ex = (int)(splitmix(s) % 121) - 60;
        const double x = ldexp(mant, ex);
        const double sq = sqrt(x), rs = 1.0 / sqrt(x), rc = 1.0 / x;
        m[0] = fmax(m[0], ulps(sqrt_v<1>(x), sq)); m[1] = fmax(m[1], ulps(sqrt_v<2>(x), sq)); m[2] = fmax(m[2], ulps(sqrt_v<3>(x), sq));
        m[3] = fmax(m[3], ulps(rsqrt_v<1>(x), rs)); m[4] = fmax(m[4], ulps(rsqrt_v<2>(x), rs)); m[5] = fmax(m[5], ulps(rsqrt_v<3>(x), rs));
        m[6] = fmax(m[6], ulps(rcp_v<1>(x), rc)); m[7] = fmax(m[7], ulps(rcp_v<2>(x), rc)); m[8] = fmax(m[8], ulps(rcp_v<3>(x), rc));
        const double a = ldexp(1.0 + (double)(splitmix(s) >> 12) * (1.0 / 4503599627370496.0), (int)(splitmix(s) % 41) - 20);
        m[11] = fmax(m[11], ulps(div_v<1>(a, x), a / x));
        mr = fmax(mr, ulps(rsqrt_r(x), rs));
        {   // exp on the spline's argument ranges: soft-clipped [-3.5, 3.5] (and the derivative's [-7, 7]), max-subtracted [-700, 0]
            const double a = -7.0 + 14.0 * ((double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0));
            const double b = -700.0 * ((double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0));
            m[13] = fmax(m[13], fmax(ulps(exp_bounded(a), exp(a)), ulps(exp_bounded(b), exp(b))));
        }
        m[12] = fmax(m[12], ulps(sqrt_g(x), sq));
        m[9] = fmax(m[9], fabs(rsqrt_seed(x) * sq - 1.0)); m[10] = fmax(m[10], fabs(rcp_seed(x) * x - 1.0));
    }
    m[14] = mr;
    for (int k = 0; k < 16; ++k) out[t * 16 + k] = m[k];
}

int main() {
    const int blocks = 592, threads = 256, per = 4096;
    double* d; cudaMalloc(&d, sizeof(double) * blocks * threads * 16);
    check<<<blocks, threads>>>(d, per);
    double* h = new double[blocks * threads * 16];
    if (cudaMemcpy(h, d, sizeof(double) * blocks * threads * 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
    double m[16] = {0};
    for (int t = 0; t < blocks * threads; ++t) for (int k = 0; k < 16; ++k) m[k] = fmax(m[k], h[t * 16 + k]);
    printf("samples %lld\n", (long long)blocks * threads * per);
    printf("sqrt  max ulp, 1/2/3 Newton steps on the rsqrt seed + residual step: %.3f %.3f %.3f\n", m[0], m[1], m[2]);
    printf("rsqrt max ulp, 1/2/3 Newton steps:                                   %.3f %.3f %.3f\n", m[3], m[4], m[5]);
    printf("rcp   max ulp, 1/2/3 Newton steps:                                   %.3f %.3f %.3f\n", m[6], m[7], m[8]);
    printf("rsqrt max ulp, 1 standard + 1 residual-form step:                    %.3f\n", m[14]);
    printf("sqrt  max ulp, coupled form (7 operations):                          %.3f\n", m[12]);
    printf("exp   max ulp, Estrin degree 13 on [-7, 7] and [-700, 0]:            %.3f\n", m[13]);
    printf("div   max ulp, 1 Newton step on the rcp seed + residual step:        %.3f\n", m[11]);
    printf("seed relative error: rsqrt %.3g (2^%.1f), rcp %.3g (2^%.1f)\n", m[9], log2(m[9]), m[10], log2(m[10]));
    return 0;
}
