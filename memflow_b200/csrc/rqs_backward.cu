// rqs_backward.cu -- vector-Jacobian product of K3/K4: gradients of the spline output and of its
// log-det with respect to the input and to the unconstrained parameters phi = [w | h | d] of every (event, feature) row,
// as autograd computes them through zuko's soft clip -> softmax -> cumsum -> gather -> rational quadratic.
// One row per thread. The 14 local partials (output and log-det w.r.t. x0, x1, y0, y1, d0, d1, v) come from one
// evaluation of the rational quadratic in forward-mode dual numbers; the softmax / cumsum part is closed form:
//   knot X_j = B(2 c_j - 1), c_j = sum_{i<j} p_i  =>  dL/da_i = clip'(a_i) p_i 2B [gX_hi ([i<=bin] - c_hi) + gX_lo ([i<bin] - c_lo)]
// so no per-row arrays are needed: passes over the row staged in shared memory (sums, bin, gradients written in place).
// Tiles come in by TMA bulk load and the gradient tile leaves by TMA bulk store; the per-row math uses library
// exp/log (accuracy first), so this kernel is compute-bound rather than at the HBM roofline.
#include "rqs.cuh"

namespace mf {

template <typename T> struct Dual {
    T v;
    T d[7];
};
template <typename T> MF_D Dual<T> dconst(T c) { Dual<T> r; r.v = c; for (int i = 0; i < 7; ++i) r.d[i] = (T)0; return r; }
template <typename T> MF_D Dual<T> dvar(T c, int k) { Dual<T> r = dconst(c); r.d[k] = (T)1; return r; }
template <typename T> MF_D Dual<T> operator+(const Dual<T>& a, const Dual<T>& b) { Dual<T> r; r.v = a.v + b.v; for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <typename T> MF_D Dual<T> operator-(const Dual<T>& a, const Dual<T>& b) { Dual<T> r; r.v = a.v - b.v; for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <typename T> MF_D Dual<T> operator-(const Dual<T>& a) { Dual<T> r; r.v = -a.v; for (int i = 0; i < 7; ++i) r.d[i] = -a.d[i]; return r; }
template <typename T> MF_D Dual<T> operator*(const Dual<T>& a, const Dual<T>& b) { Dual<T> r; r.v = a.v * b.v; for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <typename T> MF_D Dual<T> operator*(T a, const Dual<T>& b) { Dual<T> r; r.v = a * b.v; for (int i = 0; i < 7; ++i) r.d[i] = a * b.d[i]; return r; }
template <typename T> MF_D Dual<T> operator/(const Dual<T>& a, const Dual<T>& b) {
    Dual<T> r; const T ib = (T)1 / b.v; r.v = a.v * ib;
    for (int i = 0; i < 7; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <typename T> MF_D Dual<T> dsqrt(const Dual<T>& a) { Dual<T> r; r.v = sqrt_(a.v); const T h = (T)0.5 / r.v; for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] * h; return r; }
template <typename T> MF_D Dual<T> dlog(const Dual<T>& a) { Dual<T> r; r.v = log_(a.v); const T ia = (T)1 / a.v; for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] * ia; return r; }

// exact (library) soft clip and its derivative: c / (1 + |a c'|), d/da = 1 / (1 + |a c'|)^2
template <typename T> MF_D T clip_val(T a, T c, int clip) { return clip ? a / ((T)1 + fabs_(a * c)) : a; }
template <typename T> MF_D T clip_der(T a, T c, int clip) { if (!clip) return (T)1; const T t = (T)1 + fabs_(a * c); return (T)1 / (t * t); }

// exp(clip(a) - m): fp32 with the soft clip uses the 4-instruction exp-clip of the forward (bounded argument, m = 0)
template <typename T> MF_D T bw_exp(T a, T m, const RqsConst<T>& C) { return exp_(clip_val(a, C.two_over_L, C.clip) - m); }
template <> MF_D float bw_exp<float>(float a, float m, const RqsConst<float>& C) {
    return C.clip ? exp_clip_w(a, C) : FastMath<float>::exp(a - m);
}

// Row gradient, in place: on entry r[0..3K-2] = parameters of the row, on exit = their gradients. Returns g_v.
template <typename T, bool INV>
MF_D T rqs_row_backward(T* __restrict__ r, int K, const RqsConst<T>& C, T v, T go, T gl) {
    const int PW = 3 * K - 1;
    const T B = C.bound;
    // ---- pass 1: softmax sums. torch.softmax subtracts the row maximum; with the soft clip the arguments are bounded
    //      by |log slope| / 2, so in fp32 the subtraction is skipped (same value up to rounding) ----
    T mw = (T)0, mh = (T)0;
    if (!(C.clip && sizeof(T) == 4)) {
        mw = clip_val(r[0], C.two_over_L, C.clip); mh = clip_val(r[K], C.two_over_L, C.clip);
        for (int k = 1; k < K; ++k) {
            mw = fmax_(mw, clip_val(r[k], C.two_over_L, C.clip));
            mh = fmax_(mh, clip_val(r[K + k], C.two_over_L, C.clip));
        }
    }
    T Sw = (T)0, Sh = (T)0;
    for (int k = 0; k < K; ++k) {
        Sw += bw_exp(r[k], mw, C);
        Sh += bw_exp(r[K + k], mh, C);
    }
    const T iSw = (T)1 / Sw, iSh = (T)1 / Sh;
    // ---- pass 2: searchsorted on the searched coordinate (X forward / Y inverse) with the knots as the forward builds
    //      them, keeping the cumulative sums on both sides of the bin; then the other coordinate's prefix sums ----
    T cs_lo = (T)0, cs_hi = (T)0;
    int cnt = ((T)-B < v) ? 1 : 0;
    {
        const T ms = INV ? mh : mw, is = INV ? iSh : iSw;
        T c = (T)0;
        bool found = false;
        for (int k = 0; k < K; ++k) {
            const T cprev = c;
            c += bw_exp(r[(INV ? K : 0) + k], ms, C) * is;
            const bool below = B * ((T)2 * c - (T)1) < v;
            cnt += below ? 1 : 0;
            if (!below && !found) { cs_lo = cprev; cs_hi = c; found = true; }
        }
    }
    const int bin = cnt - 1;
    if (bin < 0 || bin >= K) {  // identity tails: out = v, ladj = 0
        for (int i = 0; i < PW; ++i) r[i] = (T)0;
        return go;
    }
    T co_lo = (T)0, co_hi = (T)0;
    {
        const T mo = INV ? mw : mh, io = INV ? iSw : iSh;
        T c = (T)0;
        for (int k = 0; k <= bin; ++k) {
            if (k == bin) co_lo = c;
            c += bw_exp(r[(INV ? 0 : K) + k], mo, C) * io;
        }
        co_hi = c;
    }
    const T cw_lo = INV ? co_lo : cs_lo, cw_hi = INV ? co_hi : cs_hi, ch_lo = INV ? cs_lo : co_lo, ch_hi = INV ? cs_hi : co_hi;
    const T x0v = B * ((T)2 * cw_lo - (T)1), x1v = B * ((T)2 * cw_hi - (T)1);
    const T y0v = B * ((T)2 * ch_lo - (T)1), y1v = B * ((T)2 * ch_hi - (T)1);
    const T d0v = bin > 0 ? exp_(clip_val(r[2 * K + bin - 1], C.one_over_L, C.clip)) : (T)1;
    const T d1v = bin < K - 1 ? exp_(clip_val(r[2 * K + bin], C.one_over_L, C.clip)) : (T)1;
    // ---- local partials by forward-mode duals: variables 0..6 = x0, x1, y0, y1, d0, d1, v ----
    const Dual<T> x0 = dvar(x0v, 0), x1 = dvar(x1v, 1), y0 = dvar(y0v, 2), y1 = dvar(y1v, 3), d0 = dvar(d0v, 4),
                  d1 = dvar(d1v, 5), vv = dvar(v, 6);
    const Dual<T> dx = x1 - x0, dy = y1 - y0;
    const Dual<T> s = dy / dx;
    const Dual<T> dd = d0 + d1 - (T)2 * s;
    const Dual<T> one = dconst((T)1);
    Dual<T> z, out;
    if constexpr (!INV) {
        z = (vv - x0) / dx;
        const Dual<T> zz = z * (one - z);
        out = y0 + dy * (s * (z * z) + d0 * zz) / (s + dd * zz);
    } else {
        const Dual<T> y_ = vv - y0;
        const Dual<T> a = dy * (s - d0) + y_ * dd;
        const Dual<T> b = dy * d0 - y_ * dd;
        const Dual<T> c = -(s * y_);
        z = ((T)2 * c) / (-b - dsqrt(b * b - (T)4 * (a * c)));
        out = x0 + z * dx;
    }
    const Dual<T> zz = z * (one - z);
    const Dual<T> den = s + dd * zz;
    const Dual<T> omz = one - z;
    const Dual<T> jac = (s * s) * ((T)2 * (s * zz) + d0 * (omz * omz) + d1 * (z * z)) / (den * den);
    const Dual<T> ladj = dlog(jac);
    T adj[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) adj[i] = go * out.d[i] + gl * ladj.d[i];
    // ---- pass 2: parameter gradients, overwriting the parameters ----
    const T twoB = (T)2 * B;
    const T gd0 = bin > 0 ? adj[4] * d0v * clip_der(r[2 * K + bin - 1], C.one_over_L, C.clip) : (T)0;
    const T gd1 = bin < K - 1 ? adj[5] * d1v * clip_der(r[2 * K + bin], C.one_over_L, C.clip) : (T)0;
    for (int i = 0; i < K; ++i) {
        const T aw = r[i], ah = r[K + i];
        const T pw = bw_exp(aw, mw, C) * iSw;
        const T ph = bw_exp(ah, mh, C) * iSh;
        const T iw_hi = (i <= bin) ? (T)1 : (T)0, iw_lo = (i < bin) ? (T)1 : (T)0;
        r[i] = clip_der(aw, C.two_over_L, C.clip) * pw * twoB * (adj[1] * (iw_hi - cw_hi) + adj[0] * (iw_lo - cw_lo));
        r[K + i] = clip_der(ah, C.two_over_L, C.clip) * ph * twoB * (adj[3] * (iw_hi - ch_hi) + adj[2] * (iw_lo - ch_lo));
    }
    for (int i = 0; i < K - 1; ++i) r[2 * K + i] = (i == bin - 1) ? gd0 : ((i == bin) ? gd1 : (T)0);
    return adj[6];
}

// Persistent CTAs; a tile of `rpt` rows comes in by one TMA bulk load, every thread turns its row into its gradient in
// place, and the tile leaves by one TMA bulk store (cp.async.bulk.global.shared::cta): parameters and gradients each
// cross HBM exactly once, fully coalesced.
template <typename T, bool INV>
__global__ void __launch_bounds__(128)
rqs_backward_kernel(const T* __restrict__ phi, const T* __restrict__ vin, const T* __restrict__ g_out,
                    const T* __restrict__ g_ladj, int64_t rows, int K, RqsConst<T> C, int rpt, int bulk_ok,
                    T* __restrict__ g_phi, T* __restrict__ g_v) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int PW = 3 * K - 1;
    const size_t tile_elems = (size_t)rpt * PW;
    T* sm = reinterpret_cast<T*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (tile_elems * sizeof(T) + 127) / 128 * 128);
    const int tid = threadIdx.x;
    const int64_t num_tiles = (rows + rpt - 1) / rpt;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0u;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * rpt;
        const int nrows = (int)((rows - row0) < rpt ? (rows - row0) : rpt);
        const bool bulk = bulk_ok && nrows == rpt;
        const T* src = phi + (size_t)row0 * PW;
        T* dst = g_phi + (size_t)row0 * PW;
        if (bulk) {
            if (tid == 0) {
                bulk_store_wait_read();  // the previous tile's store has finished reading the buffer
                const uint32_t bytes = (uint32_t)(tile_elems * sizeof(T));
                mbar_expect_tx(bar, bytes);
                bulk_g2s(sm, src, bytes, bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            if (tid == 0) bulk_store_wait_read();
            __syncthreads();
            const size_t ne = (size_t)nrows * PW;
            for (size_t i = tid; i < ne; i += 128) sm[i] = __ldg(src + i);
            __syncthreads();
        }
        for (int rr = tid; rr < nrows; rr += 128) {
            const int64_t row = row0 + rr;
            const T go = g_out ? g_out[row] : (T)0, gl = g_ladj ? g_ladj[row] : (T)0;
            g_v[row] = rqs_row_backward<T, INV>(sm + (size_t)rr * PW, K, C, vin[row], go, gl);
        }
        fence_proxy_async();  // generic writes of the gradients -> visible to the bulk store
        __syncthreads();
        if (bulk) {
            if (tid == 0) bulk_s2g(dst, sm, (uint32_t)(tile_elems * sizeof(T)));
        } else {
            const size_t ne = (size_t)nrows * PW;
            for (size_t i = tid; i < ne; i += 128) dst[i] = sm[i];
            __syncthreads();
        }
    }
    if (tid == 0) bulk_store_wait_read();
}

template <typename T, bool INV>
static int rqs_bwd_launch(const T* phi, const T* vin, const T* g_out, const T* g_ladj, int64_t B, int F, int K, T bound,
                          T slope, T* g_phi, T* g_v, cudaStream_t stream) {
    if (B < 0 || F < 1 || K < 1 || K > 128) return MF_E_SHAPE;
    if (B > 0 && (!phi || !vin || !g_phi || !g_v)) return MF_E_BADARG;
    const int64_t rows = B * (int64_t)F;
    if (rows == 0) return MF_OK;
    const RqsConst<T> C = make_rqs_const<T>(bound, slope);
    const int PW = 3 * K - 1;
    int dev = 0, sms = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    auto need = [&](int r) { return ((size_t)r * PW * sizeof(T) + 127) / 128 * 128 + 128; };
    int rpt = 128;
    while (rpt > 4 && need(rpt) > (size_t)smem_max / 2) rpt -= 4;   // >= 2 CTAs per SM when possible
    while (rpt > 4 && need(rpt) > (size_t)smem_max) rpt -= 4;
    if (need(rpt) > (size_t)smem_max) return MF_E_SHAPE;
    const int bulk_ok = (aligned(phi, 16) && aligned(g_phi, 16) && ((size_t)rpt * PW * sizeof(T)) % 16 == 0) ? 1 : 0;
    const size_t smem = need(rpt);
    const int64_t tiles = (rows + rpt - 1) / rpt;
    auto kern = rqs_backward_kernel<T, INV>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int ctas = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 128, smem);
    if (ctas < 1) ctas = 1;
    int64_t grid = (int64_t)sms * ctas;
    if (grid > tiles) grid = tiles;
    kern<<<(unsigned)grid, 128, smem, stream>>>(phi, vin, g_out, g_ladj, rows, K, C, rpt, bulk_ok, g_phi, g_v);
    return launch_status();
}

}  // namespace mf

extern "C" int mf_rqs_forward_bwd_f32(const float* phi, const float* x, const float* g_y, const float* g_ladj, int64_t B,
                                      int F, int K, float bound, float slope, float* g_phi, float* g_x, mf_stream_t s) {
    return mf::rqs_bwd_launch<float, false>(phi, x, g_y, g_ladj, B, F, K, bound, slope, g_phi, g_x, (cudaStream_t)s);
}
extern "C" int mf_rqs_forward_bwd_f64(const double* phi, const double* x, const double* g_y, const double* g_ladj, int64_t B,
                                      int F, int K, double bound, double slope, double* g_phi, double* g_x, mf_stream_t s) {
    return mf::rqs_bwd_launch<double, false>(phi, x, g_y, g_ladj, B, F, K, bound, slope, g_phi, g_x, (cudaStream_t)s);
}
extern "C" int mf_rqs_inverse_bwd_f32(const float* phi, const float* y, const float* g_x, const float* g_ladj, int64_t B,
                                      int F, int K, float bound, float slope, float* g_phi, float* g_y, mf_stream_t s) {
    return mf::rqs_bwd_launch<float, true>(phi, y, g_x, g_ladj, B, F, K, bound, slope, g_phi, g_y, (cudaStream_t)s);
}
extern "C" int mf_rqs_inverse_bwd_f64(const double* phi, const double* y, const double* g_x, const double* g_ladj, int64_t B,
                                      int F, int K, double bound, double slope, double* g_phi, double* g_y, mf_stream_t s) {
    return mf::rqs_bwd_launch<double, true>(phi, y, g_x, g_ladj, B, F, K, bound, slope, g_phi, g_y, (cudaStream_t)s);
}
