// rqs_backward.cu -- vector-Jacobian product of K3/K4: gradients of the spline output and of its
// log-det with respect to the input and to the unconstrained parameters phi = [w | h | d] of every (event, feature) row,
// as autograd computes them through zuko's soft clip -> softmax -> cumsum -> gather -> rational quadratic.
// The 14 local partials (output and log-det w.r.t. x0, x1, y0, y1, d0, d1, v) come from one evaluation of the
// rational quadratic in forward-mode dual numbers; the softmax / cumsum part is closed form:
//   knot X_j = B(2 c_j - 1), c_j = sum_{i<j} p_i  =>  dL/da_i = clip'(a_i) p_i 2B [gX_hi ([i<=bin] - c_hi) + gX_lo ([i<bin] - c_lo)]
// so no per-row arrays are needed: passes over the row staged in shared memory, each parameter replaced by its
// exponential in the first pass (clip' is recovered from it, clip_der_from_e) and by its gradient in the last.
// Tiles come in by TMA bulk load and the gradient tile leaves by TMA bulk store: parameters and gradients cross HBM
// exactly once. Short rows: one thread per row (rqs_backward_kernel); long rows: 2 / 4 / 8 threads per row
// (rqs_backward_coop_kernel). Measured on B200 (scripts/rqs_bwd_time.py, read + written bytes): fp32 F=20 K=64
// 4.5 TB/s, fp32 F=7 K=10 3.1 TB/s, fp64 F=10 K=40 2.1 TB/s.
#include <cstdlib>

#include "rqs.cuh"

namespace mf {

template <typename T, int NT> struct RDual {
    T v;
    T d[NT];
};
#define MF_DUAL template <typename T, int NT> MF_D RDual<T, NT>
#define MF_DLOOP _Pragma("unroll") for (int i = 0; i < NT; ++i)
MF_DUAL dconst(T c, const RDual<T, NT>&) { RDual<T, NT> r; r.v = c; MF_DLOOP r.d[i] = (T)0; return r; }
MF_DUAL operator+(const RDual<T, NT>& a, const RDual<T, NT>& b) { RDual<T, NT> r; r.v = a.v + b.v; MF_DLOOP r.d[i] = a.d[i] + b.d[i]; return r; }
MF_DUAL operator-(const RDual<T, NT>& a, const RDual<T, NT>& b) { RDual<T, NT> r; r.v = a.v - b.v; MF_DLOOP r.d[i] = a.d[i] - b.d[i]; return r; }
MF_DUAL operator-(const RDual<T, NT>& a) { RDual<T, NT> r; r.v = -a.v; MF_DLOOP r.d[i] = -a.d[i]; return r; }
MF_DUAL operator*(const RDual<T, NT>& a, const RDual<T, NT>& b) { RDual<T, NT> r; r.v = a.v * b.v; MF_DLOOP r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
MF_DUAL operator*(T a, const RDual<T, NT>& b) { RDual<T, NT> r; r.v = a * b.v; MF_DLOOP r.d[i] = a * b.d[i]; return r; }
MF_DUAL operator/(const RDual<T, NT>& a, const RDual<T, NT>& b) {
    RDual<T, NT> r; const T ib = rcp_fast(b.v); r.v = a.v * ib;   // fp64: MUFU seed + 2 Newton steps (common.cuh), <= 1 ulp
    MF_DLOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
MF_DUAL dsqrt(const RDual<T, NT>& a) { RDual<T, NT> r; r.v = sqrt_fast(a.v); const T h = (T)0.5 * rcp_fast(r.v); MF_DLOOP r.d[i] = a.d[i] * h; return r; }
MF_DUAL dlog(const RDual<T, NT>& a) { RDual<T, NT> r; r.v = log_(a.v); const T ia = rcp_fast(a.v); MF_DLOOP r.d[i] = a.d[i] * ia; return r; }

// Adjoints of go * out + gl * ladj w.r.t. variables [base, base + NT) of (x0, x1, y0, y1, d0, d1, v): one evaluation
// of the rational quadratic (forward: out = y, inverse: out = x; ladj = forward log|dy/dx| at x either way) on dual
// numbers that carry only those NT tangents, so the seven can be split over the threads that share a row.
template <typename T, bool INV, int NT>
MF_D void rq_adjoints(int base, T x0v, T x1v, T y0v, T y1v, T d0v, T d1v, T v, T go, T gl, T (&adj)[NT]) {
    using D = RDual<T, NT>;
    auto var = [&](T c, int idx) { D r; r.v = c; MF_DLOOP r.d[i] = (idx - base == i) ? (T)1 : (T)0; return r; };
    const D x0 = var(x0v, 0), x1 = var(x1v, 1), y0 = var(y0v, 2), y1 = var(y1v, 3), d0 = var(d0v, 4), d1 = var(d1v, 5),
            vv = var(v, 6);
    const D dx = x1 - x0, dy = y1 - y0;
    const D s = dy / dx;
    const D dd = d0 + d1 - (T)2 * s;
    const D one = dconst((T)1, s);
    D z, out;
    if constexpr (!INV) {
        z = (vv - x0) / dx;
        const D zz = z * (one - z);
        out = y0 + dy * (s * (z * z) + d0 * zz) / (s + dd * zz);
    } else {
        const D y_ = vv - y0;
        const D a = dy * (s - d0) + y_ * dd;
        const D b = dy * d0 - y_ * dd;
        const D c = -(s * y_);
        z = ((T)2 * c) / (-b - dsqrt(b * b - (T)4 * (a * c)));
        out = x0 + z * dx;
    }
    const D zz = z * (one - z);
    const D den = s + dd * zz;
    const D omz = one - z;
    const D jac = (s * s) * ((T)2 * (s * zz) + d0 * (omz * omz) + d1 * (z * z)) / (den * den);
    const D ladj = dlog(jac);
    MF_DLOOP adj[i] = go * out.d[i] + gl * ladj.d[i];
}

// exact (library) soft clip and its derivative: c / (1 + |a c'|), d/da = 1 / (1 + |a c'|)^2
template <typename T> MF_D T clip_val(T a, T c, int clip) { return clip ? a / ((T)1 + fabs_(a * c)) : a; }
template <typename T> MF_D T clip_der(T a, T c, int clip) { if (!clip) return (T)1; const T t = (T)1 + fabs_(a * c); return (T)1 / (t * t); }

// exp(clip(a) - m): fp32 with the soft clip uses the 4-instruction exp-clip of the forward (bounded argument, m = 0)
template <typename T> MF_D T bw_exp(T a, T m, const RqsConst<T>& C) { return exp_(clip_val(a, C.two_over_L, C.clip) - m); }
template <> MF_D float bw_exp<float>(float a, float m, const RqsConst<float>& C) {
    return C.clip ? exp_clip_w(a, C) : FastMath<float>::exp(a - m);
}

// d clip(a) / da recovered from e = exp(clip(a) - m): with u = clip(a) = a / (1 + c|a|), 1 / (1 + c|a|) = 1 - c|u|, so
// the derivative is (1 - c |ln e|)^2 (m = 0 whenever the clip is on in fp32; in fp64 m is added back). This lets the
// kernels overwrite each parameter by its exponential once and never exponentiate it again: one MUFU.LG2 per
// element in the gradient pass instead of a reciprocal, an exponential and a division.
template <typename T> MF_D T clip_der_from_e(T e, T m, const RqsConst<T>& C) {
    if (!C.clip) return (T)1;
    const T t = (T)1 - fabs_(C.two_over_L) * fabs_(log_(e) + m);
    return t * t;
}
template <> MF_D float clip_der_from_e<float>(float e, float, const RqsConst<float>& C) {
    if (!C.clip) return 1.0f;
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(e));
    const float t = fmaf(-C.k1w, fabsf(l), 1.0f);   // k1w = |2 / log slope| ln 2
    return t * t;
}

// ---- float64 with the soft clip (every MEMFlow production flow): the lean sequence of rqs.cuh. The first pass stores
// q = log2(exp(clip(a))) = a / (ln2 + |a| k1) in place instead of the exponential: 2^q is 14 FP64 instructions away
// (exp2_of_q) wherever the exponential is needed again, and the clip's derivative is (1 - k1 |q|)^2 -- no log, no division.
// Round 1 (library exp + IEEE division in the first pass, library log per element in the gradient pass): ~105 FP64-pipe
// instructions per parameter; now ~52. Non-finite / astronomically large parameters poison the row with NaN. ----
template <bool LEAN, typename T> MF_D T lean_store(T a, T m, const RqsConst<T>& C, T& e, int& hmax) {
    if constexpr (LEAN) {
        const double q = clip_log2(a, C.k1w, C.ln2, hmax);
        e = exp2_of_q(q);
        return q;
    } else {
        e = bw_exp(a, m, C);
        return e;
    }
}
template <bool LEAN, typename T> MF_D T lean_exp(T stored) {
    if constexpr (LEAN) return exp2_of_q(stored); else return stored;
}
template <bool LEAN, typename T> MF_D T lean_clip_der(T stored, T e, T m, const RqsConst<T>& C) {
    if constexpr (LEAN) { const double t = fma(-C.k1w, fabs(stored), 1.0); return t * t; }
    else return clip_der_from_e(e, m, C);
}
// exp(clip(d)) and d clip / dd for a derivative parameter of the selected bin
template <bool LEAN, typename T> MF_D void lean_deriv(T a, const RqsConst<T>& C, T& dv, T& f) {
    if constexpr (LEAN) {
        int hm = 0;
        const double q = clip_log2(a, C.k1d, C.ln2, hm);
        const double t = fma(-C.k1d, fabs(q), 1.0);
        dv = hm >= kHiAbsLimit ? __longlong_as_double(0x7ff8000000000000LL) : exp2_of_q(q);
        f = dv * (t * t);
    } else {
        dv = exp_(clip_val(a, C.one_over_L, C.clip));
        f = dv * clip_der(a, C.one_over_L, C.clip);
    }
}

// Row gradient, in place: on entry r[0..3K-2] = parameters of the row, on exit = their gradients. Returns g_v.
template <typename T, bool INV, bool LEAN>
MF_D T rqs_row_backward(T* __restrict__ r, int K, const RqsConst<T>& C, T v, T go, T gl, int rot0) {
    const int PW = 3 * K - 1;
    const T B = C.bound;
    // ---- pass 1: softmax sums. torch.softmax subtracts the row maximum; with the soft clip the arguments are bounded
    //      by |log slope| / 2, so in fp32 the subtraction is skipped (same value up to rounding) ----
    T mw = (T)0, mh = (T)0;
    if (!LEAN && !(C.clip && sizeof(T) == 4)) {
        mw = clip_val(r[0], C.two_over_L, C.clip); mh = clip_val(r[K], C.two_over_L, C.clip);
        for (int k = 1; k < K; ++k) {
            mw = fmax_(mw, clip_val(r[k], C.two_over_L, C.clip));
            mh = fmax_(mh, clip_val(r[K + k], C.two_over_L, C.clip));
        }
    }
    T Sw = (T)0, Sh = (T)0;
    // (rot0 >= 0: where this row starts inside the order-free passes; odd K makes the row pitch even and the rows of a
    //  warp would otherwise share 2-4 shared-memory banks at a given offset; -1: plain order)
    int hmax = 0;
    auto exp_at = [&](int k) {   // each parameter is replaced by its exponential (LEAN: by its log2), reused by every later pass
        T ew, eh;
        r[k] = lean_store<LEAN>(r[k], mw, C, ew, hmax);
        r[K + k] = lean_store<LEAN>(r[K + k], mh, C, eh, hmax);
        Sw += ew;
        Sh += eh;
    };
    if (rot0 >= 0) {   // uniform per launch: even K runs the plain loop
        int kk = rot0;
        for (int j = 0; j < K; ++j) { exp_at(kk); kk = (kk + 1 == K) ? 0 : kk + 1; }
    } else {
        for (int k = 0; k < K; ++k) exp_at(k);
    }
    if (LEAN && hmax >= kHiAbsLimit) {   // non-finite / astronomically large parameter: outside the lean sequence's domain
        for (int i = 0; i < PW; ++i) r[i] = (T)NAN;
        return (T)NAN;
    }
    const T iSw = (T)1 / Sw, iSh = (T)1 / Sh;
    // ---- pass 2: searchsorted on the searched coordinate (X forward / Y inverse) with the knots as the forward builds
    //      them, keeping the cumulative sums on both sides of the bin; then the other coordinate's prefix sums ----
    T cs_lo = (T)0, cs_hi = (T)0;
    int cnt = ((T)-B < v) ? 1 : 0;
    {
        const T is = INV ? iSh : iSw;
        T c = (T)0;
        bool found = false;
        for (int k = 0; k < K; ++k) {
            const T cprev = c;
            c += lean_exp<LEAN>(r[(INV ? K : 0) + k]) * is;
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
        const T io = INV ? iSw : iSh;
        T c = (T)0;
        for (int k = 0; k <= bin; ++k) {
            if (k == bin) co_lo = c;
            c += lean_exp<LEAN>(r[(INV ? 0 : K) + k]) * io;
        }
        co_hi = c;
    }
    const T cw_lo = INV ? co_lo : cs_lo, cw_hi = INV ? co_hi : cs_hi, ch_lo = INV ? cs_lo : co_lo, ch_hi = INV ? cs_hi : co_hi;
    const T x0v = B * ((T)2 * cw_lo - (T)1), x1v = B * ((T)2 * cw_hi - (T)1);
    const T y0v = B * ((T)2 * ch_lo - (T)1), y1v = B * ((T)2 * ch_hi - (T)1);
    T d0v = (T)1, d1v = (T)1, f0 = (T)0, f1 = (T)0;
    if (bin > 0) lean_deriv<LEAN>(r[2 * K + bin - 1], C, d0v, f0);
    if (bin < K - 1) lean_deriv<LEAN>(r[2 * K + bin], C, d1v, f1);
    // ---- local partials by forward-mode duals: variables 0..6 = x0, x1, y0, y1, d0, d1, v ----
    T adj[7];
    rq_adjoints<T, INV, 7>(0, x0v, x1v, y0v, y1v, d0v, d1v, v, go, gl, adj);
    // ---- pass 2: parameter gradients, overwriting the parameters ----
    const T twoB = (T)2 * B;
    const T gd0 = adj[4] * f0, gd1 = adj[5] * f1;
    auto grad_at = [&](int i) {
        const T sw = r[i], sh = r[K + i];
        const T ew = lean_exp<LEAN>(sw), eh = lean_exp<LEAN>(sh);
        const T iw_hi = (i <= bin) ? (T)1 : (T)0, iw_lo = (i < bin) ? (T)1 : (T)0;
        r[i] = lean_clip_der<LEAN>(sw, ew, mw, C) * (ew * iSw) * twoB * (adj[1] * (iw_hi - cw_hi) + adj[0] * (iw_lo - cw_lo));
        r[K + i] = lean_clip_der<LEAN>(sh, eh, mh, C) * (eh * iSh) * twoB * (adj[3] * (iw_hi - ch_hi) + adj[2] * (iw_lo - ch_lo));
    };
    if (rot0 >= 0) {
        int ii = rot0;
        for (int j = 0; j < K; ++j) { grad_at(ii); ii = (ii + 1 == K) ? 0 : ii + 1; }
    } else {
        for (int i = 0; i < K; ++i) grad_at(i);
    }
    for (int i = 0; i < K - 1; ++i) r[2 * K + i] = (i == bin - 1) ? gd0 : ((i == bin) ? gd1 : (T)0);
    return adj[6];
}

// Persistent CTAs; a tile of `rpt` rows comes in by one TMA bulk load, every thread turns its row into its gradient in
// place, and the tile leaves by one TMA bulk store (cp.async.bulk.global.shared::cta): parameters and gradients each
// cross HBM exactly once, fully coalesced.
template <typename T, bool INV, bool LEAN>
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
            g_v[row] = rqs_row_backward<T, INV, LEAN>(sm + (size_t)rr * PW, K, C, vin[row], go, gl, (PW % 2 == 0) ? rr % K : -1);
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

// ---- cooperative variant for long rows: LANES threads share a row ----
// One thread per row needs its whole row in shared memory, so a 3K-1 = 191-word row leaves room for only 8 warps per
// SM and every pass is a serial chain over K elements: the one-row-per-thread kernel above runs at 36 % issue
// utilisation for K = 64. Here a CTA of 256 threads owns 256 / LANES rows; the LANES threads of a row sit in DIFFERENT
// warps (warp = chunk of the row, lane = row), so a warp's 32 threads read 32 different rows at the same offset
// (row pitch 3K-1 words is odd: conflict-free) and the threads of a row meet through a few words of shared memory:
//   S1 chunk sums of exp      -> totals and the chunk's prefix offset
//   S2 chunk count of knots below v -> bin
//   S3 the chunk that holds the bin rebuilds its four knots
//   S4 the seven local adjoints are split over the row's threads (rq_adjoints with NT = ceil(7 / LANES) tangents)
//   S5 every thread writes the gradients of its chunk in place
// Knots are offset + running sum inside the chunk, so they may differ from the sequential cumsum in the last bit;
// the spline and its parameter gradients are continuous across a knot, so a bin flipped by that is harmless.
template <typename T, bool INV, int LANES, bool LEAN>
// fp64: 3 CTAs per SM (<= 85 registers, a few spilled words) beat 2 for LANES >= 4; LANES = 2 carries four tangents
// per thread in S4 and would spill 200 B
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 4 : (LANES >= 4 ? 3 : 2))
rqs_backward_coop_kernel(const T* __restrict__ phi, const T* __restrict__ vin, const T* __restrict__ g_out,
                         const T* __restrict__ g_ladj, int64_t rows, int K, RqsConst<T> C, int bulk_ok,
                         T* __restrict__ g_phi, T* __restrict__ g_v) {
    constexpr int RPT = 256 / LANES;               // rows per tile
    constexpr int NT = (7 + LANES - 1) / LANES;    // tangents per thread in S4
    constexpr int SLOTS = 4 * LANES + 12;          // [2 LANES max | cnt] [2 LANES sums] [4 knots] [8 adjoints]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int PW = 3 * K - 1;
    const size_t tile_elems = (size_t)RPT * PW;
    T* sm = reinterpret_cast<T*>(smem_raw);
    T* sc = reinterpret_cast<T*>(smem_raw + (tile_elems * sizeof(T) + 127) / 128 * 128);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sc + (size_t)SLOTS * RPT);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = wid % LANES, rr = (wid / LANES) * 32 + lane;
    const int CH = (K + LANES - 1) / LANES;
    const int k0 = min(K, chunk * CH), k1 = min(K, k0 + CH);
    const bool rot = (PW % 2) == 0;
    T* const r = sm + (size_t)rr * PW;
    T* const scM = sc + rr;                          // slot s at scM[s * RPT]
    T* const scS = sc + (size_t)2 * LANES * RPT + rr;
    T* const scK = sc + (size_t)4 * LANES * RPT + rr;
    T* const scA = sc + (size_t)(4 * LANES + 4) * RPT + rr;
    const T B = C.bound;
    const int64_t num_tiles = (rows + RPT - 1) / RPT;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0u;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * RPT;
        const int nrows = (int)((rows - row0) < RPT ? (rows - row0) : RPT);
        const bool bulk = bulk_ok && nrows == RPT;
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
            for (size_t i = tid; i < ne; i += 256) sm[i] = __ldg(src + i);
            __syncthreads();
        }
        const bool valid = rr < nrows;
        const int64_t row = row0 + rr;
        const T v = valid ? vin[row] : (T)0;
        const T go = (valid && g_out) ? g_out[row] : (T)0, gl = (valid && g_ladj) ? g_ladj[row] : (T)0;
        // ---- S0 (only without the soft clip): row maxima, as torch.softmax subtracts them ----
        T mw = (T)0, mh = (T)0;
        if (!C.clip) {
            T a = -INFINITY, b = -INFINITY;
            for (int k = k0; k < k1; ++k) { a = fmax_(a, r[k]); b = fmax_(b, r[K + k]); }
            scM[(2 * chunk) * RPT] = a;
            scM[(2 * chunk + 1) * RPT] = b;
            __syncthreads();
            mw = scM[0]; mh = scM[RPT];
#pragma unroll
            for (int c = 1; c < LANES; ++c) { mw = fmax_(mw, scM[(2 * c) * RPT]); mh = fmax_(mh, scM[(2 * c + 1) * RPT]); }
        }
        // ---- S1: chunk sums -> totals, prefix offsets ----
        T pre_s = (T)0;
        int start = k0;
        {
            int hmax = 0;
            T a = (T)0, b = (T)0;
            // (odd K: even row pitch, bank conflicts between the rows of a warp; the order inside the chunk is free here,
            //  so row r starts r elements in and wraps -- see rqs_coop.cu)
            if (rot && k1 > k0) {   // uniform branch: even K runs the plain loops
                start = k0 + rr % (k1 - k0);
                int kk = start;
                for (int j = k0; j < k1; ++j) {   // each parameter is replaced by its exponential (see clip_der_from_e)
                    const int k = kk;
                    kk = (kk + 1 == k1) ? k0 : kk + 1;
                    T ew, eh;
                    r[k] = lean_store<LEAN>(r[k], mw, C, ew, hmax);
                    r[K + k] = lean_store<LEAN>(r[K + k], mh, C, eh, hmax);
                    a += ew;
                    b += eh;
                    pre_s += (k < start) ? (INV ? eh : ew) : (T)0;   // searched array, elements before the row's start
                }
            } else {
#pragma unroll 2
                for (int k = k0; k < k1; ++k) {
                    T ew, eh;
                    r[k] = lean_store<LEAN>(r[k], mw, C, ew, hmax);
                    r[K + k] = lean_store<LEAN>(r[K + k], mh, C, eh, hmax);
                    a += ew;
                    b += eh;
                }
            }
            if (LEAN && hmax >= kHiAbsLimit) a = b = (T)NAN;   // outside the lean sequence's domain: NaN gradients for the row
            scS[(2 * chunk) * RPT] = a;
            scS[(2 * chunk + 1) * RPT] = b;
        }
        __syncthreads();
        T Sw = (T)0, Sh = (T)0, offw = (T)0, offh = (T)0;
#pragma unroll
        for (int c = 0; c < LANES; ++c) {
            if (c == chunk) { offw = Sw; offh = Sh; }
            Sw += scS[(2 * c) * RPT];
            Sh += scS[(2 * c + 1) * RPT];
        }
        const T iSw = (T)1 / Sw, iSh = (T)1 / Sh;
        const int sb = INV ? K : 0, ob = INV ? 0 : K;              // searched / other coordinate
        const T is = INV ? iSh : iSw, io = INV ? iSw : iSh;
        const T cs0 = (INV ? offh : offw) * is, co0 = (INV ? offw : offh) * io;
        // ---- S2: knots of this chunk below v ----
        {   // the same rotated walk as S1: from the row's start to the chunk's end on the running sum that begins at
            // offset + pre_s, then from the chunk's first element on the one that begins at the offset
            int cnt = 0;
            if (rot) {
                T c = cs0 + pre_s * is;
                int kk = start;
                for (int j = k0; j < k1; ++j) {
                    c += lean_exp<LEAN>(r[sb + kk]) * is;
                    cnt += (B * ((T)2 * c - (T)1) < v) ? 1 : 0;
                    if (++kk == k1) { kk = k0; c = cs0; }
                }
            } else {
                T c = cs0;
#pragma unroll 2
                for (int k = k0; k < k1; ++k) {
                    c += lean_exp<LEAN>(r[sb + k]) * is;
                    cnt += (B * ((T)2 * c - (T)1) < v) ? 1 : 0;
                }
            }
            scM[chunk * RPT] = (T)cnt;   // the max slots are dead by now (read before the S1 barrier)
        }
        __syncthreads();
        int bin = ((T)-B < v) ? 0 : -1;
#pragma unroll
        for (int c = 0; c < LANES; ++c) bin += (int)scM[c * RPT];
        const bool inside = bin >= 0 && bin < K;
        // ---- S3: the chunk that holds the bin rebuilds the four knots around it ----
        if (inside && bin >= k0 && bin < k1) {
            T cs = cs0, co = co0;
            for (int k = k0; k < bin; ++k) { cs += lean_exp<LEAN>(r[sb + k]) * is; co += lean_exp<LEAN>(r[ob + k]) * io; }
            const T cs_hi = cs + lean_exp<LEAN>(r[sb + bin]) * is, co_hi = co + lean_exp<LEAN>(r[ob + bin]) * io;
            scK[0] = INV ? co : cs;            // cw_lo
            scK[RPT] = INV ? co_hi : cs_hi;    // cw_hi
            scK[2 * RPT] = INV ? cs : co;      // ch_lo
            scK[3 * RPT] = INV ? cs_hi : co_hi;  // ch_hi
        }
        __syncthreads();
        const T cw_lo = scK[0], cw_hi = scK[RPT], ch_lo = scK[2 * RPT], ch_hi = scK[3 * RPT];
        // ---- S4: seven local adjoints, NT per thread. The two derivative words of the bin are read here by every
        //      thread of the row and rewritten in S5 by their owners: all reads are done before the barrier ----
        T f0 = (T)0, f1 = (T)0;
        if (inside) {
            T d0v = (T)1, d1v = (T)1;
            if (bin > 0) lean_deriv<LEAN>(r[2 * K + bin - 1], C, d0v, f0);
            if (bin < K - 1) lean_deriv<LEAN>(r[2 * K + bin], C, d1v, f1);
            if (chunk * NT < 7) {
                T part[NT];
                rq_adjoints<T, INV, NT>(chunk * NT, B * ((T)2 * cw_lo - (T)1), B * ((T)2 * cw_hi - (T)1),
                                        B * ((T)2 * ch_lo - (T)1), B * ((T)2 * ch_hi - (T)1), d0v, d1v, v, go, gl, part);
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    if (chunk * NT + j < 7) scA[(chunk * NT + j) * RPT] = part[j];
            }
        }
        __syncthreads();
        // ---- S5: gradients of this chunk, in place ----
        if (inside) {
            T adj[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) adj[j] = scA[j * RPT];
            const T twoB = (T)2 * B;
            const T gd0 = adj[4] * f0, gd1 = adj[5] * f1;
            // (the bracket takes three values per array: i < bin, i == bin, i > bin)
            const T gw_lt = twoB * iSw * (adj[1] * ((T)1 - cw_hi) + adj[0] * ((T)1 - cw_lo));
            const T gw_eq = twoB * iSw * (adj[1] * ((T)1 - cw_hi) - adj[0] * cw_lo);
            const T gw_gt = -twoB * iSw * (adj[1] * cw_hi + adj[0] * cw_lo);
            const T gh_lt = twoB * iSh * (adj[3] * ((T)1 - ch_hi) + adj[2] * ((T)1 - ch_lo));
            const T gh_eq = twoB * iSh * (adj[3] * ((T)1 - ch_hi) - adj[2] * ch_lo);
            const T gh_gt = -twoB * iSh * (adj[3] * ch_hi + adj[2] * ch_lo);
            auto grad_at = [&](int i) {
                const T sw = r[i], sh = r[K + i];
                const T ew = lean_exp<LEAN>(sw), eh = lean_exp<LEAN>(sh);
                r[i] = lean_clip_der<LEAN>(sw, ew, mw, C) * ew * (i < bin ? gw_lt : (i == bin ? gw_eq : gw_gt));
                r[K + i] = lean_clip_der<LEAN>(sh, eh, mh, C) * eh * (i < bin ? gh_lt : (i == bin ? gh_eq : gh_gt));
            };
            if (rot) {
                int ii = start;
                for (int j = k0; j < k1; ++j) { grad_at(ii); ii = (ii + 1 == k1) ? k0 : ii + 1; }
            } else {
#pragma unroll 2
                for (int i = k0; i < k1; ++i) grad_at(i);
            }
            for (int i = k0; i < min(k1, K - 1); ++i) r[2 * K + i] = (i == bin - 1) ? gd0 : ((i == bin) ? gd1 : (T)0);
            if (chunk == 0 && valid) g_v[row] = adj[6];
        } else {  // identity tails: out = v, ladj = 0
            for (int i = k0; i < k1; ++i) { r[i] = (T)0; r[K + i] = (T)0; }
            for (int i = k0; i < min(k1, K - 1); ++i) r[2 * K + i] = (T)0;
            if (chunk == 0 && valid) g_v[row] = go;
        }
        fence_proxy_async();  // generic writes of the gradients -> visible to the bulk store
        __syncthreads();
        if (bulk) {
            if (tid == 0) bulk_s2g(dst, sm, (uint32_t)(tile_elems * sizeof(T)));
        } else {
            const size_t ne = (size_t)nrows * PW;
            for (size_t i = tid; i < ne; i += 256) dst[i] = sm[i];
            __syncthreads();
        }
    }
    if (tid == 0) bulk_store_wait_read();
}

template <typename T, bool INV, int LANES, bool LEAN>
static int rqs_bwd_coop_launch(const T* phi, const T* vin, const T* g_out, const T* g_ladj, int64_t rows, int K,
                               const RqsConst<T>& C, T* g_phi, T* g_v, cudaStream_t stream) {
    constexpr int RPT = 256 / LANES, SLOTS = 4 * LANES + 12;
    const int PW = 3 * K - 1;
    const size_t smem = ((size_t)RPT * PW * sizeof(T) + 127) / 128 * 128 + (size_t)SLOTS * RPT * sizeof(T) + 16;
    int dev = 0, sms = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)smem_max) return MF_E_SHAPE;
    const int bulk_ok = (aligned(phi, 16) && aligned(g_phi, 16)) ? 1 : 0;   // RPT * PW * sizeof(T) is a multiple of 128
    auto kern = rqs_backward_coop_kernel<T, INV, LANES, LEAN>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int ctas = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 256, smem);
    if (ctas < 1) ctas = 1;
    const int64_t tiles = (rows + RPT - 1) / RPT;
    int64_t grid = (int64_t)sms * ctas;
    if (grid > tiles) grid = tiles;
    kern<<<(unsigned)grid, 256, smem, stream>>>(phi, vin, g_out, g_ladj, rows, K, C, bulk_ok, g_phi, g_v);
    return launch_status();
}

// ---------------------------------------------------------------------------------------------
// float64 with the soft clip, register-chunk form (round 2): EVERY PARAMETER IS EXPONENTIATED ONCE.
// The cooperative kernel above re-derives 2^q from the stored q wherever the exponential is needed again (S2 search, S3
// knots, S5 gradients: ~3.4 exp2 of 14 FP64 instructions per parameter, ~52 FP64-pipe instructions per parameter in all).
// Here thread (row, chunk) owns CH consecutive bins of BOTH softmax arrays: one pass computes q, e = 2^q and the
// gradient factor p = e (1 - k1|q|)^2 = d exp(clip(a)) / da, keeps p in REGISTERS (2 CH doubles) and stores the chunk's
// RUNNING SUM of e in place -- the search (count of chunk totals below the target + count inside one chunk) and the four
// knots around the bin then cost a handful of shared-memory loads (as in rqs_f64.cu), and the last pass is one
// multiplication and one select per parameter: ~25 FP64-pipe instructions per parameter.
//   S1   all threads: q, e, p; running sums in place; chunk totals -> tot[]
//   S3a  evaluator thread of the row (chunk LANES-1): exclusive prefixes of the chunk totals (back into tot[]), bin
//   S3b  four threads of the row, one task each: knots of the widths, knots of the heights, exp of the bin's left /
//        right derivative parameter -> pub[]   (one thread doing all of it was 35 % of the tile's critical path)
//   S4   seven local adjoints split over the row's threads (rq_adjoints, NT = ceil(7 / LANES) tangents each) -> adj[]
//   S5   all threads: gradients of their chunk from the registers, in place; the tile leaves by TMA bulk store
// Two tile buffers: the next tile's bulk copy is issued after S1 (once the bulk store of the tile before has finished
// reading that buffer) and lands while S3-S5 run.
// Odd K (even row pitch): one bulk copy per row into 16-byte aligned rows of pitch 2 (mod 4) doubles, as rqs_f64.cu.
// ---------------------------------------------------------------------------------------------
struct Rqs64BwdArgs {
    const double *phi, *vin, *g_out, *g_ladj;
    double *g_phi, *g_v;
    int64_t rows;
    RqsConst<double> C;
    int K, pitch, bulk_ok, chw;
};

// published per row: cwl cwh chl chh | iSw iSh | d0 d1 f0 f1 | v go gl | bin | Sw Sh | oc
constexpr int kBwdPub = 17;

template <bool INV, int LANES, int CH>
__global__ void __launch_bounds__(256, 3)
rqs64_backward_kernel(const __grid_constant__ Rqs64BwdArgs A) {
    constexpr int RPT = 256 / LANES;
    constexpr int NT = (7 + LANES - 1) / LANES;
    constexpr double kNaN = __builtin_nan("");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = A.K, PW = 3 * K - 1, PWP = A.pitch, chw = A.chw;
    const RqsConst<double>& C = A.C;
    const size_t buf_bytes = ((size_t)RPT * PWP * sizeof(double) + 127) / 128 * 128;
    double* const tot = reinterpret_cast<double*>(smem_raw + 2 * buf_bytes);   // [2 LANES][RPT]
    double* const pub = tot + (size_t)2 * LANES * RPT;    // [kBwdPub][RPT]
    double* const adjs = pub + (size_t)kBwdPub * RPT;     // [7][RPT]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(adjs + (size_t)7 * RPT);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = wid % LANES, rr = (wid / LANES) * 32 + lane;
    const bool evaluator = chunk == LANES - 1;
    const int k0 = min(K, chunk * chw), k1 = min(K, k0 + chw);
    const double Bd = C.bound, half_over_B = 0.5 / Bd;
    const int64_t num_tiles = (A.rows + RPT - 1) / RPT;
    const int64_t full_tiles = A.bulk_ok ? A.rows / RPT : 0;
    const uint32_t tile_bytes = (uint32_t)((size_t)RPT * PW * sizeof(double));
    // padded rows: one bulk copy / bulk store per row, one row per lane, issued by the warps that do NOT evaluate (the
    // evaluating warp is the tile's critical path between the first two barriers)
    constexpr int NCW = 8 - 8 / LANES;                    // copier warps
    constexpr int PER = (RPT + NCW - 1) / NCW;            // rows per copier warp
    const int myrow = (wid - wid / LANES) * PER + lane;   // the row this thread copies in / out
    const bool copier = !evaluator && lane < PER && myrow < RPT;
    auto buf = [&](int b) { return reinterpret_cast<double*>(smem_raw + (size_t)b * buf_bytes); };

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();
    // copy of a full tile: one bulk copy when the shared-memory rows are packed, one per row when they are padded. The
    // thread that issues a copy into rows first waits until its own earlier bulk stores have finished reading them.
    auto issue = [&](int64_t tile, int b) {
        const double* src = A.phi + (size_t)tile * RPT * PW;
        if (PWP == PW) {
            if (tid == 0) {
                bulk_store_wait_read();
                mbar_expect_tx(&bars[b], tile_bytes);
                bulk_g2s(buf(b), src, tile_bytes, &bars[b]);
            }
        } else {
            if (tid == 0) mbar_expect_tx(&bars[b], tile_bytes);
            if (copier) {
                bulk_store_wait_read();
                bulk_g2s(buf(b) + (size_t)myrow * PWP, src + (size_t)myrow * PW, (uint32_t)(PW * sizeof(double)), &bars[b]);
            }
        }
    };
    uint32_t parity0 = 0u, parity1 = 0u;
    int64_t tile = blockIdx.x;
    if (tile < full_tiles) issue(tile, 0);
    for (int it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        double* const sm = buf(b);
        double* const r = sm + (size_t)rr * PWP;
        const bool bulk = tile < full_tiles;
        const int64_t row0 = tile * RPT;
        const int nrows = (int)((A.rows - row0) < RPT ? (A.rows - row0) : RPT);
        const bool valid = rr < nrows;
        const int64_t row = row0 + rr;
        double v = 0.0, go = 0.0, gl = 0.0;
        if (evaluator && valid) {   // requested before the tile wait
            v = __ldg(A.vin + row);
            go = A.g_out ? __ldg(A.g_out + row) : 0.0;
            gl = A.g_ladj ? __ldg(A.g_ladj + row) : 0.0;
        }
        if (bulk) {
            mbar_wait(&bars[b], b ? parity1 : parity0);
            if (b) parity1 ^= 1u; else parity0 ^= 1u;
        } else {   // ragged last tile or unaligned pointers: plain cooperative copy into the same (padded) rows
            if (tid == 0 || (PWP != PW && copier)) bulk_store_wait_read();
            __syncthreads();
            for (int rw = wid; rw < nrows; rw += 8) {
                const double* src = A.phi + (size_t)(row0 + rw) * PW;
                for (int i = lane; i < PW; i += 32) sm[(size_t)rw * PWP + i] = __ldg(src + i);
            }
            __syncthreads();
        }
        // ---- S1: one exponential per parameter; gradient factors into registers, running sums in place ----
        double pw[CH], ph[CH];
        {
            double cw = 0.0, chh = 0.0;
            int hmax = 0;
            double* const rw_ = r + k0;
            double* const rh_ = r + K + k0;
            auto one = [&](int j) {
                const double aw = rw_[j], ah = rh_[j];
                const double qw = clip_log2(aw, C.k1w, C.ln2, hmax), qh = clip_log2(ah, C.k1w, C.ln2, hmax);
                const double ew = exp2_of_q(qw), eh = exp2_of_q(qh);
                const double tw = fma(-C.k1w, fabs(qw), 1.0), th = fma(-C.k1w, fabs(qh), 1.0);
                pw[j] = ew * (tw * tw);
                ph[j] = eh * (th * th);
                cw += ew;
                chh += eh;
                rw_[j] = cw;
                rh_[j] = chh;
            };
            if (k1 - k0 == CH) {   // warp-uniform (chunk = warp index mod LANES): the usual case, no per-bin test
#pragma unroll
                for (int j = 0; j < CH; ++j) one(j);
            } else {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    pw[j] = 0.0; ph[j] = 0.0;
                    if (k0 + j < k1) one(j);
                }
            }
            if (hmax >= kHiAbsLimit) cw = chh = kNaN;   // outside the lean sequence's domain
            tot[(size_t)chunk * RPT + rr] = cw;
            tot[(size_t)(LANES + chunk) * RPT + rr] = chh;
        }
        __syncthreads();
        {   // the other buffer: the store of the tile before has had S1 to finish reading it
            const int64_t nt = tile + gridDim.x;
            if (nt < full_tiles) issue(nt, b ^ 1);
        }
        double* const pb = pub + rr;
        // ---- S3a (one thread per row): exclusive prefixes of the chunk totals, bin ----
        if (evaluator && valid) {
            double* const tw_ = tot + rr;
            double* const th_ = tot + (size_t)LANES * RPT + rr;
            double Sw = 0.0, Sh = 0.0;
            int oc = 0;
            double tws[LANES], ths[LANES];
#pragma unroll
            for (int c = 0; c < LANES; ++c) { tws[c] = tw_[(size_t)c * RPT]; ths[c] = th_[(size_t)c * RPT]; }
#pragma unroll
            for (int c = 0; c < LANES; ++c) {   // inclusive running sums; the exclusive ones go back to shared memory
                tw_[(size_t)c * RPT] = Sw;
                th_[(size_t)c * RPT] = Sh;
                Sw += tws[c];
                Sh += ths[c];
                tws[c] = Sw;
                ths[c] = Sh;
            }
            const double target = fma(v, half_over_B, 0.5) * (INV ? Sh : Sw);   // v in units of the searched running sum
#pragma unroll
            for (int c = 0; c + 1 < LANES; ++c) oc += ((INV ? ths[c] : tws[c]) < target) ? 1 : 0;   // monotone: first chunk whose sum reaches the target
            const double off = (INV ? th_ : tw_)[(size_t)oc * RPT];
            const int sb = INV ? K : 0;
            const int c0 = min(K, oc * chw), len = min(K, c0 + chw) - c0;
            const double tloc = target - off;
            int lo = 0;
#pragma unroll
            for (int j = 0; j < CH; ++j) lo += (j < len && r[sb + c0 + j] < tloc) ? 1 : 0;
            int bin = c0 + lo - ((-Bd < v) ? 0 : 1);   // searchsorted(knots, v) - 1 with X_0 = -B
            if (!(Sw == Sw) || !(Sh == Sh)) {
                bin = -2;                          // NaN gradients for the whole row
                A.g_v[row] = kNaN;
            } else if (bin < 0 || bin >= K) {
                bin = -1;                          // identity tails: out = v, ladj = 0
                A.g_v[row] = go;
            }
            pb[10 * RPT] = v; pb[11 * RPT] = go; pb[12 * RPT] = gl;
            pb[13 * RPT] = (double)bin;
            pb[14 * RPT] = Sw; pb[15 * RPT] = Sh;
            pb[16 * RPT] = (double)oc;
        }
        __syncthreads();
        const int bin = valid ? (int)pb[13 * RPT] : -1;
        // ---- S3b: knots and the bin's derivative parameters, one task per thread of the row ----
        if (bin >= 0) {
            // (a valid bin lies in chunk oc; bin - 1 in the same chunk or the one before)
            auto knots = [&](int base, const double* ex, double S, int slot) {
                const int oc = (int)pb[16 * RPT];
                const double iS = rcp_fast(S);
                const int jm = max(bin - 1, 0);
                const int cm = (jm >= oc * chw) ? oc : oc - 1;
                const double lo_ = bin > 0 ? (ex[(size_t)max(cm, 0) * RPT] + r[base + jm]) * iS : 0.0;
                const double hi_ = (ex[(size_t)oc * RPT] + r[base + bin]) * iS;
                pb[slot * RPT] = lo_; pb[(slot + 1) * RPT] = hi_; pb[(4 + slot / 2) * RPT] = iS;
            };
            auto deriv = [&](bool has, int idx, int slot) {
                double d = 1.0, f = 0.0;   // D_0 = D_K = exp(0): constants
                if (has) lean_deriv<true>(r[2 * K + idx], C, d, f);
                pb[slot * RPT] = d; pb[(slot + 2) * RPT] = f;
            };
            if constexpr (LANES >= 4) {
                if (chunk == 0) knots(0, tot + rr, pb[14 * RPT], 0);
                else if (chunk == 1) knots(K, tot + (size_t)LANES * RPT + rr, pb[15 * RPT], 2);
                else if (chunk == 2) deriv(bin > 0, bin - 1, 6);
                else if (chunk == 3) deriv(bin < K - 1, bin, 7);
            } else {
                if (chunk == 0) { knots(0, tot + rr, pb[14 * RPT], 0); deriv(bin > 0, bin - 1, 6); }
                else { knots(K, tot + (size_t)LANES * RPT + rr, pb[15 * RPT], 2); deriv(bin < K - 1, bin, 7); }
            }
        }
        __syncthreads();
        // ---- S4: seven local adjoints, NT per thread ----
        if (bin >= 0 && chunk * NT < 7) {
            const double cwl = pb[0], cwh = pb[RPT], chl = pb[2 * RPT], chh = pb[3 * RPT];
            double part[NT];
            rq_adjoints<double, INV, NT>(chunk * NT, Bd * fma(2.0, cwl, -1.0), Bd * fma(2.0, cwh, -1.0), Bd * fma(2.0, chl, -1.0),
                                         Bd * fma(2.0, chh, -1.0), pb[6 * RPT], pb[7 * RPT], pb[10 * RPT], pb[11 * RPT],
                                         pb[12 * RPT], part);
#pragma unroll
            for (int j = 0; j < NT; ++j)
                if (chunk * NT + j < 7) adjs[(size_t)(chunk * NT + j) * RPT + rr] = part[j];
        }
        __syncthreads();
        // ---- S5: gradients of this chunk from the registers, in place. One path for every row: identity tails take zero
        //      multipliers, poisoned rows NaN ones (whatever their unwritten pub / adj slots hold is discarded by the selects) ----
        if (valid) {
            double* const rw_ = r + k0;
            double* const rh_ = r + K + k0;
            double* const rd_ = r + 2 * K + k0;
            const int nd = min(k1, K - 1) - k0;   // derivative parameters of this chunk
            const double* const ad = adjs + rr;
            const bool in = bin >= 0;
            const double fill = bin == -2 ? kNaN : 0.0;
            const double cwl = pb[0], cwh = pb[RPT], chl = pb[2 * RPT], chh = pb[3 * RPT];
            const double a0 = ad[0], a1 = ad[RPT], a2 = ad[2 * RPT], a3 = ad[3 * RPT];
            const double sw2 = (2.0 * Bd) * pb[4 * RPT], sh2 = (2.0 * Bd) * pb[5 * RPT];
            // (the bracket takes three values per array: i < bin, i == bin, i > bin)
            const double uw = a1 * (1.0 - cwh), lw = a0 * cwl, uh = a3 * (1.0 - chh), lh = a2 * chl;
            const double gw_lt = in ? sw2 * (uw + a0 * (1.0 - cwl)) : fill, gw_eq = in ? sw2 * (uw - lw) : fill;
            const double gw_gt = in ? -sw2 * (a1 * cwh + lw) : fill;
            const double gh_lt = in ? sh2 * (uh + a2 * (1.0 - chl)) : fill, gh_eq = in ? sh2 * (uh - lh) : fill;
            const double gh_gt = in ? -sh2 * (a3 * chh + lh) : fill;
            const double gd0 = in ? ad[4 * RPT] * pb[8 * RPT] : fill, gd1 = in ? ad[5 * RPT] * pb[9 * RPT] : fill;
            const int jb = in ? bin - k0 : -16;   // the bin's place in this chunk (may lie outside it)
            auto grad = [&](int j) {
                rw_[j] = pw[j] * (j < jb ? gw_lt : (j == jb ? gw_eq : gw_gt));
                rh_[j] = ph[j] * (j < jb ? gh_lt : (j == jb ? gh_eq : gh_gt));
            };
            auto dgrad = [&](int j) { rd_[j] = (j == jb - 1) ? gd0 : ((j == jb) ? gd1 : fill); };
            if (k1 - k0 == CH) {   // warp-uniform: the usual case, no per-bin test (a full chunk has CH or CH - 1 derivative slots)
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    grad(j);
                    if (j < CH - 1) dgrad(j);
                }
                if (nd == CH) dgrad(CH - 1);
            } else {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (k0 + j < k1) grad(j);
                    if (j < nd) dgrad(j);
                }
            }
            if (evaluator && in) A.g_v[row] = ad[6 * RPT];
        }
        fence_proxy_async();   // generic writes of the gradients -> visible to the bulk store
        __syncthreads();
        double* const dst = A.g_phi + (size_t)row0 * PW;
        if (bulk) {
            if (PWP == PW) {
                if (tid == 0) bulk_s2g(dst, sm, tile_bytes);
            } else if (copier) {
                bulk_s2g(dst + (size_t)myrow * PW, sm + (size_t)myrow * PWP, (uint32_t)(PW * sizeof(double)));
            }
        } else {
            for (int rw = wid; rw < nrows; rw += 8)
                for (int i = lane; i < PW; i += 32) dst[(size_t)rw * PW + i] = sm[(size_t)rw * PWP + i];
            __syncthreads();
        }
    }
    bulk_store_wait_read();
}

// MF_E_SHAPE when the shape is outside this kernel: the caller then takes rqs_backward_coop_kernel.
template <bool INV>
static int rqs64_bwd_launch(const double* phi, const double* vin, const double* g_out, const double* g_ladj, int64_t rows,
                            int K, const RqsConst<double>& C, double* g_phi, double* g_v, cudaStream_t stream) {
    static const bool enabled = [] { const char* e = getenv("MF_RQS64_BWD"); return !(e && e[0] == '0'); }();
    static const int forced = [] { const char* e = getenv("MF_RQS64_BWD_LANES"); return e ? atoi(e) : 0; }();
    if (!enabled || K < 6 || K > 64) return MF_E_SHAPE;
    // fewest threads per row whose chunk fits the register budget (5 bins; 8 for long rows)
    int lanes = forced ? forced : (K <= 10 ? 2 : (K <= 20 ? 4 : 8));
    if (lanes != 2 && lanes != 4 && lanes != 8) return MF_E_SHAPE;
    const int chw = (K + lanes - 1) / lanes;
    if (chw > 8) return MF_E_SHAPE;
    const int PW = 3 * K - 1;
    const bool pair = (K & 1) != 0;   // odd K: even row pitch -> padded, 16-byte aligned rows
    const int pitch = pair ? (((PW / 2) & 1) ? PW : PW + 2) : PW;
    const int rpt = 256 / lanes;
    const size_t smem = 2 * (((size_t)rpt * pitch * sizeof(double) + 127) / 128 * 128) +
                        (size_t)(2 * lanes + kBwdPub + 7) * rpt * sizeof(double) + 16;
    int dev = 0, sms = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)smem_max) return MF_E_SHAPE;
    Rqs64BwdArgs A{};
    A.phi = phi; A.vin = vin; A.g_out = g_out; A.g_ladj = g_ladj; A.g_phi = g_phi; A.g_v = g_v;
    A.rows = rows; A.C = C; A.K = K; A.pitch = pitch; A.chw = chw;
    // packed rows: tile = rpt * PW * 8 bytes, a multiple of 128; padded rows: PW even, every row is a multiple of 16 bytes
    A.bulk_ok = (aligned(phi, 16) && aligned(g_phi, 16)) ? 1 : 0;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int ctas = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 256, smem);
        if (ctas < 1) ctas = 1;
        const int64_t tiles = (rows + rpt - 1) / rpt;
        int64_t grid = (int64_t)sms * ctas;
        if (grid > tiles) grid = tiles;
        kern<<<(unsigned)grid, 256, smem, stream>>>(A);
        return launch_status();
    };
    // the register chunk is compiled for the exact chunk length: a chunk shorter than the template's takes the per-bin
    // tests in every pass (K = 45 with chunks of 6 in an 8-bin template ran no faster than the cooperative kernel)
    switch (lanes * 16 + chw) {
    case 2 * 16 + 3: return go(rqs64_backward_kernel<INV, 2, 3>);
    case 2 * 16 + 4: return go(rqs64_backward_kernel<INV, 2, 4>);
    case 2 * 16 + 5: return go(rqs64_backward_kernel<INV, 2, 5>);
    case 4 * 16 + 3: return go(rqs64_backward_kernel<INV, 4, 3>);
    case 4 * 16 + 4: return go(rqs64_backward_kernel<INV, 4, 4>);
    case 4 * 16 + 5: return go(rqs64_backward_kernel<INV, 4, 5>);
    case 8 * 16 + 3: return go(rqs64_backward_kernel<INV, 8, 3>);
    case 8 * 16 + 4: return go(rqs64_backward_kernel<INV, 8, 4>);
    case 8 * 16 + 5: return go(rqs64_backward_kernel<INV, 8, 5>);
    case 8 * 16 + 6: return go(rqs64_backward_kernel<INV, 8, 6>);
    case 8 * 16 + 7: return go(rqs64_backward_kernel<INV, 8, 7>);
    case 8 * 16 + 8: return go(rqs64_backward_kernel<INV, 8, 8>);
    default: break;
    }
    // (a forced thread count outside the table: the nearest template that holds the chunk)
    if (chw <= 5) {
        if (lanes == 2) return go(rqs64_backward_kernel<INV, 2, 5>);
        if (lanes == 4) return go(rqs64_backward_kernel<INV, 4, 5>);
        return go(rqs64_backward_kernel<INV, 8, 5>);
    }
    if (lanes == 2) return go(rqs64_backward_kernel<INV, 2, 8>);
    if (lanes == 4) return go(rqs64_backward_kernel<INV, 4, 8>);
    return go(rqs64_backward_kernel<INV, 8, 8>);
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
    // float64 with the soft clip: the lean exp2 sequence (MF_RQS64=0 keeps the library path, for A/B timing)
    static const bool lean_on = [] { const char* e = getenv("MF_RQS64"); return !(e && e[0] == '0'); }();
    const bool lean = sizeof(T) == 8 && lean_on && C.clip && (double)C.k1w < 1048576.0;
    if constexpr (sizeof(T) == 8) {   // register-chunk kernel: one exponential per parameter
        if (lean) {
            const int rc = rqs64_bwd_launch<INV>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
            if (rc != MF_E_SHAPE) return rc;
        }
    }
    {   // long rows: several threads per row. Thresholds from scripts/rqs_bwd_time.py sweeps on B200 (row bytes decide:
        // they set how many one-thread-per-row warps fit in shared memory). MF_RQS_BWD_LANES = 1/2/4/8 forces a variant.
        static const int forced = [] { const char* e = getenv("MF_RQS_BWD_LANES"); return e ? atoi(e) : 0; }();
        const size_t row_bytes = (size_t)PW * sizeof(T);
        int lanes = forced ? forced : (row_bytes < 420 ? 1 : (row_bytes < 700 ? 2 : (row_bytes < 2000 ? 4 : 8)));
        for (; lanes >= 2 && lanes <= 8; lanes *= 2) {   // a tile that does not fit: fewer rows per tile
            int rc = MF_E_SHAPE;
            if constexpr (sizeof(T) == 8) {
                if (lean) {
                    if (lanes == 2) rc = rqs_bwd_coop_launch<T, INV, 2, true>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
                    if (lanes == 4) rc = rqs_bwd_coop_launch<T, INV, 4, true>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
                    if (lanes == 8) rc = rqs_bwd_coop_launch<T, INV, 8, true>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
                    if (rc != MF_E_SHAPE) return rc;
                    continue;
                }
            }
            if (lanes == 2) rc = rqs_bwd_coop_launch<T, INV, 2, false>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
            if (lanes == 4) rc = rqs_bwd_coop_launch<T, INV, 4, false>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
            if (lanes == 8) rc = rqs_bwd_coop_launch<T, INV, 8, false>(phi, vin, g_out, g_ladj, rows, K, C, g_phi, g_v, stream);
            if (rc != MF_E_SHAPE) return rc;
        }
    }
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
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int ctas = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 128, smem);
        if (ctas < 1) ctas = 1;
        int64_t grid = (int64_t)sms * ctas;
        if (grid > tiles) grid = tiles;
        kern<<<(unsigned)grid, 128, smem, stream>>>(phi, vin, g_out, g_ladj, rows, K, C, rpt, bulk_ok, g_phi, g_v);
    };
    if constexpr (sizeof(T) == 8) {
        if (lean) { go(rqs_backward_kernel<T, INV, true>); return launch_status(); }
    }
    go(rqs_backward_kernel<T, INV, false>);
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
