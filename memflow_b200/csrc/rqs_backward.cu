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
