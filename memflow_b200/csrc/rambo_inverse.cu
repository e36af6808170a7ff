// rambo_inverse.cu -- K2: CM-frame final-state momenta -> unit hypercube + inverse weight.
// Replaces FlatInvertiblePhasespace.getPSpoint_batch (memflow/phasespace/rambo_generator.py:615-736)
// and get_uniform_from_x1x2 (memflow/phasespace/utils.py:287-309). One event per thread.
//
// HBM layout: momenta rows are read with one 256-bit LDG each (32 B = one sector, fp64); the
// event pitch is a parameter so that the `momenta[:, 2:]` slice of a forward output is read in
// place (the reference clones it, :628, and again per particle, :680).
// Algorithmic bytes per event: (4n + 2 + 3n-2 + 1) * sizeof(T).
#include <cstdlib>

#include "rambo.cuh"

namespace mf {

template <typename T, int NF>
__global__ void __launch_bounds__(256, (NF <= 3 ? 4 : (NF <= 4 ? 3 : 2)))
rambo_inverse_kernel(const T* __restrict__ momenta, int64_t event_stride, const T* __restrict__ x1a,
                     const T* __restrict__ x2a, int64_t N, const __grid_constant__ RamboParams<T, NF> P,
                     T* __restrict__ ps, T* __restrict__ prob, T* __restrict__ logprob,
                     int32_t* __restrict__ status, const T* __restrict__ boost_reco, uint8_t* __restrict__ mask,
                     T* __restrict__ logit_out) {
    constexpr int D = 3 * NF - 4, W = D + 2;
    // threads past the end redo the last event and write nothing: the CTA stores its rows cooperatively (flush_rows)
    const int64_t ev_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = ev_raw < N;
    const int64_t ev = valid ? ev_raw : N - 1;
    // Rows of 3n-2 values are 8-byte aligned only when 3n-2 is odd (n = 3, 5, 7): per-thread stores would be 8-byte
    // scatters over 32 sectors per instruction (n = 7 ran slower than n = 8). The CTA's 256 rows are contiguous in
    // memory, so for odd widths they are staged in shared memory and leave as 128-bit coalesced stores (n = 7: +30 %,
    // n = 5: +7 %, n = 3: +4 %); even widths keep their per-thread 128 / 256-bit stores (staging cost n = 6 30 %).
    constexpr bool STAGED = (W % 2) == 1;
    __shared__ __align__(16) T stage[STAGED ? 256 * W : 1];
    const int64_t ev0 = (int64_t)blockIdx.x * blockDim.x;
    const int nrows = (int)((N - ev0) < (int64_t)blockDim.x ? (N - ev0) : (int64_t)blockDim.x);
    auto flush_rows = [&](T* __restrict__ gbase, const T (&row)[W]) {
        if constexpr (!STAGED) {
            if (valid) store_row<W>(gbase + ev * W, row);
            return;
        }
        __syncthreads();  // the previous flush has been read out
#pragma unroll
        for (int c = 0; c < W; ++c) stage[threadIdx.x * W + c] = row[c];
        __syncthreads();
        T* g = gbase + ev0 * W;
        const int total = nrows * W;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            constexpr int V = 16 / sizeof(T);
            const int nv = total / V;
            const uint4* s4 = reinterpret_cast<const uint4*>(stage);
            uint4* g4 = reinterpret_cast<uint4*>(g);
            for (int i = threadIdx.x; i < nv; i += blockDim.x) g4[i] = s4[i];
            for (int i = nv * V + threadIdx.x; i < total; i += blockDim.x) g[i] = stage[i];
        } else {
            for (int i = threadIdx.x; i < total; i += blockDim.x) g[i] = stage[i];
        }
    };

    T Pm[NF][4];
    {
        const T* base = momenta + ev * event_stride;
#pragma unroll
        for (int k = 0; k < NF; ++k) ld4(base + 4 * P.order[k], Pm[k][0], Pm[k][1], Pm[k][2], Pm[k][3]);
    }
    // get_PS mode (memflow/unfolding_flow/utils.py:1056-1152): x1,x2 from the reconstructed boost, problems -> mask
    const bool getps = boost_reco != nullptr;
    bool problem = false;
    T x1, x2;
    if (getps) {
        T b0, b1, b2, b3;
        ld4(boost_reco + 4 * ev, b0, b1, b2, b3);
        x1 = (b0 + b3) / P.e_collider;
        x2 = (b0 - b3) / P.e_collider;
        problem = (x1 > (T)1) | (x1 < (T)0) | (x2 < (T)0) | (x1 > (T)1);  // sic (:1068): x2 > 1 is never tested
        x1 = x1 < (T)0 ? (T)0 : (x1 > (T)1 ? (T)1 : x1);
        x2 = x2 < (T)0 ? (T)0 : (x2 > (T)1 ? (T)1 : x2);
    } else {
        x1 = x1a[ev];
        x2 = x2a[ev];
    }

    if (getps || (!(P.flags & MF_FLAG_NO_CM_CHECK) && status != nullptr)) {  // :630-635 / get_PS :1079-1081
        T s1 = (T)0, s2 = (T)0, s3 = (T)0;
#pragma unroll
        for (int k = 0; k < NF; ++k) { s1 += Pm[k][1]; s2 += Pm[k][2]; s3 += Pm[k][3]; }
        T r2 = (s1 * s1 + s2 * s2) + s3 * s3;
        if (getps) problem |= (r2 > (T)1e-5);
        else if (valid && !(r2 < (T)1e-3)) atomicOr(status, MF_ST_NOT_CM);
    }

    // M_j = sqrt((sum_{k>=j} P_k)^2), K_j = M_j - sum_{k>=j} m_k   (:656-660; sums in the reference's order)
    T M[NF], K[NF];
#pragma unroll
    for (int j = NF - 1; j >= 0; --j) {
        T s[4] = {(T)0, (T)0, (T)0, (T)0};
#pragma unroll
        for (int k = j; k < NF; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += Pm[k][c];
        T sq = sub_rn(sub_rn(sub_rn(mul_rn(s[0], s[0]), mul_rn(s[1], s[1])), mul_rn(s[2], s[2])), mul_rn(s[3], s[3]));
        M[j] = sqrt_fast(sq);
        K[j] = M[j] - P.msum_rev[j];
    }

    T r[W];
    T Q[4] = {Pm[NF - 1][0], Pm[NF - 1][1], Pm[NF - 1][2], Pm[NF - 1][3]};
    // i = NF .. 2 (j = i-1): mass variable r[j-1], angles r[NF-5+2i-1], r[NF-4+2i-1]   (:665-696)
    [&]<int... I>(std::integer_sequence<int, I...>) {
        ((void)[&] {
            constexpr int i = NF - I;  // NF, NF-1, ..., 2
            constexpr int j = i - 1;
            const T uu = div_fast(K[j], K[j - 1]);
            if constexpr (i < NF)  // i == NF writes r[NF-2] = 1, overwritten by the i == 2 angle (:672,:681)
                r[j - 1] = (T)(NF + 1 - i) * ipow<NF - i>(uu) - (T)(NF - i) * ipow<NF + 1 - i>(uu);
#pragma unroll
            for (int c = 0; c < 4; ++c) Q[c] = Q[c] + Pm[j - 1][c];
            T pb[4];
            const T niq = -rcp_fast(Q[0]);
            boost(Pm[j - 1], Q[1] * niq, Q[2] * niq, Q[3] * niq, pb);
            r[NF - 5 + 2 * i - 1] = ((pb[3] * rsqrt_fast(rho2_3(pb[1], pb[2], pb[3]))) + (T)1) * (T)0.5;
            T phi = atan_(div_fast(pb[2], pb[1]));
            T dphi = (pb[2] < (T)0 && pb[1] > (T)0) ? (T)6.283185307179586 : (T)0;
            dphi += (pb[1] < (T)0) ? (T)3.141592653589793 : (T)0;
            phi += dphi;
            r[NF - 4 + 2 * i - 1] = phi * (T)0.15915494309189535;  // 1 / (2 pi)
        }(), ...);
    }(std::make_integer_sequence<int, NF - 1>{});

    // x1,x2 -> uniforms (utils.py:287-309)
    T jacinv;
    if (P.flags & MF_FLAG_X1X2_DIRECT) {
        r[D] = x1; r[D + 1] = x2; jacinv = (T)1;
    } else {
        const T ml1 = -log_(x1), ml2 = -log_(x2);
        r[D] = ((T)(-0.5) * (ml1 * ml1) + P.q * ml1) * P.inv_A;
        r[D + 1] = div_fast(ml2, P.q - ml1);
        jacinv = rcp_fast((P.A * x1) * x2);
    }
    if (getps) {  // no RAMBO weight in get_PS: returns 0 * jac_x1x2 (:1139,1152) and the problem mask
#pragma unroll
        for (int c = 0; c < D; ++c) problem |= (r[c] < (T)0) | (r[c] > (T)1);
        flush_rows(ps, r);
        if (valid && prob != nullptr) prob[ev] = (T)0 * jacinv;
        if (valid && mask != nullptr) mask[ev] = problem ? 1 : 0;
        if (logit_out != nullptr) {  // torch.logit(ps, eps) and affine scaling, fused (unfolding_flow.py:170-171)
            const T lo = P.logit_eps, hi = (T)1 - P.logit_eps;
#pragma unroll
            for (int c = 0; c < W; ++c) {
                T z = r[c] < lo ? lo : (r[c] > hi ? hi : r[c]);
                r[c] = (log_(z / ((T)1 - z)) - P.lmean[c]) / P.lscale[c];
            }
            flush_rows(logit_out, r);
        }
        return;
    }
    const T E_cm = sqrt_fast(x1 * x2) * P.e_collider;
    const bool quirks = (P.flags & MF_FLAG_REF_FP32_QUIRKS) != 0;
    T sM[NF], invM[NF];
#pragma unroll
    for (int k = 0; k <= NF - 2; ++k) {
        sM[k] = rho_num(M[k], (k == NF - 2) ? P.m[NF - 1] : M[k + 1], P.m[k]);  // :709-713 uses m_{n-1}, not M_{n-1}
        invM[k] = rcp_fast(M[k]);
    }
    sM[NF - 1] = (T)0; invM[NF - 1] = (T)0;
    const T jac = rambo_weight<T, NF>(P, E_cm, K, M, sM, invM, quirks);
    T pr;
    if (!quirks) {
        pr = rcp_fast(jac) * jacinv;
    } else {
        float pf = __fdiv_rn(1.0f, (float)jac);
        pr = (T)(float)((double)pf * (double)jacinv);
    }
    flush_rows(ps, r);
    if (valid) {
        prob[ev] = pr;
        if (logprob != nullptr) logprob[ev] = log_(pr);
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming form (fp64, every n but 4; not the get_PS / float32-quirk modes). The kernel above holds the whole event: 4n
// momenta, M[n], K[n] and the 3n-2 outputs, 124-128 registers at n = 7, 8, i.e. 2 CTAs per SM, and its FP64 chains
// stall on their own latency with 4 warps per scheduler. Everything the inverse needs at step j is the running sum
// Q_j = p_{n-1} + ... + p_j, p_j itself and (M, K) of step j+1, so here particles are consumed last to first, each
// loaded when its turn comes, the weight is accumulated on the way (the same pieces as rambo_weight) and every
// unit-cube coordinate goes straight into the CTA's shared-memory staging row (coalesced 128-bit stores at the end).
// Difference to the array form: M_j^2 comes from Q_j (summed last to first) instead of a separate first-to-last sum
// (the reference's torch.sum, :656): a rounding-level change of an ill-conditioned quantity, inside the
// conditioning-aware bound of tests/parity.py (gamma^2 of the system).
// ---------------------------------------------------------------------------------------------
// One event of the streaming form. put(c, value) receives unit-cube coordinate c; returns prob, leaves the summed final
// state in Q (for the CM check). Pr = primitive policy (common.cuh).
template <int NF, typename Pr, typename Put>
MF_D double rambo_inverse_stream_event(const RamboParams<double, NF>& P, const double* __restrict__ base, double x1, double x2,
                                       Put&& put, double (&Q)[4]) {
    using T = double;
    constexpr int D = 3 * NF - 4;
    T Mn = (T)0, Kn = (T)0;                    // M_{j+1}, K_{j+1}
    T num = (T)1, den = (T)1, f8 = (T)1, pw = (T)1;
    T pn[4];                                   // the next particle is requested one step ahead of its use
    ld4(base + 4 * P.order[NF - 1], pn[0], pn[1], pn[2], pn[3]);
    [&]<int... I>(std::integer_sequence<int, I...>) {
        ((void)[&] {
            constexpr int j = NF - 1 - I;      // NF-1, ..., 0
            const T p[4] = {pn[0], pn[1], pn[2], pn[3]};
            if constexpr (j > 0) ld4(base + 4 * P.order[j - 1], pn[0], pn[1], pn[2], pn[3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) Q[c] = (j == NF - 1) ? p[c] : Q[c] + p[c];
            const T sq = sub_rn(sub_rn(sub_rn(mul_rn(Q[0], Q[0]), mul_rn(Q[1], Q[1])), mul_rn(Q[2], Q[2])), mul_rn(Q[3], Q[3]));
            const T Mj = Pr::sqrt(sq);
            const T Kj = Mj - P.msum_rev[j];
            if constexpr (j <= NF - 2) {
                if constexpr (j < NF - 2) {    // mass variable (:665-672), exponent n-2-j
                    const T uu = Pr::div(Kn, Kj);
                    put(j, (T)(NF - 1 - j) * ipow<NF - 2 - j>(uu) - (T)(NF - 2 - j) * ipow<NF - 1 - j>(uu));
                }
                T pb[4];                       // p_j in the rest frame of Q_j -> (cos theta, phi) (:680-696)
                const T niq = -Pr::rcp(Q[0]);
                boost<Pr>(p, Q[1] * niq, Q[2] * niq, Q[3] * niq, pb);
                put(NF + 2 * j - 2, ((pb[3] * Pr::rsqrt(rho2_3(pb[1], pb[2], pb[3]))) + (T)1) * (T)0.5);
                T phi = atan_(Pr::div(pb[2], pb[1]));
                T dphi = (pb[2] < (T)0 && pb[1] > (T)0) ? (T)6.283185307179586 : (T)0;
                dphi += (pb[1] < (T)0) ? (T)3.141592653589793 : (T)0;
                put(NF + 2 * j - 1, (phi + dphi) * (T)0.15915494309189535);  // 1 / (2 pi)
                // weight pieces of the pair (j, j+1), see rambo_weight(); :709-713 uses m_{n-1}, not M_{n-1}
                const T sM = rho_num<Pr>(Mj, (j == NF - 2) ? P.m[NF - 1] : Mn, P.m[j]);
                const T invM = Pr::rcp(Mj);
                if constexpr (j == NF - 2) {
                    f8 = sM * (invM * invM);
                } else {
                    num *= (sM * (invM * invM)) * ((Kj * Kj) * Mn);
                    den *= fabs_(sub_rn(mul_rn(Kj, Kj), mul_rn(Kn, Kn))) * Kn;
                }
                if constexpr (j == 0) pw = ipow<2 * NF - 4>(Kj * invM);
            }
            Mn = Mj;
            Kn = Kj;
        }(), ...);
    }(std::make_integer_sequence<int, NF>{});
    T jacinv;  // x1, x2 -> uniforms (utils.py:287-309)
    if (P.flags & MF_FLAG_X1X2_DIRECT) {
        put(D, x1); put(D + 1, x2); jacinv = (T)1;
    } else {
        const T ml1 = -log_(x1), ml2 = -log_(x2);
        put(D, ((T)(-0.5) * (ml1 * ml1) + P.q * ml1) * P.inv_A);
        put(D + 1, Pr::div(ml2, P.q - ml1));
        jacinv = Pr::rcp((P.A * x1) * x2);
    }
    const T E_cm = Pr::sqrt(x1 * x2) * P.e_collider;
    T w = P.flat_c * (ipow<NF - 2>(E_cm * E_cm) / P.flat_ff);
    w *= f8;
    w *= Pr::div(num, den);
    w *= pw;
    return Pr::rcp(w) * jacinv;
}

// The event again on the guarded primitives, out of line (see rambo_forward_redo in rambo_forward.cu)
template <int NF>
__device__ __noinline__ double rambo_inverse_stream_redo(const RamboParams<double, NF>& P, const double* base, double x1, double x2,
                                                         double* row, double* Qout) {
    double Q[4] = {0.0, 0.0, 0.0, 0.0};
    const double pr = rambo_inverse_stream_event<NF, PrimIEEE>(P, base, x1, x2, [&](int c, double v) { row[c] = v; }, Q);
    Qout[0] = Q[1]; Qout[1] = Q[2]; Qout[2] = Q[3];
    return pr;
}

// 4 CTAs per SM (64 registers, no spills) up to n = 7: measured +3 % (n = 5), +19 % (n = 6), +5 % (n = 7) on one box; neutral at n = 8
#ifndef MF_K2_MINB
#define MF_K2_MINB (NF <= 7 ? 4 : 3)
#endif
template <int NF>
__global__ void __launch_bounds__(256, MF_K2_MINB)
rambo_inverse_stream_kernel(const double* __restrict__ momenta, int64_t event_stride, const double* __restrict__ x1a,
                            const double* __restrict__ x2a, int64_t N, const __grid_constant__ RamboParams<double, NF> P,
                            double* __restrict__ ps, double* __restrict__ prob, double* __restrict__ logprob,
                            int32_t* __restrict__ status) {
    using T = double;
    constexpr int D = 3 * NF - 4, W = D + 2;
    // row pitch: W = 16 doubles (n = 6) would start every row of a warp in the same bank (32-way conflicts on every store)
    constexpr int WP = (W % 16 == 0) ? W + 1 : W;
    __shared__ __align__(16) T stage[256 * WP];
    const int64_t ev_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = ev_raw < N;
    const int64_t ev = valid ? ev_raw : N - 1;  // threads past the end redo the last event and write nothing
    T* const row = stage + threadIdx.x * WP;
    const T* base = momenta + ev * event_stride;
    const T x1 = x1a[ev], x2 = x2a[ev];

    // speculative pass on the unguarded primitives: unit-cube coordinates lie in [0, 1] (exponent field <= 0x3ff), so the OR
    // of their high words has an all-ones exponent only if one of them is NaN / inf (never a false negative)
    T Q[4] = {(T)0, (T)0, (T)0, (T)0};
    unsigned hi_or = 0u;
    T pr = rambo_inverse_stream_event<NF, PrimRaw>(
        P, base, x1, x2, [&](int c, T v) { hi_or |= (unsigned)__double2hiint(v); row[c] = v; }, Q);
    if ((hi_or & 0x7ff00000u) == 0x7ff00000u || not_finite_bits(pr)) {
        T q3[3];
        pr = rambo_inverse_stream_redo<NF>(P, base, x1, x2, row, q3);
        Q[1] = q3[0]; Q[2] = q3[1]; Q[3] = q3[2];
    }
    if (!(P.flags & MF_FLAG_NO_CM_CHECK) && status != nullptr) {  // :630-635: Q_0 is the summed final state
        const T r2 = (Q[1] * Q[1] + Q[2] * Q[2]) + Q[3] * Q[3];
        if (valid && !(r2 < (T)1e-3)) atomicOr(status, MF_ST_NOT_CM);
    }
    if (valid) {
        prob[ev] = pr;
        if (logprob != nullptr) logprob[ev] = log_(pr);
    }
    // the CTA's rows are contiguous in memory: coalesced 128-bit stores
    __syncthreads();
    const int64_t ev0 = (int64_t)blockIdx.x * blockDim.x;
    const int nrows = (int)((N - ev0) < (int64_t)blockDim.x ? (N - ev0) : (int64_t)blockDim.x);
    T* g = ps + ev0 * W;
    const int total = nrows * W;
    if constexpr (WP != W) {
        int r = threadIdx.x / W, c = threadIdx.x - r * W;
        for (int i = threadIdx.x; i < total; i += 256) {
            g[i] = stage[r * WP + c];
            r += 256 / W;
            c += 256 % W;
            if (c >= W) { c -= W; ++r; }
        }
    } else if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int nv = total / 2;
        const uint4* s4 = reinterpret_cast<const uint4*>(stage);
        uint4* g4 = reinterpret_cast<uint4*>(g);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) g4[i] = s4[i];
        for (int i = nv * 2 + threadIdx.x; i < total; i += blockDim.x) g[i] = stage[i];
    } else {
        for (int i = threadIdx.x; i < total; i += blockDim.x) g[i] = stage[i];
    }
}

template <typename T, int NF>
static int launch_inverse(const T* momenta, int64_t event_stride, const T* x1, const T* x2, int64_t N,
                          const int32_t* order, const T* masses, const T* target_mass, T e_collider,
                          uint32_t flags, T* ps, T* prob, T* logprob, int32_t* status, cudaStream_t stream,
                          const T* boost = nullptr, uint8_t* mask = nullptr, bool permute_target = false,
                          T logit_eps = (T)0, const T* lmean = nullptr, const T* lscale = nullptr, T* logit_out = nullptr) {
    RamboParams<T, NF> P{};
    P.logit_eps = logit_eps;
    for (int c = 0; c < 3 * NF - 2; ++c) { P.lmean[c] = lmean ? lmean[c] : (T)0; P.lscale[c] = lscale ? lscale[c] : (T)1; }
    bool seen[NF] = {};
    for (int i = 0; i < NF; ++i) {
        int o = order ? order[i] : i;
        if (o < 0 || o >= NF || seen[o]) return MF_E_SHAPE;
        seen[o] = true;
        P.order[i] = o;
        P.m[i] = target_mass ? target_mass[permute_target ? o : i] : masses[o];  // :649-652 (get_PS permutes, :1074)
    }
    for (int j = 0; j < NF; ++j) {  // torch.sum(masses_t[:, j:]) in forward order (:660)
        T s = (T)0;
        for (int k = j; k < NF; ++k) s += P.m[k];
        P.msum_rev[j] = s;
    }
    {
        T s = (T)0;
        for (int i = 0; i < NF; ++i) s += (permute_target && target_mass) ? target_mass[i] : masses[i];  // tot_final_state_masses / target_mass.sum()
        P.msum = s;
        T ratio = P.msum / e_collider;
        T logtau = (sizeof(T) == 4) ? (T)logf((float)(ratio * ratio)) : (T)std::log((double)(ratio * ratio));
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
    P.flags = flags;
    if (N == 0) return MF_OK;
    const int threads = 256;
    const int64_t blocks = (N + threads - 1) / threads;
    if constexpr (sizeof(T) == 8 && NF != 4) {   // measured vs the array form: n = 3 +3 %, 4 -3 %, 5 +17 %, 6 +1 %, 7 +16 %, 8 +23 %
        // MF_K2_STREAM=0 keeps the array-form kernel (A/B timing)
        static const bool stream_form = [] { const char* e = getenv("MF_K2_STREAM"); return !(e && e[0] == '0'); }();
        if (stream_form && boost == nullptr && !(flags & MF_FLAG_REF_FP32_QUIRKS)) {
            rambo_inverse_stream_kernel<NF><<<(unsigned)blocks, threads, 0, stream>>>(momenta, event_stride, x1, x2, N, P, ps,
                                                                                      prob, logprob, status);
            return launch_status();
        }
    }
    rambo_inverse_kernel<T, NF><<<(unsigned)blocks, threads, 0, stream>>>(momenta, event_stride, x1, x2, N, P, ps,
                                                                          prob, logprob, status, boost, mask, logit_out);
    return launch_status();
}

template <typename T>
static int inverse_dispatch(const T* momenta, int64_t event_stride, const T* x1, const T* x2, int64_t N, int n,
                            const int32_t* order, const T* masses, const T* target_mass, T e_collider,
                            uint32_t flags, T* ps, T* prob, T* logprob, int32_t* status, cudaStream_t stream) {
    if (N < 0 || masses == nullptr) return MF_E_BADARG;
    if (n < MF_MIN_FINAL || n > MF_MAX_FINAL) return MF_E_NFINAL;
    if (N > 0 && (!momenta || !x1 || !x2 || !ps || !prob)) return MF_E_BADARG;
    if (event_stride < 4 * n || event_stride % 4 != 0) return MF_E_SHAPE;
    const int W = 3 * n - 2;
    const size_t row_align = (W % 4 == 0) ? 4 * sizeof(T) : ((W % 2 == 0) ? 2 * sizeof(T) : sizeof(T));
    if (!aligned(momenta, 4 * sizeof(T)) || !aligned(ps, row_align)) return MF_E_ALIGN;
    if (N > (int64_t)0x7fffffff * 256) return MF_E_BADARG;
    if ((flags & MF_FLAG_ZERO_STATUS) && status != nullptr) cudaMemsetAsync(status, 0, sizeof(int32_t), stream);
    switch (n) {
#define MF_CASE(NF) \
    case NF: return launch_inverse<T, NF>(momenta, event_stride, x1, x2, N, order, masses, target_mass, e_collider, flags, ps, prob, logprob, status, stream);
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return MF_E_NFINAL;
}

template <typename T>
static int getps_dispatch(const T* momenta, int64_t event_stride, const T* boost, int64_t N, int n, const int32_t* order,
                          const T* target_mass, T e_collider, T* ps, T* jac0, uint8_t* mask, T logit_eps, const T* lmean,
                          const T* lscale, T* logit_out, cudaStream_t stream) {
    if (N < 0 || target_mass == nullptr) return MF_E_BADARG;
    if (n < MF_MIN_FINAL || n > MF_MAX_FINAL) return MF_E_NFINAL;
    if (N > 0 && (!momenta || !boost || !ps)) return MF_E_BADARG;
    if (event_stride < 4 * n || event_stride % 4 != 0) return MF_E_SHAPE;
    const int W = 3 * n - 2;
    const size_t row_align = (W % 4 == 0) ? 4 * sizeof(T) : ((W % 2 == 0) ? 2 * sizeof(T) : sizeof(T));
    if (!aligned(momenta, 4 * sizeof(T)) || !aligned(boost, 4 * sizeof(T)) || !aligned(ps, row_align)) return MF_E_ALIGN;
    switch (n) {
#define MF_CASE(NF) \
    case NF: return launch_inverse<T, NF>(momenta, event_stride, nullptr, nullptr, N, order, target_mass, target_mass, e_collider, 0u, ps, jac0, nullptr, nullptr, stream, boost, mask, true, logit_eps, lmean, lscale, logit_out);
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return MF_E_NFINAL;
}

}  // namespace mf

extern "C" int mf_get_ps_f64(const double* momenta, int64_t event_stride, const double* boost, int64_t N, int n_final,
                             const int32_t* order, const double* target_mass, double e_collider, double* ps,
                             double* jac0, uint8_t* mask, double logit_eps, const double* logit_mean,
                             const double* logit_scale, double* logit_scaled, mf_stream_t stream) {
    if (logit_scaled && !mf::aligned(logit_scaled, 8)) return MF_E_ALIGN;
    return mf::getps_dispatch<double>(momenta, event_stride, boost, N, n_final, order, target_mass, e_collider, ps, jac0,
                                      mask, logit_eps, logit_mean, logit_scale, logit_scaled, (cudaStream_t)stream);
}

extern "C" int mf_rambo_inverse_f64(const double* momenta, int64_t event_stride, const double* x1, const double* x2,
                                    int64_t N, int n_final, const int32_t* order, const double* masses,
                                    const double* target_mass, double e_collider, uint32_t flags, double* ps,
                                    double* prob, double* logprob, int32_t* status, mf_stream_t stream) {
    return mf::inverse_dispatch<double>(momenta, event_stride, x1, x2, N, n_final, order, masses, target_mass,
                                        e_collider, flags, ps, prob, logprob, status, (cudaStream_t)stream);
}
extern "C" int mf_rambo_inverse_f32(const float* momenta, int64_t event_stride, const float* x1, const float* x2,
                                    int64_t N, int n_final, const int32_t* order, const float* masses,
                                    const float* target_mass, float e_collider, uint32_t flags, float* ps,
                                    float* prob, float* logprob, int32_t* status, mf_stream_t stream) {
    return mf::inverse_dispatch<float>(momenta, event_stride, x1, x2, N, n_final, order, masses, target_mass,
                                       e_collider, flags, ps, prob, logprob, status, (cudaStream_t)stream);
}
