// is_chain.cu -- fused importance-sampling chain, one launch per chunk of events:
//   Philox base sample z ~ U[-B,B]^F  ->  T spline inverses (K4)  ->  u = (x+B)/2B  ->  RAMBO forward (K1)
//   ->  w = jac / (2 x1 x2 s) / q(u)  ->  estimator accumulators (K5).
// Replaces the notebook sequence flow().sample / flow().log_prob / rambo.get_momenta_from_ps / weight /
// sums (notebooks/TestFlows_ImportanceSampling.ipynb cells [45,50-54]); squared matrix element and PDF
// are external to the reference's hot path and set to 1.
//
// Mapping: a CTA of 256 threads owns a tile of E events. Spline phase: one thread per (event, feature)
// row (rows of one transform are independent once the parameters are given), parameters staged by a
// double-buffered 1-D TMA bulk copy per (tile, transform). Phase-space phase: the first E threads, one
// event each, fp64, registers only. Nothing but the 3 accumulators is written to HBM unless the
// optional debug outputs are requested. Algorithmic bytes per event: T F (3K-1) 4.
#include <cstdlib>

#include "philox.cuh"
#include "rambo.cuh"
#include "reduce.cuh"
#include "rqs.cuh"

namespace mf {

constexpr int kChainThreads = 256;

template <int NF, int MINB>
__global__ void __launch_bounds__(kChainThreads, MINB)
is_chain_kernel(const float* __restrict__ phi, int64_t N, int T, int K, RqsConst<float> C, int E /*events per tile*/,
                int bulk_ok, uint64_t seed, uint64_t event_offset, const __grid_constant__ RamboParams<double, NF> P,
                double inv_2s, double* __restrict__ acc, void* workspace, double* __restrict__ u_out,
                double* __restrict__ logq_out, double* __restrict__ momenta, double* __restrict__ weight,
                double* __restrict__ x1o, double* __restrict__ x2o) {
    constexpr int F = 3 * NF - 2, NP = NF + 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int PW = 3 * K - 1;
    const int rows = E * F;
    const size_t tile_elems = (size_t)rows * PW;
    const size_t tile_bytes_al = (tile_elems * sizeof(float) + 127) / 128 * 128;
    auto stage_ptr = [&](int st) { return reinterpret_cast<float*>(smem_raw + (size_t)st * tile_bytes_al); };
    unsigned char* tail = smem_raw + 2 * tile_bytes_al;
    float* val = reinterpret_cast<float*>(tail);                // [kChainThreads] current x per row
    float* lsum = val + kChainThreads;                          // [kChainThreads] sum_t ladj per row
    uint64_t* bars = reinterpret_cast<uint64_t*>(lsum + kChainThreads);

    const int tid = threadIdx.x;
    const int64_t num_tiles = (N + E - 1) / E;
    const size_t transform_pitch = (size_t)N * F * PW;  // elements between phi[t] and phi[t+1]
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();

    // work items are (tile, transform) pairs in order; item j of this CTA -> stage j & 1
    auto tile_is_bulk = [&](int64_t tile) { return bulk_ok && (tile + 1) * (int64_t)E <= N; };
    auto issue = [&](int64_t tile, int t, int st) {
        if (tid == 0 && tile < num_tiles && tile_is_bulk(tile)) {
            const uint32_t bytes = (uint32_t)(tile_elems * sizeof(float));
            mbar_expect_tx(&bars[st], bytes);
            bulk_g2s(stage_ptr(st), phi + (size_t)t * transform_pitch + (size_t)tile * tile_elems, bytes, &bars[st]);
        }
    };

    Kahan k1, k2;
    double cnt = 0.0;
    uint32_t parity0 = 0u, parity1 = 0u;
    int item = 0;
    int64_t tile = blockIdx.x;
    issue(tile, T - 1, 0);  // sampling applies the transforms in reverse order: f_T^-1 first
    for (; tile < num_tiles; tile += gridDim.x) {
        const int64_t ev0 = tile * E;
        const int nev = (int)((N - ev0) < E ? (N - ev0) : E);
        const int nrows = nev * F;
        const int e = tid / F, f = tid - e * F;
        const bool row_active = tid < nrows;
        float x = 0.0f, ladj_acc = 0.0f;
        if (row_active) {  // base sample: Philox counter = (global event index, feature block), key = seed
            const uint64_t gev = event_offset + (uint64_t)(ev0 + e);
            Philox4 r4 = philox4x32_10((uint32_t)gev, (uint32_t)(gev >> 32), (uint32_t)(f >> 2), 0u, (uint32_t)seed,
                                       (uint32_t)(seed >> 32));
            const uint32_t bits = (f & 3) == 0 ? r4.x : ((f & 3) == 1 ? r4.y : ((f & 3) == 2 ? r4.z : r4.w));
            x = C.bound * (2.0f * u01_from_bits(bits) - 1.0f);
        }
        for (int tt = 0; tt < T; ++tt, ++item) {
            const int t = T - 1 - tt;
            const int st = item & 1;
            // prefetch the next (tile, transform) item into the other stage
            if (tt + 1 < T) issue(tile, t - 1, st ^ 1); else issue(tile + gridDim.x, T - 1, st ^ 1);
            float* sm = stage_ptr(st);
            if (tile_is_bulk(tile)) {
                mbar_wait(&bars[st], st ? parity1 : parity0);
                if (st) parity1 ^= 1u; else parity0 ^= 1u;
            } else {
                const float* src = phi + (size_t)t * transform_pitch + (size_t)tile * tile_elems;
                const size_t ne = (size_t)nrows * PW;
                for (size_t i = tid; i < ne; i += kChainThreads) sm[i] = __ldg(src + i);
                __syncthreads();
            }
            if (row_active) {
                float out, ladj;
                int bin;
                float* rowp = sm + (size_t)tid * PW;
                switch (K) {  // register-resident fast path for the common small bin counts
                case 8: rqs_row_reg<8, true>(rowp, C, x, out, ladj, bin); break;
                case 10: rqs_row_reg<10, true>(rowp, C, x, out, ladj, bin); break;
                case 12: rqs_row_reg<12, true>(rowp, C, x, out, ladj, bin); break;
                case 16: rqs_row_reg<16, true>(rowp, C, x, out, ladj, bin); break;
                default: rqs_row_1lane<float, true>(rowp, K, C, x, out, ladj, bin); break;
                }
                x = out;
                ladj_acc += ladj;
            }
            fence_proxy_async();
            __syncthreads();
        }
        val[tid] = x;
        lsum[tid] = ladj_acc;
        __syncthreads();
        if (tid < nev) {  // phase-space phase: one event per thread, fp64
            double r[F];
            double logq = 0.0;
#pragma unroll
            for (int j = 0; j < F; ++j) {
                r[j] = ((double)val[tid * F + j] + (double)C.bound) / (2.0 * (double)C.bound);
                logq += (double)lsum[tid * F + j];
            }
            double out[NP][4], jac, x1, x2;
            rambo_forward_event<double, NF>(r, P, out, jac, x1, x2);
            const double target = inv_2s / (x1 * x2);  // 1 / (2 x1 x2 s): ME = pdf = 1
            const double w = target * jac / exp(logq);
            if (target > 0.0 && jac == jac && isfinite(w)) { k1.add(w); k2.add(w * w); cnt += 1.0; }
            const int64_t ge = ev0 + tid;
            if (u_out) {
#pragma unroll
                for (int j = 0; j < F; ++j) u_out[ge * F + j] = r[j];
            }
            if (logq_out) logq_out[ge] = logq;
            if (momenta) {
#pragma unroll
                for (int k = 0; k < NP; ++k) st4(momenta + ge * (NP * 4) + 4 * k, out[k][0], out[k][1], out[k][2], out[k][3]);
            }
            if (weight) weight[ge] = jac;
            if (x1o) x1o[ge] = x1;
            if (x2o) x2o[ge] = x2;
        }
        __syncthreads();  // val/lsum reused by the next tile
    }
    grid_accumulate<kChainThreads>(k1.value(), k2.value(), cnt, acc, workspace);
}

// ---------------------------------------------------------------------------------------------
// Split variant: two launches per chunk.
//   A  flow_sample_kernel : Philox -> T spline inverses, one thread per (event, feature) row, ~40 registers ->
//      high occupancy; writes u[N,F] (fp64) and log q[N]   (+ (F+1)*8 B/event of HBM traffic)
//   B  rambo_is_kernel<NF>: K1 + weight + K5 fused, one event per thread, fp64 registers.
// The fused kernel above keeps everything on chip but carries the fp64 phase's register budget into the
// fp32 spline phase; which one is faster is measured and selected by MF_CHAIN_MODE.
// ---------------------------------------------------------------------------------------------
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
flow_sample_kernel(const float* __restrict__ phi, int64_t N, int F, int T, int K, RqsConst<float> C, int E, int stages,
                   int bulk_ok, uint64_t seed, uint64_t event_offset, double* __restrict__ u_out,
                   double* __restrict__ logq_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int PW = 3 * K - 1;
    const int rows = E * F;
    const size_t tile_elems = (size_t)rows * PW;
    const size_t tile_bytes_al = (tile_elems * sizeof(float) + 127) / 128 * 128;
    auto stage_ptr = [&](int st) { return reinterpret_cast<float*>(smem_raw + (size_t)st * tile_bytes_al); };
    unsigned char* tail = smem_raw + (size_t)stages * tile_bytes_al;
    float* lsum = reinterpret_cast<float*>(tail);  // [THREADS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(lsum + THREADS);

    const int tid = threadIdx.x;
    const int64_t num_tiles = (N + E - 1) / E;
    const size_t transform_pitch = (size_t)N * F * PW;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();
    auto tile_is_bulk = [&](int64_t tile) { return bulk_ok && (tile + 1) * (int64_t)E <= N; };
    auto issue = [&](int64_t tile, int t, int st) {
        if (tid == 0 && tile < num_tiles && tile_is_bulk(tile)) {
            const uint32_t bytes = (uint32_t)(tile_elems * sizeof(float));
            mbar_expect_tx(&bars[st], bytes);
            bulk_g2s(stage_ptr(st), phi + (size_t)t * transform_pitch + (size_t)tile * tile_elems, bytes, &bars[st]);
        }
    };
    uint32_t parity0 = 0u, parity1 = 0u;
    int item = 0;
    int64_t tile = blockIdx.x;
    const double inv_2b = 1.0 / (2.0 * (double)C.bound);  // exact for the usual bound = 1
    issue(tile, T - 1, 0);
    for (; tile < num_tiles; tile += gridDim.x) {
        const int64_t ev0 = tile * E;
        const int nev = (int)((N - ev0) < E ? (N - ev0) : E);
        const int nrows = nev * F;
        const int e = tid / F, f = tid - e * F;
        const bool row_active = tid < nrows;
        float x = 0.0f, ladj_acc = 0.0f;
        if (row_active) {
            const uint64_t gev = event_offset + (uint64_t)(ev0 + e);
            Philox4 r4 = philox4x32_10((uint32_t)gev, (uint32_t)(gev >> 32), (uint32_t)(f >> 2), 0u, (uint32_t)seed,
                                       (uint32_t)(seed >> 32));
            const uint32_t bits = (f & 3) == 0 ? r4.x : ((f & 3) == 1 ? r4.y : ((f & 3) == 2 ? r4.z : r4.w));
            x = C.bound * (2.0f * u01_from_bits(bits) - 1.0f);
        }
        for (int tt = 0; tt < T; ++tt, ++item) {
            const int t = T - 1 - tt;
            const int st = (stages == 2) ? (item & 1) : 0;
            if (stages == 2) { if (tt + 1 < T) issue(tile, t - 1, st ^ 1); else issue(tile + gridDim.x, T - 1, st ^ 1); }
            float* sm = stage_ptr(st);
            if (tile_is_bulk(tile)) {
                mbar_wait(&bars[st], st ? parity1 : parity0);
                if (st) parity1 ^= 1u; else parity0 ^= 1u;
            } else {
                const float* src = phi + (size_t)t * transform_pitch + (size_t)tile * tile_elems;
                const size_t ne = (size_t)nrows * PW;
                for (size_t i = tid; i < ne; i += THREADS) sm[i] = __ldg(src + i);
                __syncthreads();
            }
            if (row_active) {
                float out, ladj;
                int bin;
                float* rowp = sm + (size_t)tid * PW;
                switch (K) {
                case 8: rqs_row_reg<8, true>(rowp, C, x, out, ladj, bin); break;
                case 10: rqs_row_reg<10, true>(rowp, C, x, out, ladj, bin); break;
                case 12: rqs_row_reg<12, true>(rowp, C, x, out, ladj, bin); break;
                case 16: rqs_row_reg<16, true>(rowp, C, x, out, ladj, bin); break;
                default: rqs_row_1lane<float, true>(rowp, K, C, x, out, ladj, bin); break;
                }
                x = out;
                ladj_acc += ladj;
            }
            fence_proxy_async();
            __syncthreads();
            if (stages == 1) { if (tt + 1 < T) issue(tile, t - 1, 0); else issue(tile + gridDim.x, T - 1, 0); }
        }
        lsum[tid] = ladj_acc;
        if (row_active) u_out[ev0 * F + tid] = ((double)x + (double)C.bound) * inv_2b;
        __syncthreads();
        if (tid < nev) {
            double lq = 0.0;
            for (int j = 0; j < F; ++j) lq += (double)lsum[tid * F + j];
            logq_out[ev0 + tid] = lq;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Flow-sample kernel for long rows (compile-time K = 32..64, e.g. the production flow F=10, K=40, T=6 of BASELINE
// config 5): same two-sided register scheme as rqs_reg2_kernel (rqs.cu) -- thread t < 64 owns the searched (height)
// array of row t, thread 64 + t the width array, both exponentiated once into registers -- chained over the T
// transforms of a tile: the new x goes back to the searching side through shared memory.
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128, MF_REG2_MINB(K))
flow_sample_reg2_kernel(const float* __restrict__ phi, int64_t N, int F, int T, RqsConst<float> C, int E, int bulk_ok,
                        uint64_t seed, uint64_t event_offset, double* __restrict__ u_out, double* __restrict__ logq_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PW = 3 * K - 1;
    const int rows = E * F;  // <= 64
    const size_t tile_elems = (size_t)rows * PW;
    const size_t tile_bytes_al = (tile_elems * sizeof(float) + 127) / 128 * 128;
    float* sm = reinterpret_cast<float*>(smem_raw);
    float4* xch = reinterpret_cast<float4*>(smem_raw + tile_bytes_al);  // [64] {bin, k0, k1, -}
    float* xnext = reinterpret_cast<float*>(xch + 64);                   // [64] x after the current transform
    float* lsum = xnext + 64;                                            // [64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(lsum + 64);
    const int tid = threadIdx.x, side = tid >> 6, row = tid & 63;
    const int64_t num_tiles = (N + E - 1) / E;
    const size_t transform_pitch = (size_t)N * F * PW;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    auto tile_is_bulk = [&](int64_t tile) { return bulk_ok && (tile + 1) * (int64_t)E <= N; };
    auto issue = [&](int64_t tile, int t) {
        if (tid == 0 && tile < num_tiles && tile_is_bulk(tile)) {
            const uint32_t bytes = (uint32_t)(tile_elems * sizeof(float));
            mbar_expect_tx(bar, bytes);
            bulk_g2s(sm, phi + (size_t)t * transform_pitch + (size_t)tile * tile_elems, bytes, bar);
        }
    };
    uint32_t parity = 0u;
    const double inv_2b = 1.0 / (2.0 * (double)C.bound);
    int64_t tile = blockIdx.x;
    issue(tile, T - 1);
    for (; tile < num_tiles; tile += gridDim.x) {
        const int64_t ev0 = tile * E;
        const int nev = (int)((N - ev0) < E ? (N - ev0) : E);
        const int nrows = nev * F;
        const bool active = row < nrows;
        const int e = row / F, f = row - e * F;
        float x = 0.0f, ladj_acc = 0.0f;
        if (active) {  // both sides draw the same base sample (counter = global event index)
            const uint64_t gev = event_offset + (uint64_t)(ev0 + e);
            Philox4 r4 = philox4x32_10((uint32_t)gev, (uint32_t)(gev >> 32), (uint32_t)(f >> 2), 0u, (uint32_t)seed,
                                       (uint32_t)(seed >> 32));
            const uint32_t bits = (f & 3) == 0 ? r4.x : ((f & 3) == 1 ? r4.y : ((f & 3) == 2 ? r4.z : r4.w));
            x = C.bound * (2.0f * u01_from_bits(bits) - 1.0f);
        }
        for (int tt = 0; tt < T; ++tt) {
            const int t = T - 1 - tt;
            if (tile_is_bulk(tile)) {
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {
                const float* src = phi + (size_t)t * transform_pitch + (size_t)tile * tile_elems;
                const size_t ne = (size_t)nrows * PW;
                for (size_t i = tid; i < ne; i += 128) sm[i] = __ldg(src + i);
                __syncthreads();
            }
            float ev_[K];
            float S = 1.0f;
            const float* r = sm + (size_t)row * PW;
            if (active) {
                S = rqs_exp_row<K>(r + (side == 0 ? K : 0), C, ev_);  // inverse: heights are searched (side 0)
                if (side == 0) {
                    float k0, k1;
                    const int bin = rqs_search_row<K>(ev_, S, C.bound, x, k0, k1);
                    xch[row] = make_float4(__int_as_float(bin), k0, k1, 0.0f);
                }
            }
            __syncthreads();
            if (side == 1 && active) {
                const float4 q = xch[row];
                const int bin = __float_as_int(q.x);
                float out = x, ladj = 0.0f;
                if (bin >= 0 && bin < K) {
                    float o0, o1;
                    rqs_prefix_row<K>(ev_, S, C.bound, bin, o0, o1);
                    rqs_eval_fast<true>(r + 2 * K, K, C, bin, o0, o1, q.y, q.z, x, out, ladj);
                }
                x = out;
                ladj_acc += ladj;
                xnext[row] = x;
            }
            fence_proxy_async();
            __syncthreads();
            if (tt + 1 < T) issue(tile, t - 1); else issue(tile + gridDim.x, T - 1);
            if (side == 0 && active) x = xnext[row];
        }
        if (side == 1) {
            lsum[row] = ladj_acc;
            if (active) u_out[ev0 * F + row] = ((double)x + (double)C.bound) * inv_2b;
        }
        __syncthreads();
        if (tid < nev) {
            double lq = 0.0;
            for (int j = 0; j < F; ++j) lq += (double)lsum[tid * F + j];
            logq_out[ev0 + tid] = lq;
        }
        // lsum / xnext / xch are rewritten only after the next transform's first barrier
    }
}

// The event again on the guarded primitives (PrimIEEE, common.cuh), out of line; taken when the speculative pass of
// rambo_is_kernel produced a non-finite weight or momentum. Reloads the event's uniforms from global memory.
template <int NF>
__device__ __noinline__ void rambo_is_redo(const RamboParams<double, NF>& P, const double* urow, double* mo, double* wx) {
    double jac, x1, x2;
    rambo_forward_stream<double, NF, PrimIEEE>(
        P, [&](int i) { return urow[i]; },
        [&](int j, const double (&p)[4]) { if (mo) st4(mo + 4 * j, p[0], p[1], p[2], p[3]); }, jac, x1, x2);
    wx[0] = jac; wx[1] = x1; wx[2] = x2;
}

#ifndef MF_CHAINB_MINB
#define MF_CHAINB_MINB 3
#endif
template <int NF, int THREADS>
__global__ void __launch_bounds__(THREADS, MF_CHAINB_MINB * 256 / THREADS)
rambo_is_kernel(const double* __restrict__ u, const double* __restrict__ logq, int64_t N,
                const __grid_constant__ RamboParams<double, NF> P, double inv_2s, double* __restrict__ acc,
                void* workspace, double* __restrict__ momenta, double* __restrict__ weight, double* __restrict__ x1o,
                double* __restrict__ x2o) {
    constexpr int F = 3 * NF - 2, NP = NF + 2;
    Kahan k1, k2;
    double cnt = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ev < N; ev += stride) {
        // all of the event's inputs are requested up front: the map's dependency chains would otherwise wait for each
        // global load where it is first used
        double uc[F];
#pragma unroll
        for (int i = 0; i < F; ++i) uc[i] = ld1(u + ev * F + i);
        const double lq = __ldg(logq + ev);
        double* mo = momenta ? momenta + ev * (NP * 4) : nullptr;
        double jac, x1, x2;
        // streaming core (rambo.cuh): O(1) live state; momenta leave the registers only if they were asked for
        // speculative pass on the unguarded primitives (see rambo_forward_kernel)
        bool redo = false;
        rambo_forward_stream<double, NF, PrimRaw>(
            P, [&](int i) { return uc[i]; },
            [&](int j, const double (&p)[4]) { if (mo) { redo |= not_finite_bits(p[0]); st4(mo + 4 * j, p[0], p[1], p[2], p[3]); } },
            jac, x1, x2, redo);
        if (redo | not_finite_bits(jac)) {
            double wx[3];
            rambo_is_redo<NF>(P, u + ev * F, mo, wx);
            jac = wx[0]; x1 = wx[1]; x2 = wx[2];
        }
        const double target = inv_2s * rcp_fast(x1 * x2);  // 1 / (2 x1 x2 s): ME = pdf = 1
        const double w = target * jac * exp(-lq);
        if (target > 0.0 && jac == jac && isfinite(w)) { k1.add(w); k2.add(w * w); cnt += 1.0; }
        if (weight) weight[ev] = jac;
        if (x1o) x1o[ev] = x1;
        if (x2o) x2o[ev] = x2;
    }
    grid_accumulate<THREADS>(k1.value(), k2.value(), cnt, acc, workspace);
}

struct ChainGeom { int E, bulk_ok, sms, smem_max; };
static ChainGeom chain_geometry(const float* phi, int64_t N, int F, int T, int K, int stages, int threads = kChainThreads) {
    ChainGeom g{};
    const int PW = 3 * K - 1;
    int dev = 0;
    g.sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g.smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int E = threads / F;
    if (E >= 32) E = E / 32 * 32;
    const size_t extra = 2 * threads * sizeof(float) + 2 * sizeof(uint64_t) + 256;
    auto need = [&](int e) { return (((size_t)e * F * PW * sizeof(float) + 127) / 128 * 128) * stages + extra; };
    while (E > 1 && need(E) > (size_t)g.smem_max / 2) --E;
    while (E > 1 && need(E) > (size_t)g.smem_max) --E;
    if (need(E) > (size_t)g.smem_max) { g.E = 0; return g; }
    g.bulk_ok = aligned(phi, 16) && (((size_t)N * F * PW * sizeof(float)) % 16 == 0 || T == 1) ? 1 : 0;
    if (g.bulk_ok) {
        int e = E;
        while (e > 0 && ((size_t)e * F * PW * sizeof(float)) % 16 != 0) --e;
        if (e > 0) E = e; else g.bulk_ok = 0;
    }
    g.E = E;
    return g;
}

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// ---- launch A, compile-time K (8, 10, 12, 16), single stage: what bench.py times ----
// Same algorithm as flow_sample_kernel; every thread does ONE row per (tile, transform), so whatever is executed per
// tile is paid per row. The source-level profile of the generic kernel showed 45 % of its instructions outside the row
// math: 64-bit tile / transform address arithmetic redone at every use, the runtime stage / K dispatch, and one full
// Philox4x32-10 (102 instructions) per ROW although it yields four variates. Here the tile base pointer and the
// bulk / ragged decision are taken once per tile, only thread 0 forms copy addresses, and one thread in four runs
// Philox and hands the other variates over through shared memory.
template <int K>
__global__ void __launch_bounds__(256, 8)
flow_sample_fast_kernel(const float* __restrict__ phi, int64_t N, int F, int T, RqsConst<float> C, int E, int bulk_ok,
                        uint64_t seed, uint64_t event_offset, double* __restrict__ u_out, double* __restrict__ logq_out) {
    constexpr int PW = 3 * K - 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int rows = E * F;
    const uint32_t tile_elems = (uint32_t)rows * PW;
    const uint32_t tile_bytes = tile_elems * (uint32_t)sizeof(float);
    float* const sm = reinterpret_cast<float*>(smem_raw);
    float* const lsum = reinterpret_cast<float*>(smem_raw + ((size_t)tile_bytes + 127) / 128 * 128);  // [256]
    uint32_t* const rnd = reinterpret_cast<uint32_t*>(lsum);  // the same words, before the transforms
    uint64_t* const bar = reinterpret_cast<uint64_t*>(lsum + 256);
    const int tid = threadIdx.x;
    const int e = tid / F, f = tid - e * F;
    float* const rowp = sm + (size_t)tid * PW;
    const int64_t num_tiles = (N + E - 1) / E, full_tiles = bulk_ok ? N / E : 0;  // tiles below full_tiles come by TMA
    const size_t pitch = (size_t)N * F * PW;                                       // one transform's parameters
    const float* const last = phi + (size_t)(T - 1) * pitch;                       // transforms run T-1 ... 0
    const double inv_2b = 1.0 / (2.0 * (double)C.bound);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0u;
    int64_t tile = blockIdx.x;
    if (tid == 0 && tile < full_tiles) {
        mbar_expect_tx(bar, tile_bytes);
        bulk_g2s(sm, last + (size_t)tile * tile_elems, tile_bytes, bar);
    }
    for (; tile < num_tiles; tile += gridDim.x) {
        const bool bulk = tile < full_tiles;
        const int64_t ev0 = tile * E;
        const int nev = (int)((N - ev0) < E ? (N - ev0) : E);
        const int nrows = nev * F;
        const bool row_active = tid < nrows;
        if (row_active && (f & 3) == 0) {   // one Philox call serves features f .. f+3 of the event
            const uint64_t gev = event_offset + (uint64_t)(ev0 + e);
            const Philox4 r4 = philox4x32_10((uint32_t)gev, (uint32_t)(gev >> 32), (uint32_t)(f >> 2), 0u, (uint32_t)seed,
                                             (uint32_t)(seed >> 32));
            rnd[tid] = r4.x;
            if (f + 1 < F) rnd[tid + 1] = r4.y;
            if (f + 2 < F) rnd[tid + 2] = r4.z;
            if (f + 3 < F) rnd[tid + 3] = r4.w;
        }
        __syncthreads();
        float x = row_active ? C.bound * (2.0f * u01_from_bits(rnd[tid]) - 1.0f) : 0.0f;
        float ladj_acc = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            if (bulk) {
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {  // ragged last tile or unaligned phi: plain cooperative copy
                const float* src = phi + (size_t)t * pitch + (size_t)tile * tile_elems;
                const uint32_t ne = (uint32_t)nrows * PW;
                for (uint32_t i = tid; i < ne; i += 256) sm[i] = __ldg(src + i);
                __syncthreads();
            }
            if (row_active) {
                float out, ladj;
                int bin;
                rqs_row_reg<K, true>(rowp, C, x, out, ladj, bin);
                x = out;
                ladj_acc += ladj;
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {   // refill the (single) stage: next transform of this tile, else the next tile's first
                const int64_t nt = t > 0 ? tile : tile + gridDim.x;
                if (nt < full_tiles) {
                    mbar_expect_tx(bar, tile_bytes);
                    bulk_g2s(sm, phi + (size_t)(t > 0 ? t - 1 : T - 1) * pitch + (size_t)nt * tile_elems, tile_bytes, bar);
                }
            }
        }
        lsum[tid] = ladj_acc;
        if (row_active) u_out[ev0 * F + tid] = ((double)x + (double)C.bound) * inv_2b;
        __syncthreads();
        if (tid < nev) {
            double lq = 0.0;
            for (int j = 0; j < F; ++j) lq += (double)lsum[tid * F + j];
            logq_out[ev0 + tid] = lq;
        }
        __syncthreads();
    }
}


// A: Philox base sample + T spline inverses -> u [N,F] fp64, log q [N]
static int flow_sample_launch(const float* phi, int64_t N, int F, int T, int K, float bound, float slope, uint64_t seed,
                              uint64_t event_offset, double* u, double* logq, cudaStream_t stream) {
    static const int threads = env_int("MF_CHAIN_THREADS", 256), stages_env = env_int("MF_CHAIN_STAGES", 1);
    const int stages = stages_env == 2 ? 2 : 1;
    if (F > 64) return MF_E_SHAPE;
    if (K == 32 || K == 40 || K == 48 || K == 64) {  // long rows: two-sided register kernel
        const int PWl = 3 * K - 1;
        int dev = 0, sms = 148, smem_max = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const size_t extra = 64 * sizeof(float4) + 128 * sizeof(float) + sizeof(uint64_t) + 128;
        auto need = [&](int e) { return ((size_t)e * F * PWl * sizeof(float) + 127) / 128 * 128 + extra; };
        int E = 64 / F;
        while (E > 1 && need(E) > (size_t)smem_max / 4) --E;
        int bulk_ok = aligned(phi, 16) && (((size_t)N * F * PWl * sizeof(float)) % 16 == 0 || T == 1) ? 1 : 0;
        if (bulk_ok) {
            int e = E;
            while (e > 0 && ((size_t)e * F * PWl * sizeof(float)) % 16 != 0) --e;
            if (e > 0) E = e; else bulk_ok = 0;
        }
        const size_t smem_l = need(E);
        const RqsConst<float> Cl = make_rqs_const<float>(bound, slope);
        const int64_t tiles = (N + E - 1) / E;
        auto gol = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            int ctas = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 128, smem_l);
            if (ctas < 1) ctas = 1;
            int64_t grid = (int64_t)sms * ctas;
            if (grid > tiles) grid = tiles;
            kern<<<(unsigned)grid, 128, smem_l, stream>>>(phi, N, F, T, Cl, E, bulk_ok, seed, event_offset, u, logq);
        };
        switch (K) {
        case 32: gol(flow_sample_reg2_kernel<32>); break;
        case 40: gol(flow_sample_reg2_kernel<40>); break;
        case 48: gol(flow_sample_reg2_kernel<48>); break;
        default: gol(flow_sample_reg2_kernel<64>); break;
        }
        return launch_status();
    }
    const int thr = (threads == 256 || F > 32) ? 256 : (threads == 64 && F <= 16 ? 64 : 128);
    const ChainGeom g = chain_geometry(phi, N, F, T, K, stages, thr);
    if (g.E == 0) return MF_E_SHAPE;
    const int PW = 3 * K - 1;
    const RqsConst<float> C = make_rqs_const<float>(bound, slope);
    const int64_t num_tiles = (N + g.E - 1) / g.E;
    const size_t smem = (((size_t)g.E * F * PW * sizeof(float) + 127) / 128 * 128) * stages + thr * sizeof(float) +
                        2 * sizeof(uint64_t) + 128;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int ctas = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, thr, smem);
        if (ctas < 1) ctas = 1;
        int64_t grid = (int64_t)g.sms * ctas;
        if (grid > num_tiles) grid = num_tiles;
        kern<<<(unsigned)grid, thr, smem, stream>>>(phi, N, F, T, K, C, g.E, stages, g.bulk_ok, seed, event_offset, u, logq);
    };
    // 32 registers per thread -> 2048 resident threads per SM for every CTA size
    static const int minb = env_int("MF_CHAIN_MINB", 8);
    static const int fast = env_int("MF_CHAIN_FAST", 1);
    if (fast && thr == 256 && minb == 8 && stages == 1 && (K == 8 || K == 10 || K == 12 || K == 16)) {
        auto gof = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            int ctas = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 256, smem);
            if (ctas < 1) ctas = 1;
            static const int ctas_env = env_int("MF_CHAINA_CTAS", 0);   // fewer resident CTAs: room for launch B of the previous chunk
            if (ctas_env > 0 && ctas_env < ctas) ctas = ctas_env;
            int64_t grid = (int64_t)g.sms * ctas;
            if (grid > num_tiles) grid = num_tiles;
            kern<<<(unsigned)grid, 256, smem, stream>>>(phi, N, F, T, C, g.E, g.bulk_ok, seed, event_offset, u, logq);
        };
        if (K == 8) gof(flow_sample_fast_kernel<8>);
        else if (K == 10) gof(flow_sample_fast_kernel<10>);
        else if (K == 12) gof(flow_sample_fast_kernel<12>);
        else gof(flow_sample_fast_kernel<16>);
        return launch_status();
    }
    if (thr == 256 && minb == 6) go(flow_sample_kernel<256, 6>);
    else if (thr == 256 && minb == 5) go(flow_sample_kernel<256, 5>);
    else if (thr == 256) go(flow_sample_kernel<256, 8>);
    else if (thr == 64) go(flow_sample_kernel<64, 32>);
    else go(flow_sample_kernel<128, 16>);
    return launch_status();
}

// B: K1 + weight + K5 on given u, log q
template <int NF>
static int rambo_is_launch(const double* u, const double* logq, int64_t N, const double* masses, double e_collider,
                           uint32_t flags, double* acc, void* workspace, double* momenta, double* weight, double* x1,
                           double* x2, cudaStream_t stream) {
    const RamboParams<double, NF> P = make_forward_params<double, NF>(masses, e_collider, nullptr, flags & ~MF_FLAG_CUTS);
    const double inv_2s = 1.0 / (2.0 * e_collider * e_collider);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // MF_CHAINB_THREADS=128 with MF_CHAINB_CTAS=<CTAs per SM>: a small-footprint launch that can share the SMs with launch A of
    // the next chunk (scripts/overlap_ab.py); default: 256 threads, as many CTAs as fit
    static const int thr = env_int("MF_CHAINB_THREADS", 256);
    static const int ctas_env = env_int("MF_CHAINB_CTAS", 0);
    auto go = [&](auto kern, int threads) {
        int ctas = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, threads, 0);
        if (ctas < 1) ctas = 1;
        if (ctas_env > 0 && ctas_env < ctas) ctas = ctas_env;
        int64_t grid = (int64_t)sms * ctas;
        const int64_t want = (N + threads - 1) / threads;
        if (grid > want) grid = want;
        if (grid > kReduceMaxBlocks) grid = kReduceMaxBlocks;
        kern<<<(unsigned)grid, threads, 0, stream>>>(u, logq, N, P, inv_2s, acc, workspace, momenta, weight, x1, x2);
    };
    if (thr == 96) go(rambo_is_kernel<NF, 96>, 96);
    else if (thr == 128) go(rambo_is_kernel<NF, 128>, 128);
    else go(rambo_is_kernel<NF, 256>, 256);
    return launch_status();
}

template <int NF>
static int chain_launch(const float* phi, int64_t N, int T, int K, float bound, float slope, uint64_t seed,
                        uint64_t event_offset, const double* masses, double e_collider, uint32_t flags, double* acc,
                        void* workspace, double* u_out, double* logq, double* momenta, double* weight, double* x1,
                        double* x2, cudaStream_t stream) {
    constexpr int F = 3 * NF - 2;
    static const int fused = []() { const char* e = getenv("MF_CHAIN_MODE"); return (e && e[0] == 'f') ? 1 : 0; }();
    if (!fused) {
        // default: two launches on the same stream, A (flow sample) then B (phase space + weight + reduce)
        double *u_buf = u_out, *lq_buf = logq;
        void* scratch = nullptr;
        if (!u_buf || !lq_buf) {
            const size_t bytes = (size_t)N * (F + 1) * sizeof(double);
            if (cudaMallocAsync(&scratch, bytes, stream) != cudaSuccess) return launch_status();
            if (!u_buf) u_buf = reinterpret_cast<double*>(scratch);
            if (!lq_buf) lq_buf = reinterpret_cast<double*>(scratch) + (size_t)N * F;
        }
        int rc = flow_sample_launch(phi, N, F, T, K, bound, slope, seed, event_offset, u_buf, lq_buf, stream);
        if (rc == MF_OK)
            rc = rambo_is_launch<NF>(u_buf, lq_buf, N, masses, e_collider, flags, acc, workspace, momenta, weight, x1, x2, stream);
        if (scratch) cudaFreeAsync(scratch, stream);
        return rc;
    }
    // MF_CHAIN_MODE=fused: single launch, nothing but the accumulators leaves the SM (slower: DESIGN.md section 4)
    static const int minb = env_int("MF_CHAIN_FUSED_MINB", 4);
    const ChainGeom g = chain_geometry(phi, N, F, T, K, 2);
    if (g.E == 0) return MF_E_SHAPE;
    const int PW = 3 * K - 1;
    const size_t smem = (((size_t)g.E * F * PW * sizeof(float) + 127) / 128 * 128) * 2 + 2 * kChainThreads * sizeof(float) +
                        2 * sizeof(uint64_t) + 256;
    const RqsConst<float> C = make_rqs_const<float>(bound, slope);
    const RamboParams<double, NF> P = make_forward_params<double, NF>(masses, e_collider, nullptr, flags & ~MF_FLAG_CUTS);
    const double inv_2s = 1.0 / (2.0 * e_collider * e_collider);
    const int64_t num_tiles = (N + g.E - 1) / g.E;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int ctas_per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kChainThreads, smem);
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        int64_t grid = (int64_t)g.sms * ctas_per_sm;
        if (grid > num_tiles) grid = num_tiles;
        if (grid > kReduceMaxBlocks) grid = kReduceMaxBlocks;
        kern<<<(unsigned)grid, kChainThreads, smem, stream>>>(phi, N, T, K, C, g.E, g.bulk_ok, seed, event_offset, P, inv_2s,
                                                              acc, workspace, u_out, logq, momenta, weight, x1, x2);
    };
    switch (minb) {
    case 1: go(is_chain_kernel<NF, 1>); break;
    case 2: go(is_chain_kernel<NF, 2>); break;
    case 3: go(is_chain_kernel<NF, 3>); break;
    default: go(is_chain_kernel<NF, 4>); break;
    }
    return launch_status();
}

}  // namespace mf

extern "C" int mf_is_chain_f64(const float* phi, int64_t N, int n_final, int T, int K, float bound, float slope,
                               uint64_t seed, uint64_t event_offset, const double* masses, double e_collider,
                               uint32_t flags, double* acc, void* workspace, double* u_out, double* logq,
                               double* momenta, double* weight, double* x1, double* x2, mf_stream_t stream) {
    if (N < 0 || T < 1 || !masses || !acc || !workspace || (N > 0 && !phi)) return MF_E_BADARG;
    if (K < 1 || K > 128) return MF_E_SHAPE;
    if (n_final < MF_MIN_FINAL || n_final > MF_MAX_FINAL) return MF_E_NFINAL;
    if (momenta && !mf::aligned(momenta, 32)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    switch (n_final) {
#define MF_CASE(NF) \
    case NF: return mf::chain_launch<NF>(phi, N, T, K, bound, slope, seed, event_offset, masses, e_collider, flags, acc, workspace, u_out, logq, momenta, weight, x1, x2, (cudaStream_t)stream);
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return MF_E_NFINAL;
}

extern "C" int mf_flow_sample_f32(const float* phi, int64_t N, int F, int T, int K, float bound, float slope, uint64_t seed,
                                  uint64_t event_offset, double* u, double* logq, mf_stream_t stream) {
    if (N < 0 || T < 1 || F < 1 || F > mf::kChainThreads || (N > 0 && (!phi || !u || !logq))) return MF_E_BADARG;
    if (K < 1 || K > 128) return MF_E_SHAPE;
    if (N == 0) return MF_OK;
    return mf::flow_sample_launch(phi, N, F, T, K, bound, slope, seed, event_offset, u, logq, (cudaStream_t)stream);
}

// ---- launch A with float64 flow parameters (the dtype of every MEMFlow config): the same chain built from the fp64
// spline kernel. Philox base point (the fp32 chain's draw, widened) -> T x rqs_coop_kernel<double, inverse> in place on
// the u buffer, the per-row log-dets summed per event in feature order and accumulated over the transforms
// (deterministic) -> u = (x + B) / 2B. FP64-latency bound at ~2.6 TB/s of parameters (profiles/: rqs64f). ----
namespace mf {
__global__ void __launch_bounds__(256)
philox_base_f64_kernel(int64_t N, int F, double bound, uint64_t seed, uint64_t event_offset, double* __restrict__ x,
                       double* __restrict__ logq) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    const uint64_t gev = event_offset + (uint64_t)ev;
    for (int blk = 0; 4 * blk < F; ++blk) {
        const Philox4 r4 = philox4x32_10((uint32_t)gev, (uint32_t)(gev >> 32), (uint32_t)blk, 0u, (uint32_t)seed,
                                         (uint32_t)(seed >> 32));
        const uint32_t b[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * blk + j < F)  // the fp32 chain's base point bound * (2 u01 - 1), rounded in fp32 like there
                x[ev * F + 4 * blk + j] = (double)((float)bound * (2.0f * u01_from_bits(b[j]) - 1.0f));
    }
    logq[ev] = 0.0;
}
__global__ void __launch_bounds__(256)
ladj_accumulate_kernel(const double* __restrict__ ladj, int64_t N, int F, double* __restrict__ logq) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += ladj[ev * F + f];
    logq[ev] += s;
}
__global__ void __launch_bounds__(256) to_unit_kernel(double* __restrict__ x, int64_t n, double bound, double inv_2b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (x[i] + bound) * inv_2b;
}
}  // namespace mf

extern "C" int mf_flow_sample_f64(const double* phi, int64_t N, int F, int T, int K, double bound, double slope,
                                  uint64_t seed, uint64_t event_offset, double* u, double* logq, mf_stream_t stream) {
    if (N < 0 || T < 1 || F < 1 || (N > 0 && (!phi || !u || !logq))) return MF_E_BADARG;
    if (K < 1 || K > 128) return MF_E_SHAPE;
    if (N == 0) return MF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {   // one launch: Philox base point -> T inverses -> u, log q (rqs_f64.cu)
        const int rc1 = mf::rqs64_launch(true, true, phi, nullptr, N, F, K, T, bound, slope, seed, event_offset, u, nullptr, logq,
                                         nullptr, st);
        if (rc1 != MF_E_SHAPE) return rc1;
    }
    // shapes outside that kernel (no soft clip, very wide events): one launch per transform on the general kernels,
    // ping-ponging between u and a scratch buffer (the kernels' input and output pointers are __restrict__)
    const unsigned grid = (unsigned)((N + 255) / 256);
    const int64_t n = N * (int64_t)F;
    double* scratch = nullptr;
    if (cudaMallocAsync((void**)&scratch, (size_t)2 * n * sizeof(double), st) != cudaSuccess) return -(int)cudaErrorMemoryAllocation;
    double* ladj = scratch + n;
    double* cur = (T % 2 == 0) ? u : scratch;   // T swaps later the result sits in u
    double* nxt = (T % 2 == 0) ? scratch : u;
    mf::philox_base_f64_kernel<<<grid, 256, 0, st>>>(N, F, bound, seed, event_offset, cur, logq);
    const size_t pitch = (size_t)N * F * (3 * K - 1);
    int rc = MF_OK;
    for (int t = T - 1; t >= 0 && rc == MF_OK; --t) {
        rc = mf::rqs_coop_launch<double, true>(phi + (size_t)t * pitch, cur, N, F, K, bound, slope, nxt, ladj, nullptr, nullptr, st,
                                                (int64_t)event_offset * F);
        if (rc == MF_OK) mf::ladj_accumulate_kernel<<<grid, 256, 0, st>>>(ladj, N, F, logq);
        double* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (rc == MF_OK) {   // cur == u
        mf::to_unit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(u, n, bound, 1.0 / (2.0 * bound));
        rc = mf::launch_status();
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

extern "C" int mf_rambo_is_f64(const double* u, const double* logq, int64_t N, int n_final, const double* masses,
                               double e_collider, uint32_t flags, double* acc, void* workspace, double* momenta,
                               double* weight, double* x1, double* x2, mf_stream_t stream) {
    if (N < 0 || !masses || !acc || !workspace || (N > 0 && (!u || !logq))) return MF_E_BADARG;
    if (n_final < MF_MIN_FINAL || n_final > MF_MAX_FINAL) return MF_E_NFINAL;
    const int W = 3 * n_final - 2;
    const size_t row_align = (W % 4 == 0) ? 32 : ((W % 2 == 0) ? 16 : 8);
    if (!mf::aligned(u, row_align) || (momenta && !mf::aligned(momenta, 32))) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    switch (n_final) {
#define MF_CASE(NF) \
    case NF: return mf::rambo_is_launch<NF>(u, logq, N, masses, e_collider, flags, acc, workspace, momenta, weight, x1, x2, (cudaStream_t)stream);
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return MF_E_NFINAL;
}
