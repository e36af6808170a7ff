// rqs.cu -- K3/K4: monotonic rational-quadratic spline forward / inverse / log-det
// (zuko.transforms.MonotonicRQSTransform; call sites memflow/unfolding_flow/unfolding_flow.py:88-97,
// custom_spline_flow_ps.py:16-33). The reference runs softmax x2, cumsum x2, pad, exp, searchsorted,
// 6 gathers and ~40 elementwise kernels per transform, each a round trip through HBM.
//
// Here: one launch. The parameter tensor phi[B, F, 3K-1] -- 99 % of the bytes -- is streamed exactly
// once: a persistent CTA per tile of whole events, the tile's contiguous parameter block is brought
// into shared memory by ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx), double-buffered so
// the copy of tile i+1 overlaps the math of tile i. LANES threads per (event, feature) row turn the row
// into knots in place (rqs.cuh); row pitch 3K-1 words is odd for even K -> conflict-free.
// Algorithmic bytes per event and transform: F(3K-1) s + 2 F s (+ s for ladj_sum).
#include <cstdlib>

#include "rqs.cuh"

namespace mf {

constexpr int kRqsThreads = 128;

template <typename T, bool INV, int LANES>
__global__ void __launch_bounds__(kRqsThreads)
rqs_kernel(const T* __restrict__ phi, const T* __restrict__ vin, int64_t B, int F, int K, RqsConst<T> C,
           int ept /*events per tile*/, int stages, int bulk_ok, T* __restrict__ vout, T* __restrict__ ladj_out,
           T* __restrict__ ladj_sum, int32_t* __restrict__ bin_out, int64_t event_stride /*elements between events*/) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int PW = 3 * K - 1;
    const int rows_per_tile = ept * F;
    const size_t tile_elems = (size_t)rows_per_tile * PW;
    const size_t tile_bytes_al = (tile_elems * sizeof(T) + 127) / 128 * 128;
    auto stage_ptr = [&](int st) { return reinterpret_cast<T*>(smem_raw + (size_t)st * tile_bytes_al); };
    unsigned char* tail = smem_raw + tile_bytes_al * stages;
    T* scratch = reinterpret_cast<T*>(tail);                         // [kRqsThreads] per-row ladj
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + kRqsThreads * sizeof(T));  // [2]

    const int tid = threadIdx.x;
    const int64_t num_tiles = (B + ept - 1) / ept;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto tile_is_bulk = [&](int64_t tile) { return bulk_ok && (tile + 1) * (int64_t)ept <= B; };
    auto issue = [&](int64_t tile, int st) {
        if (tid == 0 && tile < num_tiles && tile_is_bulk(tile)) {
            const uint32_t bytes = (uint32_t)(tile_elems * sizeof(T));
            mbar_expect_tx(&bars[st], bytes);
            bulk_g2s(stage_ptr(st), phi + (size_t)tile * tile_elems, bytes, &bars[st]);
        }
    };

    uint32_t parity0 = 0u, parity1 = 0u;
    int64_t tile = blockIdx.x;
    issue(tile, 0);
    for (int it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
        const int st = (stages == 2) ? (it & 1) : 0;
        if (stages == 2) issue(tile + gridDim.x, st ^ 1);
        const int64_t ev0 = tile * ept;
        const int nev = (int)((B - ev0) < ept ? (B - ev0) : ept);
        const int nrows = nev * F;
        T* sm = stage_ptr(st);
        if (tile_is_bulk(tile)) {
            mbar_wait(&bars[st], st ? parity1 : parity0);
            if (st) parity1 ^= 1u; else parity0 ^= 1u;
        } else if (event_stride == (int64_t)F * PW) {  // ragged last tile or unaligned phi: plain cooperative copy
            const T* src = phi + (size_t)tile * tile_elems;
            const size_t ne = (size_t)nrows * PW;
            for (size_t i = tid; i < ne; i += kRqsThreads) sm[i] = __ldg(src + i);
            __syncthreads();
        } else {  // a feature subset of a wider parameter tensor (autoregressive sampling passes): the F rows of an event are
                  // contiguous, events are event_stride elements apart
            const int epw = F * PW;
            for (int e = 0; e < nev; ++e) {
                const T* src = phi + (size_t)(ev0 + e) * event_stride;
                for (int i = tid; i < epw; i += kRqsThreads) sm[e * epw + i] = __ldg(src + i);
            }
            __syncthreads();
        }

        // ---- math: LANES threads per row ----
        const int row = tid / LANES, sub = tid % LANES;
        const bool active = row < nrows;
        T v = (T)0, out = (T)0, ladj = (T)0;
        int bin = -1;
        if (active) {
            T* r = sm + (size_t)row * PW;
            v = __ldg(vin + ev0 * F + row);
            if constexpr (LANES == 1) {
                if constexpr (sizeof(T) == 4) {
                    switch (K) {  // register-resident rows for small bin counts (rqs.cuh)
                    case 8: rqs_row_reg<8, INV>(r, C, v, out, ladj, bin); break;
                    case 10: rqs_row_reg<10, INV>(r, C, v, out, ladj, bin); break;
                    case 12: rqs_row_reg<12, INV>(r, C, v, out, ladj, bin); break;
                    case 16: rqs_row_reg<16, INV>(r, C, v, out, ladj, bin); break;
                    default: rqs_row_1lane<T, INV>(r, K, C, v, out, ladj, bin); break;
                    }
                } else {
                    rqs_row_1lane<T, INV>(r, K, C, v, out, ladj, bin);
                }
            } else {
                // lane 0 -> widths/X knots, lane 1 -> heights/Y knots; the searched array gets v
                const bool searching = INV ? (sub == 1) : (sub == 0);
                int cnt = rqs_knots_inplace(r + sub * K, K, C, searching ? v : Limits<T>::huge());
                bin = cnt - 1;
            }
        }
        if constexpr (LANES == 2) {
            __syncwarp();
            const int other = __shfl_xor_sync(0xffffffffu, bin, 1);
            if (INV ? (sub == 0) : (sub == 1)) bin = other;
            if (active && sub == 0) {
                const T* r = sm + (size_t)row * PW;
                rqs_eval<T, INV>(r, r + K, r + 2 * K, K, C, bin, v, out, ladj);
            }
        }
        if (active && sub == 0) {
            const int64_t gi = ev0 * F + row;
            vout[gi] = out;
            if (ladj_out) ladj_out[gi] = ladj;
            if (bin_out) bin_out[gi] = bin;
        }
        if (ladj_sum != nullptr) {  // sum over the F features of each event (AutoregressiveTransform)
            if (active && sub == 0) scratch[row] = ladj;
            __syncthreads();
            if (tid < nev) {
                T s = (T)0;
                for (int f = 0; f < F; ++f) s += scratch[tid * F + f];
                ladj_sum[ev0 + tid] = s;
            }
        }
        fence_proxy_async();  // in-place generic writes to this stage happen-before the next TMA refill
        __syncthreads();
        if (stages == 1) issue(tile + gridDim.x, 0);
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant for long rows (compile-time K = 32..64, fp32): 128 threads per CTA, thread t < 64 owns
// the SEARCHED softmax array of row t, thread 64 + t the other array of the same row. Each array is read from
// shared memory once into registers (rows are 3K-1 words apart -> consecutive rows hit distinct banks), the bin is
// found on the unnormalised running sum, and only {bin, knot pair} crosses between the two sides. Single-stage
// TMA tile per CTA, several CTAs per SM overlap each other's copies.
// ---------------------------------------------------------------------------------------------
template <int K, bool INV>
__global__ void __launch_bounds__(128, MF_REG2_MINB(K))
rqs_reg2_kernel(const float* __restrict__ phi, const float* __restrict__ vin, int64_t B, int F, RqsConst<float> C, int ept,
                int bulk_ok, float* __restrict__ vout, float* __restrict__ ladj_out, float* __restrict__ ladj_sum,
                int32_t* __restrict__ bin_out, int64_t event_stride /*elements between events*/) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PW = 3 * K - 1;
    const int rows_per_tile = ept * F;
    const size_t tile_elems = (size_t)rows_per_tile * PW;
    const size_t tile_bytes_al = (tile_elems * sizeof(float) + 127) / 128 * 128;
    float* sm = reinterpret_cast<float*>(smem_raw);
    float4* xch = reinterpret_cast<float4*>(smem_raw + tile_bytes_al);  // [64] {bin, k0, k1, -}
    float* lad = reinterpret_cast<float*>(xch + 64);                     // [64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(lad + 64);
    const int tid = threadIdx.x, side = tid >> 6, row = tid & 63;
    const int64_t num_tiles = (B + ept - 1) / ept;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0u;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t ev0 = tile * ept;
        const int nev = (int)((B - ev0) < ept ? (B - ev0) : ept);
        const int nrows = nev * F;
        const bool bulk = bulk_ok && nev == ept;
        const float* src = phi + (size_t)tile * tile_elems;
        if (bulk) {
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)(tile_elems * sizeof(float));
                mbar_expect_tx(bar, bytes);
                bulk_g2s(sm, src, bytes, bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else if (event_stride == (int64_t)F * PW) {
            const size_t ne = (size_t)nrows * PW;
            for (size_t i = tid; i < ne; i += 128) sm[i] = __ldg(src + i);
            __syncthreads();
        } else {  // a feature subset of a wider parameter tensor: the event's F rows are contiguous, events event_stride apart
            const int epw = F * PW;
            for (int i = tid; i < nev * epw; i += 128) {
                const int e = i / epw;
                sm[i] = __ldg(phi + (size_t)(ev0 + e) * event_stride + (i - e * epw));
            }
            __syncthreads();
        }
        const bool active = row < nrows;
        const int64_t gi = ev0 * F + row;
        float e[K];
        float S = 1.0f, v = 0.0f;
        const float* r = sm + (size_t)row * PW;
        if (active) {
            v = __ldg(vin + gi);
            const int which = (side == 0) ? (INV ? 1 : 0) : (INV ? 0 : 1);  // 0 = widths, 1 = heights
            S = rqs_exp_row<K>(r + which * K, C, e);
            if (side == 0) {
                float k0, k1;
                const int bin = rqs_search_row<K>(e, S, C.bound, v, k0, k1);
                xch[row] = make_float4(__int_as_float(bin), k0, k1, 0.0f);
            }
        }
        __syncthreads();
        if (side == 1) {
            float out = v, ladj = 0.0f;
            int bin = -1;
            if (active) {
                const float4 x = xch[row];
                bin = __float_as_int(x.x);
                if (bin >= 0 && bin < K) {
                    float o0, o1;
                    rqs_prefix_row<K>(e, S, C.bound, bin, o0, o1);
                    const float x0 = INV ? o0 : x.y, x1 = INV ? o1 : x.z, y0 = INV ? x.y : o0, y1 = INV ? x.z : o1;
                    rqs_eval_fast<INV>(r + 2 * K, K, C, bin, x0, x1, y0, y1, v, out, ladj);
                }
                vout[gi] = out;
                if (ladj_out) ladj_out[gi] = ladj;
                if (bin_out) bin_out[gi] = bin;
            }
            lad[row] = ladj;
        }
        if (ladj_sum != nullptr) {
            __syncthreads();
            if (tid < nev) {
                float s = 0.0f;
                for (int f = 0; f < F; ++f) s += lad[tid * F + f];
                ladj_sum[ev0 + tid] = s;
            }
        }
        fence_proxy_async();
        __syncthreads();
    }
}

template <typename T> struct SearchConst {};

template <typename T>
__global__ void __launch_bounds__(256)
searchsorted_kernel(const T* __restrict__ knots, const T* __restrict__ v, int64_t BF, int len, int32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BF) return;
    const T* seq = knots + i * len;
    const T x = v[i];
    int lo = 0, hi = len;  // first j with seq[j] >= x  (torch.searchsorted, right=False)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(seq + mid) < x) lo = mid + 1; else hi = mid;
    }
    idx[i] = lo;
}

template <typename T, bool INV>
static int rqs_launch(const T* phi, const T* vin, int64_t B, int F, int K, T bound, T slope, T* vout, T* ladj,
                      T* ladj_sum, int32_t* bin, cudaStream_t stream, int64_t event_stride = 0) {
    if (B < 0 || F < 1 || K < 1) return MF_E_SHAPE;
    if (K > 128) return MF_E_SHAPE;
    if (B > 0 && (!phi || !vin || !vout)) return MF_E_BADARG;
    if (B == 0) return MF_OK;
    const int PW = 3 * K - 1;
    const size_t row_bytes = (size_t)PW * sizeof(T);
    if (event_stride == 0) event_stride = (int64_t)F * PW;
    if (event_stride < (int64_t)F * PW) return MF_E_SHAPE;
    const bool strided = event_stride != (int64_t)F * PW;   // a feature subset: only the kernels with a row-wise copy path
    if constexpr (sizeof(T) == 4) {
        if ((K == 32 || K == 40 || K == 48 || K == 64) && F <= 64) {
            int dev = 0, sms = 148, smem_max = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            const size_t extra = 64 * sizeof(float4) + 64 * sizeof(float) + sizeof(uint64_t) + 128;
            auto need = [&](int e) { return ((size_t)e * F * row_bytes + 127) / 128 * 128 + extra; };
            int ept = 64 / F;
            while (ept > 1 && need(ept) > (size_t)smem_max / 4) --ept;  // aim for 4 CTAs per SM
            int bulk_ok = (aligned(phi, 16) && !strided) ? 1 : 0;
            if (bulk_ok) {
                int e = ept;
                while (e > 0 && ((size_t)e * F * row_bytes) % 16 != 0) --e;
                if (e > 0) ept = e; else bulk_ok = 0;
            }
            const size_t smem = need(ept);
            const RqsConst<float> C = make_rqs_const<float>((float)bound, (float)slope);
            const int64_t num_tiles = (B + ept - 1) / ept;
            auto go = [&](auto kern) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                int ctas = 1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 128, smem);
                if (ctas < 1) ctas = 1;
                int64_t grid = (int64_t)sms * ctas;
                if (grid > num_tiles) grid = num_tiles;
                kern<<<(unsigned)grid, 128, smem, stream>>>((const float*)phi, (const float*)vin, B, F, C, ept, bulk_ok,
                                                            (float*)vout, (float*)ladj, (float*)ladj_sum, bin, event_stride);
            };
            switch (K) {
            case 32: go(rqs_reg2_kernel<32, INV>); break;
            case 40: go(rqs_reg2_kernel<40, INV>); break;
            case 48: go(rqs_reg2_kernel<48, INV>); break;
            default: go(rqs_reg2_kernel<64, INV>); break;
            }
            return launch_status();
        }
    }
    if constexpr (sizeof(T) == 8) {   // float64 with the soft clip: rqs_f64.cu
        const int rc = rqs64_launch(INV, false, phi, vin, B, F, K, 1, bound, slope, 0, 0, vout, ladj, ladj_sum, bin, stream,
                                    event_stride);
        if (rc != MF_E_SHAPE) return rc;
    }
    if (!strided) {   // every other shape (fp64 without the clip, fp32 with bin counts the register kernels do not cover): several threads per
        // row (rqs_coop.cu). MF_RQS_COOP=0 keeps the generic kernel below, for A/B timing.
        static const bool coop = [] { const char* e = getenv("MF_RQS_COOP"); return !(e && e[0] == '0'); }();
        if (coop) {
            const int rc = rqs_coop_launch<T, INV>(phi, vin, B, F, K, bound, slope, vout, ladj, ladj_sum, bin, stream);
            if (rc != MF_E_SHAPE) return rc;
        }
    }
    // 2 lanes per row when rows are long (more warps for the same shared-memory footprint)
    const int lanes = (K >= 24) ? 2 : 1;
    const int rows_max = kRqsThreads / lanes;
    if (F > rows_max) return MF_E_SHAPE;  // one event must fit one tile
    int dev = 0, sms = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t extra = kRqsThreads * sizeof(T) + 2 * sizeof(uint64_t) + 256;
    int ept = rows_max / F;
    int stages = 2;
    auto need = [&](int e, int st) { return (((size_t)e * F * row_bytes + 127) / 128 * 128) * st + extra; };
    // prefer >= 2 CTAs per SM with double buffering; else 1 CTA double-buffered; else single buffer
    while (ept > 1 && need(ept, 2) > (size_t)smem_max) --ept;
    if (need(ept, 2) > (size_t)smem_max) stages = 1;
    if (need(ept, stages) > (size_t)smem_max) return MF_E_SHAPE;
    // tile byte size must be a multiple of 16 for the bulk copy; shrink ept until it is (or give up bulk)
    int bulk_ok = (aligned(phi, 16) && !strided) ? 1 : 0;
    if (bulk_ok) {
        int e = ept;
        while (e > 0 && ((size_t)e * F * row_bytes) % 16 != 0) --e;
        if (e > 0) ept = e; else bulk_ok = 0;
    }
    const size_t smem = need(ept, stages);
    const RqsConst<T> C = make_rqs_const<T>(bound, slope);
    const int64_t num_tiles = (B + ept - 1) / ept;
    int ctas_per_sm = (int)((size_t)smem_max / smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    int64_t grid = (int64_t)sms * ctas_per_sm;
    if (grid > num_tiles) grid = num_tiles;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kern<<<(unsigned)grid, kRqsThreads, smem, stream>>>(phi, vin, B, F, K, C, ept, stages, bulk_ok, vout, ladj,
                                                            ladj_sum, bin, event_stride);
    };
    if (lanes == 2) go(rqs_kernel<T, INV, 2>); else go(rqs_kernel<T, INV, 1>);
    return launch_status();
}

template <typename T>
static int search_launch(const T* knots, const T* v, int64_t BF, int len, int32_t* idx, cudaStream_t stream) {
    if (BF < 0 || len < 1 || (BF > 0 && (!knots || !v || !idx))) return MF_E_BADARG;
    if (BF == 0) return MF_OK;
    searchsorted_kernel<T><<<(unsigned)((BF + 255) / 256), 256, 0, stream>>>(knots, v, BF, len, idx);
    return launch_status();
}

}  // namespace mf

extern "C" int mf_rqs_forward_f32(const float* phi, const float* x, int64_t B, int F, int K, float bound, float slope,
                                  float* y, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<float, false>(phi, x, B, F, K, bound, slope, y, ladj, ladj_sum, bin, (cudaStream_t)s);
}
extern "C" int mf_rqs_forward_f64(const double* phi, const double* x, int64_t B, int F, int K, double bound,
                                  double slope, double* y, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<double, false>(phi, x, B, F, K, bound, slope, y, ladj, ladj_sum, bin, (cudaStream_t)s);
}
extern "C" int mf_rqs_inverse_f32(const float* phi, const float* y, int64_t B, int F, int K, float bound, float slope,
                                  float* x, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<float, true>(phi, y, B, F, K, bound, slope, x, ladj, ladj_sum, bin, (cudaStream_t)s);
}
extern "C" int mf_rqs_inverse_f64(const double* phi, const double* y, int64_t B, int F, int K, double bound,
                                  double slope, double* x, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<double, true>(phi, y, B, F, K, bound, slope, x, ladj, ladj_sum, bin, (cudaStream_t)s);
}
// K3 / K4 on a FEATURE SUBSET of a wider parameter tensor: the F rows of an event are contiguous, consecutive events are
// event_stride elements apart (phi points at the subset's first row). What an autoregressive sampler needs: pass j of
// MaskedAutoregressiveTransform._inverse determines feature j only, i.e. 3K-1 of the event's F_total (3K-1) parameters.
extern "C" int mf_rqs_forward_strided_f32(const float* phi, int64_t event_stride, const float* x, int64_t B, int F, int K, float bound,
                                          float slope, float* y, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<float, false>(phi, x, B, F, K, bound, slope, y, ladj, ladj_sum, bin, (cudaStream_t)s, event_stride);
}
extern "C" int mf_rqs_forward_strided_f64(const double* phi, int64_t event_stride, const double* x, int64_t B, int F, int K, double bound,
                                          double slope, double* y, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<double, false>(phi, x, B, F, K, bound, slope, y, ladj, ladj_sum, bin, (cudaStream_t)s, event_stride);
}
extern "C" int mf_rqs_inverse_strided_f32(const float* phi, int64_t event_stride, const float* y, int64_t B, int F, int K, float bound,
                                          float slope, float* x, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<float, true>(phi, y, B, F, K, bound, slope, x, ladj, ladj_sum, bin, (cudaStream_t)s, event_stride);
}
extern "C" int mf_rqs_inverse_strided_f64(const double* phi, int64_t event_stride, const double* y, int64_t B, int F, int K, double bound,
                                          double slope, double* x, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t s) {
    return mf::rqs_launch<double, true>(phi, y, B, F, K, bound, slope, x, ladj, ladj_sum, bin, (cudaStream_t)s, event_stride);
}
extern "C" int mf_searchsorted_f32(const float* knots, const float* v, int64_t BF, int len, int32_t* idx, mf_stream_t s) {
    return mf::search_launch<float>(knots, v, BF, len, idx, (cudaStream_t)s);
}
extern "C" int mf_searchsorted_f64(const double* knots, const double* v, int64_t BF, int len, int32_t* idx, mf_stream_t s) {
    return mf::search_launch<double>(knots, v, BF, len, idx, (cudaStream_t)s);
}
