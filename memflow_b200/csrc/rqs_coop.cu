// rqs_coop.cu -- K3 / K4 for the shapes the register-resident fp32 kernels (rqs.cu) do not cover: every fp64 call
// (all of MEMFlow's configs run the flows in float64) and fp32 with other bin counts.
//
// One thread per row needs the whole row in shared memory: a K = 40 fp64 row is 952 B, so only ~230 rows = 7 warps fit
// on an SM and each thread walks a serial chain of K exponentials -- the generic rqs_kernel ran at 1.1 TB/s there.
// Here a CTA of 256 threads owns up to 256 / LANES rows and the LANES threads of a row sit in DIFFERENT warps (warp =
// chunk of the row, lane = row): a warp reads 32 rows at the same offset (odd row pitch: no bank conflicts) and the
// threads of a row meet through a few words of shared memory:
//   S0 (no soft clip only) chunk maxima                         S1 chunk sums of exp, each exp stored in place
//   S2 chunk count of knots below v (offset + running sum) -> bin
//   S3 the row's chunk-0 thread rebuilds the four knots around the bin (the row is in shared memory) and evaluates the
//      rational quadratic: one fully occupied warp per 32 rows instead of every warp at 1 / LANES lane occupancy
// Knots are chunk offset + running sum, so they can differ from a sequential cumsum in the last bit (torch's CUDA
// cumsum is a parallel scan with the same property); searchsorted on given knots (mf_searchsorted_*) stays exact.
#include <cstdlib>

#include "rqs.cuh"

namespace mf {

template <typename T> MF_D T fw_exp(T a, T m, const RqsConst<T>& C);
// (library exp: a hand-written Cody-Waite + degree-13 Estrin exp with an integer exponent add was 12 % SLOWER here,
//  scripts/fastmath_check.cu keeps it for the record)
template <> MF_D double fw_exp<double>(double a, double m, const RqsConst<double>& C) {
    if (C.clip) a = div_fast(a, 1.0 + fabs(a * C.two_over_L));
    return exp(a - m);
}
template <> MF_D float fw_exp<float>(float a, float m, const RqsConst<float>& C) {
    return C.clip ? exp_clip_w(a, C) : FastMath<float>::exp(a - m);
}

// rational quadratic on explicit knots: same operation order as rqs_eval (rqs.cuh)
template <typename T, bool INV>
MF_D void rq_on_knots(T x0, T x1, T y0, T y1, T d0, T d1, T v, T& out, T& ladj) {
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

template <typename T, bool INV, int LANES>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 4 : 3)
rqs_coop_kernel(const T* __restrict__ phi, const T* __restrict__ vin, int64_t rows, int K, RqsConst<T> C, int bulk_ok,
                T* __restrict__ vout, T* __restrict__ ladj_out, int32_t* __restrict__ bin_out, int64_t rot_base) {
    constexpr int RPT = 256 / LANES;         // rows per tile
    constexpr int SLOTS = 4 * LANES;         // [2 LANES max | cnt] [2 LANES sums]
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
    T* const scM = sc + rr;
    T* const scS = sc + (size_t)2 * LANES * RPT + rr;
    const T Bd = C.bound;
    const int64_t num_tiles = (rows + RPT - 1) / RPT;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0u;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * RPT;
        const int nrows = (int)((rows - row0) < RPT ? (rows - row0) : RPT);
        const bool bulk = bulk_ok && nrows == RPT;
        const T* src = phi + (size_t)tile * tile_elems;
        if (bulk) {
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)(tile_elems * sizeof(T));
                mbar_expect_tx(bar, bytes);
                bulk_g2s(sm, src, bytes, bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {  // ragged last tile or unaligned phi: plain cooperative copy
            const size_t ne = (size_t)nrows * PW;
            for (size_t i = tid; i < ne; i += 256) sm[i] = __ldg(src + i);
            __syncthreads();
        }
        const bool valid = rr < nrows;
        // rows past the tile's end alias row 0: same work, nothing written
        T* const r = sm + (size_t)(valid ? rr : 0) * PW;
        const int64_t gi = row0 + rr;
        const T v = valid ? __ldg(vin + gi) : (T)0;
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
        // ---- S1: chunk sums; each parameter is replaced by its exponential ----
        T pre_s = (T)0;          // searched array: unnormalised sum of the chunk's elements before this row's start
        int start = k0;
        {
            T a = (T)0, b = (T)0;
            if (valid) {
                // Odd K makes the row pitch 3K-1 even: the 32 rows of a warp then share 2-4 banks at a given offset
                // (K = 35, the most common bin count of MEMFlow's configs: 8-way conflicts, -26 % in fp64, 2x in fp32).
                // This pass does not care about the order inside the chunk, so row r starts r elements in and
                // wraps: the effective stride is 3K, odd.
                const int len = k1 - k0;
                if (rot && len > 0) {   // (uniform branch: even K pays nothing for the rotation)
                    // (the start depends on the row's GLOBAL index, rot_base + row: the chain passes its event offset, so
                    //  a row is summed in the same order whatever chunk it arrives in)
                    start = k0 + (int)((rot_base + gi) % len);
                    int kk = start;
#pragma unroll 4
                    for (int j = 0; j < len; ++j) {
                        const int k = kk;
                        kk = (kk + 1 == k1) ? k0 : kk + 1;
                        const T ew = fw_exp(r[k], mw, C), eh = fw_exp(r[K + k], mh, C);
                        r[k] = ew;
                        r[K + k] = eh;
                        a += ew;
                        b += eh;
                        pre_s += (k < start) ? (INV ? eh : ew) : (T)0;
                    }
                } else {
#pragma unroll 4   // eight independent exponentials in flight: the fp64 pipe stalls on its own latency otherwise
                    for (int k = k0; k < k1; ++k) {
                        const T ew = fw_exp(r[k], mw, C), eh = fw_exp(r[K + k], mh, C);
                        r[k] = ew;
                        r[K + k] = eh;
                        a += ew;
                        b += eh;
                    }
                }
            }
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
        const T iSw = FastMath<T>::rcp(Sw), iSh = FastMath<T>::rcp(Sh);
        const int sb = INV ? K : 0, ob = INV ? 0 : K;              // searched / other coordinate
        const T is = INV ? iSh : iSw, io = INV ? iSw : iSh;
        const T cs0 = (INV ? offh : offw) * is;
        // ---- S2: knots of this chunk below v ----
        {   // same rotated walk as S1 (conflict-free for odd K): from the row's start to the end of the chunk on the running
            // sum that begins at offset + pre_s, then from the chunk's first element on the one that begins at offset
            int cnt = 0;
            if (rot) {
                T c = cs0 + pre_s * is;
                int kk = start;
                for (int j = k0; j < k1; ++j) {
                    c += r[sb + kk] * is;
                    cnt += (Bd * ((T)2 * c - (T)1) < v) ? 1 : 0;
                    if (++kk == k1) { kk = k0; c = cs0; }
                }
            } else {
                T c = cs0;
                for (int k = k0; k < k1; ++k) {
                    c += r[sb + k] * is;
                    cnt += (Bd * ((T)2 * c - (T)1) < v) ? 1 : 0;
                }
            }
            scM[chunk * RPT] = (T)cnt;   // the max slots are dead by now (read before the S1 barrier)
        }
        __syncthreads();
        int bin = ((T)-Bd < v) ? 0 : -1;
#pragma unroll
        for (int c = 0; c < LANES; ++c) bin += (int)scM[c * RPT];
        const bool inside = bin >= 0 && bin < K;
        // ---- S3: evaluation, by the row's chunk-0 thread only: the whole row is in shared memory, so any thread can
        //      rebuild the four knots around the bin (prefix of the chunk that holds it + running sum, the same values
        //      the search used). Leaving it to the thread whose chunk holds the bin would run this, the most expensive
        //      and divergent part, in every warp of the row group at 1 / LANES lane occupancy. ----
        if (valid && chunk == 0) {
            T out = v, ladj = (T)0;
            if (inside) {
                const int oc = bin / CH;                      // chunk that holds the bin
                T offs = (T)0, offo = (T)0;
#pragma unroll
                for (int c = 0; c < LANES; ++c) {
                    if (c < oc) {
                        offs += scS[(2 * c + (INV ? 1 : 0)) * RPT];
                        offo += scS[(2 * c + (INV ? 0 : 1)) * RPT];
                    }
                }
                T cs = offs * is, co = offo * io;
                for (int k = oc * CH; k < bin; ++k) { cs += r[sb + k] * is; co += r[ob + k] * io; }
                const T cs_hi = cs + r[sb + bin] * is, co_hi = co + r[ob + bin] * io;
                const T cw_lo = INV ? co : cs, cw_hi = INV ? co_hi : cs_hi, ch_lo = INV ? cs : co, ch_hi = INV ? cs_hi : co_hi;
                T d0 = (T)1, d1 = (T)1;  // D_0 = D_K = exp(0)
                if (bin > 0) { T t = r[2 * K + bin - 1]; if (C.clip) t = soft_clip(t, C.one_over_L); d0 = exp_(t); }
                if (bin < K - 1) { T t = r[2 * K + bin]; if (C.clip) t = soft_clip(t, C.one_over_L); d1 = exp_(t); }
                rq_on_knots<T, INV>(Bd * ((T)2 * cw_lo - (T)1), Bd * ((T)2 * cw_hi - (T)1), Bd * ((T)2 * ch_lo - (T)1),
                                    Bd * ((T)2 * ch_hi - (T)1), d0, d1, v, out, ladj);
            }
            vout[gi] = out;
            if (ladj_out) ladj_out[gi] = ladj;
            if (bin_out) bin_out[gi] = bin;
        }
        fence_proxy_async();  // in-place generic writes to the tile happen-before the next TMA refill
        __syncthreads();
    }
}

// sum over the F features of each event (AutoregressiveTransform), in feature order: deterministic
template <typename T>
__global__ void __launch_bounds__(256) ladj_sum_kernel(const T* __restrict__ ladj, int64_t B, int F, T* __restrict__ ladj_sum) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= B) return;
    T s = (T)0;
    for (int f = 0; f < F; ++f) s += ladj[ev * F + f];
    ladj_sum[ev] = s;
}

template <typename T, bool INV, int LANES>
static int coop_launch_lanes(const T* phi, const T* vin, int64_t rows, int K, const RqsConst<T>& C, T* vout, T* ladj,
                             int32_t* bin, cudaStream_t stream, bool must_fit_third, int64_t rot_base) {
    constexpr int RPT = 256 / LANES, SLOTS = 4 * LANES;
    const size_t row_bytes = (size_t)(3 * K - 1) * sizeof(T);
    int dev = 0, sms = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t smem = ((size_t)RPT * row_bytes + 127) / 128 * 128 + (size_t)SLOTS * RPT * sizeof(T) + 16;
    if (smem > (size_t)smem_max / (must_fit_third ? 3 : 1)) return MF_E_SHAPE;
    const int bulk_ok = aligned(phi, 16) ? 1 : 0;  // RPT * row_bytes is a multiple of 128
    auto kern = rqs_coop_kernel<T, INV, LANES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int ctas = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 256, smem);
    if (ctas < 1) ctas = 1;
    const int64_t num_tiles = (rows + RPT - 1) / RPT;
    int64_t grid = (int64_t)sms * ctas;
    if (grid > num_tiles) grid = num_tiles;
    kern<<<(unsigned)grid, 256, smem, stream>>>(phi, vin, rows, K, C, bulk_ok, vout, ladj, bin, rot_base);
    return launch_status();
}

// ---- fp32, K = 8, 10, 12, 16: one thread per row with the row in registers (rqs_row_reg, the row math of the chain's
// launch A), 256-row single-stage TMA tiles, 8 CTAs per SM. The generic rqs_kernel serves these shapes with 128-thread
// double-buffered CTAs (28 warps per SM) and whole-event tiles: 81 % of the HBM roofline at F = 7, K = 10. ----
template <int K, bool INV>
__global__ void __launch_bounds__(256, 8)
rqs_small_kernel(const float* __restrict__ phi, const float* __restrict__ vin, int64_t rows, RqsConst<float> C, int bulk_ok,
                 float* __restrict__ vout, float* __restrict__ ladj_out, int32_t* __restrict__ bin_out) {
    constexpr int PW = 3 * K - 1, RPT = 256;
    constexpr uint32_t tile_bytes = RPT * PW * (uint32_t)sizeof(float);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* const sm = reinterpret_cast<float*>(smem_raw);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem_raw + (tile_bytes + 127) / 128 * 128);
    const int tid = threadIdx.x;
    float* const rowp = sm + (size_t)tid * PW;
    const int64_t num_tiles = (rows + RPT - 1) / RPT, full_tiles = bulk_ok ? rows / RPT : 0;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0u;
    int64_t tile = blockIdx.x;
    if (tid == 0 && tile < full_tiles) {
        mbar_expect_tx(bar, tile_bytes);
        bulk_g2s(sm, phi + (size_t)tile * RPT * PW, tile_bytes, bar);
    }
    for (; tile < num_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * RPT;
        const int nrows = (int)((rows - row0) < RPT ? (rows - row0) : RPT);
        if (tile < full_tiles) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {  // ragged last tile or unaligned phi: plain cooperative copy
            const float* src = phi + (size_t)row0 * PW;
            for (int i = tid; i < nrows * PW; i += RPT) sm[i] = __ldg(src + i);
            __syncthreads();
        }
        if (tid < nrows) {
            float out, ladj;
            int bin;
            rqs_row_reg<K, INV>(rowp, C, __ldg(vin + row0 + tid), out, ladj, bin);
            vout[row0 + tid] = out;
            if (ladj_out) ladj_out[row0 + tid] = ladj;
            if (bin_out) bin_out[row0 + tid] = bin;
        }
        fence_proxy_async();  // in-place generic writes to the tile happen-before the next TMA refill
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < full_tiles) {
            mbar_expect_tx(bar, tile_bytes);
            bulk_g2s(sm, phi + (size_t)(tile + gridDim.x) * RPT * PW, tile_bytes, bar);
        }
    }
}

template <int K, bool INV>
static int small_launch(const float* phi, const float* vin, int64_t rows, const RqsConst<float>& C, float* vout, float* ladj,
                        int32_t* bin, cudaStream_t stream) {
    constexpr int PW = 3 * K - 1;
    const size_t smem = ((size_t)256 * PW * sizeof(float) + 127) / 128 * 128 + 16;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = rqs_small_kernel<K, INV>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int ctas = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 256, smem);
    if (ctas < 1) ctas = 1;
    const int64_t num_tiles = (rows + 255) / 256;
    int64_t grid = (int64_t)sms * ctas;
    if (grid > num_tiles) grid = num_tiles;
    kern<<<(unsigned)grid, 256, smem, stream>>>(phi, vin, rows, C, aligned(phi, 16) ? 1 : 0, vout, ladj, bin);
    return launch_status();
}

// Returns MF_E_SHAPE when no variant fits (the caller falls back to the generic kernel).
template <typename T, bool INV>
int rqs_coop_launch(const T* phi, const T* vin, int64_t B, int F, int K, T bound, T slope, T* vout, T* ladj, T* ladj_sum,
                    int32_t* bin, cudaStream_t stream, int64_t row_offset) {
    static const int forced = [] { const char* e = getenv("MF_RQS_LANES"); return e ? atoi(e) : 0; }();
    const RqsConst<T> C = make_rqs_const<T>(bound, slope);
    const size_t row_bytes = (size_t)(3 * K - 1) * sizeof(T);
    const int64_t rows = B * (int64_t)F;
    // tiles are RPT rows whatever F is; the per-event sum of the log-dets is a second, tiny launch over the per-row
    // values (8 B per row against the row's 3K-1 parameters), which needs them in memory
    T* ladj_rows = ladj;
    if (ladj_sum && !ladj_rows) {
        if (cudaMallocAsync((void**)&ladj_rows, (size_t)rows * sizeof(T), stream) != cudaSuccess) return -(int)cudaErrorMemoryAllocation;
    }
    const int pref = forced ? forced : (row_bytes < 420 ? 1 : (row_bytes < 700 ? 2 : (row_bytes < 2000 ? 4 : 8)));
    int rc = MF_E_SHAPE;
    if constexpr (sizeof(T) == 4) {   // register-row kernel for the small bin counts
        if (!forced) {
            if (K == 8) rc = small_launch<8, INV>(phi, vin, rows, C, vout, ladj_rows, bin, stream);
            if (K == 10) rc = small_launch<10, INV>(phi, vin, rows, C, vout, ladj_rows, bin, stream);
            if (K == 12) rc = small_launch<12, INV>(phi, vin, rows, C, vout, ladj_rows, bin, stream);
            if (K == 16) rc = small_launch<16, INV>(phi, vin, rows, C, vout, ladj_rows, bin, stream);
        }
    }
    for (int pass = 0; pass < 2 && rc == MF_E_SHAPE; ++pass) {  // first try for >= 3 CTAs per SM, then anything that fits
        for (int lanes = pref; lanes <= 8 && rc == MF_E_SHAPE; lanes *= 2) {
            const bool third = pass == 0;
            if (lanes == 1) rc = coop_launch_lanes<T, INV, 1>(phi, vin, rows, K, C, vout, ladj_rows, bin, stream, third, row_offset);
            if (lanes == 2) rc = coop_launch_lanes<T, INV, 2>(phi, vin, rows, K, C, vout, ladj_rows, bin, stream, third, row_offset);
            if (lanes == 4) rc = coop_launch_lanes<T, INV, 4>(phi, vin, rows, K, C, vout, ladj_rows, bin, stream, third, row_offset);
            if (lanes == 8) rc = coop_launch_lanes<T, INV, 8>(phi, vin, rows, K, C, vout, ladj_rows, bin, stream, third, row_offset);
        }
    }
    if (rc == MF_OK && ladj_sum) {
        ladj_sum_kernel<T><<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(ladj_rows, B, F, ladj_sum);
        rc = launch_status();
    }
    if (ladj_rows != ladj) cudaFreeAsync(ladj_rows, stream);
    return rc;
}

template int rqs_coop_launch<float, false>(const float*, const float*, int64_t, int, int, float, float, float*, float*, float*, int32_t*, cudaStream_t, int64_t);
template int rqs_coop_launch<float, true>(const float*, const float*, int64_t, int, int, float, float, float*, float*, float*, int32_t*, cudaStream_t, int64_t);
template int rqs_coop_launch<double, false>(const double*, const double*, int64_t, int, int, double, double, double*, double*, double*, int32_t*, cudaStream_t, int64_t);
template int rqs_coop_launch<double, true>(const double*, const double*, int64_t, int, int, double, double, double*, double*, double*, int32_t*, cudaStream_t, int64_t);

}  // namespace mf
