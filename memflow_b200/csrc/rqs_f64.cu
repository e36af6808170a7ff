// rqs_f64.cu -- K3 / K4 and the flow stage of the sampling chain in FLOAT64 (the dtype of every MEMFlow production flow:
// `.double()` at memflow/transfer_flow/transfer_flow_paper.py:116-127, configs/*/flow_config_*.yaml), soft clip on
// (zuko's default slope = 1e-3; slope = None goes to rqs_coop.cu).
//
// Round 1 ran this shape through rqs_coop_kernel<double>: library exp + a range-checked division per bin, ~95 issued
// instructions per parameter (IMAD / BRA / BSSY overhead of the library paths), FP64 pipe at 45 %, 0.37-0.45 of HBM.
// What changed here (ncu: profiles/r2/):
//   * exp(a / (1 + |a| c)) is ONE branch-free sequence of 19 FP64 instructions: the log2(e) factor is folded into the
//     divisor (q = a / (ln2 + |a| c ln2)), the reciprocal is MUFU.RCP64H + a cubic correction (r0 (1 + e + e^2), 4
//     operations, <= 1 ulp), 2^q = 2^k * P11(q - k) with a degree-11 minimax polynomial (3.2e-18) and an integer exponent
//     add. The soft clip bounds |q| <= |log2 slope| / 2, so no range check is needed; non-finite / astronomically large
//     parameters are detected by one integer max per element and send the segment to an out-of-line exact path.
//   * each thread walks ONE contiguous segment of the row and stores the RUNNING SUM in place (the sum is computed
//     anyway): the bin search is then a binary search on stored prefix sums by one thread per row (~5 dependent
//     shared-memory loads) instead of a second pass of 4 FP64 instructions per bin by every thread, and the knots
//     around the bin are 4 loads. One barrier less per tile.
//   * the per-row evaluation (2 exp, 3 reciprocals, sqrt, log) is split: S3a gathers its ten operands into registers,
//     then the tile is released and the NEXT tile's TMA copy is issued; S3b (pure register math + the global stores)
//     runs while that copy is in flight.
//   * odd bin counts (K = 35 is the most common one in MEMFlow's configs) make the row pitch even: rows are then
//     copied by one bulk copy EACH into 16-byte aligned rows whose pitch is 2 (mod 4) words and are walked with 128-bit
//     shared-memory accesses (conflict-free, half the LSU instructions); no rotation, so results never depend on
//     where a row sits in the batch.
//   * tiles are whole events, so the per-event sum of log-dets needs no second launch, and the same kernel runs the
//     whole float64 flow stage of the sampling chain (Philox base point -> T inverses -> u, log q) in ONE launch.
#include <cstdlib>

#include "philox.cuh"
#include "rqs.cuh"

namespace mf {

// exact path for such a segment, from the untouched global copy of the row (never taken for sane parameters)
static __device__ __noinline__ double segment_exact(const double* __restrict__ g, double* __restrict__ r, int lo, int hi,
                                                    double two_over_L) {
    double c = 0.0;
    for (int k = lo; k < hi; ++k) {
        const double a = __ldg(g + k);
        c += exp(a / (1.0 + fabs(a * two_over_L)));
        r[k] = c;
    }
    return c;
}

// One segment [lo, hi) of a row: each parameter is replaced by the running sum of the exponentials from lo on.
template <bool PAIR>
MF_D double scan_segment(double* __restrict__ r, int lo, int hi, const RqsConst<double>& C, int& hi_abs_max) {
    double c = 0.0;
    int k = lo;
    if constexpr (PAIR) {   // rows are 16-byte aligned: 128-bit accesses on even offsets, a single at either end
        if ((k & 1) && k < hi) {
            c += exp2_clip(r[k], C.k1w, C.ln2, hi_abs_max);
            r[k] = c;
            ++k;
        }
#pragma unroll 2
        for (; k + 1 < hi; k += 2) {
            const double2 a = *reinterpret_cast<const double2*>(r + k);
            const double e0 = exp2_clip(a.x, C.k1w, C.ln2, hi_abs_max), e1 = exp2_clip(a.y, C.k1w, C.ln2, hi_abs_max);
            double2 o;
            o.x = c + e0;
            o.y = o.x + e1;
            c = o.y;
            *reinterpret_cast<double2*>(r + k) = o;
        }
        if (k < hi) {
            c += exp2_clip(r[k], C.k1w, C.ln2, hi_abs_max);
            r[k] = c;
        }
    } else {
#pragma unroll 4
        for (; k < hi; ++k) {
            c += exp2_clip(r[k], C.k1w, C.ln2, hi_abs_max);
            r[k] = c;
        }
    }
    return c;
}

struct Rqs64Args {
    const double* phi;      // [T][B][F][3K-1]
    const double* vin;      // [B][F] (not CHAIN)
    double* vout;           // [B][F]: y / x (not CHAIN), u in [0,1] (CHAIN)
    double* ladj_out;       // [B][F] or null
    double* ladj_sum;       // [B] or null; CHAIN: log q
    int32_t* bin_out;       // [B][F] or null
    int64_t B;
    size_t transform_pitch; // elements between phi[t] and phi[t+1]
    int64_t event_stride;   // elements between consecutive events (F (3K-1) unless the rows are a feature subset of a wider tensor)
    uint64_t seed, event_offset;
    RqsConst<double> C;
    int F, K, T;
    int rows_per_tile;      // whole events (a multiple of F: per-event sums in the kernel) or any even count <= 256 / LANES
    int whole_events;
    int pitch /*shared-memory row pitch in elements*/, bulk_ok, chw /*chunk length*/, nsteps /*binary search*/;
    int stages;             // tile buffers: 2 = the next (tile, transform) is copied while this one is processed
};

#ifndef MF_RQS64_MINB
#define MF_RQS64_MINB(NW) (24 / (NW))   // (seven 128-thread CTAs: 72 registers with spills, measured 1-2 % slower than six)
#endif
// NW = warps per CTA: 8 (three CTAs per SM) or 4 (six CTAs of half the rows: the same 24 warps per SM in twice as many
// independent tile pipelines and barriers that span four warps instead of eight)
template <bool INV, int LANES, bool PAIR, bool CHAIN, int NW>
__global__ void __launch_bounds__(NW * 32, MF_RQS64_MINB(NW))
rqs64_kernel(const __grid_constant__ Rqs64Args A) {
    constexpr int RPT = NW * 32 / LANES;          // rows per tile (capacity)
    constexpr int HC = LANES >= 2 ? LANES / 2 : 1;  // chunks per softmax array
    constexpr int NS = 2 * HC;                    // segments per row
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = A.K, PW = 3 * K - 1, PWP = A.pitch, F = A.F, T = A.T;
    const RqsConst<double>& C = A.C;
    const size_t buf_bytes = ((size_t)RPT * PWP * sizeof(double) + 127) / 128 * 128;
    const int stages = A.stages;
    double* const tot = reinterpret_cast<double*>(smem_raw + (size_t)stages * buf_bytes);  // [NS][RPT]
    double* const lad = tot + (size_t)NS * RPT;   // [RPT] per-row log-det (sums over the event's features)
    uint64_t* const bars = reinterpret_cast<uint64_t*>(lad + RPT);
    auto buf = [&](int st) { return reinterpret_cast<double*>(smem_raw + (size_t)st * buf_bytes); };
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = wid % LANES, group = wid / LANES, rr = group * 32 + lane;
    const bool evaluator = chunk == (group % LANES);   // the thread of the row that searches and evaluates (spread over the schedulers)
    const int rows_per_tile = A.rows_per_tile;
    const int64_t total_rows = A.B * (int64_t)F;
    const int64_t num_tiles = (total_rows + rows_per_tile - 1) / rows_per_tile;
    const int64_t full_tiles = A.bulk_ok ? total_rows / rows_per_tile : 0;
    const size_t tile_elems = (size_t)rows_per_tile * PW;
    const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(double));
    const double Bd = C.bound, half_over_B = 0.5 / Bd;
    const bool want_sum = A.ladj_sum != nullptr && A.whole_events;

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();
    // bulk copy of (tile, transform): one copy when the shared-memory rows are packed, one per row when they are padded
    auto issue = [&](int64_t tile, int t, int st) {
        const double* src = A.phi + (size_t)t * A.transform_pitch + (size_t)tile * tile_elems;
        double* const sm = buf(st);
        uint64_t* const bar = &bars[st];
        if (PWP == PW) {
            if (tid == 0) {
                mbar_expect_tx(bar, tile_bytes);
                bulk_g2s(sm, src, tile_bytes, bar);
            }
        } else {
            // (a bulk copy is issued from uniform registers: the lanes of a warp take turns, so the rows are spread over
            //  all 8 warps; complete_tx may overtake thread 0's expect_tx, the pending arrival keeps the phase open)
            if (tid == 0) mbar_expect_tx(bar, tile_bytes);
            const int per = (rows_per_tile + NW - 1) / NW;
            const int row = wid * per + lane;
            if (lane < per && row < rows_per_tile)
                bulk_g2s(sm + (size_t)row * PWP, src + (size_t)row * PW, (uint32_t)(PW * sizeof(double)), bar);
        }
    };
    uint32_t parity0 = 0u, parity1 = 0u;
    int step = 0;   // (tile, transform) steps done by this CTA: step & 1 is the buffer of a two-stage run
    int64_t tile = blockIdx.x;
    if (tile < full_tiles) issue(tile, T - 1, 0);
    for (; tile < num_tiles; tile += gridDim.x) {
        const bool bulk = tile < full_tiles;
        const int64_t row0 = tile * rows_per_tile;
        const int nrows = (int)((total_rows - row0) < rows_per_tile ? (total_rows - row0) : rows_per_tile);
        const bool valid = rr < nrows;
        const int64_t gi = row0 + rr;
        double v = 0.0, ladj_acc = 0.0;
        if (evaluator && valid) {
            if constexpr (CHAIN) {   // base point: Philox counter = (global event index, feature block), key = seed
                const int64_t e = gi / F;
                const int f = (int)(gi - e * F);
                const uint64_t gev = A.event_offset + (uint64_t)e;
                const Philox4 r4 = philox4x32_10((uint32_t)gev, (uint32_t)(gev >> 32), (uint32_t)(f >> 2), 0u, (uint32_t)A.seed,
                                                 (uint32_t)(A.seed >> 32));
                const uint32_t bits = (f & 3) == 0 ? r4.x : ((f & 3) == 1 ? r4.y : ((f & 3) == 2 ? r4.z : r4.w));
                v = (double)((float)Bd * (2.0f * u01_from_bits(bits) - 1.0f));   // the fp32 chain's base point, widened
            } else {
                v = __ldg(A.vin + gi);
            }
        }
        const bool packed = A.event_stride == (int64_t)F * PW;
        // global address of tile row `row` (packed: rows follow each other; feature subset: events are event_stride apart)
        auto row_src = [&](const double* base, int row) {
            if (packed) return base + (size_t)(row0 + row) * PW;
            const int64_t g = row0 + row, e = g / F;
            return base + (size_t)e * A.event_stride + (size_t)(g - e * F) * PW;
        };
        for (int t = T - 1; t >= 0; --t, ++step) {   // sampling applies the transforms in reverse order (T = 1 for K3 / K4)
            const double* tbase = A.phi + (size_t)t * A.transform_pitch;
            const int st = stages == 2 ? (step & 1) : 0;
            double* const sm = buf(st);
            double* const r = sm + (size_t)rr * PWP;
            // two stages: the other buffer was released by the second barrier of the step before, refill it now
            const int64_t nt = t > 0 ? tile : tile + gridDim.x;
            if (stages == 2 && nt < full_tiles) issue(nt, t > 0 ? t - 1 : T - 1, st ^ 1);
            if (bulk) {
                mbar_wait(&bars[st], st ? parity1 : parity0);
                if (st) parity1 ^= 1u; else parity0 ^= 1u;
            } else {  // ragged last tile or unaligned phi: plain cooperative copy into the same (padded) rows
                for (int row = wid; row < nrows; row += NW) {
                    const double* src = row_src(tbase, row);
                    for (int i = lane; i < PW; i += 32) sm[(size_t)row * PWP + i] = __ldg(src + i);
                }
                __syncthreads();
            }
            // ---- S1: running sums of exp(clip(.)) over this thread's segment(s), in place ----
            if (valid) {
#pragma unroll
                for (int s = chunk; s < NS; s += LANES) {
                    const int arr = s / HC, ch = s - arr * HC;
                    const int lo = arr * K + min(K, ch * A.chw), hi = arr * K + min(K, (ch + 1) * A.chw);
                    int hmax = 0;
                    double total = scan_segment<PAIR>(r, lo, hi, C, hmax);
                    if (hmax >= kHiAbsLimit) total = segment_exact(row_src(tbase, rr), r, lo, hi, C.two_over_L);
                    tot[s * RPT + rr] = total;
                }
            }
            __syncthreads();
            // ---- S3a (one thread per row): bin by binary search on the stored running sums, operands into registers ----
            double Sw = 1.0, Sh = 1.0, s_lo = 0.0, s_hi = 1.0, o_lo = 0.0, o_hi = 1.0, dr0 = 0.0, dr1 = 0.0;
            int bin = -1;
            if (evaluator && valid) {
                double tw[HC], th[HC];
                Sw = 0.0;
                Sh = 0.0;
#pragma unroll
                for (int c = 0; c < HC; ++c) {
                    tw[c] = tot[c * RPT + rr];
                    th[c] = tot[(HC + c) * RPT + rr];
                    Sw += tw[c];
                    Sh += th[c];
                }
                const double target = fma(v, half_over_B, 0.5) * (INV ? Sh : Sw);  // v in units of the searched running sum
                int oc = 0;          // chunk of the searched array that holds the first sum >= target
                double off = 0.0;
#pragma unroll
                for (int c = 0; c + 1 < HC; ++c) {
                    const double ts = INV ? th[c] : tw[c];
                    if (oc == c && off + ts < target) { off += ts; oc = c + 1; }
                }
                const int sb = INV ? K : 0;
                const int c0 = min(K, oc * A.chw), len = min(K, c0 + A.chw) - c0;
                const double tloc = target - off;
                int lo = 0, hi = len;
                for (int step = 0; step < A.nsteps; ++step) {
                    const int mid = (lo + hi) >> 1;
                    const bool go = lo < hi;
                    const bool below = r[sb + c0 + min(mid, len - 1)] < tloc;
                    lo = (go && below) ? mid + 1 : lo;
                    hi = (go && !below) ? mid : hi;
                }
                bin = c0 + lo - ((-Bd < v) ? 0 : 1);   // searchsorted(knots, v) - 1 with X_0 = -B
                if (bin >= 0 && bin < K) {
                    // true prefix sums at bin - 1 and bin of both arrays: chunk offset + stored running sum
                    auto pref = [&](int base, const double (&tt)[HC], int j) {
                        double o = 0.0;
#pragma unroll
                        for (int c = 0; c + 1 < HC; ++c) o += (j >= (c + 1) * A.chw) ? tt[c] : 0.0;
                        return o + r[base + j];
                    };
                    const int jm = max(bin - 1, 0);
                    const double slo = INV ? pref(K, th, jm) : pref(0, tw, jm), olo = INV ? pref(0, tw, jm) : pref(K, th, jm);
                    s_hi = INV ? pref(K, th, bin) : pref(0, tw, bin);
                    o_hi = INV ? pref(0, tw, bin) : pref(K, th, bin);
                    s_lo = bin > 0 ? slo : 0.0;
                    o_lo = bin > 0 ? olo : 0.0;
                    dr0 = bin > 0 ? r[2 * K + bin - 1] : 0.0;       // D_0 = D_K = exp(0)
                    dr1 = bin < K - 1 ? r[2 * K + bin] : 0.0;
                }
            }
            fence_proxy_async();  // in-place generic writes to the tile happen-before the next TMA refill
            __syncthreads();
            // single stage: refill it with the next transform of this tile, else the next tile's first
            if (stages == 1 && nt < full_tiles) issue(nt, t > 0 ? t - 1 : T - 1, 0);
            // ---- S3b: the rational quadratic on the bin's four knots, registers only, while the next copy is in flight ----
            if (evaluator && valid) {
                double out = v, ladj = 0.0;
                if (bin >= 0 && bin < K) {
                    int hm = 0;
                    double d0 = exp2_clip(dr0, C.k1d, C.ln2, hm), d1 = exp2_clip(dr1, C.k1d, C.ln2, hm);
                    if (hm >= kHiAbsLimit) {   // non-finite / astronomically large derivative parameter: exact path
                        d0 = exp(dr0 / (1.0 + fabs(dr0 * C.one_over_L)));
                        d1 = exp(dr1 / (1.0 + fabs(dr1 * C.one_over_L)));
                    }
                    const double isw = rcp_fast(Sw), ish = rcp_fast(Sh);
                    const double cw_lo = INV ? o_lo : s_lo, cw_hi = INV ? o_hi : s_hi;
                    const double ch_lo = INV ? s_lo : o_lo, ch_hi = INV ? s_hi : o_hi;
                    const double x0 = Bd * fma(2.0 * cw_lo, isw, -1.0), x1 = Bd * fma(2.0 * cw_hi, isw, -1.0);
                    const double y0 = Bd * fma(2.0 * ch_lo, ish, -1.0), y1 = Bd * fma(2.0 * ch_hi, ish, -1.0);
                    const double dx = x1 - x0, dy = y1 - y0;
                    const double s = div_fast(dy, dx);
                    const double dd = d0 + d1 - 2.0 * s;
                    double z;
                    if constexpr (!INV) {
                        z = div_fast(v - x0, dx);
                        const double zz = z * (1.0 - z);
                        out = y0 + div_fast(dy * (s * (z * z) + d0 * zz), s + dd * zz);
                    } else {
                        const double y_ = v - y0;
                        const double a = dy * (s - d0) + y_ * dd;
                        const double b = dy * d0 - y_ * dd;
                        const double c = -s * y_;
                        z = div_fast(2.0 * c, -b - sqrt_fast(b * b - 4.0 * a * c));
                        out = x0 + z * dx;
                    }
                    const double zz = z * (1.0 - z);
                    const double den = s + dd * zz;
                    const double omz = 1.0 - z;
                    ladj = log(div_fast((s * s) * (2.0 * s * zz + d0 * (omz * omz) + d1 * (z * z)), den * den));
                }
                if constexpr (CHAIN) {
                    v = out;           // the next transform's input stays in this thread's register
                    ladj_acc += ladj;
                    if (t == 0) {
                        A.vout[gi] = (out + Bd) * half_over_B;   // u = (x + B) / 2B
                        if (A.ladj_out) A.ladj_out[gi] = ladj_acc;  // row tiles: the per-event sum is a second, tiny launch
                    }
                } else {
                    A.vout[gi] = out;
                    if (A.ladj_out) A.ladj_out[gi] = ladj;
                    if (A.bin_out) A.bin_out[gi] = bin;
                    ladj_acc = ladj;
                }
            }
        }
        if (want_sum) {   // sum over the F features of each event (AutoregressiveTransform), in feature order: deterministic
            if (evaluator && valid) lad[rr] = ladj_acc;
            __syncthreads();
            if (tid * F < nrows) {
                double s = 0.0;
                for (int f = 0; f < F; ++f) s += lad[tid * F + f];
                A.ladj_sum[row0 / F + tid] = s;
            }
            // lad is rewritten only after the next tile's barriers
        }
    }
}


// sum over the F features of each event, in feature order (row tiles: events straddle tiles)
__global__ void __launch_bounds__(256) rqs64_event_sum_kernel(const double* __restrict__ ladj, int64_t B, int F, double* __restrict__ out) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= B) return;
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += __ldg(ladj + ev * F + f);
    out[ev] = s;
}

template <bool INV, int LANES, bool PAIR, bool CHAIN, int NW>
static int rqs64_go(Rqs64Args& A, size_t smem, int sms, cudaStream_t stream) {
    auto kern = rqs64_kernel<INV, LANES, PAIR, CHAIN, NW>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int ctas = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, NW * 32, smem);
    if (ctas < 1) ctas = 1;
    const int64_t total_rows = A.B * (int64_t)A.F;
    const int64_t num_tiles = (total_rows + A.rows_per_tile - 1) / A.rows_per_tile;
    int64_t grid = (int64_t)sms * ctas;
    if (grid > num_tiles) grid = num_tiles;
    kern<<<(unsigned)grid, NW * 32, smem, stream>>>(A);
    return launch_status();
}

template <bool INV, bool CHAIN>
static int rqs64_pick(Rqs64Args& A, int lanes, bool pair, int nw, size_t smem, int sms, cudaStream_t stream) {
#define MF_GO(L, W) (pair ? rqs64_go<INV, L, true, CHAIN, W>(A, smem, sms, stream) : rqs64_go<INV, L, false, CHAIN, W>(A, smem, sms, stream))
    if (nw == 4) {
        switch (lanes) {
        case 2: return MF_GO(2, 4);
        default: return MF_GO(4, 4);
        }
    }
    switch (lanes) {
    case 1: return MF_GO(1, 8);
    case 2: return MF_GO(2, 8);
    case 4: return MF_GO(4, 8);
    default: return MF_GO(8, 8);
    }
#undef MF_GO
}

// Returns MF_E_SHAPE when the shape is outside this kernel (no soft clip, a row that does not fit shared memory): the
// caller then takes the general kernels of rqs_coop.cu / rqs.cu.
int rqs64_launch(bool inverse, bool chain, const double* phi, const double* vin, int64_t B, int F, int K, int T, double bound,
                 double slope, uint64_t seed, uint64_t event_offset, double* vout, double* ladj, double* ladj_sum,
                 int32_t* bin, cudaStream_t stream, int64_t event_stride) {
    static const bool enabled = [] { const char* e = getenv("MF_RQS64"); return !(e && e[0] == '0'); }();
    static const int forced = [] { const char* e = getenv("MF_RQS64_LANES"); return e ? atoi(e) : 0; }();
    static const int tiles_env = [] { const char* e = getenv("MF_RQS64_TILES"); return e ? atoi(e) : 0; }();  // 1: whole events, 2: row tiles
    if (!enabled || !(slope > 0.0) || K < 1 || F < 1 || T < 1) return MF_E_SHAPE;
    Rqs64Args A{};
    A.C = make_rqs_const<double>(bound, slope);
    if (!(A.C.k1w < 1048576.0)) return MF_E_SHAPE;   // slope within ~1e-6 of 1: the fast reciprocal's range argument no longer holds
    const int PW = 3 * K - 1;
    const bool pair = (K & 1) != 0;                   // odd K: even row pitch, 16-byte aligned rows
    const int pitch = pair ? (((PW / 2) & 1) ? PW : PW + 2) : PW;
    int dev = 0, sms = 148, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t row_bytes = (size_t)pitch * sizeof(double);
    // MF_RQS64_WARPS = 4 / 8 forces 128- / 256-thread CTAs (128: 2 or 4 threads per row only)
    static const int warps_env = [] { const char* e = getenv("MF_RQS64_WARPS"); return e ? atoi(e) : 0; }();
    int nw = 8;
    auto need = [&](int lanes, int stages) {
        const int rpt = nw * 32 / lanes, ns = lanes >= 2 ? lanes : 2;
        return (((size_t)rpt * row_bytes + 127) / 128 * 128) * stages + (size_t)(ns + 1) * rpt * sizeof(double) + 16;
    };
    // MF_RQS64_STAGES=2: two tile buffers (the next (tile, transform) is copied while this one is processed). Measured on
    // B200 (scripts/rqs64_ab.py) and NOT the default: two 61 KB buffers leave one CTA per SM at 4 threads per row, and 8
    // threads per row (32-row tiles, three CTAs) put the per-row evaluation of 32 rows on one warp: F=10 K=40 0.270 ms
    // single stage / 4 threads per row, 0.364 ms two stages / 8, 0.374 ms single stage / 8.
    static const int stages_env = [] { const char* e = getenv("MF_RQS64_STAGES"); return e ? atoi(e) : 0; }();
    const int pref = forced ? forced : (row_bytes < 420 ? 1 : (row_bytes < 700 ? 2 : (row_bytes < 2000 ? 4 : 8)));
    // odd K (one bulk copy per row) gains 3-5 % from the smaller CTAs (scripts/rqs64_ab.py: F=10 K=35 0.298 -> 0.290 ms, the
    // chain's flow stage 1.489 -> 1.418 ms), even K loses 0-4 %
    if ((warps_env == 4 || (warps_env == 0 && pair)) && (pref == 2 || pref == 4)) nw = 4;
    int lanes = 0, stages = 1;
    if (stages_env == 2)
        for (int l = pref; l <= 8 && !lanes; l *= 2)
            if (need(l, 2) <= (size_t)smem_max) { lanes = l; stages = 2; }
    for (int pass = 0; pass < 2 && !lanes; ++pass)   // single stage: first aim at 3 CTAs per SM, then at anything that fits
        for (int l = pref; l <= 8 && !lanes; l *= 2)
            if (need(l, 1) <= (size_t)smem_max / (pass == 0 ? 24 / nw : 1)) lanes = l;
    if (!lanes) return MF_E_SHAPE;
    if (nw == 4 && lanes != 2 && lanes != 4) nw = 8;
    const int rpt = nw * 32 / lanes, hc = lanes >= 2 ? lanes / 2 : 1;
    // Tiles of whole events keep the per-event log-det sum in the kernel; when the events pack a tile badly (F = 20 in
    // 32 rows: 37 % of the lanes idle) or no sum is asked for, tiles are plain runs of rows and the sum is a second,
    // tiny launch over 8 B per row.
    const bool want_sum = ladj_sum != nullptr;
    int ept = rpt / F;
    bool whole = want_sum && ept >= 1 && (double)(ept * F) >= 0.9 * rpt;
    if (tiles_env == 1 && ept >= 1) whole = true;
    if (tiles_env == 2) whole = false;
    A.transform_pitch = (size_t)B * F * PW;
    A.event_stride = event_stride > 0 ? event_stride : (int64_t)F * PW;
    const bool strided = A.event_stride != (int64_t)F * PW;   // feature subset: rows are gathered by plain loads, no bulk copy
    if (strided && (chain || T != 1)) return MF_E_SHAPE;
    int bulk_ok = !strided && aligned(phi, 16) && (T == 1 || (A.transform_pitch * sizeof(double)) % 16 == 0) ? 1 : 0;
    int rows_per_tile = whole ? ept * F : rpt;
    if (bulk_ok && pitch == PW && ((size_t)rows_per_tile * PW * sizeof(double)) % 16 != 0) {  // one copy per tile: 16-byte multiples
        if (whole) {
            int e = ept;
            while (e > 0 && ((size_t)e * F * PW * sizeof(double)) % 16 != 0) --e;
            if (e > 0) rows_per_tile = e * F; else bulk_ok = 0;
        } else {
            rows_per_tile -= 1;   // rpt is even
        }
    }
    double* ladj_rows = ladj;
    if (want_sum && !whole && !ladj_rows) {
        if (cudaMallocAsync((void**)&ladj_rows, (size_t)B * F * sizeof(double), stream) != cudaSuccess) return -(int)cudaErrorMemoryAllocation;
    }
    A.phi = phi; A.vin = vin; A.vout = vout; A.ladj_out = ladj_rows; A.ladj_sum = ladj_sum; A.bin_out = bin;
    A.B = B; A.seed = seed; A.event_offset = event_offset;
    A.F = F; A.K = K; A.T = T; A.rows_per_tile = rows_per_tile; A.whole_events = whole ? 1 : 0; A.pitch = pitch; A.bulk_ok = bulk_ok;
    A.chw = (K + hc - 1) / hc;
    A.stages = stages;
    A.nsteps = 0;
    while ((1 << A.nsteps) < A.chw + 1) ++A.nsteps;
    const size_t smem = need(lanes, stages);
    int rc;
    if (chain) rc = rqs64_pick<true, true>(A, lanes, pair, nw, smem, sms, stream);
    else rc = inverse ? rqs64_pick<true, false>(A, lanes, pair, nw, smem, sms, stream)
                      : rqs64_pick<false, false>(A, lanes, pair, nw, smem, sms, stream);
    if (rc == MF_OK && want_sum && !whole) {
        rqs64_event_sum_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(ladj_rows, B, F, ladj_sum);
        rc = launch_status();
    }
    if (ladj_rows != ladj) cudaFreeAsync(ladj_rows, stream);
    return rc;
}

}  // namespace mf
