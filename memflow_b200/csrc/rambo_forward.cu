// rambo_forward.cu -- K1: unit hypercube -> CM-frame momenta + Jacobian weight, one event per thread.
// Replaces FlatInvertiblePhasespace.generateKinematics_batch (memflow/phasespace/rambo_generator.py:168-398),
// which runs ~250-600 torch elementwise launches plus a 180-level batch bisection per call.
//
// HBM layout: the reference's AoS tensors are kept as they are.
//   in : u[N][3n-2]            read once, row per thread, widest vector the row pitch allows; the rows of
//                              a warp are contiguous so every DRAM sector is fetched once (L1 absorbs
//                              the per-thread stride).
//   out: momenta[N][n+2][4]    one 4-momentum = 32 B (fp64) = one DRAM sector = one 256-bit STG.
//        weight/x1/x2/logw[N]  coalesced scalar stores.
// Algorithmic bytes per event: (3n-2 + 4(n+2) + 3) * sizeof(T) (+ sizeof(T) with logweight).
#include "rambo.cuh"
#include "rqs.cuh"  // mbarrier / TMA bulk-copy helpers

namespace mf {

// The event again on the guarded primitives (common.cuh, PrimIEEE), out of line: taken only when the speculative pass below
// stored a non-finite value (an operand of a reciprocal / square root was 0, denormal, inf or NaN: exact thresholds,
// cos(theta) = +-1, garbage input). `src` is the event's row of uniforms (generic pointer: shared or global memory).
// A NaN among the inputs always reaches the outputs, so the NaN-input status bit is raised here as well and the fast
// path carries no per-input test.
template <typename T, int NF>
__device__ __noinline__ void rambo_forward_redo(const RamboParams<T, NF>& P, const T* src, T* mo, T* wx, int32_t* status) {
    T wgt, x1, x2;
    bool bad = false;
    rambo_forward_stream<T, NF, PrimIEEE>(
        P, [&](int i) { const T v = src[i]; bad |= (v != v); return v; },
        [&](int j, const T (&p)[4]) { st4(mo + 4 * j, p[0], p[1], p[2], p[3]); }, wgt, x1, x2);
    if (status != nullptr && bad) atomicOr(status, MF_ST_NAN_INPUT);
    wx[0] = wgt; wx[1] = x1; wx[2] = x2;
}

#ifndef MF_K1_MINB
#define MF_K1_MINB (NF <= 4 ? 4 : 3)
#endif
// EPILOGUE == false: plain CM-frame output, every 4-momentum is stored the moment it is final (streaming core).
// EPILOGUE == true : lab-frame output and / or cuts need the whole event first (array interface).
template <typename T, int NF, bool EPILOGUE>
__global__ void __launch_bounds__(256, (EPILOGUE ? (NF <= 4 ? 3 : 2) : MF_K1_MINB))
rambo_forward_kernel(const T* __restrict__ u, int64_t N, const __grid_constant__ RamboParams<T, NF> P,
                     T* __restrict__ momenta, T* __restrict__ weight, T* __restrict__ logweight,
                     T* __restrict__ x1o, T* __restrict__ x2o, int32_t* __restrict__ status) {
    constexpr int W = 3 * NF - 2, NP = NF + 2;
    // n >= 6, streaming core: the CTA's 256 input rows (contiguous) come in with one TMA bulk copy, so the uniforms the
    // dependency chains wait for are shared-memory reads instead of global loads issued one by one along the chain
    constexpr bool PREFETCH = !EPILOGUE && NF >= 6 && sizeof(T) == 8;  // n = 6: +15 %, n = 7, 8: +4-7 %, n = 5: neutral
    // row pitch in shared memory: W = 16 doubles (n = 6) would put every row of a warp in the same banks (32-way conflicts on
    // each read); those rows are spaced 18 doubles apart and filled by 16-byte asynchronous copies
    constexpr int WP = (W % 16 == 0) ? W + 2 : W;
    __shared__ __align__(128) T su[PREFETCH ? 256 * WP : 1];
    __shared__ uint64_t bar;
    if constexpr (PREFETCH) {
        const int64_t ev0 = (int64_t)blockIdx.x * blockDim.x;
        const int nrows = (int)((N - ev0) < (int64_t)blockDim.x ? (N - ev0) : (int64_t)blockDim.x);
        const T* src = u + ev0 * W;
        const uint32_t bytes = (uint32_t)nrows * W * (uint32_t)sizeof(T);
        const bool bulk = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (bytes & 15) == 0;
        if (bulk && WP != W) {
            // padded rows (n = 6): 16-byte asynchronous copies (LDGSTS), coalesced over the CTA's contiguous rows, each
            // landing at its padded place. (One bulk copy per row, 256 per CTA, was what made n = 6 the slowest map per byte.)
            constexpr int CPR = W * (int)sizeof(T) / 16;   // 16-byte chunks per row
            for (int c = threadIdx.x; c < nrows * CPR; c += blockDim.x) {
                const int row = c / CPR, ch = c - row * CPR;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(su + row * WP) + 16u * ch;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const char*>(src) + 16 * (size_t)c) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        } else if (bulk) {
            if (threadIdx.x == 0) {
                mbar_init(&bar, 1);
                fence_mbar_init();
                mbar_expect_tx(&bar, bytes);
                bulk_g2s(su, src, bytes, &bar);
            }
            __syncthreads();
            mbar_wait(&bar, 0u);
        } else {
            for (int i = threadIdx.x; i < nrows * W; i += blockDim.x) su[(i / W) * WP + (i % W)] = src[i];
            __syncthreads();
        }
    }
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    const T* __restrict__ urow = u + ev * W;
    T* mo = momenta + ev * (NP * 4);
    T wgt, x1, x2;
    bool bad = false;
    if constexpr (!EPILOGUE) {
        // speculative pass on the unguarded primitives; every stored energy depends on all components of its row, the
        // last row on the whole chain: one exponent test per row (+ the weight) sees any operand that left their domain
        const T* __restrict__ srow = su + threadIdx.x * WP;
        bool redo = false;
        // (Fetching the first decay's five uniforms with plain loads and waiting for the tile's bulk copy only before the
        // second decay was tried: 2-9 % slower at n = 6, 7, 8 on the same box.)
        rambo_forward_stream<T, NF, PrimRaw>(
            P, [&](int i) { return PREFETCH ? srow[i] : ld1(urow + i); },
            [&](int j, const T (&p)[4]) { redo |= not_finite_bits(p[0]); st4(mo + 4 * j, p[0], p[1], p[2], p[3]); }, wgt, x1, x2, redo);
        if (redo | not_finite_bits(wgt)) {
            T wx[3];
            rambo_forward_redo<T, NF>(P, PREFETCH ? srow : urow, mo, wx, status);
            wgt = wx[0]; x1 = wx[1]; x2 = wx[2];
        }
    } else {
        T r[W];
        load_row<W>(urow, r);
#pragma unroll
        for (int i = 0; i < W; ++i) bad |= (r[i] != r[i]);
        T out[NP][4];
        rambo_forward_event<T, NF>(r, P, out, wgt, x1, x2);
#pragma unroll
        for (int k = 0; k < NP; ++k) st4(mo + 4 * k, out[k][0], out[k][1], out[k][2], out[k][3]);
    }
    if (status != nullptr && bad) atomicOr(status, MF_ST_NAN_INPUT);
    weight[ev] = wgt;
    x1o[ev] = x1;
    x2o[ev] = x2;
    if (logweight != nullptr) logweight[ev] = log_(wgt);
}

template <typename T, int NF>
static int launch_forward(const T* u, int64_t N, const T* masses, T e_collider, const T* cuts, uint32_t flags,
                          T* momenta, T* weight, T* logweight, T* x1, T* x2, int32_t* status, cudaStream_t stream) {
    const RamboParams<T, NF> P = make_forward_params<T, NF>(masses, e_collider, cuts, flags);
    if (N == 0) return MF_OK;
    const int threads = 256;
    const int64_t blocks = (N + threads - 1) / threads;
    if (P.flags & (MF_FLAG_CUTS | MF_FLAG_LAB_FRAME))
        rambo_forward_kernel<T, NF, true><<<(unsigned)blocks, threads, 0, stream>>>(u, N, P, momenta, weight, logweight,
                                                                                    x1, x2, status);
    else
        rambo_forward_kernel<T, NF, false><<<(unsigned)blocks, threads, 0, stream>>>(u, N, P, momenta, weight, logweight,
                                                                                     x1, x2, status);
    return launch_status();
}

template <typename T>
static int forward_dispatch(const T* u, int64_t N, int n, const T* masses, T e_collider, const T* cuts,
                            uint32_t flags, T* momenta, T* weight, T* logweight, T* x1, T* x2, int32_t* status,
                            cudaStream_t stream) {
    if (N < 0 || masses == nullptr) return MF_E_BADARG;
    if (n < MF_MIN_FINAL || n > MF_MAX_FINAL) return MF_E_NFINAL;
    if (N > 0 && (!u || !momenta || !weight || !x1 || !x2)) return MF_E_BADARG;
    if ((flags & MF_FLAG_CUTS) && cuts == nullptr) return MF_E_BADARG;
    const int W = 3 * n - 2;
    const size_t row_align = (W % 4 == 0) ? 4 * sizeof(T) : ((W % 2 == 0) ? 2 * sizeof(T) : sizeof(T));
    if (!aligned(u, row_align) || !aligned(momenta, 4 * sizeof(T))) return MF_E_ALIGN;
    if (N > (int64_t)0x7fffffff * 256) return MF_E_BADARG;
    if ((flags & MF_FLAG_ZERO_STATUS) && status != nullptr) cudaMemsetAsync(status, 0, sizeof(int32_t), stream);
    switch (n) {
#define MF_CASE(NF) \
    case NF: return launch_forward<T, NF>(u, N, masses, e_collider, cuts, flags, momenta, weight, logweight, x1, x2, status, stream);
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return MF_E_NFINAL;
}

template <int NF>
__global__ void __launch_bounds__(256) massvar_kernel(const double* __restrict__ v, int64_t N, double* __restrict__ u) {
    constexpr int NE = NF - 2;
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    [&]<int... I>(std::integer_sequence<int, I...>) {
        ((u[ev * NE + I] = solve_massvar<NF - 2 - I>(v[ev * NE + I])), ...);
    }(std::make_integer_sequence<int, NE>{});
}

}  // namespace mf

extern "C" int mf_massvar_solve_f64(const double* v, int64_t N, int n_final, double* u, mf_stream_t stream) {
    if (N < 0 || (N > 0 && (!v || !u))) return MF_E_BADARG;
    if (n_final < MF_MIN_FINAL || n_final > MF_MAX_FINAL) return MF_E_NFINAL;
    if (N == 0) return MF_OK;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    switch (n_final) {
#define MF_CASE(NF) case NF: mf::massvar_kernel<NF><<<blocks, 256, 0, (cudaStream_t)stream>>>(v, N, u); break;
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return mf::launch_status();
}

extern "C" int mf_rambo_forward_f64(const double* u, int64_t N, int n_final, const double* masses, double e_collider,
                                    const double* cuts, uint32_t flags, double* momenta, double* weight,
                                    double* logweight, double* x1, double* x2, int32_t* status, mf_stream_t stream) {
    return mf::forward_dispatch<double>(u, N, n_final, masses, e_collider, cuts, flags, momenta, weight, logweight,
                                        x1, x2, status, (cudaStream_t)stream);
}
extern "C" int mf_rambo_forward_f32(const float* u, int64_t N, int n_final, const float* masses, float e_collider,
                                    const float* cuts, uint32_t flags, float* momenta, float* weight,
                                    float* logweight, float* x1, float* x2, int32_t* status, mf_stream_t stream) {
    return mf::forward_dispatch<float>(u, N, n_final, masses, e_collider, cuts, flags, momenta, weight, logweight,
                                       x1, x2, status, (cudaStream_t)stream);
}
