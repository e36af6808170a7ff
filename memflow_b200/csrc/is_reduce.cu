// is_reduce.cu -- K5: importance-sampling weight + MC estimator accumulators.
// Replaces notebooks/TestFlows_ImportanceSampling.ipynb cells [50-54] (mask of cells [35,43]):
//   w = target * rambo_jac / exp(logq);   acc += {sum w, sum w^2, n_valid}
// Reads 24 B/event (16 B with target == NULL), coalesced; fixed persistent grid -> deterministic.
#include "reduce.cuh"

namespace mf {

constexpr int kIsThreads = 256;

// One event into the accumulators. Mask of notebook cells [35,43] (target > 0, Jacobian not NaN) and a FINITE weight: one
// NaN / inf log q must not poison the accumulators of a whole sharded run (n_valid would still look healthy).
MF_D void is_accumulate(double t, double j, double lq, Kahan& k1, Kahan& k2, double& cnt) {
    const double w = t * j * exp(-lq);
    if (t > 0.0 && j == j && isfinite(w)) {
        k1.add(w);
        k2.add(w * w);
        cnt += 1.0;
    }
}

// VEC: 16-byte loads, two events per thread and iteration, two iterations in flight (ncu of the scalar form: 8.9
// long-scoreboard stalls per issue, 4.35 TB/s = 67 % of the measured HBM rate).
template <bool VEC>
__global__ void __launch_bounds__(kIsThreads)
is_reduce_kernel(const double* __restrict__ target, const double* __restrict__ jac, const double* __restrict__ logq,
                 int64_t N, double* __restrict__ acc, void* workspace) {
    Kahan k1, k2;
    double cnt = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kIsThreads;
    const int64_t gtid = (int64_t)blockIdx.x * kIsThreads + threadIdx.x;
    if constexpr (VEC) {
        const int64_t npair = N >> 1;
        for (int64_t p = gtid; p < npair; p += 2 * stride) {
            const int64_t p2 = p + stride;
            const bool has2 = p2 < npair;
            double t0 = 1.0, t1 = 1.0, t2 = 1.0, t3 = 1.0, j0, j1, j2 = 0.0, j3 = 0.0, l0, l1, l2 = 0.0, l3 = 0.0;
            ld2(jac + 2 * p, j0, j1);
            ld2(logq + 2 * p, l0, l1);
            if (target) ld2(target + 2 * p, t0, t1);
            if (has2) {
                ld2(jac + 2 * p2, j2, j3);
                ld2(logq + 2 * p2, l2, l3);
                if (target) ld2(target + 2 * p2, t2, t3);
            }
            is_accumulate(t0, j0, l0, k1, k2, cnt);
            is_accumulate(t1, j1, l1, k1, k2, cnt);
            if (has2) {
                is_accumulate(t2, j2, l2, k1, k2, cnt);
                is_accumulate(t3, j3, l3, k1, k2, cnt);
            }
        }
        if ((N & 1) && gtid == 0) is_accumulate(target ? target[N - 1] : 1.0, jac[N - 1], logq[N - 1], k1, k2, cnt);
    } else {
        for (int64_t i = gtid; i < N; i += stride)
            is_accumulate(target ? __ldg(target + i) : 1.0, __ldg(jac + i), __ldg(logq + i), k1, k2, cnt);
    }
    grid_accumulate<kIsThreads>(k1.value(), k2.value(), cnt, acc, workspace);
}

}  // namespace mf

extern "C" int64_t mf_is_reduce_workspace_bytes(void) { return mf::kReduceWorkspaceBytes; }

extern "C" int mf_is_reduce_f64(const double* target, const double* jac, const double* logq, int64_t N, double* acc,
                                void* workspace, mf_stream_t stream) {
    if (N < 0 || !acc || !workspace || (N > 0 && (!jac || !logq))) return MF_E_BADARG;
    if (N == 0) return MF_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (N / 2 + mf::kIsThreads - 1) / mf::kIsThreads + 1;
    int blocks = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    if (blocks > mf::kReduceMaxBlocks) blocks = mf::kReduceMaxBlocks;
    const bool vec = mf::aligned(jac, 16) && mf::aligned(logq, 16) && (!target || mf::aligned(target, 16));
    if (vec) mf::is_reduce_kernel<true><<<blocks, mf::kIsThreads, 0, (cudaStream_t)stream>>>(target, jac, logq, N, acc, workspace);
    else mf::is_reduce_kernel<false><<<blocks, mf::kIsThreads, 0, (cudaStream_t)stream>>>(target, jac, logq, N, acc, workspace);
    return mf::launch_status();
}
