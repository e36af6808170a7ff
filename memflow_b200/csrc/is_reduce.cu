// is_reduce.cu -- K5: importance-sampling weight + MC estimator accumulators.
// Replaces notebooks/TestFlows_ImportanceSampling.ipynb cells [50-54] (mask of cells [35,43]):
//   w = target * rambo_jac / exp(logq);   acc += {sum w, sum w^2, n_valid}
// Reads 24 B/event (16 B with target == NULL), coalesced; fixed persistent grid -> deterministic.
#include "reduce.cuh"

namespace mf {

constexpr int kIsThreads = 256;

__global__ void __launch_bounds__(kIsThreads)
is_reduce_kernel(const double* __restrict__ target, const double* __restrict__ jac, const double* __restrict__ logq,
                 int64_t N, double* __restrict__ acc, void* workspace) {
    Kahan k1, k2;
    double cnt = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kIsThreads;
    for (int64_t i = (int64_t)blockIdx.x * kIsThreads + threadIdx.x; i < N; i += stride) {
        const double t = target ? __ldg(target + i) : 1.0;
        const double j = __ldg(jac + i);
        const double w = t * j / exp(__ldg(logq + i));
        // mask of notebook cells [35,43] (target > 0, Jacobian not NaN) and a FINITE weight: one NaN / inf log q must not
        // poison the accumulators of a whole sharded run (n_valid would still look healthy)
        if (t > 0.0 && j == j && isfinite(w)) {
            k1.add(w);
            k2.add(w * w);
            cnt += 1.0;
        }
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
    int64_t want = (N + mf::kIsThreads - 1) / mf::kIsThreads;
    int blocks = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    if (blocks > mf::kReduceMaxBlocks) blocks = mf::kReduceMaxBlocks;
    mf::is_reduce_kernel<<<blocks, mf::kIsThreads, 0, (cudaStream_t)stream>>>(target, jac, logq, N, acc, workspace);
    return mf::launch_status();
}
