// reduce.cuh -- deterministic block/grid reduction of the MC estimator accumulators
// (sum w, sum w^2, n_valid): per-thread Kahan, warp-shuffle tree, smem across warps, per-block
// partials in a workspace, the last block to finish folds the partials in index order.
#pragma once
#include "common.cuh"

namespace mf {

constexpr int kReduceMaxBlocks = 2048;
constexpr int64_t kReduceWorkspaceBytes = 64 + (int64_t)kReduceMaxBlocks * 3 * sizeof(double);

struct Kahan {
    double s = 0.0, c = 0.0;
    MF_D void add(double x) {
        double y = __dsub_rn(x, c);
        double t = __dadd_rn(s, y);
        c = __dsub_rn(__dsub_rn(t, s), y);
        s = t;
    }
    MF_D double value() const { return s - c; }
};

MF_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Every thread of the block calls this once with its partial sums. acc[0..2] += totals (done by the
// last block). workspace: [0] = uint32 ticket counter (zero on entry, left zero), +64 B: partials.
template <int THREADS>
MF_D void grid_accumulate(double w1, double w2, double cnt, double* __restrict__ acc, void* workspace) {
    __shared__ double sm[3][THREADS / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    w1 = warp_sum(w1); w2 = warp_sum(w2); cnt = warp_sum(cnt);
    if (lane == 0) { sm[0][warp] = w1; sm[1][warp] = w2; sm[2][warp] = cnt; }
    __syncthreads();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 64);
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) { a += sm[0][i]; b += sm[1][i]; c += sm[2][i]; }
        partials[3 * blockIdx.x + 0] = a;
        partials[3 * blockIdx.x + 1] = b;
        partials[3 * blockIdx.x + 2] = c;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        Kahan k1, k2, k3;
        const volatile double* vp = partials;
        for (unsigned int i = lane; i < gridDim.x; i += 32) {
            k1.add(vp[3 * i + 0]); k2.add(vp[3 * i + 1]); k3.add(vp[3 * i + 2]);
        }
        double a = warp_sum(k1.value()), b = warp_sum(k2.value()), c = warp_sum(k3.value());
        if (lane == 0) {
            acc[0] += a; acc[1] += b; acc[2] += c;
            *ticket = 0u;  // leave the workspace ready for the next launch on this stream
        }
    }
}

}  // namespace mf
