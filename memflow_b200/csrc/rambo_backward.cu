// rambo_backward.cu -- vector-Jacobian product of K1, with the gradient semantics of the reference's
// autograd graph (generateKinematics_batch(..., requires_grad=True), memflow/phasespace/rambo_generator.py:168-398):
//   * the n-2 mass-variable columns get NO gradient (the bisection output is built from constants, :411-433) and the
//     intermediate masses do not depend on x1, x2 in the graph (E_cm is written into M under torch.no_grad(), :282-283);
//   * the angle columns get the true derivative of every final-state momentum through the chain of two-body decays and
//     boosts (:315-345, utils.py:88-118);
//   * the two x1/x2 columns reach only x1, x2, the weight (x1x2 Jacobian and the E_cm^(2n-4) volume) and the incoming
//     rows (:223-228, :112-140, :511-551).
// One event per thread. The reverse sweep rebuilds Q_k = sum_{j>=k} p_j from the saved forward momenta, so only the
// masses and decay momenta q_k are recomputed.
#include "rambo_bwd.cuh"

namespace mf {

template <int NF>
__global__ void __launch_bounds__(256, 2)
rambo_forward_bwd_kernel(const double* __restrict__ u, const double* __restrict__ momenta, const double* __restrict__ weight,
                         const double* __restrict__ x1a, const double* __restrict__ x2a, const double* __restrict__ g_mom,
                         const double* __restrict__ g_w, const double* __restrict__ g_x1, const double* __restrict__ g_x2,
                         int64_t N, const __grid_constant__ RamboParams<double, NF> P, double* __restrict__ grad_u) {
    constexpr int W = 3 * NF - 2, NP = NF + 2;
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double r[W], g[W];
    load_row<W>(u + ev * W, r);
    rambo_forward_bwd_event<NF>(r, momenta + ev * (NP * 4), weight[ev], x1a[ev], x2a[ev], g_mom ? g_mom + ev * (NP * 4) : nullptr,
                                g_w ? g_w[ev] : 0.0, g_x1 ? g_x1[ev] : 0.0, g_x2 ? g_x2[ev] : 0.0, P, g);
    store_row<W>(grad_u + ev * W, g);
}

}  // namespace mf

extern "C" int mf_rambo_forward_bwd_f64(const double* u, const double* momenta, const double* weight, const double* x1,
                                        const double* x2, const double* g_momenta, const double* g_weight,
                                        const double* g_x1, const double* g_x2, int64_t N, int n_final, const double* masses,
                                        double e_collider, uint32_t flags, double* grad_u, mf_stream_t stream) {
    if (N < 0 || !masses || (N > 0 && (!u || !momenta || !weight || !x1 || !x2 || !grad_u))) return MF_E_BADARG;
    if (n_final < MF_MIN_FINAL || n_final > MF_MAX_FINAL) return MF_E_NFINAL;
    if (flags & (MF_FLAG_LAB_FRAME | MF_FLAG_CUTS)) return MF_E_BADARG;  // no reference gradient exists for these outputs
    const int W = 3 * n_final - 2;
    const size_t row_align = (W % 4 == 0) ? 32 : ((W % 2 == 0) ? 16 : 8);
    if (!mf::aligned(u, row_align) || !mf::aligned(grad_u, row_align) || !mf::aligned(momenta, 32)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    switch (n_final) {
#define MF_CASE(NF)                                                                                                          \
    case NF: {                                                                                                               \
        const mf::RamboParams<double, NF> P = mf::make_forward_params<double, NF>(masses, e_collider, nullptr, flags);       \
        mf::rambo_forward_bwd_kernel<NF><<<blocks, 256, 0, (cudaStream_t)stream>>>(u, momenta, weight, x1, x2, g_momenta,    \
                                                                                   g_weight, g_x1, g_x2, N, P, grad_u);      \
        break;                                                                                                               \
    }
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return mf::launch_status();
}
