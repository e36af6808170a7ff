// decay_backward.cu -- vector-Jacobian product of mf_decay_partons_f64 (decay.cu): gradients of
// sum(out * g_out) w.r.t. the scaled propagators, the five decay-angle pairs and the scaled boost, i.e. what autograd
// gives through Compute_ParticlesTensor.get_decayPartons_fromlab_propagators_angles
// (memflow/unfolding_flow/utils.py:826-935) in the sampling loss of unfolding_flow_withDecay_v2.py:376-417.
//
// The event factorises: rows 0-2 depend on (propagator 0, higgs angles) = 5 inputs, rows 3-6 on (propagator 2, thad
// angles) = 7, rows 7-10 on (propagator 1, tlep angles) = 7, and the ISR row on the three propagators' momenta and the
// boost pz only through the 3-vector g = boost - partons. Each block is re-evaluated once on dual numbers carrying
// its 5 / 7 input tangents (decay.cuh: the same templates as the forward kernel, so both stay in step), and each
// finished row's tangents are contracted with its incoming gradient on the fly. The ISR row is differentiated first
// w.r.t. g (3 tangents); its adjoint enters every block through the propagator's (px, py, pz) tangents.
#include "decay_bwd.cuh"

namespace mf {

__global__ void __launch_bounds__(128)
decay_partons_bwd_kernel(const double* __restrict__ prop, const double* __restrict__ higgs_angles,
                         const double* __restrict__ thad_b_angles, const double* __restrict__ thad_w_angles,
                         const double* __restrict__ tlep_b_angles, const double* __restrict__ tlep_w_angles,
                         const double* __restrict__ boost, int64_t N, const __grid_constant__ mf_decay_partons_cfg C,
                         const double* __restrict__ g_out, double* __restrict__ g_prop, double* __restrict__ g_higgs,
                         double* __restrict__ g_thad_b, double* __restrict__ g_thad_w, double* __restrict__ g_tlep_b,
                         double* __restrict__ g_tlep_w, double* __restrict__ g_boost) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double gb = 0.0;
    decay_partons_bwd_event(prop + ev * 9, higgs_angles + 2 * ev, thad_b_angles + 2 * ev, thad_w_angles + 2 * ev,
                            tlep_b_angles + 2 * ev, tlep_w_angles + 2 * ev, boost[2 * ev + 1], C, g_out + ev * 48,
                            g_prop + ev * 9, g_higgs + 2 * ev, g_thad_b + 2 * ev, g_thad_w + 2 * ev, g_tlep_b + 2 * ev,
                            g_tlep_w + 2 * ev, gb);
    st2(g_boost + 2 * ev, 0.0, gb);
}

}  // namespace mf

extern "C" int mf_decay_partons_bwd_f64(const double* propagators, const double* higgs_angles, const double* thad_b_angles,
                                        const double* thad_w_angles, const double* tlep_b_angles,
                                        const double* tlep_w_angles, const double* boost, int64_t N,
                                        const mf_decay_partons_cfg* cfg, const double* g_out, double* g_propagators,
                                        double* g_higgs_angles, double* g_thad_b_angles, double* g_thad_w_angles,
                                        double* g_tlep_b_angles, double* g_tlep_w_angles, double* g_boost,
                                        mf_stream_t stream) {
    if (N < 0 || !cfg) return MF_E_BADARG;
    const void* ptrs[] = {propagators, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, boost, g_out,
                          g_propagators, g_higgs_angles, g_thad_b_angles, g_thad_w_angles, g_tlep_b_angles, g_tlep_w_angles,
                          g_boost};
    for (const void* p : ptrs)
        if (N > 0 && !p) return MF_E_BADARG;
    if (cfg->final_scaling && !cfg->ptetaphi) return MF_E_BADARG;
    for (int i = 1; i < 15; ++i)
        if (i != 8 && !mf::aligned(ptrs[i], i == 7 ? 32 : 16)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    mf::decay_partons_bwd_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        propagators, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, boost, N, *cfg, g_out,
        g_propagators, g_higgs_angles, g_thad_b_angles, g_thad_w_angles, g_tlep_b_angles, g_tlep_w_angles, g_boost);
    return mf::launch_status();
}
