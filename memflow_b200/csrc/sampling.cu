// sampling.cu -- the sampling branch of the reference's training loss in ONE forward and ONE backward launch
// (memflow/unfolding_flow/unfolding_flow_withDecay_v2.py:376-417):
//   ps [N, 3n-2]  ->  rambo.get_momenta_from_ps(ps, requires_grad=True)                       K1, CM frame     (:377)
//                 ->  boost_Epz = (E (x1+x2)/2, 0, 0, E (x1-x2)/2)                                              (:380)
//                 ->  boost_tt(momenta[:, -4:-1], boostVector_t(boost_Epz))                     lab frame        (:384-387)
//                 ->  get_ptetaphi_comp_batch -> log(pt + 1) -> (x - mean) / std                scaled (pt,eta,phi) (:390-393)
//                 ->  boost_scaled = ((log(E_b + 1), pz_b) - mean) / std                                         (:396-398)
//                 ->  get_decayPartons_fromlab_propagators_angles(...)                          12 partons       (:402-417)
// The reference runs this as K1's ~250 launches plus ~60 elementwise kernels plus the decay chain; round 1 of this repo ran
// K1, ~20 torch glue launches and the fused decay kernel. Here the propagators never leave the registers between the
// phase-space map and the decay chain. The backward kernel composes the three vector-Jacobian products in reverse: decay
// chain (dual numbers, decay_bwd.cuh) -> glue (the same template on dual numbers carrying the 4 + 2 tangents of one
// propagator and the boost) -> K1 (rambo_bwd.cuh, with the reference graph's semantics: no gradient through the mass
// variables). Saved between the two launches: the CM momenta, x1, x2 (what the reference's graph holds as well) and the
// scaled propagators.
#include "decay_bwd.cuh"
#include "rambo_bwd.cuh"

namespace mf {

template <int NF>
__global__ void __launch_bounds__(128)
sampling_partons_kernel(const double* __restrict__ ps, int64_t N, const __grid_constant__ RamboParams<double, NF> P,
                        const double* __restrict__ higgs_angles, const double* __restrict__ thad_b_angles,
                        const double* __restrict__ thad_w_angles, const double* __restrict__ tlep_b_angles,
                        const double* __restrict__ tlep_w_angles, const __grid_constant__ mf_decay_partons_cfg C,
                        double* __restrict__ out, double* __restrict__ boost_scaled, double* __restrict__ momenta,
                        double* __restrict__ weight, double* __restrict__ x1o, double* __restrict__ x2o,
                        double* __restrict__ prop, int32_t* __restrict__ status) {
    constexpr int W = 3 * NF - 2, NP = NF + 2;
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double r[W];
    load_row<W>(ps + ev * W, r);
    bool bad = false;
#pragma unroll
    for (int i = 0; i < W; ++i) bad |= (r[i] != r[i]);
    if (status != nullptr && bad) atomicOr(status, MF_ST_NAN_INPUT);
    double mom[NP][4], wgt, x1, x2;
    rambo_forward_event<double, NF>(r, P, mom, wgt, x1, x2);
    double* mo = momenta + ev * (NP * 4);
#pragma unroll
    for (int k = 0; k < NP; ++k) st4(mo + 4 * k, mom[k][0], mom[k][1], mom[k][2], mom[k][3]);
    weight[ev] = wgt;
    x1o[ev] = x1;
    x2o[ev] = x2;
    const double bE = P.e_collider * (x1 + x2) / 2.0, bpz = P.e_collider * (x1 - x2) / 2.0;   // :380
    const double bs0 = (log(bE + 1.0) - C.mean_boost[0]) / C.std_boost[0], bs1 = (bpz - C.mean_boost[1]) / C.std_boost[1];
    st2(boost_scaled + 2 * ev, bs0, bs1);
    double k[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // momenta[:, -4:-1]: the three final-state rows before the last one
        double ki[3];
        lab_scaled_propagator(mom[NP - 4 + i], bE, bpz, C, ki);
        k[3 * i] = ki[0]; k[3 * i + 1] = ki[1]; k[3 * i + 2] = ki[2];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) prop[ev * 9 + i] = k[i];
    decay_partons_event(k, higgs_angles + 2 * ev, thad_b_angles + 2 * ev, thad_w_angles + 2 * ev, tlep_b_angles + 2 * ev,
                        tlep_w_angles + 2 * ev, bs1, C, out + ev * 48);
}

template <int NF>
__global__ void __launch_bounds__(128)
sampling_partons_bwd_kernel(const double* __restrict__ ps, int64_t N, const __grid_constant__ RamboParams<double, NF> P,
                            const double* __restrict__ higgs_angles, const double* __restrict__ thad_b_angles,
                            const double* __restrict__ thad_w_angles, const double* __restrict__ tlep_b_angles,
                            const double* __restrict__ tlep_w_angles, const __grid_constant__ mf_decay_partons_cfg C,
                            const double* __restrict__ momenta, const double* __restrict__ weight,
                            const double* __restrict__ x1a, const double* __restrict__ x2a, const double* __restrict__ prop,
                            const double* __restrict__ boost_scaled, const double* __restrict__ g_out,
                            const double* __restrict__ g_boost_scaled, double* __restrict__ g_ps, double* __restrict__ g_higgs,
                            double* __restrict__ g_thad_b, double* __restrict__ g_thad_w, double* __restrict__ g_tlep_b,
                            double* __restrict__ g_tlep_w) {
    constexpr int W = 3 * NF - 2, NP = NF + 2;
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    // ---- decay chain: adjoints of the scaled propagators and of the scaled boost pz ----
    double gk[9], gb1 = 0.0;
    decay_partons_bwd_event(prop + ev * 9, higgs_angles + 2 * ev, thad_b_angles + 2 * ev, thad_w_angles + 2 * ev,
                            tlep_b_angles + 2 * ev, tlep_w_angles + 2 * ev, boost_scaled[2 * ev + 1], C, g_out + ev * 48, gk,
                            g_higgs + 2 * ev, g_thad_b + 2 * ev, g_thad_w + 2 * ev, g_tlep_b + 2 * ev, g_tlep_w + 2 * ev, gb1);
    double gbs0 = 0.0, gbs1 = gb1;
    if (g_boost_scaled != nullptr) { gbs0 += g_boost_scaled[2 * ev]; gbs1 += g_boost_scaled[2 * ev + 1]; }
    const double x1 = x1a[ev], x2 = x2a[ev];
    const double bE = P.e_collider * (x1 + x2) / 2.0, bpz = P.e_collider * (x1 - x2) / 2.0;
    double g_bE = gbs0 / C.std_boost[0] / (bE + 1.0), g_bpz = gbs1 / C.std_boost[1];
    // ---- glue: one propagator at a time on dual numbers (tangents: E, px, py, pz of the CM propagator, boost E, boost pz) ----
    const double* mo = momenta + ev * (NP * 4);
    double gm[NP * 4];
#pragma unroll
    for (int i = 0; i < NP * 4; ++i) gm[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        using D = Dual<6>;
        double p[4];
        ld4(mo + 4 * (NP - 4 + i), p[0], p[1], p[2], p[3]);
        const D pd[4] = {seed<6>(p[0], 0), seed<6>(p[1], 1), seed<6>(p[2], 2), seed<6>(p[3], 3)};
        D kd[3];
        lab_scaled_propagator(pd, seed<6>(bE, 4), seed<6>(bpz, 5), C, kd);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            gm[4 * (NP - 4 + i) + c] = (gk[3 * i] * kd[0].d[c] + gk[3 * i + 1] * kd[1].d[c]) + gk[3 * i + 2] * kd[2].d[c];
        g_bE += (gk[3 * i] * kd[0].d[4] + gk[3 * i + 1] * kd[1].d[4]) + gk[3 * i + 2] * kd[2].d[4];
        g_bpz += (gk[3 * i] * kd[0].d[5] + gk[3 * i + 1] * kd[1].d[5]) + gk[3 * i + 2] * kd[2].d[5];
    }
    const double gx1 = 0.5 * P.e_collider * (g_bE + g_bpz), gx2 = 0.5 * P.e_collider * (g_bE - g_bpz);
    // ---- K1 ----
    double r[W], g[W];
    load_row<W>(ps + ev * W, r);
    rambo_forward_bwd_event<NF>(r, mo, weight[ev], x1, x2, gm, 0.0, gx1, gx2, P, g);
    store_row<W>(g_ps + ev * W, g);
}

}  // namespace mf

extern "C" int mf_sampling_partons_f64(const double* ps, int64_t N, int n_final, const double* masses, double e_collider,
                                       const double* higgs_angles, const double* thad_b_angles, const double* thad_w_angles,
                                       const double* tlep_b_angles, const double* tlep_w_angles,
                                       const mf_decay_partons_cfg* cfg, double* out, double* boost_scaled, double* momenta,
                                       double* weight, double* x1, double* x2, double* propagators, int32_t* status,
                                       mf_stream_t stream) {
    if (N < 0 || !cfg || !masses) return MF_E_BADARG;
    if (n_final < 4 || n_final > MF_MAX_FINAL) return MF_E_NFINAL;   // momenta[:, -4:-1] needs three rows before the last
    const void* ptrs[] = {ps, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, out, boost_scaled,
                          momenta, weight, x1, x2, propagators};
    for (const void* p : ptrs)
        if (N > 0 && !p) return MF_E_BADARG;
    if (cfg->final_scaling && !cfg->ptetaphi) return MF_E_BADARG;
    const int W = 3 * n_final - 2;
    const size_t row_align = (W % 4 == 0) ? 32 : ((W % 2 == 0) ? 16 : 8);
    if (!mf::aligned(ps, row_align) || !mf::aligned(out, 32) || !mf::aligned(momenta, 32) || !mf::aligned(boost_scaled, 16))
        return MF_E_ALIGN;
    for (int i = 1; i < 6; ++i)
        if (!mf::aligned(ptrs[i], 16)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    const unsigned blocks = (unsigned)((N + 127) / 128);
    switch (n_final) {
#define MF_CASE(NF)                                                                                                       \
    case NF: {                                                                                                            \
        const mf::RamboParams<double, NF> P = mf::make_forward_params<double, NF>(masses, e_collider, nullptr, 0u);       \
        mf::sampling_partons_kernel<NF><<<blocks, 128, 0, (cudaStream_t)stream>>>(                                       \
            ps, N, P, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, *cfg, out, boost_scaled,  \
            momenta, weight, x1, x2, propagators, status);                                                                \
        break;                                                                                                            \
    }
        MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return mf::launch_status();
}

extern "C" int mf_sampling_partons_bwd_f64(const double* ps, int64_t N, int n_final, const double* masses, double e_collider,
                                           const double* higgs_angles, const double* thad_b_angles,
                                           const double* thad_w_angles, const double* tlep_b_angles,
                                           const double* tlep_w_angles, const mf_decay_partons_cfg* cfg,
                                           const double* momenta, const double* weight, const double* x1, const double* x2,
                                           const double* propagators, const double* boost_scaled, const double* g_out,
                                           const double* g_boost_scaled, double* g_ps, double* g_higgs_angles,
                                           double* g_thad_b_angles, double* g_thad_w_angles, double* g_tlep_b_angles,
                                           double* g_tlep_w_angles, mf_stream_t stream) {
    if (N < 0 || !cfg || !masses) return MF_E_BADARG;
    if (n_final < 4 || n_final > MF_MAX_FINAL) return MF_E_NFINAL;
    const void* ptrs[] = {ps, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, momenta, weight, x1, x2,
                          propagators, boost_scaled, g_out, g_ps, g_higgs_angles, g_thad_b_angles, g_thad_w_angles,
                          g_tlep_b_angles, g_tlep_w_angles};
    for (const void* p : ptrs)
        if (N > 0 && !p) return MF_E_BADARG;
    if (cfg->final_scaling && !cfg->ptetaphi) return MF_E_BADARG;
    const int W = 3 * n_final - 2;
    const size_t row_align = (W % 4 == 0) ? 32 : ((W % 2 == 0) ? 16 : 8);
    if (!mf::aligned(ps, row_align) || !mf::aligned(g_ps, row_align) || !mf::aligned(momenta, 32) || !mf::aligned(g_out, 32))
        return MF_E_ALIGN;
    for (int i : {1, 2, 3, 4, 5, 14, 15, 16, 17, 18})
        if (!mf::aligned(ptrs[i], 16)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    const unsigned blocks = (unsigned)((N + 127) / 128);
    switch (n_final) {
#define MF_CASE(NF)                                                                                                          \
    case NF: {                                                                                                               \
        const mf::RamboParams<double, NF> P = mf::make_forward_params<double, NF>(masses, e_collider, nullptr, 0u);          \
        mf::sampling_partons_bwd_kernel<NF><<<blocks, 128, 0, (cudaStream_t)stream>>>(                                      \
            ps, N, P, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, *cfg, momenta, weight, x1,   \
            x2, propagators, boost_scaled, g_out, g_boost_scaled, g_ps, g_higgs_angles, g_thad_b_angles, g_thad_w_angles,    \
            g_tlep_b_angles, g_tlep_w_angles);                                                                               \
        break;                                                                                                               \
    }
        MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return mf::launch_status();
}
