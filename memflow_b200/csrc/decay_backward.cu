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
#include "decay.cuh"

namespace mf {

template <int NT>
MF_D void contract_row(const Dual<NT> (&p)[4], bool is_prop, bool pad, const mf_decay_partons_cfg& C,
                       const double* __restrict__ g_row, double (&acc)[NT]) {
    Dual<NT> r[4];
    finish_row(p, is_prop, pad, C, r);
    double g[4];
    ld4(g_row, g[0], g[1], g[2], g[3]);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j] += g[c] * r[c].d[j];
}

template <int NT>
MF_D Dual<NT> seed(double v, int j) {
    Dual<NT> r(v);
    r.d[j] = 1.0;
    return r;
}

// one decay block: TOP = false -> H -> bb (rows row0..row0+2), TOP = true -> t -> b q q' (rows row0..row0+3)
template <bool TOP>
MF_D void block_backward(const double* __restrict__ k_in, const double* __restrict__ ang_b, const double* __restrict__ ang_w,
                         double mass, double w_mass, const mf_decay_partons_cfg& C, const double* __restrict__ g_rows,
                         const double (&gl)[3], double* __restrict__ g_k, double* __restrict__ g_b, double* __restrict__ g_w) {
    constexpr int NT = TOP ? 7 : 5;
    using D = Dual<NT>;
    const D k[3] = {seed<NT>(k_in[0], 0), seed<NT>(k_in[1], 1), seed<NT>(k_in[2], 2)};
    D P[4];
    propagator_cartesian(k, mass, C, P);
    double acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] = 0.0;
    // the ISR row's adjoint: g_c = (boost) - sum over propagators of P_c
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j] -= gl[c] * P[c + 1].d[j];
    double a0, a1;
    ld2(ang_b, a0, a1);
    const D eb = seed<NT>(a0, 3), pb = seed<NT>(a1, 4);
    contract_row(P, true, false, C, g_rows, acc);
    if constexpr (!TOP) {
        D l1[4], l2[4];
        higgs_decay_rows(P, eb, pb, mass, l1, l2);
        contract_row(l1, false, false, C, g_rows + 4, acc);
        contract_row(l2, false, false, C, g_rows + 8, acc);
    } else {
        double w0, w1;
        ld2(ang_w, w0, w1);
        const D ew = seed<NT>(w0, 5), pw = seed<NT>(w1, 6);
        D bl[4], q1l[4], q2l[4];
        top_decay_rows(P, eb, pb, ew, pw, top_Eb(mass, w_mass, C.b_mass), w_mass, C.b_mass, bl, q1l, q2l);
        contract_row(bl, false, false, C, g_rows + 4, acc);
        contract_row(q1l, false, false, C, g_rows + 8, acc);
        contract_row(q2l, false, false, C, g_rows + 12, acc);
        st2(g_w, acc[5], acc[6]);
    }
    g_k[0] = acc[0];
    g_k[1] = acc[1];
    g_k[2] = acc[2];
    st2(g_b, acc[3], acc[4]);
}

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
    const double mass[3] = {C.higgs_mass, C.tlep_mass, C.thad_mass};
    const double* go = g_out + ev * 48;
    // ---- ISR row: value from the plain propagators, adjoint w.r.t. its 3-vector ----
    double gl[3] = {0.0, 0.0, 0.0};
    {
        double P[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double k[3] = {prop[ev * 9 + 3 * i], prop[ev * 9 + 3 * i + 1], prop[ev * 9 + 3 * i + 2]};
            propagator_cartesian(k, mass[i], C, P[i]);
        }
        const double boost_pz = boost[2 * ev + 1] * C.std_boost[1] + C.mean_boost[1];
        using D3 = Dual<3>;
        D3 g[4];
        g[1] = seed<3>(-((P[0][1] + P[2][1]) + P[1][1]), 0);
        g[2] = seed<3>(-((P[0][2] + P[2][2]) + P[1][2]), 1);
        g[3] = seed<3>(boost_pz - ((P[0][3] + P[2][3]) + P[1][3]), 2);
        g[0] = dk::sqrt((g[1] * g[1] + g[2] * g[2]) + g[3] * g[3]) + C.eps;
        contract_row(g, false, gluon_pad(g, C), C, go + 44, gl);
    }
    st2(g_boost + 2 * ev, 0.0, gl[2] * C.std_boost[1]);  // only pz of the boost enters (:857-862, :884)
    block_backward<false>(prop + ev * 9, higgs_angles + 2 * ev, nullptr, C.higgs_mass, 0.0, C, go, gl, g_prop + ev * 9,
                          g_higgs + 2 * ev, nullptr);
    block_backward<true>(prop + ev * 9 + 6, thad_b_angles + 2 * ev, thad_w_angles + 2 * ev, C.thad_mass, C.w_had_mass, C,
                         go + 12, gl, g_prop + ev * 9 + 6, g_thad_b + 2 * ev, g_thad_w + 2 * ev);
    block_backward<true>(prop + ev * 9 + 3, tlep_b_angles + 2 * ev, tlep_w_angles + 2 * ev, C.tlep_mass, C.w_lep_mass, C,
                         go + 28, gl, g_prop + ev * 9 + 3, g_tlep_b + 2 * ev, g_tlep_w + 2 * ev);
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
