// decay_bwd.cuh -- device functions of the decay chain's vector-Jacobian product, shared by decay_backward.cu and the fused
// sampling branch (sampling.cu). See decay_backward.cu for the derivation.
#pragma once
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

// One event of the vector-Jacobian product: inputs as in decay_partons_event (prop = this event's 9 scaled components),
// go = this event's incoming gradient [12][4] (global memory); writes the angle gradients (global memory, this event's)
// and returns the adjoints of the 9 propagator components and of the scaled boost pz.
MF_D void decay_partons_bwd_event(const double* __restrict__ prop, const double* __restrict__ higgs_angles,
                                  const double* __restrict__ thad_b_angles, const double* __restrict__ thad_w_angles,
                                  const double* __restrict__ tlep_b_angles, const double* __restrict__ tlep_w_angles,
                                  double boost_pz_scaled, const mf_decay_partons_cfg& C, const double* __restrict__ go,
                                  double* __restrict__ g_prop, double* __restrict__ g_higgs, double* __restrict__ g_thad_b,
                                  double* __restrict__ g_thad_w, double* __restrict__ g_tlep_b, double* __restrict__ g_tlep_w,
                                  double& g_boost_pz_scaled) {
    const double mass[3] = {C.higgs_mass, C.tlep_mass, C.thad_mass};
    // ---- ISR row: value from the plain propagators, adjoint w.r.t. its 3-vector ----
    double gl[3] = {0.0, 0.0, 0.0};
    {
        double P[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double k[3] = {prop[3 * i], prop[3 * i + 1], prop[3 * i + 2]};
            propagator_cartesian(k, mass[i], C, P[i]);
        }
        const double boost_pz = boost_pz_scaled * C.std_boost[1] + C.mean_boost[1];
        using D3 = Dual<3>;
        D3 g[4];
        g[1] = seed<3>(-((P[0][1] + P[2][1]) + P[1][1]), 0);
        g[2] = seed<3>(-((P[0][2] + P[2][2]) + P[1][2]), 1);
        g[3] = seed<3>(boost_pz - ((P[0][3] + P[2][3]) + P[1][3]), 2);
        g[0] = dk::sqrt((g[1] * g[1] + g[2] * g[2]) + g[3] * g[3]) + C.eps;
        contract_row(g, false, gluon_pad(g, C), C, go + 44, gl);
    }
    g_boost_pz_scaled = gl[2] * C.std_boost[1];  // only pz of the boost enters (:857-862, :884)
    block_backward<false>(prop, higgs_angles, nullptr, C.higgs_mass, 0.0, C, go, gl, g_prop, g_higgs, nullptr);
    block_backward<true>(prop + 6, thad_b_angles, thad_w_angles, C.thad_mass, C.w_had_mass, C, go + 12, gl, g_prop + 6, g_thad_b,
                         g_thad_w);
    block_backward<true>(prop + 3, tlep_b_angles, tlep_w_angles, C.tlep_mass, C.w_lep_mass, C, go + 28, gl, g_prop + 3, g_tlep_b,
                         g_tlep_w);
}

}  // namespace mf
