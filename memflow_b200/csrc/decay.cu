// decay.cu -- two-body decay epilogues of the sampled propagators:
//   H -> b b              Compute_ParticlesTensor.higgs_decayProducts, memflow/unfolding_flow/utils.py:745-770
//   t -> b W, W -> q q'   Compute_ParticlesTensor.top_decayProducts,   memflow/unfolding_flow/utils.py:772-815
// Daughters are built from (eta, phi) in the parent's rest frame with E = m/2 (or the massive two-body energies), then
// boosted by the parent's velocity (utils.boost_t) -- the W daughters by the boosted W. One event per thread; the
// parent 4-momentum is one 256-bit load (row stride given, so a row of a K1 output is read in place), every
// output row one 256-bit store. Optionally converted to (E, pt, eta, phi) like the reference's ptetaphi=True.
#include "decay.cuh"

namespace mf {

MF_D void emit_row(double* dst, const double (&p)[4], int as_ptetaphi) {
    if (as_ptetaphi) {
        double o[4];
        ptetaphi(p, o);
        st4(dst, o[0], o[1], o[2], o[3]);
    } else {
        st4(dst, p[0], p[1], p[2], p[3]);
    }
}

__global__ void __launch_bounds__(256)
higgs_decay_kernel(const double* __restrict__ parent, int64_t parent_stride, const double* __restrict__ angles, int64_t N,
                   double mass, int as_ptetaphi, double* __restrict__ out) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double H[4], eta, phi, l1[4], l2[4];
    ld4(parent + ev * parent_stride, H[0], H[1], H[2], H[3]);
    ld2(angles + 2 * ev, eta, phi);
    higgs_decay_rows(H, eta, phi, mass, l1, l2);
    double* o = out + ev * 12;
    emit_row(o, H, as_ptetaphi);
    emit_row(o + 4, l1, as_ptetaphi);
    emit_row(o + 8, l2, as_ptetaphi);
}

__global__ void __launch_bounds__(256)
top_decay_kernel(const double* __restrict__ parent, int64_t parent_stride, const double* __restrict__ b_angles,
                 const double* __restrict__ w_angles, int64_t N, double E_b, double w_mass, double b_mass,
                 int as_ptetaphi, double* __restrict__ out) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double t[4], eb, pb, ew, pw, bl[4], q1l[4], q2l[4];
    ld4(parent + ev * parent_stride, t[0], t[1], t[2], t[3]);
    ld2(b_angles + 2 * ev, eb, pb);
    ld2(w_angles + 2 * ev, ew, pw);
    top_decay_rows(t, eb, pb, ew, pw, E_b, w_mass, b_mass, bl, q1l, q2l);
    double* o = out + ev * 16;
    emit_row(o, t, as_ptetaphi);
    emit_row(o + 4, bl, as_ptetaphi);
    emit_row(o + 8, q1l, as_ptetaphi);
    emit_row(o + 12, q2l, as_ptetaphi);
}

// ---- the whole parton-level event from the three lab-frame propagators, the five decay angle pairs and the boost:
//      Compute_ParticlesTensor.get_decayPartons_fromlab_propagators_angles, memflow/unfolding_flow/utils.py:826-935
//      (call site: unfolding_flow_withDecay_v2.py:399-417). One event per thread: unscale -> pt = exp|.|-1 -> cartesian
//      -> H->bb, t->bqq', t->blv -> ISR gluon from the momentum balance -> (E,pt,eta,phi) -> log(pt+1) + scaling.
//      Output rows: H b1 b2 | thad b q1 q2 | tlep b l nu | ISR. The reference's naming quirk is kept: propagator 1
//      takes (tlep_mass, tlep angles), propagator 2 takes (thad_mass, thad angles) (:866-873). ----
__global__ void __launch_bounds__(128)
decay_partons_kernel(const double* __restrict__ prop, const double* __restrict__ higgs_angles,
                     const double* __restrict__ thad_b_angles, const double* __restrict__ thad_w_angles,
                     const double* __restrict__ tlep_b_angles, const double* __restrict__ tlep_w_angles,
                     const double* __restrict__ boost, int64_t N, const __grid_constant__ mf_decay_partons_cfg C,
                     double* __restrict__ out) {
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = prop[ev * 9 + i];
    decay_partons_event(k, higgs_angles + 2 * ev, thad_b_angles + 2 * ev, thad_w_angles + 2 * ev, tlep_b_angles + 2 * ev,
                        tlep_w_angles + 2 * ev, boost[2 * ev + 1], C, out + ev * 48);
}

}  // namespace mf

extern "C" int mf_higgs_decay_f64(const double* parent, int64_t parent_stride, const double* angles, int64_t N, double mass,
                                  int as_ptetaphi, double* out, mf_stream_t stream) {
    if (N < 0 || (N > 0 && (!parent || !angles || !out))) return MF_E_BADARG;
    if (parent_stride < 4 || parent_stride % 4 != 0) return MF_E_SHAPE;
    if (!mf::aligned(parent, 32) || !mf::aligned(angles, 16) || !mf::aligned(out, 32)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    mf::higgs_decay_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parent, parent_stride, angles, N,
                                                                                          mass, as_ptetaphi, out);
    return mf::launch_status();
}

extern "C" int mf_top_decay_f64(const double* parent, int64_t parent_stride, const double* b_angles, const double* w_angles,
                                int64_t N, double top_mass, double w_mass, double b_mass, int as_ptetaphi, double* out,
                                mf_stream_t stream) {
    if (N < 0 || (N > 0 && (!parent || !b_angles || !w_angles || !out))) return MF_E_BADARG;
    if (parent_stride < 4 || parent_stride % 4 != 0) return MF_E_SHAPE;
    if (!mf::aligned(parent, 32) || !mf::aligned(b_angles, 16) || !mf::aligned(w_angles, 16) || !mf::aligned(out, 32))
        return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    const double E_b = mf::top_Eb(top_mass, w_mass, b_mass);
    mf::top_decay_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        parent, parent_stride, b_angles, w_angles, N, E_b, w_mass, b_mass, as_ptetaphi, out);
    return mf::launch_status();
}

extern "C" int mf_decay_partons_f64(const double* propagators, const double* higgs_angles, const double* thad_b_angles,
                                    const double* thad_w_angles, const double* tlep_b_angles, const double* tlep_w_angles,
                                    const double* boost, int64_t N, const mf_decay_partons_cfg* cfg, double* out,
                                    mf_stream_t stream) {
    if (N < 0 || !cfg) return MF_E_BADARG;
    if (N > 0 && (!propagators || !higgs_angles || !thad_b_angles || !thad_w_angles || !tlep_b_angles || !tlep_w_angles ||
                  !boost || !out))
        return MF_E_BADARG;
    // the reference's cartesian final scaling reads an undefined name (utils.py:913): there is nothing to reproduce
    if (cfg->final_scaling && !cfg->ptetaphi) return MF_E_BADARG;
    if (!mf::aligned(higgs_angles, 16) || !mf::aligned(thad_b_angles, 16) || !mf::aligned(thad_w_angles, 16) ||
        !mf::aligned(tlep_b_angles, 16) || !mf::aligned(tlep_w_angles, 16) || !mf::aligned(boost, 16) || !mf::aligned(out, 32))
        return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    mf::decay_partons_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        propagators, higgs_angles, thad_b_angles, thad_w_angles, tlep_b_angles, tlep_w_angles, boost, N, *cfg, out);
    return mf::launch_status();
}
