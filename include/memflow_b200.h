/*
 * memflow_b200.h -- C ABI of the B200-native (sm_100a) MEMFlow phase-space hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. Every entry point
 * replaces one reference interface, cited as file:line relative to valsdav/MEMFlow.
 * The Python host layer (memflow_b200/) binds these with ctypes and mirrors the reference's
 * classes; INTEGRATION.md shows the stub a MEMFlow maintainer would add.
 *
 * Conventions
 *   - All array pointers are DEVICE pointers unless the comment says "host".
 *   - Row-major, contiguous. Momenta are AoS [event][particle][E,px,py,pz] exactly as the
 *     reference returns them (rambo_generator.py:327-331).
 *   - Kernels are enqueued on `stream` (a cudaStream_t); nothing synchronises. Re-entrant:
 *     no global mutable state, masses / constants travel with each launch.
 *   - Return value: 0 = MF_OK, >0 = MF_E_* argument errors detected on the host before any
 *     launch, <0 = -(cudaError_t) from the launch. No exceptions cross this boundary.
 *   - Data-dependent failures of the reference (NaN scan rambo_generator.py:191-199, CM check
 *     :630-635) are reported through the optional device word `status` (bit-or of MF_ST_*),
 *     never by a host sync.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MEMFLOW_B200_H
#define MEMFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mf_stream_t; /* == cudaStream_t */

#define MF_ABI_VERSION 1
#define MF_MAX_FINAL 8 /* n_final supported: 3..8 (the reference itself fails for n=2, :402-403,455) */
#define MF_MIN_FINAL 3

/* return codes */
#define MF_OK 0
#define MF_E_BADARG 1   /* null pointer / negative size */
#define MF_E_NFINAL 2   /* n_final outside [MF_MIN_FINAL, MF_MAX_FINAL] */
#define MF_E_ALIGN 3    /* pointer not aligned for the vector width used (32 B f64 / 16 B f32 momenta rows) */
#define MF_E_SHAPE 4    /* spline K/F out of range, bad order permutation ... */
#define MF_E_NODEVICE 5 /* no CUDA device / wrong architecture (needs sm_100) */

/* flags (bit-or) */
#define MF_FLAG_REF_FP32_QUIRKS 1u /* reproduce the reference's float32 volume/weight chain and float32 \
                                      incoming rows (rambo_generator.py:132,489-507,533); default = clean fp64 */
#define MF_FLAG_LAB_FRAME 2u       /* forward: return momenta boosted to the lab frame (utils.py:180-202) */
#define MF_FLAG_CUTS 4u            /* forward: apply cuts[3] = {pT_min, dR_min, |eta|_max} (:350-393) */
#define MF_FLAG_X1X2_DIRECT 8u     /* uniform_x1x2=False: last two columns ARE x1,x2, no Jacobian (:229-234) */
#define MF_FLAG_NO_CM_CHECK 16u    /* inverse: ensure_CM=False (:630) */
#define MF_FLAG_ZERO_STATUS 32u    /* forward / inverse: *status is cleared (cudaMemsetAsync on the call's stream) before the \
                                      launch, so the caller may hand in uninitialised memory */

/* status bits written with atomicOr into *status (device int32, may be NULL) */
#define MF_ST_NAN_INPUT 1 /* forward: a random variable was NaN (:191-199 raises) */
#define MF_ST_NOT_CM 2    /* inverse: |sum p|^2 >= 1e-3 for some event (:634-635 raises) */

int mf_abi_version(void);
const char* mf_error_string(int code);
/* device probe: fills SM count and compute capability of the current device */
int mf_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * K1  rambo_forward : unit hypercube -> CM-frame 4-momenta + Jacobian weight
 * replaces FlatInvertiblePhasespace.generateKinematics_batch, rambo_generator.py:168-398
 * (with get_x1x2_from_uniform utils.py:260-284, bisect_vec_batch :400-446 -> closed form /
 * Newton root solve, generateIntermediatesMassive_batch :475-509, setInitialStateMomenta_batch
 * :511-551, boost_t utils.py:88-118).
 *   u        [N, 3n-2]    3n-4 RAMBO variables, then (u_x1, u_x2)
 *   masses   host [n]     final-state masses
 *   cuts     host [3]     or NULL (only read with MF_FLAG_CUTS)
 *   momenta  [N, n+2, 4]  rows in1, in2, final_0..final_{n-1}    (32 B-aligned f64 / 16 B f32)
 *   weight   [N]          |det J(u->p)| * x1x2 Jacobian * cut mask
 *   logweight[N] or NULL  ln(weight) (new; avoids the reference's float32 overflow for n>=7)
 *   x1, x2   [N]
 * ---------------------------------------------------------------------------------------- */
int mf_rambo_forward_f64(const double* u, int64_t N, int n_final, const double* masses, double e_collider,
                         const double* cuts, uint32_t flags, double* momenta, double* weight,
                         double* logweight, double* x1, double* x2, int32_t* status, mf_stream_t stream);
int mf_rambo_forward_f32(const float* u, int64_t N, int n_final, const float* masses, float e_collider,
                         const float* cuts, uint32_t flags, float* momenta, float* weight, float* logweight,
                         float* x1, float* x2, int32_t* status, mf_stream_t stream);

/* Vector-Jacobian product of K1 with the gradient semantics of the reference's autograd graph
 * (generateKinematics_batch(..., requires_grad=True)): grad_u [N, 3n-2] from the cotangents of the four
 * outputs (each may be NULL = zero). Mass-variable columns get 0; angle columns the true derivative through the decay
 * chain; the x1/x2 columns reach x1, x2, the weight and the incoming rows only. u / momenta / weight / x1 / x2 are the
 * forward's input and outputs. Not available with MF_FLAG_LAB_FRAME / MF_FLAG_CUTS. */
int mf_rambo_forward_bwd_f64(const double* u, const double* momenta, const double* weight, const double* x1,
                             const double* x2, const double* g_momenta, const double* g_weight, const double* g_x1,
                             const double* g_x2, int64_t N, int n_final, const double* masses, double e_collider,
                             uint32_t flags, double* grad_u, mf_stream_t stream);

/* Vector-Jacobian product of K2 / get_PS: gradient of ps [N, 3n-2] w.r.t. the final-state momenta (g_momenta [N, n, 4],
 * packed, original particle order) and x1, x2 (g_x1, g_x2 [N], both or neither NULL), as torch autograd computes it on the
 * reference graph (rambo_generator.py:615-736; memflow/unfolding_flow/utils.py:1056-1152 -- pass the x1, x2 get_PS derived
 * from the boost). order / masses / target_mass as in the forward; permute_target != 0 = get_PS convention
 * (target_mass[order]). `prob` has no gradient. */
int mf_rambo_inverse_bwd_f64(const double* momenta, int64_t event_stride, const double* x1, const double* x2,
                             const double* g_ps, int64_t N, int n_final, const int32_t* order, const double* masses,
                             const double* target_mass, int permute_target, double e_collider, uint32_t flags,
                             double* g_momenta, double* g_x1, double* g_x2, mf_stream_t stream);

/* Mass-variable root solve alone: u[i, j] solves v = u^e ((e+1) - e u), e = n_final-2-j.
 * replaces FlatInvertiblePhasespace.bisect_vec_batch, rambo_generator.py:400-446. v, u: [N, n_final-2]. */
int mf_massvar_solve_f64(const double* v, int64_t N, int n_final, double* u, mf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  rambo_inverse : CM-frame final-state momenta -> unit hypercube + 1/weight
 * replaces FlatInvertiblePhasespace.getPSpoint_batch, rambo_generator.py:615-736
 * (with get_uniform_from_x1x2 utils.py:287-309).
 *   momenta      final-state rows only; event e, particle k at momenta + e*event_stride + 4k
 *                (event_stride in ELEMENTS, multiple of 4; pass 4n for a packed [N,n,4], or
 *                4(n+2) with momenta pointing at row 2 of a forward output -- no copy needed)
 *   order        host [n] permutation (:624-625)
 *   masses       host [n] generator masses
 *   target_mass  host [n] or NULL. NULL = ensure_onShell=True (masses[order], :649-650); otherwise
 *                used as given, unpermuted (ensure_onShell=False, :651-652; the reference only
 *                handles a [1,n] tensor there)
 *   ps [N, 3n-2], prob [N] = jac_x1x2 / rambo_jac (:732-734), logprob [N] or NULL
 * ---------------------------------------------------------------------------------------- */
int mf_rambo_inverse_f64(const double* momenta, int64_t event_stride, const double* x1, const double* x2,
                         int64_t N, int n_final, const int32_t* order, const double* masses,
                         const double* target_mass, double e_collider, uint32_t flags, double* ps,
                         double* prob, double* logprob, int32_t* status, mf_stream_t stream);
int mf_rambo_inverse_f32(const float* momenta, int64_t event_stride, const float* x1, const float* x2,
                         int64_t N, int n_final, const int32_t* order, const float* masses,
                         const float* target_mass, float e_collider, uint32_t flags, float* ps, float* prob,
                         float* logprob, int32_t* status, mf_stream_t stream);

/* K2 with the semantics of Compute_ParticlesTensor.get_PS, memflow/unfolding_flow/utils.py:1056-1152 (called every
 * training step, unfolding_flow.py:163): x1,x2 = (E +- pz)/e_collider of the reconstructed boost [N,4], clamped to [0,1];
 * masses = target_mass[order] (host [n]); no RAMBO weight: jac0 [N] or NULL receives 0 * jac_x1x2 like the reference;
 * mask [N] (uint8) or NULL = x out of range | |sum p|^2 > 1e-5 | any RAMBO variable outside [0,1].
 * Fused epilogue of the caller (unfolding_flow.py:170-171, Dataset_Parton_Level.py:428-431): if logit_scaled != NULL,
 * logit_scaled[N, 3n-2] = (torch.logit(ps, eps=logit_eps) - logit_mean) / logit_scale (host [3n-2] each; NULL = 0 / 1). */
int mf_get_ps_f64(const double* momenta, int64_t event_stride, const double* boost, int64_t N, int n_final,
                  const int32_t* order, const double* target_mass, double e_collider, double* ps, double* jac0,
                  uint8_t* mask, double logit_eps, const double* logit_mean, const double* logit_scale,
                  double* logit_scaled, mf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3 / K4  monotonic rational-quadratic spline, zuko.transforms.MonotonicRQSTransform
 * (third-party; constructed at memflow/unfolding_flow/unfolding_flow.py:88-97 and
 * memflow/unfolding_flow/custom_spline_flow_ps.py:16-33).
 *   phi  [B, F, 3K-1]  per (event, feature): widths[K] | heights[K] | derivatives[K-1], unconstrained
 *   in   [B, F]        x (forward) or y (inverse)
 *   out  [B, F]
 *   ladj [B, F] or NULL    forward: log|dy/dx| at x;  inverse: the same FORWARD log|dy/dx| evaluated
 *                          at the recovered x (so sample + log q need one pass)
 *   ladj_sum [B] or NULL   sum over F (what AutoregressiveTransform.call_and_ladj returns)
 *   bin  [B, F] or NULL    searchsorted(knots, v) - 1 (int32; -1 / K = identity tails)
 *   slope <= 0 disables the soft clip (zuko <= 0.1). 1 <= K <= 128.
 * ---------------------------------------------------------------------------------------- */
int mf_rqs_forward_f32(const float* phi, const float* x, int64_t B, int F, int K, float bound, float slope,
                       float* y, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t stream);
int mf_rqs_forward_f64(const double* phi, const double* x, int64_t B, int F, int K, double bound, double slope,
                       double* y, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t stream);
int mf_rqs_inverse_f32(const float* phi, const float* y, int64_t B, int F, int K, float bound, float slope,
                       float* x, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t stream);
int mf_rqs_inverse_f64(const double* phi, const double* y, int64_t B, int F, int K, double bound, double slope,
                       double* x, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t stream);
/* K3 / K4 on a FEATURE SUBSET of a wider parameter tensor, read in place: phi points at the subset's first row, the F rows of
 * an event are contiguous and consecutive events are event_stride ELEMENTS apart (event_stride >= F (3K-1); equality = the
 * calls above). This is what an autoregressive sampler needs: pass j of zuko's MaskedAutoregressiveTransform inverse
 * (call sites memflow/unfolding_flow/unfolding_flow.py:97,186) determines feature j only, so it has to read 3K-1 of the
 * event's F_total (3K-1) parameters -- phi + j (3K-1), F = 1, event_stride = F_total (3K-1) -- instead of all of them.
 * x / y / ladj / bin are packed [B, F]. */
int mf_rqs_forward_strided_f32(const float* phi, int64_t event_stride, const float* x, int64_t B, int F, int K, float bound,
                               float slope, float* y, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t stream);
int mf_rqs_forward_strided_f64(const double* phi, int64_t event_stride, const double* x, int64_t B, int F, int K, double bound,
                               double slope, double* y, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t stream);
int mf_rqs_inverse_strided_f32(const float* phi, int64_t event_stride, const float* y, int64_t B, int F, int K, float bound,
                               float slope, float* x, float* ladj, float* ladj_sum, int32_t* bin, mf_stream_t stream);
int mf_rqs_inverse_strided_f64(const double* phi, int64_t event_stride, const double* y, int64_t B, int F, int K, double bound,
                               double slope, double* x, double* ladj, double* ladj_sum, int32_t* bin, mf_stream_t stream);

/* Vector-Jacobian products of K3 / K4: from the cotangents g_out [B,F] of the output and g_ladj [B,F] of the
 * per-element log-det (either may be NULL = zero; for a cotangent of ladj_sum broadcast it over F), the gradients
 * g_phi [B,F,3K-1] w.r.t. the unconstrained parameters and g_in [B,F] w.r.t. the input, through soft clip, softmax, cumsum,
 * gather and the rational quadratic. forward_bwd differentiates mf_rqs_forward (in = x), inverse_bwd mf_rqs_inverse
 * (in = y; ladj = forward log-det at the recovered x). */
int mf_rqs_forward_bwd_f32(const float* phi, const float* x, const float* g_y, const float* g_ladj, int64_t B, int F, int K,
                           float bound, float slope, float* g_phi, float* g_x, mf_stream_t stream);
int mf_rqs_forward_bwd_f64(const double* phi, const double* x, const double* g_y, const double* g_ladj, int64_t B, int F, int K,
                           double bound, double slope, double* g_phi, double* g_x, mf_stream_t stream);
int mf_rqs_inverse_bwd_f32(const float* phi, const float* y, const float* g_x, const float* g_ladj, int64_t B, int F, int K,
                           float bound, float slope, float* g_phi, float* g_y, mf_stream_t stream);
int mf_rqs_inverse_bwd_f64(const double* phi, const double* y, const double* g_x, const double* g_ladj, int64_t B, int F, int K,
                           double bound, double slope, double* g_phi, double* g_y, mf_stream_t stream);

/* bin search alone, knots given: idx = torch.searchsorted(knots[b,f,:], v[b,f], right=False) (bit-exact) */
int mf_searchsorted_f32(const float* knots, const float* v, int64_t BF, int len, int32_t* idx, mf_stream_t stream);
int mf_searchsorted_f64(const double* knots, const double* v, int64_t BF, int len, int32_t* idx, mf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5  importance-sampling weight + Monte-Carlo estimator reduction
 * replaces notebooks/TestFlows_ImportanceSampling.ipynb cells [50-54] (mask of cells [35,43]):
 *   w_i = target_i * jac_i / exp(logq_i);  valid_i = target_i > 0 and jac_i is not NaN
 *   acc[0] += sum w, acc[1] += sum w^2, acc[2] += n_valid      (acc: device double[3])
 * target may be NULL (== 1). Deterministic: fixed grid, per-block Kahan partials, the last block
 * folds them in index order. workspace: device, mf_is_reduce_workspace_bytes() bytes, zeroed once
 * by the caller before first use (the kernel leaves it zeroed).
 * ---------------------------------------------------------------------------------------- */
int64_t mf_is_reduce_workspace_bytes(void);
int mf_is_reduce_f64(const double* target, const double* jac, const double* logq, int64_t N, double* acc,
                     void* workspace, mf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Importance-sampling chain (TestFlows_ImportanceSampling.ipynb [45,50-54] in one call, two launches A+B below):
 *   Philox4x32-10 base sample z ~ U[-bound,bound]^F (counter = global event index, key = seed)
 *   -> T spline inverses (K4, parameters phi[t] given)  -> u = (x + bound) / (2 bound)
 *   -> K1 (n_final bodies, F = 3n-2)  -> w = jac / (2 x1 x2 s) / q(u)  -> K5 accumulation.
 *   phi      [T, N, F, 3K-1] float32 (one contiguous [N, F, 3K-1] block per transform)
 *   acc/workspace as K5. Optional outputs (NULL to skip): u_out [N,F] f64, logq [N] f64,
 *   momenta [N,n+2,4], weight [N] (RAMBO weight), x1, x2.
 * ---------------------------------------------------------------------------------------- */
int mf_is_chain_f64(const float* phi, int64_t N, int n_final, int T, int K, float bound, float slope,
                    uint64_t seed, uint64_t event_offset, const double* masses, double e_collider,
                    uint32_t flags, double* acc, void* workspace, double* u_out, double* logq,
                    double* momenta, double* weight, double* x1, double* x2, mf_stream_t stream);

/* Two-body decay epilogues of the sampled propagators, cartesian or (E, pt, eta, phi) output:
 *   mf_higgs_decay_f64: Compute_ParticlesTensor.higgs_decayProducts, memflow/unfolding_flow/utils.py:745-770
 *                       parent [N] rows of 4 (row stride parent_stride elements), angles [N,2] = (eta, phi) of b1 in
 *                       the Higgs rest frame -> out [N,3,4] = (H, b1, b2)
 *   mf_top_decay_f64  : Compute_ParticlesTensor.top_decayProducts, memflow/unfolding_flow/utils.py:772-815
 *                       b_angles / w_angles [N,2] (b in the top frame, q1 in the W frame) -> out [N,4,4] = (t, b, q1, q2) */
int mf_higgs_decay_f64(const double* parent, int64_t parent_stride, const double* angles, int64_t N, double mass,
                       int as_ptetaphi, double* out, mf_stream_t stream);
int mf_top_decay_f64(const double* parent, int64_t parent_stride, const double* b_angles, const double* w_angles,
                     int64_t N, double top_mass, double w_mass, double b_mass, int as_ptetaphi, double* out,
                     mf_stream_t stream);

/* The whole parton-level event from the lab-frame propagators, the decay angles and the boost:
 *   Compute_ParticlesTensor.get_decayPartons_fromlab_propagators_angles, memflow/unfolding_flow/utils.py:826-935
 *   (call site memflow/unfolding_flow/unfolding_flow_withDecay_v2.py:399-417), fused into one launch:
 *   unscale -> pt = exp|log pt| - 1 -> cartesian -> H->bb, t->bqq', t->blv -> ISR gluon = boost - partons
 *   -> (E, pt, eta, phi) -> log(pt + 1) and affine scaling.
 *   propagators [N,3,3] = scaled (log(pt+1), eta, phi) of H, thad, tlep (NOT modified; the reference unscales its
 *   argument in place); *_angles [N,2] = (eta, phi) in the parent rest frame; boost [N,2] = scaled (log(E+1), pz);
 *   out [N,12,4] rows: H b1 b2 | thad b q1 q2 | tlep b l nu | ISR.
 *   The reference pairs propagator 1 with (tlep_mass, tlep angles) and propagator 2 with (thad_mass, thad angles)
 *   (:866-873); kept. final_scaling && !ptetaphi reads an undefined name in the reference (:913): MF_E_BADARG. */
typedef struct mf_decay_partons_cfg {
    double mean_prop[3], std_prop[3];   /* log_mean/std_parton_Hthadtlep: (log(pt+1), eta, phi) */
    double mean_lab[3], std_lab[3];     /* log_mean/std_parton_lab:       (log(pt+1), eta, phi) */
    double mean_boost[2], std_boost[2]; /* log_mean/std_boost:            (log(E+1), pz) */
    double higgs_mass, thad_mass, tlep_mass, w_had_mass, w_lep_mass, b_mass, eps;
    int32_t ptetaphi, unscale_phi, final_scaling, reserved;
} mf_decay_partons_cfg;
int mf_decay_partons_f64(const double* propagators, const double* higgs_angles, const double* thad_b_angles,
                         const double* thad_w_angles, const double* tlep_b_angles, const double* tlep_w_angles,
                         const double* boost, int64_t N, const mf_decay_partons_cfg* cfg, double* out,
                         mf_stream_t stream);

/* Vector-Jacobian product of mf_decay_partons_f64: g_out [N,12,4] -> gradients w.r.t. every input (same shapes).
 * Replaces the autograd graph of get_decayPartons_fromlab_propagators_angles in the sampling loss
 * (memflow/unfolding_flow/unfolding_flow_withDecay_v2.py:376-417). g_boost[:,0] = 0: only pz of the boost enters. */
int mf_decay_partons_bwd_f64(const double* propagators, const double* higgs_angles, const double* thad_b_angles,
                             const double* thad_w_angles, const double* tlep_b_angles, const double* tlep_w_angles,
                             const double* boost, int64_t N, const mf_decay_partons_cfg* cfg, const double* g_out,
                             double* g_propagators, double* g_higgs_angles, double* g_thad_b_angles,
                             double* g_thad_w_angles, double* g_tlep_b_angles, double* g_tlep_w_angles, double* g_boost,
                             mf_stream_t stream);

/* The sampling branch of the reference's training loss as ONE launch (memflow/unfolding_flow/unfolding_flow_withDecay_v2.py:
 * 376-417: rambo.get_momenta_from_ps(ps, requires_grad=True) -> boost_tt(momenta[:, -4:-1], boostVector_t(boost_Epz)) ->
 * get_ptetaphi_comp_batch -> log(pt+1) and scaling -> get_decayPartons_fromlab_propagators_angles), n_final >= 4.
 *   ps [N, 3n-2], the five angle pairs [N,2], cfg as for mf_decay_partons_f64 (mean/std_prop scale the propagators both
 *   ways, mean/std_boost the boost)  ->  out [N,12,4] (the 12-parton event), boost_scaled [N,2] = scaled (log(E_b+1), pz_b).
 *   Saved for the backward launch (and returned to the caller like the reference's tensors): momenta [N,n+2,4] (CM frame),
 *   weight, x1, x2 [N] (K1's outputs), propagators [N,3,3] (scaled lab-frame (log(pt+1), eta, phi) of H, thad, tlep). */
int mf_sampling_partons_f64(const double* ps, int64_t N, int n_final, const double* masses, double e_collider,
                            const double* higgs_angles, const double* thad_b_angles, const double* thad_w_angles,
                            const double* tlep_b_angles, const double* tlep_w_angles, const mf_decay_partons_cfg* cfg,
                            double* out, double* boost_scaled, double* momenta, double* weight, double* x1, double* x2,
                            double* propagators, int32_t* status, mf_stream_t stream);
/* Its vector-Jacobian product, ONE launch: g_out [N,12,4], g_boost_scaled [N,2] (may be NULL) -> g_ps [N,3n-2] and the
 * gradients of the five angle pairs. Gradient semantics of the reference's autograd graph (see mf_rambo_forward_bwd_f64:
 * the mass-variable columns of ps receive none). */
int mf_sampling_partons_bwd_f64(const double* ps, int64_t N, int n_final, const double* masses, double e_collider,
                                const double* higgs_angles, const double* thad_b_angles, const double* thad_w_angles,
                                const double* tlep_b_angles, const double* tlep_w_angles, const mf_decay_partons_cfg* cfg,
                                const double* momenta, const double* weight, const double* x1, const double* x2,
                                const double* propagators, const double* boost_scaled, const double* g_out,
                                const double* g_boost_scaled, double* g_ps, double* g_higgs_angles, double* g_thad_b_angles,
                                double* g_thad_w_angles, double* g_tlep_b_angles, double* g_tlep_w_angles, mf_stream_t stream);

/* The two stages of the chain as separate entry points (mf_is_chain_f64 runs them back to back):
 *  A  mf_flow_sample_f32: Philox base sample + T spline inverses -> u [N,F] (fp64, in [0,1]) and log q [N]
 *                         (flow().sample + flow().log_prob of TestFlows_ImportanceSampling.ipynb [45,50] in one pass)
 *  B  mf_rambo_is_f64   : K1 + w = jac / (2 x1 x2 s) / exp(log q) + K5 accumulation, on given u / log q
 *                         (cells [50-54]); optional momenta / weight / x1 / x2 outputs. */
int mf_flow_sample_f32(const float* phi, int64_t N, int F, int T, int K, float bound, float slope, uint64_t seed,
                       uint64_t event_offset, double* u, double* logq, mf_stream_t stream);
/* A with float64 flow parameters (phi [T, N, F, 3K-1] double: the dtype of MEMFlow's configs). Same base point and
 * outputs as mf_flow_sample_f32; built from the fp64 spline kernel (one launch per transform). */
int mf_flow_sample_f64(const double* phi, int64_t N, int F, int T, int K, double bound, double slope, uint64_t seed,
                       uint64_t event_offset, double* u, double* logq, mf_stream_t stream);
int mf_rambo_is_f64(const double* u, const double* logq, int64_t N, int n_final, const double* masses,
                    double e_collider, uint32_t flags, double* acc, void* workspace, double* momenta, double* weight,
                    double* x1, double* x2, mf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MEMFLOW_B200_H */
