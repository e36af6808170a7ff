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
#include "rambo.cuh"

namespace mf {

template <int NF>
__global__ void __launch_bounds__(256, 2)
rambo_forward_bwd_kernel(const double* __restrict__ u, const double* __restrict__ momenta, const double* __restrict__ weight,
                         const double* __restrict__ x1a, const double* __restrict__ x2a, const double* __restrict__ g_mom,
                         const double* __restrict__ g_w, const double* __restrict__ g_x1, const double* __restrict__ g_x2,
                         int64_t N, const __grid_constant__ RamboParams<double, NF> P, double* __restrict__ grad_u) {
    constexpr int D = 3 * NF - 4, W = D + 2, NE = NF - 2, NP = NF + 2;
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double r[W], g[W];
    load_row<W>(u + ev * W, r);
#pragma unroll
    for (int i = 0; i < W; ++i) g[i] = 0.0;
    const double x1 = x1a[ev], x2 = x2a[ev], w = weight[ev];
    const double E_cm = sqrt(x1 * x2) * P.e_collider;
    const double* gm = g_mom ? g_mom + ev * (NP * 4) : nullptr;
    const double* mo = momenta + ev * (NP * 4);

    // ---- x1 / x2 columns ----
    {
        double gE = 0.0;
        if (gm) gE = 0.5 * ((gm[0] + gm[3]) + (gm[4] - gm[7]));  // incoming rows (E/2, 0, 0, +-E/2)
        const double gw = g_w ? g_w[ev] : 0.0, gx1 = g_x1 ? g_x1[ev] : 0.0, gx2 = g_x2 ? g_x2[ev] : 0.0;
        if (P.flags & MF_FLAG_X1X2_DIRECT) {
            const double common = (double)(NF - 2) * gw * w + 0.5 * gE * E_cm;  // weight ~ E^(2n-4) only
            g[D] = gx1 + common / x1;
            g[D + 1] = gx2 + common / x2;
        } else {
            const double common = (double)(NF - 1) * gw * w + 0.5 * gE * E_cm;  // d ln w = (n-1) (d ln x1 + d ln x2)
            const double adj1 = gx1 * x1 + common, adj2 = gx2 * x2 + common;   // adjoints of ln x1, ln x2
            const double s1 = sqrt(1.0 - r[D]);
            const double a = P.q * (1.0 - s1);          // -ln x1
            const double da = P.q / (2.0 * s1);         // d(-ln x1) / d u_x1
            g[D] = -adj1 * da + adj2 * (r[D + 1] * da);  // -ln x2 = u_x2 (q - a)
            g[D + 1] = -adj2 * (P.q - a);
        }
    }

    if (gm) {
        // ---- masses and decay momenta (same as the forward) ----
        double M[NF], qk[NF];
        {
            const double K0 = E_cm - P.msum;
            double prod = 1.0, Kc[NF];
            Kc[0] = K0;
            [&]<int... I>(std::integer_sequence<int, I...>) {
                ((prod *= solve_massvar<NF - 2 - I>(r[I]), Kc[I + 1] = K0 * prod), ...);
            }(std::make_integer_sequence<int, NE>{});
#pragma unroll
            for (int i = 0; i <= NF - 2; ++i) M[i] = Kc[i] + P.msum_rev[i];
            M[NF - 1] = P.m[NF - 1];
#pragma unroll
            for (int k = 0; k <= NF - 2; ++k) qk[k] = rho_num(M[k], M[k + 1], P.m[k]) * (0.5 / M[k]);
        }
        // ---- reverse sweep over the decays ----
        double Q[4], gQ[3];
        ld4(mo + 4 * (NP - 1), Q[0], Q[1], Q[2], Q[3]);
        {
            const double* go = gm + 4 * (NP - 1);
            const double eo = go[0] / Q[0];  // last row: E = sqrt(|Q|^2 + m^2)
            gQ[0] = go[1] + eo * Q[1]; gQ[1] = go[2] + eo * Q[2]; gQ[2] = go[3] + eo * Q[3];
        }
#pragma unroll
        for (int k = NF - 2; k >= 0; --k) {
            double pb[4];
            ld4(mo + 4 * (2 + k), pb[0], pb[1], pb[2], pb[3]);
            const double* go = gm + 4 * (2 + k);
            const double eo = go[0] / pb[0];
            double gpb[3] = {go[1] + eo * pb[1] - gQ[0], go[2] + eo * pb[2] - gQ[1], go[3] + eo * pb[3] - gQ[2]};
            // Q_k = Q_{k+1} + p_k
#pragma unroll
            for (int c = 0; c < 4; ++c) Q[c] += pb[c];
            // rest-frame decay momentum from the angles
            const double ct = 2.0 * r[NE + 2 * k] - 1.0;
            const double st = sqrt(1.0 - ct * ct);
            const double phi = 6.283185307179586 * r[NE + 2 * k + 1];
            const double cp = cos(phi);
            const double sq = sqrt(1.0 - cp * cp);
            const double sp = (phi > 3.141592653589793) ? -sq : sq;
            const double q = qk[k];
            const double p[3] = {(q * st) * cp, (q * st) * sp, q * ct};
            double gp[3];
            if (k > 0) {
                const double Ep = sqrt(q * q + P.m[k] * P.m[k]);
                const double iq = 1.0 / Q[0];
                const double b[3] = {Q[1] * iq, Q[2] * iq, Q[3] * iq};
                const double b2 = (b[0] * b[0] + b[1] * b[1]) + b[2] * b[2];
                const double gamma = 1.0 / sqrt(1.0 - b2);
                const double g3 = gamma * gamma * gamma;
                double gamma2, c2;
                if (b2 > 1e-3) {
                    gamma2 = (gamma - 1.0) / b2;
                    c2 = (g3 * b2 - 2.0 * (gamma - 1.0)) / (b2 * b2);
                } else {  // series: no cancellation for slow parents
                    gamma2 = 0.5 + b2 * (0.375 + b2 * (0.3125 + b2 * 0.2734375));
                    c2 = 0.75 + b2 * (1.25 + b2 * 1.640625);
                }
                const double bp = (b[0] * p[0] + b[1] * p[1]) + b[2] * p[2];
                const double f = gamma2 * bp + gamma * Ep;
                const double s = (gpb[0] * b[0] + gpb[1] * b[1]) + gpb[2] * b[2];
                const double cb = bp * c2 + Ep * g3;
                double gb[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    gp[c] = gpb[c] + s * (gamma2 * b[c] + (gamma / Ep) * p[c]);
                    gb[c] = f * gpb[c] + s * (gamma2 * p[c] + cb * b[c]);
                }
                const double gbq = ((gb[0] * Q[1] + gb[1] * Q[2]) + gb[2] * Q[3]) * iq * iq * iq;
#pragma unroll
                for (int c = 0; c < 3; ++c) gQ[c] += gb[c] * iq - gbq * Q[c + 1];
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) gp[c] = gpb[c];
            }
            // d p / d cos(theta) = q (-ct/st cphi, -ct/st sphi, 1);  d p / d phi = q st (-sphi, cphi, 0)
            const double cot = ct / st;
            g[NE + 2 * k] = 2.0 * q * ((-cot * cp) * gp[0] + (-cot * sp) * gp[1] + gp[2]);
            g[NE + 2 * k + 1] = 6.283185307179586 * (q * st) * (cp * gp[1] - sp * gp[0]);
        }
    }
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
