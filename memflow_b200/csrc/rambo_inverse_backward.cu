// rambo_inverse_backward.cu -- vector-Jacobian product of K2: gradient of the unit-hypercube point
// ps[N, 3n-2] returned by getPSpoint_batch (memflow/phasespace/rambo_generator.py:615-736) / get_PS
// (memflow/unfolding_flow/utils.py:1056-1152) with respect to the final-state momenta and x1, x2, as torch autograd
// computes it on the reference graph (the conditioning vector of the unfolding flow is built from ps, so this gradient
// reaches the regression network at every training step, memflow/unfolding_flow/unfolding_flow.py:163-175).
// `prob` carries no gradient here (get_PS returns 0 * jac; the dataset path runs without autograd).
// One event per thread: forward recompute of the suffix sums S_j, masses M_j, K_j; adjoints of S_j are collected while
// walking the decays backwards and folded into the particles by a prefix sum (S_j = sum_{k>=j} P_k).
#include "rambo.cuh"

namespace mf {

template <int NF>
__global__ void __launch_bounds__(256)
rambo_inverse_bwd_kernel(const double* __restrict__ momenta, int64_t event_stride, const double* __restrict__ x1a,
                         const double* __restrict__ x2a, const double* __restrict__ g_ps, int64_t N,
                         const __grid_constant__ RamboParams<double, NF> P, double* __restrict__ g_mom /*[N, NF, 4]*/,
                         double* __restrict__ g_x1, double* __restrict__ g_x2) {
    constexpr int D = 3 * NF - 4, W = D + 2, NE = NF - 2;
    const int64_t ev = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ev >= N) return;
    double Pm[NF][4], g[W];
    {
        const double* base = momenta + ev * event_stride;
#pragma unroll
        for (int k = 0; k < NF; ++k) ld4(base + 4 * P.order[k], Pm[k][0], Pm[k][1], Pm[k][2], Pm[k][3]);
    }
    load_row<W>(g_ps + ev * W, g);

    // ---- x1, x2: utils.py:287-309 ----
    if (g_x1 != nullptr) {
        const double x1 = x1a[ev], x2 = x2a[ev];
        if (P.flags & MF_FLAG_X1X2_DIRECT) {
            g_x1[ev] = g[D];
            g_x2[ev] = g[D + 1];
        } else {
            const double a1 = -log(x1), b = -log(x2), qa = P.q - a1;
            g_x1[ev] = -(g[D] * qa / (P.A * x1)) - g[D + 1] * b / (qa * qa * x1);
            g_x2[ev] = -g[D + 1] / (qa * x2);
        }
    }

    // ---- walk j = NF-1 .. 0: S_j = S_{j+1} + P_j, M_j, K_j; adjoints gS_j ----
    double gS[NF][4];
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) gS[j][c] = 0.0;
    double gPl[NF][4];  // local (boost-input) adjoints of P_j
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) gPl[j][c] = 0.0;

    // forward quantities needed by the mass variables: M_j, K_j for all j (O(n^2) sums in the reference's order)
    double Mj[NF], Kj[NF], Sj[NF][4];
#pragma unroll
    for (int j = NF - 1; j >= 0; --j) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = j; k < NF; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += Pm[k][c];
#pragma unroll
        for (int c = 0; c < 4; ++c) Sj[j][c] = s[c];
        Mj[j] = sqrt(((s[0] * s[0] - s[1] * s[1]) - s[2] * s[2]) - s[3] * s[3]);
        Kj[j] = Mj[j] - P.msum_rev[j];
    }
    double gK[NF];
#pragma unroll
    for (int j = 0; j < NF; ++j) gK[j] = 0.0;

    [&]<int... I>(std::integer_sequence<int, I...>) {
        ((void)[&] {
            constexpr int a = I;  // decay index: particle a in the rest frame of S_a; angle columns NE+2a, NE+2a+1
            // mass variable r[a] = (e+1) t^e - e t^(e+1), t = K_{a+1} / K_a, e = NF-2-a   (a <= NF-3)
            if constexpr (a <= NF - 3) {
                constexpr int e = NF - 2 - a;
                const double t = Kj[a + 1] / Kj[a];
                const double dr = (double)(e * (e + 1)) * ipow<e - 1>(t) * (1.0 - t);
                const double gt = g[a] * dr;
                gK[a + 1] += gt / Kj[a];
                gK[a] -= gt * t / Kj[a];
            }
            // angles of particle a boosted by B = -S_a / S_a0
            const double S0 = Sj[a][0];
            const double iS = 1.0 / S0;
            const double B[3] = {-Sj[a][1] * iS, -Sj[a][2] * iS, -Sj[a][3] * iS};
            const double b2 = (B[0] * B[0] + B[1] * B[1]) + B[2] * B[2];
            const double gamma = 1.0 / sqrt(1.0 - b2);
            const double g3 = gamma * gamma * gamma;
            double gamma2, c2;
            if (b2 > 1e-3) {
                gamma2 = (gamma - 1.0) / b2;
                c2 = (g3 * b2 - 2.0 * (gamma - 1.0)) / (b2 * b2);
            } else if (b2 > 0.0) {
                gamma2 = 0.5 + b2 * (0.375 + b2 * (0.3125 + b2 * 0.2734375));
                c2 = 0.75 + b2 * (1.25 + b2 * 1.640625);
            } else {  // reference: gamma2 = 0 where b2 == 0 (utils.py:108-110)
                gamma2 = 0.0;
                c2 = 0.0;
            }
            const double* Pa = Pm[a];
            const double bp = (B[0] * Pa[1] + B[1] * Pa[2]) + B[2] * Pa[3];
            const double f = gamma2 * bp + gamma * Pa[0];
            const double ps_[3] = {Pa[1] + f * B[0], Pa[2] + f * B[1], Pa[3] + f * B[2]};
            const double r2 = (ps_[0] * ps_[0] + ps_[1] * ps_[1]) + ps_[2] * ps_[2];
            const double ir = 1.0 / sqrt(r2);
            const double c = ps_[2] * ir;
            const double gth = 0.5 * g[NE + 2 * a] * ir;            // d((c+1)/2)/dp* = (z - c p^) / (2|p*|)
            const double pt2 = ps_[0] * ps_[0] + ps_[1] * ps_[1];
            const double gph = g[NE + 2 * a + 1] * 0.15915494309189535 / pt2;  // d(phi/2pi)/dp* = (-py, px, 0)/(2pi pt^2)
            double gp[3] = {-gth * c * ps_[0] * ir - gph * ps_[1], -gth * c * ps_[1] * ir + gph * ps_[0],
                            gth * (1.0 - c * c)};
            const double s = (gp[0] * B[0] + gp[1] * B[1]) + gp[2] * B[2];
            const double cb = bp * c2 + Pa[0] * g3;
            double gB[3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                gPl[a][cc + 1] += gp[cc] + s * gamma2 * B[cc];
                gB[cc] = f * gp[cc] + s * (gamma2 * Pa[cc + 1] + cb * B[cc]);
            }
            gPl[a][0] += s * gamma;
            // B = -S_vec / S0
            gS[a][0] += ((gB[0] * Sj[a][1] + gB[1] * Sj[a][2]) + gB[2] * Sj[a][3]) * iS * iS;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) gS[a][cc + 1] -= gB[cc] * iS;
        }(), ...);
    }(std::make_integer_sequence<int, NF - 1>{});

    // M_j = sqrt(S_j . S_j): gS_j += gK_j (S0, -S_vec) / M_j. j = NF-1 is skipped: K_{NF-1} only feeds a value the
    // reference overwrites (:672,:681), its adjoint is structurally zero -- and M_{NF-1} = sqrt(E^2 - p^2) of a
    // near-massless body is NaN for half of the events (the reference's own gradient is NaN there: 0 * NaN).
#pragma unroll
    for (int j = 0; j < NF - 1; ++j) {
        const double w = gK[j] / Mj[j];
        gS[j][0] += w * Sj[j][0];
#pragma unroll
        for (int c = 1; c < 4; ++c) gS[j][c] -= w * Sj[j][c];
    }
    // S_j = sum_{k>=j} P_k  ->  gP_k = local_k + sum_{j<=k} gS_j ; scatter back through the permutation
    double run[4] = {0.0, 0.0, 0.0, 0.0};
    double* out = g_mom + ev * (NF * 4);
#pragma unroll
    for (int k = 0; k < NF; ++k) {
#pragma unroll
        for (int c = 0; c < 4; ++c) run[c] += gS[k][c];
        st4(out + 4 * P.order[k], gPl[k][0] + run[0], gPl[k][1] + run[1], gPl[k][2] + run[2], gPl[k][3] + run[3]);
    }
}

}  // namespace mf

extern "C" int mf_rambo_inverse_bwd_f64(const double* momenta, int64_t event_stride, const double* x1, const double* x2,
                                        const double* g_ps, int64_t N, int n_final, const int32_t* order,
                                        const double* masses, const double* target_mass, int permute_target,
                                        double e_collider, uint32_t flags, double* g_momenta, double* g_x1, double* g_x2,
                                        mf_stream_t stream) {
    if (N < 0 || !masses || (N > 0 && (!momenta || !g_ps || !g_momenta))) return MF_E_BADARG;
    if ((g_x1 == nullptr) != (g_x2 == nullptr)) return MF_E_BADARG;
    if (g_x1 && (!x1 || !x2)) return MF_E_BADARG;
    if (n_final < MF_MIN_FINAL || n_final > MF_MAX_FINAL) return MF_E_NFINAL;
    if (event_stride < 4 * n_final || event_stride % 4 != 0) return MF_E_SHAPE;
    const int W = 3 * n_final - 2;
    const size_t row_align = (W % 4 == 0) ? 32 : ((W % 2 == 0) ? 16 : 8);
    if (!mf::aligned(momenta, 32) || !mf::aligned(g_momenta, 32) || !mf::aligned(g_ps, row_align)) return MF_E_ALIGN;
    if (N == 0) return MF_OK;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    switch (n_final) {
#define MF_CASE(NF)                                                                                                        \
    case NF: {                                                                                                             \
        mf::RamboParams<double, NF> P{};                                                                                   \
        bool seen[NF] = {};                                                                                                \
        for (int i = 0; i < NF; ++i) {                                                                                     \
            int o = order ? order[i] : i;                                                                                  \
            if (o < 0 || o >= NF || seen[o]) return MF_E_SHAPE;                                                            \
            seen[o] = true;                                                                                                \
            P.order[i] = o;                                                                                                \
            P.m[i] = target_mass ? target_mass[permute_target ? o : i] : masses[o];                                        \
        }                                                                                                                  \
        for (int j = 0; j < NF; ++j) {                                                                                     \
            double s = 0.0;                                                                                                \
            for (int k = j; k < NF; ++k) s += P.m[k];                                                                      \
            P.msum_rev[j] = s;                                                                                             \
        }                                                                                                                  \
        double ms = 0.0;                                                                                                   \
        for (int i = 0; i < NF; ++i) ms += (permute_target && target_mass) ? target_mass[i] : masses[i];                   \
        const double ratio = ms / e_collider;                                                                              \
        P.q = -std::log(ratio * ratio);                                                                                    \
        P.A = -0.5 * (P.q * P.q) + (P.q * P.q);                                                                            \
        P.e_collider = e_collider;                                                                                         \
        P.flags = flags;                                                                                                   \
        mf::rambo_inverse_bwd_kernel<NF><<<blocks, 256, 0, (cudaStream_t)stream>>>(momenta, event_stride, x1, x2, g_ps, N, \
                                                                                   P, g_momenta, g_x1, g_x2);              \
        break;                                                                                                             \
    }
        MF_CASE(3) MF_CASE(4) MF_CASE(5) MF_CASE(6) MF_CASE(7) MF_CASE(8)
#undef MF_CASE
    }
    return mf::launch_status();
}
