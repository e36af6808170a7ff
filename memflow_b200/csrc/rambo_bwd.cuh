// rambo_bwd.cuh -- one event of K1's vector-Jacobian product (see rambo_backward.cu for the gradient semantics), shared by
// rambo_backward.cu and the fused sampling branch (sampling.cu).
#pragma once
#include "rambo.cuh"

namespace mf {

// r = the event's uniforms, mo = its saved forward momenta [NF+2][4] (global memory), (w, x1, x2) = its saved outputs;
// gm = incoming gradient of the momenta [NF+2][4] (any address space; nullptr = none), gw / gx1 / gx2 = of weight, x1, x2.
template <int NF>
MF_D void rambo_forward_bwd_event(const double (&r)[3 * NF - 2], const double* __restrict__ mo, double w, double x1, double x2,
                                  const double* gm, double gw, double gx1, double gx2, const RamboParams<double, NF>& P,
                                  double (&g)[3 * NF - 2]) {
    constexpr int D = 3 * NF - 4, W = D + 2, NE = NF - 2, NP = NF + 2;
#pragma unroll
    for (int i = 0; i < W; ++i) g[i] = 0.0;
    const double E_cm = sqrt(x1 * x2) * P.e_collider;

    // ---- x1 / x2 columns ----
    {
        double gE = 0.0;
        if (gm) gE = 0.5 * ((gm[0] + gm[3]) + (gm[4] - gm[7]));  // incoming rows (E/2, 0, 0, +-E/2)
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
}

}  // namespace mf
