// Ionic models as host/device inline functions (also compiled for the host by tests/hostmath to
// check the arithmetic against the oracle without a GPU).
//   FHN      src/modeling/cells/fhn.jl:6-34
//   PCG2019  src/modeling/cells/pcg2019.jl:4-133
//   Aliev-Panfilov  src/modeling/cells/aliev-panfilov.jl:1-34 (recovery variable FIRST: phi_m is state index 2 there, 1 here)
// Operation order follows the reference expression by expression (m*m*m*h*h, sigmoid as
// 1/(1+exp(sign*(phi-E)/k))); the library is compiled with -fmad=false because Julia does not
// contract a*b+c, so FHN is bitwise the reference and PCG2019 differs only through exp's last ulp.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

struct tb_cell_params {
    double p[36];
};

template <int MODEL> struct tb_cell_traits;
template <> struct tb_cell_traits<0> { static constexpr int NS = 2; static constexpr int NP = 6; static constexpr int PHI = 0; };
template <> struct tb_cell_traits<1> { static constexpr int NS = 7; static constexpr int NP = 36; static constexpr int PHI = 0; };
template <> struct tb_cell_traits<2> { static constexpr int NS = 2; static constexpr int NP = 6; static constexpr int PHI = 1; };

namespace tbpcg {
enum { g_Na, E_m, k_m, tau_m, E_h, k_h, delta_h, tau_h0, g_K1, E_z, k_z, g_to, E_r, k_r, E_s, k_s, tau_s, g_CaL, E_d, k_d,
       E_f, k_f, tau_f, g_Kr, E_xr, k_xr, tau_xr, E_y, k_y, g_Ks, E_xs, k_xs, tau_xs, E_Na, E_K, E_Ca };
}

TB_HD double tb_sigmoid(double phi, double E, double k, double sign) { return 1.0 / (1.0 + exp(sign * (phi - E) / k)); }

template <int MODEL> TB_HD void tb_cell_rhs(const tb_cell_params &prm, const double *u, double t, double *du);

template <> TB_HD void tb_cell_rhs<0>(const tb_cell_params &prm, const double *u, double t, double *du) {
    const double a = prm.p[0], b = prm.p[1], c = prm.p[2], d = prm.p[3], e = prm.p[4], f = prm.p[5];
    const double phi = u[0], s = u[1];
    du[0] = f * (phi * (1 - phi) * (phi - a) - s);
    du[1] = e * (b * phi - c * s - d);
}

// aliev-panfilov.jl:15-34; prm = c_t, k, a, eps0, mu1, mu2; u = (s, phi)
template <> TB_HD void tb_cell_rhs<2>(const tb_cell_params &prm, const double *u, double t, double *du) {
    const double ct = prm.p[0], k = prm.p[1], a = prm.p[2], e0 = prm.p[3], mu1 = prm.p[4], mu2 = prm.p[5];
    const double phi = u[1], s = u[0];
    const double eps = e0 + s * mu1 / (phi + mu2);
    du[1] = ct * (k * phi * (phi - 1.0) * (phi - a) - phi * s);
    du[0] = ct * eps * (-s - k * phi * (phi - a - 1.0));
}

template <> TB_HD void tb_cell_rhs<1>(const tb_cell_params &prm, const double *u, double t, double *du) {
    using namespace tbpcg;
    const double *p = prm.p;
    const double C_m = 1.0;   // hard-coded in the reference (pcg2019.jl:55)
    const double phi = u[0], h = u[1], m = u[2], f = u[3], s = u[4], xs = u[5], xr = u[6];
    // instantaneous gates
    const double r_inf = tb_sigmoid(phi, p[E_r], p[k_r], -1.0);
    const double d_inf = tb_sigmoid(phi, p[E_d], p[k_d], -1.0);
    const double z_inf = tb_sigmoid(phi, p[E_z], p[k_z], 1.0);
    const double y_inf = tb_sigmoid(phi, p[E_y], p[k_y], 1.0);
    // currents
    const double I_Na = p[g_Na] * m * m * m * h * h * (phi - p[E_Na]);
    const double I_K1 = p[g_K1] * z_inf * (phi - p[E_K]);
    const double I_to = p[g_to] * r_inf * s * (phi - p[E_K]);
    const double I_CaL = p[g_CaL] * d_inf * f * (phi - p[E_Ca]);
    const double I_Kr = p[g_Kr] * xr * y_inf * (phi - p[E_K]);
    const double I_Ks = p[g_Ks] * xs * (phi - p[E_K]);
    const double I_total = I_Na + I_K1 + I_to + I_CaL + I_Kr + I_Ks;
    du[0] = -I_total / C_m;
    const double tau_h_ = (2.0 * p[tau_h0] * exp(p[delta_h] * (phi - p[E_h]) / p[k_h])) / (1.0 + exp((phi - p[E_h]) / p[k_h]));
    const double h_inf = tb_sigmoid(phi, p[E_h], p[k_h], 1.0);
    du[1] = (h_inf - h) / tau_h_;
    const double m_inf = tb_sigmoid(phi, p[E_m], p[k_m], -1.0);
    du[2] = (m_inf - m) / p[tau_m];
    const double f_inf = tb_sigmoid(phi, p[E_f], p[k_f], 1.0);
    du[3] = (f_inf - f) / p[tau_f];
    const double s_inf = tb_sigmoid(phi, p[E_s], p[k_s], 1.0);
    du[4] = (s_inf - s) / p[tau_s];
    const double xs_inf = tb_sigmoid(phi, p[E_xs], p[k_xs], -1.0);
    du[5] = (xs_inf - xs) / p[tau_xs];
    const double xr_inf = tb_sigmoid(phi, p[E_xr], p[k_xr], -1.0);
    du[6] = (xr_inf - xr) / p[tau_xr];
}

// One node, one outer step.  ADAPTIVE = false: ForwardEulerCellSolver (partitioned_solver.jl:80-99);
// true: AdaptiveForwardEulerSubstepper (:196-234).  The node's state stays in registers across all
// sub-steps; nothing but the final state goes back to memory.  Returns the phi component of the last
// rhs evaluation (what the reference leaves in cache.du for the ReactionTangentController).
template <int MODEL, bool ADAPTIVE>
TB_HD double tb_cell_node_step(const tb_cell_params &prm, double *u, double t, double dt, int substeps, double thr) {
    constexpr int NS = tb_cell_traits<MODEL>::NS;
    constexpr int PHI = tb_cell_traits<MODEL>::PHI;
    double du[NS];
    tb_cell_rhs<MODEL>(prm, u, t, du);
    if (!ADAPTIVE || fabs(du[PHI]) < thr) {
#pragma unroll
        for (int j = 0; j < NS; j++) u[j] += dt * du[j];
    } else {
        const double dts = dt / substeps;
#pragma unroll
        for (int j = 0; j < NS; j++) u[j] += dts * du[j];
        for (int k = 2; k <= substeps; k++) {
            const double ts = t + (k - 1) * dts;
            tb_cell_rhs<MODEL>(prm, u, ts, du);
#pragma unroll
            for (int j = 0; j < NS; j++) u[j] += dts * du[j];
        }
    }
    return du[PHI];
}
