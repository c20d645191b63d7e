// Element kernels of the three weak forms, as host/device inline templates:
//   mass       Me[i][j] += rho (Ni Nj) dOmega                    src/modeling/core/mass.jl:28-43
//   diffusion  Ke[i][j] -= ((gradNj . D) . gradNi) dOmega        src/modeling/core/diffusion.jl:28-50, utils.jl:409
//   source     be[j]    += f(x_q,t) Nj dOmega                    src/modeling/core/analytical_coefficient.jl:80-101
// with the geometry mapping of Ferrite's reinit! (device twin src/ferrite-addons/PR883.jl:254-291,367-387):
//   J = sum_a x_a (x) dM_a/dxi,  dOmega = det(J) w_q,  gradN_a = dN_a/dxi . J^-1.
// Quadrature/shape tables follow Ferrite 1.6.0 (QuadratureRule, Lagrange{refshape,1}); see tb_build_tables.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

#define TB_MAXQ 64

struct tb_elem_tables {
    int nv, dim, nq, pad;
    double w[TB_MAXQ];
    double N[TB_MAXQ * 8];        // N[q*nv + a]
    double dN[TB_MAXQ * 8 * 3];   // dN[(q*nv + a)*dim + d]
    double xi[TB_MAXQ * 3];       // reference points (for tb_quadrature)
};

// what the element kernels read: the first nq entries of the tables (a kernel stages just these in smem)
struct tb_tables_view {
    int nq;
    const double *w, *N, *dN;
};
static inline tb_tables_view tb_view_of(const tb_elem_tables *T) { return tb_tables_view{T->nq, T->w, T->N, T->dN}; }

// ---- host-side table construction --------------------------------------------------------------
static inline int tb_gauss_1d(int order, double *p, double *w) {
    switch (order) {
    case 1: p[0] = 0.0; w[0] = 2.0; return 1;
    case 2: p[0] = -0.5773502691896257645; p[1] = 0.5773502691896257645; w[0] = w[1] = 1.0; return 2;
    case 3:
        p[0] = -0.7745966692414833770; p[1] = 0.0; p[2] = 0.7745966692414833770;
        w[0] = w[2] = 0.5555555555555555556; w[1] = 0.8888888888888888889; return 3;
    case 4:
        p[0] = -0.8611363115940525752; p[1] = -0.3399810435848562648;
        p[2] = 0.3399810435848562648;  p[3] = 0.8611363115940525752;
        w[0] = w[3] = 0.3478548451374538574; w[1] = w[2] = 0.6521451548625461426; return 4;
    default: return 0;
    }
}

// celltype: 0 quad, 1 hex, 2 tri, 3 tet.  Returns 0 on success.
static inline int tb_build_tables(int celltype, int qorder, tb_elem_tables *T) {
    const int dim = (celltype == 0 || celltype == 2) ? 2 : 3;
    const int nv = celltype == 0 ? 4 : celltype == 1 ? 8 : celltype == 2 ? 3 : 4;
    T->nv = nv;
    T->dim = dim;
    T->pad = 0;
    int nq = 0;
    if (celltype == 0 || celltype == 1) {
        double p1[4], w1[4];
        const int n = tb_gauss_1d(qorder, p1, w1);
        if (!n) return 1;
        for (int k = 0; k < (dim == 3 ? n : 1); k++)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++) {   // first coordinate fastest
                    T->xi[nq * dim + 0] = p1[i];
                    T->xi[nq * dim + 1] = p1[j];
                    if (dim == 3) T->xi[nq * dim + 2] = p1[k];
                    T->w[nq] = dim == 3 ? w1[i] * w1[j] * w1[k] : w1[i] * w1[j];
                    nq++;
                }
    } else if (celltype == 3) {
        if (qorder == 1) {
            T->xi[0] = T->xi[1] = T->xi[2] = 0.25;
            T->w[0] = 1.0 / 6.0;
            nq = 1;
        } else if (qorder == 2) {
            const double a = 0.1381966011250105, b = 0.5854101966249685;
            const double P[4][3] = {{a, a, a}, {a, a, b}, {a, b, a}, {b, a, a}};
            for (int q = 0; q < 4; q++) {
                for (int d = 0; d < 3; d++) T->xi[q * 3 + d] = P[q][d];
                T->w[q] = 1.0 / 24.0;
            }
            nq = 4;
        } else return 1;
    } else {
        if (qorder == 1) {
            T->xi[0] = T->xi[1] = 1.0 / 3.0;
            T->w[0] = 0.5;
            nq = 1;
        } else if (qorder == 2) {
            const double P[3][2] = {{1.0 / 6.0, 1.0 / 6.0}, {2.0 / 3.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0}};
            for (int q = 0; q < 3; q++) {
                T->xi[q * 2] = P[q][0];
                T->xi[q * 2 + 1] = P[q][1];
                T->w[q] = 1.0 / 6.0;
            }
            nq = 3;
        } else return 1;
    }
    T->nq = nq;
    for (int q = 0; q < nq; q++) {
        const double *xi = T->xi + q * dim;
        double *N = T->N + q * nv, *dN = T->dN + q * nv * dim;
        if (celltype == 0) {
            const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
            for (int a = 0; a < 4; a++) {
                N[a] = 0.25 * (1 + sx[a] * xi[0]) * (1 + sy[a] * xi[1]);
                dN[a * 2 + 0] = 0.25 * sx[a] * (1 + sy[a] * xi[1]);
                dN[a * 2 + 1] = 0.25 * (1 + sx[a] * xi[0]) * sy[a];
            }
        } else if (celltype == 1) {
            const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1},
                         sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
            for (int a = 0; a < 8; a++) {
                const double fx = 1 + sx[a] * xi[0], fy = 1 + sy[a] * xi[1], fz = 1 + sz[a] * xi[2];
                N[a] = 0.125 * fx * fy * fz;
                dN[a * 3 + 0] = 0.125 * sx[a] * fy * fz;
                dN[a * 3 + 1] = 0.125 * fx * sy[a] * fz;
                dN[a * 3 + 2] = 0.125 * fx * fy * sz[a];
            }
        } else if (celltype == 3) {
            N[0] = 1 - xi[0] - xi[1] - xi[2]; N[1] = xi[0]; N[2] = xi[1]; N[3] = xi[2];
            const double g[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            for (int a = 0; a < 4; a++)
                for (int d = 0; d < 3; d++) dN[a * 3 + d] = g[a][d];
        } else {
            N[0] = xi[0]; N[1] = xi[1]; N[2] = 1 - xi[0] - xi[1];
            const double g[3][2] = {{1, 0}, {0, 1}, {-1, -1}};
            for (int a = 0; a < 3; a++)
                for (int d = 0; d < 2; d++) dN[a * 2 + d] = g[a][d];
        }
    }
    return 0;
}

// ---- geometry at one quadrature point --------------------------------------------------------------
// X[(a*DIM+d)*XS]: vertex coordinates (XS = element stride of the staging buffer, 1 on the host).
// dNq: dN/dxi of this point [a*DIM+d].  G (nullable): spatial gradients [a*DIM+d].  Returns det J.
template <int NV, int DIM, int XS, bool WANT_G>
TB_HD double tb_map_qp(const double *X, const double *dNq, double *G) {
    double J[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM * DIM; i++) J[i] = 0.0;
#pragma unroll
    for (int a = 0; a < NV; a++)
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            const double xa = X[(a * DIM + i) * XS];
#pragma unroll
            for (int j = 0; j < DIM; j++) J[i * DIM + j] += xa * dNq[a * DIM + j];
        }
    double det;
    double Ji[DIM * DIM];
    if constexpr (DIM == 2) {
        det = J[0] * J[3] - J[1] * J[2];
        if (WANT_G) {
            const double dinv = 1.0 / det;   // Tensors.jl inv(::Tensor{2,dim}): dinv = 1 / det(t), every cofactor TIMES dinv
            Ji[0] = J[3] * dinv; Ji[1] = -J[1] * dinv; Ji[2] = -J[2] * dinv; Ji[3] = J[0] * dinv;
        }
    } else {
        const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
        det = J[0] * c00 + J[1] * c01 + J[2] * c02;
        if (WANT_G) {
            const double dinv = 1.0 / det;
            Ji[0] = c00 * dinv; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * dinv; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * dinv;
            Ji[3] = c01 * dinv; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * dinv; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * dinv;
            Ji[6] = c02 * dinv; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * dinv; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * dinv;
        }
    }
    if (WANT_G) {
#pragma unroll
        for (int a = 0; a < NV; a++)
#pragma unroll
            for (int i = 0; i < DIM; i++) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < DIM; j++) s += dNq[a * DIM + j] * Ji[j * DIM + i];
                G[a * DIM + i] = s;
            }
    }
    return det;
}

// index of (i,j), i <= j, in the packed upper triangle
template <int NV> TB_HD constexpr int tb_sym(int i, int j) { return i * NV - (i * (i - 1)) / 2 + (j - i); }

// ---- diffusion tensor at a quadrature point ---------------------------------------------------------
// kind 0 scalar, 1 constant tensor, 2 spectral (lambda[3] + per cell per node f,s,n); see tbolt_b200.h.
// D = kappa/(Cm*chi) (coefficients.jl:152-162).  Spectral: FieldCoefficient interpolation
// (coefficients.jl:88-99), orthogonalize_system (microstructure.jl:176-187, utils.jl:131-139),
// sum lambda_i v_i (x) v_i (microstructure.jl:136-138).
template <int NV, int DIM>
TB_HD void tb_eval_D(int kind, const double *data, double cmchi, int64_t cell, const double *Nq, double *D) {
    if (kind == 0) {
#pragma unroll
        for (int i = 0; i < DIM * DIM; i++) D[i] = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; i++) D[i * DIM + i] = data[0] / cmchi;
    } else if (kind == 1) {
#pragma unroll
        for (int i = 0; i < DIM * DIM; i++) D[i] = data[i] / cmchi;
    } else if (kind == 3) {   // one tensor per cell (a piecewise-constant AnalyticalCoefficient / per-subdomain conductivity)
#pragma unroll
        for (int i = 0; i < DIM * DIM; i++) D[i] = data[cell * (DIM * DIM) + i] / cmchi;
    } else if constexpr (DIM == 3) {
        const double *lam = data;
        const double *fsn = data + 3 + cell * NV * 9;
        double v[3][3];
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int d = 0; d < 3; d++) v[m][d] = 0.0;
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int a = 0; a < NV; a++)
#pragma unroll
                for (int d = 0; d < 3; d++) v[m][d] += Nq[a] * fsn[a * 9 + m * 3 + d];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const double nrm = sqrt(v[m][0] * v[m][0] + v[m][1] * v[m][1] + v[m][2] * v[m][2]);
#pragma unroll
            for (int d = 0; d < 3; d++) v[m][d] /= nrm;
        }
        double w1[3], w2[3], w3[3];
#pragma unroll
        for (int d = 0; d < 3; d++) w1[d] = v[0][d];
        const double d12 = w1[0] * v[1][0] + w1[1] * v[1][1] + w1[2] * v[1][2];
#pragma unroll
        for (int d = 0; d < 3; d++) w2[d] = v[1][d] - d12 * w1[d];
        const double d13 = w1[0] * v[2][0] + w1[1] * v[2][1] + w1[2] * v[2][2];
        const double d23 = w2[0] * v[2][0] + w2[1] * v[2][1] + w2[2] * v[2][2];
#pragma unroll
        for (int d = 0; d < 3; d++) w3[d] = v[2][d] - d13 * w1[d] - d23 * w2[d];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                D[i * DIM + j] = (lam[0] * w1[i] * w1[j] + lam[1] * w2[i] * w2[j] + lam[2] * w3[i] * w3[j]) / cmchi;
    }
}

// ---- element matrices (packed upper triangle, NV*(NV+1)/2 entries) -----------------------------------
template <int NV, int DIM, int XS>
TB_HD void tb_element_mass(const tb_tables_view T, const double *X, double rho, double *acc) {
#pragma unroll
    for (int i = 0; i < NV * (NV + 1) / 2; i++) acc[i] = 0.0;
    for (int q = 0; q < T.nq; q++) {
        const double *Nq = T.N + q * NV;
        const double dO = tb_map_qp<NV, DIM, XS, false>(X, T.dN + q * NV * DIM, nullptr) * T.w[q];
#pragma unroll
        for (int i = 0; i < NV; i++)
#pragma unroll
            for (int j = i; j < NV; j++) acc[tb_sym<NV>(i, j)] += rho * (Nq[i] * Nq[j]) * dO;
    }
}

template <int NV, int DIM, int XS>
TB_HD void tb_element_diffusion(const tb_tables_view T, const double *X, int kind, const double *data, double cmchi,
                                int64_t cell, double *acc) {
#pragma unroll
    for (int i = 0; i < NV * (NV + 1) / 2; i++) acc[i] = 0.0;
    double D[DIM * DIM];
    if (kind < 2) tb_eval_D<NV, DIM>(kind, data, cmchi, cell, T.N, D);   // constant coefficients: same value at every point
    for (int q = 0; q < T.nq; q++) {
        double G[NV * DIM];
        const double dO = tb_map_qp<NV, DIM, XS, true>(X, T.dN + q * NV * DIM, G) * T.w[q];
        if (kind >= 2) tb_eval_D<NV, DIM>(kind, data, cmchi, cell, T.N + q * NV, D);
#pragma unroll
        for (int j = 0; j < NV; j++) {
            if (kind == 0) {
                // _inner_product_helper(a, B::AbstractFloat, c) = a . c * B
#pragma unroll
                for (int i = 0; i <= j; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) s += G[j * DIM + d] * G[i * DIM + d];
                    acc[tb_sym<NV>(i, j)] -= (s * D[0]) * dO;
                }
            } else {
                double gD[DIM];   // gradNj . D
#pragma unroll
                for (int l = 0; l < DIM; l++) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; k++) s += G[j * DIM + k] * D[k * DIM + l];
                    gD[l] = s;
                }
#pragma unroll
                for (int i = 0; i <= j; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < DIM; l++) s += gD[l] * G[i * DIM + l];
                    acc[tb_sym<NV>(i, j)] -= s * dO;
                }
            }
        }
    }
}

// Full NV x NV diffusion matrix, every (i, j) pair evaluated on its own exactly as the reference's double
// loop does (diffusion.jl:40-47: Ke[i,j] -= ((gradNj . D) . gradNi) dOmega); Ke is NOT bitwise symmetric in
// floating point, and the gather assembly (tb_assembly.cu) reproduces the CPU path bit for bit with this.
template <int NV, int DIM, int XS>
TB_HD void tb_element_diffusion_full(const tb_tables_view T, const double *X, int kind, const double *data, double cmchi,
                                     int64_t cell, double *Ke, const double *Dconst = nullptr) {
#pragma unroll
    for (int i = 0; i < NV * NV; i++) Ke[i] = 0.0;
    // constant coefficients (kind 0, 1) have the same value at every point: evaluated once -- by the caller into shared
    // memory (Dconst; no registers held across the loop) or here
    double Dloc[DIM * DIM];
    if (kind < 2 && !Dconst) tb_eval_D<NV, DIM>(kind, data, cmchi, cell, T.N, Dloc);
    for (int q = 0; q < T.nq; q++) {
        double G[NV * DIM];
        const double dO = tb_map_qp<NV, DIM, XS, true>(X, T.dN + q * NV * DIM, G) * T.w[q];
        if (kind >= 2) tb_eval_D<NV, DIM>(kind, data, cmchi, cell, T.N + q * NV, Dloc);
        const double *D = (kind < 2 && Dconst) ? Dconst : Dloc;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            if (kind == 0) {
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; d++) s += G[j * DIM + d] * G[i * DIM + d];
                    Ke[i * NV + j] -= (s * D[0]) * dO;
                }
            } else {
                double gD[DIM];
#pragma unroll
                for (int l = 0; l < DIM; l++) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; k++) s += G[j * DIM + k] * D[k * DIM + l];
                    gD[l] = s;
                }
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < DIM; l++) s += gD[l] * G[i * DIM + l];
                    Ke[i * NV + j] -= s * dO;
                }
            }
        }
    }
}

// ---- source programs: a closure f(x, t) traced into postfix code (tbolt_b200.h TB_SRC_PROGRAM) ---------------
// A stimulus closure cannot cross a C ABI, but its expression can: the host binding calls f once with tracing numbers
// and ships the resulting postfix program (<= TB_PROG_MAXCODE instructions, <= TB_PROG_MAXCONST constants); the
// element kernels evaluate it at every quadrature point.  All arithmetic is plain IEEE fp64 in program order (no
// contraction), so +,-,*,/,sqrt,abs,min,max and the comparisons give the bits the closure itself would give on the
// host; exp/log/sin/cos/tanh/pow agree to the ulp of the two math libraries.
#define TB_PROG_MAXCODE 96
#define TB_PROG_MAXCONST 24
#define TB_PROG_MAXSTACK 16
enum tb_prog_op {
    TB_OP_X = 0,      // push x[arg]
    TB_OP_T,          // push t
    TB_OP_CONST,      // push c[arg]
    TB_OP_ADD, TB_OP_SUB, TB_OP_MUL, TB_OP_DIV, TB_OP_MIN, TB_OP_MAX, TB_OP_POW,        // binary: a = pop-under, b = top
    TB_OP_LT, TB_OP_LE, TB_OP_GT, TB_OP_GE, TB_OP_EQ, TB_OP_NE, TB_OP_AND, TB_OP_OR,    // binary, result 1.0 / 0.0
    TB_OP_NEG, TB_OP_ABS, TB_OP_SQRT, TB_OP_EXP, TB_OP_LOG, TB_OP_SIN, TB_OP_COS, TB_OP_TANH, TB_OP_NOT,   // unary
    TB_OP_SELECT,     // ternary: cond, a, b -> cond != 0 ? a : b
    TB_OP_COUNT
};
struct tb_src_program {
    int n;
    unsigned char op[TB_PROG_MAXCODE];
    unsigned char arg[TB_PROG_MAXCODE];
    double c[TB_PROG_MAXCONST];
};

// Validates code (op | arg << 8 per instruction) and fills P; returns 0, or a negative code: -1 too long, -2 unknown
// opcode, -3 bad operand, -4 stack underflow/overflow, -5 does not leave exactly one value.
static inline int tb_program_build(const int32_t *code, int ncode, const double *consts, int nconsts, int dim,
                                   tb_src_program *P) {
    if (ncode < 1 || ncode > TB_PROG_MAXCODE || nconsts < 0 || nconsts > TB_PROG_MAXCONST) return -1;
    int sp = 0;
    for (int i = 0; i < ncode; i++) {
        const int op = code[i] & 0xff, arg = (code[i] >> 8) & 0xff;
        if (op < 0 || op >= TB_OP_COUNT) return -2;
        if (op == TB_OP_X && arg >= dim) return -3;
        if (op == TB_OP_CONST && arg >= nconsts) return -3;
        const int pops = op <= TB_OP_CONST ? 0 : op <= TB_OP_OR ? 2 : op <= TB_OP_NOT ? 1 : 3;
        if (sp < pops) return -4;
        sp += 1 - pops;
        if (sp > TB_PROG_MAXSTACK) return -4;
        P->op[i] = (unsigned char)op;
        P->arg[i] = (unsigned char)arg;
    }
    if (sp != 1) return -5;
    P->n = ncode;
    for (int i = 0; i < nconsts; i++) P->c[i] = consts[i];
    for (int i = nconsts; i < TB_PROG_MAXCONST; i++) P->c[i] = 0.0;
    return 0;
}

template <int DIM> TB_HD double tb_program_eval(const tb_src_program &P, const double *x, double t) {
    double st[TB_PROG_MAXSTACK];
    int sp = 0;
    for (int pc = 0; pc < P.n; pc++) {
        const int op = P.op[pc];
        if (op <= TB_OP_CONST) {
            st[sp++] = op == TB_OP_X ? x[P.arg[pc] < DIM ? P.arg[pc] : 0] : op == TB_OP_T ? t : P.c[P.arg[pc]];
        } else if (op <= TB_OP_OR) {
            const double b = st[--sp], a = st[sp - 1];
            double r;
            switch (op) {
            case TB_OP_ADD: r = a + b; break;
            case TB_OP_SUB: r = a - b; break;
            case TB_OP_MUL: r = a * b; break;
            case TB_OP_DIV: r = a / b; break;
            case TB_OP_MIN: r = b < a ? b : a; break;
            case TB_OP_MAX: r = a < b ? b : a; break;
            case TB_OP_POW: r = pow(a, b); break;
            case TB_OP_LT: r = a < b ? 1.0 : 0.0; break;
            case TB_OP_LE: r = a <= b ? 1.0 : 0.0; break;
            case TB_OP_GT: r = a > b ? 1.0 : 0.0; break;
            case TB_OP_GE: r = a >= b ? 1.0 : 0.0; break;
            case TB_OP_EQ: r = a == b ? 1.0 : 0.0; break;
            case TB_OP_NE: r = a != b ? 1.0 : 0.0; break;
            case TB_OP_AND: r = (a != 0.0 && b != 0.0) ? 1.0 : 0.0; break;
            default: r = (a != 0.0 || b != 0.0) ? 1.0 : 0.0; break;
            }
            st[sp - 1] = r;
        } else if (op <= TB_OP_NOT) {
            const double a = st[sp - 1];
            double r;
            switch (op) {
            case TB_OP_NEG: r = -a; break;
            case TB_OP_ABS: r = fabs(a); break;
            case TB_OP_SQRT: r = sqrt(a); break;
            case TB_OP_EXP: r = exp(a); break;
            case TB_OP_LOG: r = log(a); break;
            case TB_OP_SIN: r = sin(a); break;
            case TB_OP_COS: r = cos(a); break;
            case TB_OP_TANH: r = tanh(a); break;
            default: r = a == 0.0 ? 1.0 : 0.0; break;
            }
            st[sp - 1] = r;
        } else {
            const double b = st[--sp], a = st[--sp], c = st[sp - 1];
            st[sp - 1] = c != 0.0 ? a : b;
        }
    }
    return st[0];
}

// ---- built-in stimulus families (see tbolt_b200.h TB_SRC_*) --------------------------------------------
template <int DIM> TB_HD double tb_source_eval(int kind, const double *prm, const double *x, double t) {
    double n2 = 0.0, mx = -INFINITY;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        n2 += x[d] * x[d];
        mx = x[d] > mx ? x[d] : mx;
    }
    switch (kind) {
    case 1: return (mx < prm[0] && t < prm[1]) ? prm[2] : 0.0;
    case 2: return (sqrt(n2) < prm[0] && t < prm[1]) ? prm[2] : 0.0;
    case 3: return cos(2.0 * 3.14159265358979323846 * t) * exp(-n2);
    case 4: return sqrt(n2) + t;
    case 5: return (t <= prm[1] && x[0] < prm[0]) ? prm[2] / prm[3] * exp(t / prm[3]) : 0.0;
    default: return 0.0;
    }
}

// fq != nullptr: host-evaluated f at this cell's quadrature points; prog != nullptr: traced closure
template <int NV, int DIM, int XS>
TB_HD void tb_element_source(const tb_tables_view T, const double *X, int kind, const double *prm, double t,
                             const double *fq, double *be, const tb_src_program *prog = nullptr) {
#pragma unroll
    for (int j = 0; j < NV; j++) be[j] = 0.0;
    for (int q = 0; q < T.nq; q++) {
        const double *Nq = T.N + q * NV;
        const double dO = tb_map_qp<NV, DIM, XS, false>(X, T.dN + q * NV * DIM, nullptr) * T.w[q];
        double fx;
        if (fq) {
            fx = fq[q];
        } else {
            double x[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) x[d] = 0.0;
#pragma unroll
            for (int a = 0; a < NV; a++)
#pragma unroll
                for (int d = 0; d < DIM; d++) x[d] += Nq[a] * X[(a * DIM + d) * XS];
            fx = prog ? tb_program_eval<DIM>(*prog, x, t) : tb_source_eval<DIM>(kind, prm, x, t);
        }
#pragma unroll
        for (int j = 0; j < NV; j++) be[j] += fx * Nq[j] * dO;
    }
}
