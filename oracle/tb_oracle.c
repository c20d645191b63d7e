/*
 * tb_oracle.c -- CPU restatement of Thunderbolt.jl's monodomain hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (thunderbolt.jl_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference is pure Julia and cannot be executed in this image
 * (no julia binary), and most of the arithmetic on this path lives in un-vendored
 * packages (Ferrite 1.6.0, FerriteOperators 0.3.6, LinearSolve 3.87.0 / Krylov 0.10.9,
 * OrdinaryDiffEqOperatorSplitting 0.4.0).  What IS pinned by the reference's own tests
 * (and reproduced in tests/test_oracle_known_answers.py):
 *   - substepper clock closed form        test/test_time_integrator.jl:282-294
 *   - backward-Euler steady state         test/test_time_integrator.jl:13-41
 *   - index sets / SoA layout             test/test_solution_variables.jl:76-127
 *   - PCG2019 default initial state       src/modeling/cells/pcg2019.jl:137-152
 *   - FE vs adaptive agree 1e-2/differ 1e-8  test/integration/test_electrophysiology.jl:90-95
 *   - distorted-hex geometry fixture      test/test_coefficients.jl:239-279
 * PARITY UNPINNED (restated from the published algorithms of the packages above, no
 * golden vectors exist in the reference): absolute phi_m trajectories, DoF numbering,
 * CSR pattern, quadrature point order, CG iteration counts.
 *
 * Conventions: all indices 0-based here (the reference is 1-based); fp64; compile
 * with -ffp-contract=off because Julia does not contract a*b+c into fma.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_QUAD4 0
#define ORC_HEX8 1
#define ORC_TRI3 2
#define ORC_TET4 3

#define ORC_MAXNV 8
#define ORC_MAXQ 64

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static int cell_nv(int ct) { return ct == ORC_QUAD4 ? 4 : ct == ORC_HEX8 ? 8 : ct == ORC_TRI3 ? 3 : 4; }
static int cell_dim(int ct) { return (ct == ORC_QUAD4 || ct == ORC_TRI3) ? 2 : 3; }
int orc_cell_nv(int ct) { return cell_nv(ct); }
int orc_cell_dim(int ct) { return cell_dim(ct); }

/* ------------------------------------------------------------------------------------------
 * Mesh generators: Ferrite 1.6.0 generate_grid (called at src/mesh/generators.jl:942).
 * Nodes x-fastest; Quadrilateral loops `for j, i`, Hexahedron `for k, j, i`; Triangle = 2 per
 * quad; Tetrahedron = 6 per hex with the split listed in SURVEY.md section 8c.
 * ---------------------------------------------------------------------------------------- */
void orc_grid_sizes(int ct, const int64_t *nel, int64_t *ncells, int64_t *nnodes) {
    if (cell_dim(ct) == 2) {
        *nnodes = (nel[0] + 1) * (nel[1] + 1);
        *ncells = nel[0] * nel[1] * (ct == ORC_TRI3 ? 2 : 1);
    } else {
        *nnodes = (nel[0] + 1) * (nel[1] + 1) * (nel[2] + 1);
        *ncells = nel[0] * nel[1] * nel[2] * (ct == ORC_TET4 ? 6 : 1);
    }
}

/* coordinate of grid line i of n cells between l and r.  Ferrite uses range(l, stop=r, length=n+1)
 * in 3D and a bilinear corner blend in 2D; for an axis aligned box both reduce to l + i*(r-l)/n up
 * to the last bit.  We fix the formula below for oracle AND product (coordinates are inputs of the
 * path, not results, so only self-consistency matters). */
static double grid_coord(double l, double r, int64_t i, int64_t n) {
    if (i == n) return r;
    return l + ((double)i * (r - l)) / (double)n;
}

void orc_generate_grid(int ct, const int64_t *nel, const double *left, const double *right, int64_t *conn,
                       double *coords) {
    int dim = cell_dim(ct);
    int64_t nx = nel[0] + 1, ny = nel[1] + 1, nz = dim == 3 ? nel[2] + 1 : 1;
    for (int64_t k = 0; k < nz; k++)
        for (int64_t j = 0; j < ny; j++)
            for (int64_t i = 0; i < nx; i++) {
                int64_t n = (k * ny + j) * nx + i;
                coords[n * dim + 0] = grid_coord(left[0], right[0], i, nel[0]);
                coords[n * dim + 1] = grid_coord(left[1], right[1], j, nel[1]);
                if (dim == 3) coords[n * dim + 2] = grid_coord(left[2], right[2], k, nel[2]);
            }
#define ND(i, j, k) ((((int64_t)(k)) * ny + (j)) * nx + (i))
    int64_t c = 0;
    if (dim == 2) {
        for (int64_t j = 0; j < nel[1]; j++)
            for (int64_t i = 0; i < nel[0]; i++) {
                if (ct == ORC_QUAD4) {
                    conn[c * 4 + 0] = ND(i, j, 0);
                    conn[c * 4 + 1] = ND(i + 1, j, 0);
                    conn[c * 4 + 2] = ND(i + 1, j + 1, 0);
                    conn[c * 4 + 3] = ND(i, j + 1, 0);
                    c++;
                } else {
                    conn[c * 3 + 0] = ND(i, j, 0);
                    conn[c * 3 + 1] = ND(i + 1, j, 0);
                    conn[c * 3 + 2] = ND(i, j + 1, 0);
                    c++;
                    conn[c * 3 + 0] = ND(i + 1, j, 0);
                    conn[c * 3 + 1] = ND(i + 1, j + 1, 0);
                    conn[c * 3 + 2] = ND(i, j + 1, 0);
                    c++;
                }
            }
    } else {
        static const int tets[6][4] = {{0, 1, 3, 7}, {0, 4, 1, 7}, {1, 2, 3, 7}, {1, 6, 2, 7}, {1, 4, 5, 7}, {1, 5, 6, 7}};
        for (int64_t k = 0; k < nel[2]; k++)
            for (int64_t j = 0; j < nel[1]; j++)
                for (int64_t i = 0; i < nel[0]; i++) {
                    int64_t h[8] = {ND(i, j, k),         ND(i + 1, j, k),         ND(i + 1, j + 1, k),
                                    ND(i, j + 1, k),     ND(i, j, k + 1),         ND(i + 1, j, k + 1),
                                    ND(i + 1, j + 1, k + 1), ND(i, j + 1, k + 1)};
                    if (ct == ORC_HEX8) {
                        for (int a = 0; a < 8; a++) conn[c * 8 + a] = h[a];
                        c++;
                    } else {
                        for (int s = 0; s < 6; s++) {
                            for (int a = 0; a < 4; a++) conn[c * 4 + a] = h[tets[s][a]];
                            c++;
                        }
                    }
                }
    }
#undef ND
}

/* ------------------------------------------------------------------------------------------
 * DoF numbering: Ferrite 1.6.0 DofHandler close! for one scalar Lagrange-1 field
 * (src/discretization/fem.jl:180-182): cells in order, local vertices in order, a vertex gets the
 * next free id on first touch.  Returns ndofs.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_close_dofs(int64_t ncells, int nv, const int64_t *conn, int64_t nnodes, int64_t *celldofs,
                       int64_t *node2dof) {
    for (int64_t n = 0; n < nnodes; n++) node2dof[n] = -1;
    int64_t next = 0;
    for (int64_t c = 0; c < ncells; c++)
        for (int a = 0; a < nv; a++) {
            int64_t n = conn[c * nv + a];
            if (node2dof[n] < 0) node2dof[n] = next++;
            celldofs[c * nv + a] = node2dof[n];
        }
    return next;
}

/* ------------------------------------------------------------------------------------------
 * Sparsity pattern: Ferrite allocate_matrix(dh) -> CSC with sorted rows, all dof pairs sharing a
 * cell incl. the diagonal; transposed into CSR by src/solver/interface.jl:159-168 (the pattern is
 * symmetric so rowptr = colptr, colval = rowval).  Two calls: colidx == NULL counts.
 * ---------------------------------------------------------------------------------------- */
static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : x > y;
}

/* dof -> cells adjacency (cells ascending per dof).  adjptr: ndofs+1, adj: ncells*nv. */
static void dof_cell_adjacency(int64_t ndofs, int64_t ncells, int nv, const int64_t *celldofs, int64_t *adjptr, int64_t *adj) {
    for (int64_t d = 0; d <= ndofs; d++) adjptr[d] = 0;
    for (int64_t p = 0; p < ncells * nv; p++) adjptr[celldofs[p] + 1]++;
    for (int64_t d = 0; d < ndofs; d++) adjptr[d + 1] += adjptr[d];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ndofs + 1));
    memcpy(cur, adjptr, sizeof(int64_t) * (size_t)ndofs);
    for (int64_t c = 0; c < ncells; c++)
        for (int a = 0; a < nv; a++) adj[cur[celldofs[c * nv + a]]++] = c * nv + a;
    free(cur);
}

/* sorted union of the dofs of the cells around dof d; returns its size */
static int64_t pattern_row(int64_t d, int nv, const int64_t *celldofs, const int64_t *adjptr, const int64_t *adj, int64_t **buf,
                           int64_t *cap, int64_t *out) {
    int64_t m = 0;
    const int64_t need = (adjptr[d + 1] - adjptr[d]) * nv;
    if (need > *cap) {
        *cap = need * 2;
        *buf = (int64_t *)realloc(*buf, sizeof(int64_t) * (size_t)*cap);
    }
    int64_t *b = *buf;
    for (int64_t q = adjptr[d]; q < adjptr[d + 1]; q++)
        for (int a = 0; a < nv; a++) b[m++] = celldofs[(adj[q] / nv) * nv + a];
    qsort(b, (size_t)m, sizeof(int64_t), cmp_i64);
    int64_t u = 0;
    for (int64_t q = 0; q < m; q++)
        if (q == 0 || b[q] != b[q - 1]) {
            if (out) out[u] = b[q];
            u++;
        }
    return u;
}

/* Rows are independent, so the two passes (count, fill) run threads over rows; the result does not depend on the
 * thread count. */
int64_t orc_pattern(int64_t ndofs, int64_t ncells, int nv, const int64_t *celldofs, int64_t *rowptr, int64_t *colidx) {
    int64_t *adjptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ndofs + 1));
    int64_t *adj = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncells * nv + 1));
    dof_cell_adjacency(ndofs, ncells, nv, celldofs, adjptr, adj);
    rowptr[0] = 0;
#pragma omp parallel
    {
        int64_t cap = 1024;
        int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
#pragma omp for schedule(static)
        for (int64_t d = 0; d < ndofs; d++) rowptr[d + 1] = pattern_row(d, nv, celldofs, adjptr, adj, &buf, &cap, NULL);
        free(buf);
    }
    for (int64_t d = 0; d < ndofs; d++) rowptr[d + 1] += rowptr[d];
    const int64_t nnz = rowptr[ndofs];
    if (colidx) {
#pragma omp parallel
        {
            int64_t cap = 1024;
            int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
#pragma omp for schedule(static)
            for (int64_t d = 0; d < ndofs; d++) pattern_row(d, nv, celldofs, adjptr, adj, &buf, &cap, colidx + rowptr[d]);
            free(buf);
        }
    }
    free(adj);
    free(adjptr);
    return nnz;
}

/* ------------------------------------------------------------------------------------------
 * Quadrature + Lagrange-1 tables (Ferrite 1.6.0 QuadratureRule / Lagrange; used through
 * CellValues at src/modeling/core/mass.jl:45-55, diffusion.jl:52-60).
 *   hypercube: Gauss-Legendre `order` points per direction on [-1,1], first coordinate fastest
 *   tet order 1: centroid, w=1/6; order 2: 4 points a=0.1381966011250105 b=0.5854101966249685 w=1/24
 *   tri order 1: centroid w=1/2;  order 2: (1/6,1/6),(2/3,1/6),(1/6,2/3) w=1/6
 * Reference shapes: quad/hex [-1,1]^d, vertex order (--,+-,++,-+) and bottom-then-top;
 * RefTetrahedron vertices (0,0,0),(1,0,0),(0,1,0),(0,0,1) with N=(1-x-y-z, x, y, z);
 * RefTriangle vertices (1,0),(0,1),(0,0) with N=(x, y, 1-x-y).
 * ---------------------------------------------------------------------------------------- */
static int gauss_1d(int order, double *p, double *w) {
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

/* returns nq (0 on unsupported); pts is nq x dim, w is nq */
int orc_quadrature(int ct, int order, double *pts, double *w) {
    int dim = cell_dim(ct);
    if (ct == ORC_QUAD4 || ct == ORC_HEX8) {
        double p1[8], w1[8];
        int n = gauss_1d(order, p1, w1);
        if (!n) return 0;
        int q = 0;
        for (int k = 0; k < (dim == 3 ? n : 1); k++)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++) {
                    pts[q * dim + 0] = p1[i];
                    pts[q * dim + 1] = p1[j];
                    if (dim == 3) pts[q * dim + 2] = p1[k];
                    w[q] = dim == 3 ? w1[i] * w1[j] * w1[k] : w1[i] * w1[j];
                    q++;
                }
        return q;
    }
    if (ct == ORC_TET4) {
        if (order == 1) {
            pts[0] = pts[1] = pts[2] = 0.25; w[0] = 1.0 / 6.0; return 1;
        }
        if (order == 2) {
            const double a = 0.1381966011250105, b = 0.5854101966249685;
            const double P[4][3] = {{a, a, a}, {a, a, b}, {a, b, a}, {b, a, a}};
            for (int q = 0; q < 4; q++) {
                for (int d = 0; d < 3; d++) pts[q * 3 + d] = P[q][d];
                w[q] = 1.0 / 24.0;
            }
            return 4;
        }
        return 0;
    }
    if (ct == ORC_TRI3) {
        if (order == 1) {
            pts[0] = pts[1] = 1.0 / 3.0; w[0] = 0.5; return 1;
        }
        if (order == 2) {
            const double P[3][2] = {{1.0 / 6.0, 1.0 / 6.0}, {2.0 / 3.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0}};
            for (int q = 0; q < 3; q++) {
                pts[q * 2] = P[q][0]; pts[q * 2 + 1] = P[q][1];
                w[q] = 1.0 / 6.0;
            }
            return 3;
        }
        return 0;
    }
    return 0;
}

/* N[a], dN[a][d] at reference point xi */
void orc_shape(int ct, const double *xi, double *N, double *dN) {
    if (ct == ORC_QUAD4) {
        static const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
        for (int a = 0; a < 4; a++) {
            N[a] = 0.25 * (1 + sx[a] * xi[0]) * (1 + sy[a] * xi[1]);
            dN[a * 2 + 0] = 0.25 * sx[a] * (1 + sy[a] * xi[1]);
            dN[a * 2 + 1] = 0.25 * (1 + sx[a] * xi[0]) * sy[a];
        }
    } else if (ct == ORC_HEX8) {
        static const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1},
                            sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        for (int a = 0; a < 8; a++) {
            double fx = 1 + sx[a] * xi[0], fy = 1 + sy[a] * xi[1], fz = 1 + sz[a] * xi[2];
            N[a] = 0.125 * fx * fy * fz;
            dN[a * 3 + 0] = 0.125 * sx[a] * fy * fz;
            dN[a * 3 + 1] = 0.125 * fx * sy[a] * fz;
            dN[a * 3 + 2] = 0.125 * fx * fy * sz[a];
        }
    } else if (ct == ORC_TET4) {
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

/* Geometry mapping at one qp (Ferrite reinit! / src/ferrite-addons/PR883.jl:254-291,367-387):
 * J = sum_a x_a (x) dM_a/dxi, detJ, gradN_a = dN_a/dxi . J^-1.  X is nv x dim. */
static double map_qp(int nv, int dim, const double *X, const double *dN, double *G /* nv x dim or NULL */) {
    double J[9] = {0}, Ji[9];
    for (int a = 0; a < nv; a++)
        for (int i = 0; i < dim; i++)
            for (int j = 0; j < dim; j++) J[i * dim + j] += X[a * dim + i] * dN[a * dim + j];
    double det;
    if (dim == 2) {
        det = J[0] * J[3] - J[1] * J[2];
        if (G) { /* Tensors.jl inv(::Tensor{2,dim}): dinv = 1 / det(t), every cofactor TIMES dinv */
            double dinv = 1.0 / det;
            Ji[0] = J[3] * dinv; Ji[1] = -J[1] * dinv; Ji[2] = -J[2] * dinv; Ji[3] = J[0] * dinv;
        }
    } else {
        double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
        det = J[0] * c00 + J[1] * c01 + J[2] * c02;
        if (G) {
            double dinv = 1.0 / det;
            Ji[0] = c00 * dinv; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * dinv; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * dinv;
            Ji[3] = c01 * dinv; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * dinv; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * dinv;
            Ji[6] = c02 * dinv; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * dinv; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * dinv;
        }
    }
    if (G)
        for (int a = 0; a < nv; a++)
            for (int i = 0; i < dim; i++) {
                double s = 0;
                for (int j = 0; j < dim; j++) s += dN[a * dim + j] * Ji[j * dim + i];
                G[a * dim + i] = s;
            }
    return det;
}

/* exported for the geometry fixture test (test/test_coefficients.jl:239-279) */
double orc_map_qp(int ct, const double *X, const double *xi, double *N, double *G) {
    double dN[ORC_MAXNV * 3];
    orc_shape(ct, xi, N, dN);
    return map_qp(cell_nv(ct), cell_dim(ct), X, dN, G);
}

/* ------------------------------------------------------------------------------------------
 * Diffusion coefficient D = kappa/(Cm*chi) (src/modeling/core/coefficients.jl:152-162).
 *   kind 0: scalar  data[0]                       (ConstantCoefficient of a number)
 *   kind 1: constant symmetric tensor, data = dim x dim row-major   (coefficients.jl:106-120)
 *   kind 2: SpectralTensorCoefficient over an OrthotropicMicrostructureModel of three
 *           FieldCoefficients: data = lambda[3] then per cell, per local node a, f[3] s[3] n[3]
 *           (coefficients.jl:36-99,451-488; microstructure.jl:136-187; utils.jl:131-139)
 * cmchi = Cm*chi.  D out is dim x dim row-major.
 * ---------------------------------------------------------------------------------------- */
static void eval_D(int kind, const double *data, double cmchi, int dim, int nv, int64_t cell, const double *N, double *D) {
    if (kind == 0) {
        for (int i = 0; i < dim * dim; i++) D[i] = 0;
        for (int i = 0; i < dim; i++) D[i * dim + i] = data[0] / cmchi;
    } else if (kind == 1) {
        for (int i = 0; i < dim * dim; i++) D[i] = data[i] / cmchi;
    } else if (kind == 3) { /* one tensor per cell: a piecewise-constant coefficient field (e.g. the AnalyticalCoefficient of test_ecg.jl:27-33) */
        for (int i = 0; i < dim * dim; i++) D[i] = data[cell * dim * dim + i] / cmchi;
    } else {
        const double *lam = data;
        const double *fsn = data + 3 + cell * nv * 9;
        double v[3][3] = {{0}};
        for (int m = 0; m < 3; m++)
            for (int a = 0; a < nv; a++)
                for (int d = 0; d < 3; d++) v[m][d] += N[a] * fsn[a * 9 + m * 3 + d];
        /* orthogonalize_system: normalise each, then Gram-Schmidt without renormalising */
        for (int m = 0; m < 3; m++) {
            double nrm = sqrt(v[m][0] * v[m][0] + v[m][1] * v[m][1] + v[m][2] * v[m][2]);
            for (int d = 0; d < 3; d++) v[m][d] /= nrm;
        }
        double w1[3], w2[3], w3[3];
        for (int d = 0; d < 3; d++) w1[d] = v[0][d];
        double d12 = w1[0] * v[1][0] + w1[1] * v[1][1] + w1[2] * v[1][2];
        for (int d = 0; d < 3; d++) w2[d] = v[1][d] - d12 * w1[d];
        double d13 = w1[0] * v[2][0] + w1[1] * v[2][1] + w1[2] * v[2][2];
        double d23 = w2[0] * v[2][0] + w2[1] * v[2][1] + w2[2] * v[2][2];
        for (int d = 0; d < 3; d++) w3[d] = v[2][d] - d13 * w1[d] - d23 * w2[d];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                D[i * 3 + j] = (lam[0] * w1[i] * w1[j] + lam[1] * w2[i] * w2[j] + lam[2] * w3[i] * w3[j]) / cmchi;
    }
}

/* Element kernels.
 * mass:      Me[i][j] += rho*(Ni*Nj)*dOmega                 src/modeling/core/mass.jl:28-43
 * diffusion: Ke[i][j] -= ((gradNj . D) . gradNi)*dOmega     src/modeling/core/diffusion.jl:28-50, utils.jl:409
 */
void orc_element_mass(int ct, int qorder, const double *X, double rho, double *Me) {
    int nv = cell_nv(ct), dim = cell_dim(ct);
    double pts[ORC_MAXQ * 3], w[ORC_MAXQ], N[ORC_MAXNV], dN[ORC_MAXNV * 3];
    int nq = orc_quadrature(ct, qorder, pts, w);
    for (int i = 0; i < nv * nv; i++) Me[i] = 0;
    for (int q = 0; q < nq; q++) {
        orc_shape(ct, pts + q * dim, N, dN);
        double dO = map_qp(nv, dim, X, dN, NULL) * w[q];
        for (int i = 0; i < nv; i++)
            for (int j = 0; j < nv; j++) Me[i * nv + j] += rho * (N[i] * N[j]) * dO;
    }
}

void orc_element_diffusion(int ct, int qorder, const double *X, int kind, const double *data, double cmchi,
                           int64_t cell, double *Ke) {
    int nv = cell_nv(ct), dim = cell_dim(ct);
    double pts[ORC_MAXQ * 3], w[ORC_MAXQ], N[ORC_MAXNV], dN[ORC_MAXNV * 3], G[ORC_MAXNV * 3], D[9];
    int nq = orc_quadrature(ct, qorder, pts, w);
    for (int i = 0; i < nv * nv; i++) Ke[i] = 0;
    for (int q = 0; q < nq; q++) {
        orc_shape(ct, pts + q * dim, N, dN);
        double dO = map_qp(nv, dim, X, dN, G) * w[q];
        eval_D(kind, data, cmchi, dim, nv, cell, N, D);
        for (int i = 0; i < nv; i++)
            for (int j = 0; j < nv; j++) {
                double s = 0;
                if (kind == 0) {
                    for (int d = 0; d < dim; d++) s += G[j * dim + d] * G[i * dim + d];
                    s = s * D[0];
                } else {
                    double gD[3];
                    for (int l = 0; l < dim; l++) {
                        gD[l] = 0;
                        for (int k = 0; k < dim; k++) gD[l] += G[j * dim + k] * D[k * dim + l];
                    }
                    for (int l = 0; l < dim; l++) s += gD[l] * G[i * dim + l];
                }
                Ke[i * nv + j] -= s * dO;
            }
    }
}

/* Built-in stimulus families f(x,t) (closures cannot cross a C ABI).  prm layout documented in
 * include/tbolt_b200.h (TB_SRC_*).  The examples they restate:
 *   1 box:   max(x) < p0 && t < p1 ? p2 : 0      bak/examples/conduction-velocity-benchmark.jl:47-50
 *   2 ball:  norm(x) < p0 && t < p1 ? p2 : 0     test/integration/test_electrophysiology.jl:83
 *   3 cos(2 pi t) exp(-|x|^2)                    benchmarks/benchmarks-cuda-linear-form.jl:4-18
 *   4 |x| + t                                    benchmarks/benchmarks-linear-form.jl:16-21
 *   5 t <= p1 && x[0] < p0 ? p2/p3*exp(t/p3) : 0 docs/src/literate-tutorials/ep04_geselowitz-ecg.jl:15-26
 */
double orc_source_eval(int kind, const double *prm, int dim, const double *x, double t) {
    double n2 = 0, mx = -INFINITY;
    for (int d = 0; d < dim; d++) {
        n2 += x[d] * x[d];
        if (x[d] > mx) mx = x[d];
    }
    switch (kind) {
    case 1: return (mx < prm[0] && t < prm[1]) ? prm[2] : 0.0;
    case 2: return (sqrt(n2) < prm[0] && t < prm[1]) ? prm[2] : 0.0;
    case 3: return cos(2.0 * M_PI * t) * exp(-n2);
    case 4: return sqrt(n2) + t;
    case 5: return (t <= prm[1] && x[0] < prm[0]) ? prm[2] / prm[3] * exp(t / prm[3]) : 0.0;
    default: return 0.0;
    }
}

/* be[j] += f(x_qp,t)*Nj*dOmega, x_qp = sum_a M_a x_a
 * (src/modeling/core/analytical_coefficient.jl:80-101; coefficients.jl:279-292).
 * If fq != NULL it holds host-evaluated f at the nq points of this cell (general closure path). */
void orc_element_source(int ct, int qorder, const double *X, int kind, const double *prm, double t, const double *fq,
                        double *be) {
    int nv = cell_nv(ct), dim = cell_dim(ct);
    double pts[ORC_MAXQ * 3], w[ORC_MAXQ], N[ORC_MAXNV], dN[ORC_MAXNV * 3];
    int nq = orc_quadrature(ct, qorder, pts, w);
    for (int i = 0; i < nv; i++) be[i] = 0;
    for (int q = 0; q < nq; q++) {
        orc_shape(ct, pts + q * dim, N, dN);
        double dO = map_qp(nv, dim, X, dN, NULL) * w[q];
        double x[3] = {0, 0, 0};
        for (int a = 0; a < nv; a++)
            for (int d = 0; d < dim; d++) x[d] += N[a] * X[a * dim + d];
        double fx = fq ? fq[q] : orc_source_eval(kind, prm, dim, x, t);
        for (int j = 0; j < nv; j++) be[j] += fx * N[j] * dO;
    }
}

/* ------------------------------------------------------------------------------------------
 * Element loop + scatter into the fixed CSR pattern (FerriteOperators update_operator!,
 * sequential strategy; glue src/solver/interface.jl:66-94): for cell: Ke = 0; assemble_element!;
 * A[dofs,dofs] += Ke.  op: 0 mass, 1 diffusion.
 * ---------------------------------------------------------------------------------------- */
static int64_t find_col(const int64_t *colidx, int64_t lo, int64_t hi, int64_t col) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (colidx[mid] < col) lo = mid + 1; else hi = mid;
    }
    return lo;
}

void orc_assemble_bilinear(int op, int ct, int qorder, int64_t ncells, const int64_t *conn, const double *coords,
                           const int64_t *celldofs, double rho, int kind, const double *data, double cmchi,
                           const int64_t *rowptr, const int64_t *colidx, double *vals) {
    int nv = cell_nv(ct), dim = cell_dim(ct);
    int64_t nnz_guard = 0; (void)nnz_guard;
    double X[ORC_MAXNV * 3], Ke[ORC_MAXNV * ORC_MAXNV];
    for (int64_t c = 0; c < ncells; c++) {
        for (int a = 0; a < nv; a++)
            for (int d = 0; d < dim; d++) X[a * dim + d] = coords[conn[c * nv + a] * dim + d];
        if (op == 0) orc_element_mass(ct, qorder, X, rho, Ke);
        else orc_element_diffusion(ct, qorder, X, kind, data, cmchi, c, Ke);
        for (int i = 0; i < nv; i++) {
            int64_t r = celldofs[c * nv + i];
            for (int j = 0; j < nv; j++) {
                int64_t p = find_col(colidx, rowptr[r], rowptr[r + 1], celldofs[c * nv + j]);
                vals[p] += Ke[i * nv + j];
            }
        }
    }
}

/* Threaded twin of orc_assemble_bilinear with BITWISE the same result: the sequential element loop adds the
 * contributions to one stored entry in ascending cell order, so (1) the element matrices of a chunk of consecutive
 * cells are computed by all threads into a scratch array, (2) threads over rows add, for every adjacent cell of the
 * chunk in ascending order, row `a` of that cell's matrix to the row's entries.  Chunks are processed in ascending
 * order, hence every entry still sees its contributions in the order of the sequential loop.  Test infrastructure
 * only (large parity cases and the CPU baseline's setup); tests/test_oracle_independent.py checks it against the
 * sequential loop. */
void orc_assemble_bilinear_par(int op, int ct, int qorder, int64_t ncells, const int64_t *conn, const double *coords,
                               const int64_t *celldofs, int64_t ndofs, double rho, int kind, const double *data, double cmchi,
                               const int64_t *rowptr, const int64_t *colidx, double *vals) {
    const int nv = cell_nv(ct), dim = cell_dim(ct);
    int64_t *adjptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ndofs + 1));
    int64_t *adj = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncells * nv + 1));
    dof_cell_adjacency(ndofs, ncells, nv, celldofs, adjptr, adj);
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ndofs + 1));
    memcpy(cur, adjptr, sizeof(int64_t) * (size_t)(ndofs + 1));
    const int64_t chunk = 1 << 20;
    double *EA = (double *)malloc(sizeof(double) * (size_t)(chunk * nv * nv));
    for (int64_t c0 = 0; c0 < ncells; c0 += chunk) {
        const int64_t c1 = c0 + chunk < ncells ? c0 + chunk : ncells;
#pragma omp parallel for schedule(static)
        for (int64_t c = c0; c < c1; c++) {
            double X[ORC_MAXNV * 3];
            for (int a = 0; a < nv; a++)
                for (int d = 0; d < dim; d++) X[a * dim + d] = coords[conn[c * nv + a] * dim + d];
            double *Ke = EA + (c - c0) * nv * nv;
            if (op == 0) orc_element_mass(ct, qorder, X, rho, Ke);
            else orc_element_diffusion(ct, qorder, X, kind, data, cmchi, c, Ke);
        }
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t r = 0; r < ndofs; r++) {
            int64_t q = cur[r];
            for (; q < adjptr[r + 1] && adj[q] / nv < c1; q++) {
                const int64_t c = adj[q] / nv;
                const int i = (int)(adj[q] % nv);
                const double *Ke = EA + (c - c0) * nv * nv;
                for (int j = 0; j < nv; j++) {
                    int64_t p = find_col(colidx, rowptr[r], rowptr[r + 1], celldofs[c * nv + j]);
                    vals[p] += Ke[i * nv + j];
                }
            }
            cur[r] = q;
        }
    }
    free(EA);
    free(cur);
    free(adj);
    free(adjptr);
}

/* b is zeroed first, as update_operator! of a LinearFerriteOperator does */
void orc_assemble_source(int ct, int qorder, int64_t ncells, const int64_t *conn, const double *coords,
                         const int64_t *celldofs, int kind, const double *prm, double t, const double *fq_all,
                         int64_t ndofs, double *b) {
    int nv = cell_nv(ct), dim = cell_dim(ct);
    double pts[ORC_MAXQ * 3], w[ORC_MAXQ];
    int nq = orc_quadrature(ct, qorder, pts, w);
    double X[ORC_MAXNV * 3], be[ORC_MAXNV];
    for (int64_t i = 0; i < ndofs; i++) b[i] = 0;
    for (int64_t c = 0; c < ncells; c++) {
        for (int a = 0; a < nv; a++)
            for (int d = 0; d < dim; d++) X[a * dim + d] = coords[conn[c * nv + a] * dim + d];
        orc_element_source(ct, qorder, X, kind, prm, t, fq_all ? fq_all + c * nq : NULL, be);
        for (int j = 0; j < nv; j++) b[celldofs[c * nv + j]] += be[j];
    }
}

/* ------------------------------------------------------------------------------------------
 * Pseudo-ECG after Plonsey (1964) with the Gauss theorem applied, Plonsey1964ECGGaussCache
 * (src/modeling/electrophysiology/ecg.jl:1-160):
 *   update_ecg!    flux[cell][q] = sum_i (D(x_q) . gradN_i) * phi[dof_i]            (:14-38, no dOmega here)
 *   evaluate_ecg   phi_e(x) = -1/(4 pi kappa_t) * sum_cells sum_q ((flux . (xq - x)) / |xq - x|^3) * dOmega  (:86-148)
 * D is the coefficient of the diffusion operator the cache was built on.  flux: ncells*nq*dim doubles (scratch).
 * ---------------------------------------------------------------------------------------- */
void orc_ecg_plonsey(int ct, int qorder, int64_t ncells, const int64_t *conn, const double *coords, const int64_t *celldofs,
                     int kind, const double *data, double cmchi, const double *phi, int ne, const double *electrodes,
                     double kappa_t, double *flux, double *out) {
    int nv = cell_nv(ct), dim = cell_dim(ct);
    double pts[ORC_MAXQ * 3], w[ORC_MAXQ], N[ORC_MAXNV], dN[ORC_MAXNV * 3], G[ORC_MAXNV * 3], D[9], X[ORC_MAXNV * 3];
    int nq = orc_quadrature(ct, qorder, pts, w);
    /* update_ecg!: fill!(0) then compute_quadrature_fluxes! */
    for (int64_t c = 0; c < ncells; c++) {
        for (int a = 0; a < nv; a++)
            for (int d = 0; d < dim; d++) X[a * dim + d] = coords[conn[c * nv + a] * dim + d];
        for (int q = 0; q < nq; q++) {
            double *f = flux + (c * nq + q) * dim;
            for (int d = 0; d < dim; d++) f[d] = 0.0;
            orc_shape(ct, pts + q * dim, N, dN);
            map_qp(nv, dim, X, dN, G);
            eval_D(kind, data, cmchi, dim, nv, c, N, D);
            for (int i = 0; i < nv; i++) {
                double ui = phi[celldofs[c * nv + i]];
                for (int r = 0; r < dim; r++) {
                    double s = 0;
                    for (int k = 0; k < dim; k++) s += D[r * dim + k] * G[i * dim + k];
                    f[r] += s * ui;
                }
            }
        }
    }
    for (int e = 0; e < ne; e++) {
        const double *xe = electrodes + e * dim;
        double phie = 0.0;
        for (int64_t c = 0; c < ncells; c++) {
            for (int a = 0; a < nv; a++)
                for (int d = 0; d < dim; d++) X[a * dim + d] = coords[conn[c * nv + a] * dim + d];
            double local = 0.0;
            for (int q = 0; q < nq; q++) {
                orc_shape(ct, pts + q * dim, N, dN);
                double dO = map_qp(nv, dim, X, dN, NULL) * w[q];
                double dvec[3], n2 = 0, fd = 0;
                const double *f = flux + (c * nq + q) * dim;
                for (int d = 0; d < dim; d++) {
                    double xq = 0;
                    for (int a = 0; a < nv; a++) xq += N[a] * X[a * dim + d];
                    dvec[d] = xq - xe[d];
                }
                for (int d = 0; d < dim; d++) { n2 += dvec[d] * dvec[d]; fd += f[d] * dvec[d]; }
                double n = sqrt(n2);
                local += fd / (n * n * n) * dO;
            }
            phie += local;
        }
        out[e] = -phie / (4 * M_PI * kappa_t);
    }
}

/* ------------------------------------------------------------------------------------------
 * SpMV  y = A x  (src/utils.jl:210-231): threads over rows, sequential left-to-right row sum.
 * ---------------------------------------------------------------------------------------- */
void orc_spmv(int64_t n, const int64_t *rowptr, const int64_t *colidx, const double *vals, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        double v = 0.0;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) v += vals[k] * x[colidx[k]];
        y[r] = v;
    }
}

/* nz(A) = nz(M) - dt*nz(K)   (src/solver/time/euler.jl:104-116) */
void orc_axpby_values(int64_t nnz, const double *M, const double *K, double dt, double *A) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nnz; k++) A[k] = M[k] - dt * K[k];
}

/* ------------------------------------------------------------------------------------------
 * CG: Krylov.jl 0.10.9 cg! as driven by LinearSolve.KrylovJL_CG (src/solver/time/euler.jl:10,94):
 * x0 = 0, r = b, p = r, gamma = r.r, eps = atol + rtol*sqrt(gamma);
 * loop { Ap; alpha = gamma/(p.Ap); x += alpha p; r -= alpha Ap; gamma' = r.r;
 *        solved = sqrt(gamma') <= eps; if !solved { beta = gamma'/gamma; p = r + beta p }; iter++ }
 * until solved or iter >= itmax.  threaded_blas1 = 0 mirrors Krylov's serial vector ops on
 * Vector{Float64}; = 1 threads them (stronger CPU baseline).  work = 3n doubles (r,p,Ap).
 * Returns iterations; *rnorm = final residual norm; *converged = solved.
 * ---------------------------------------------------------------------------------------- */
static double dot_serial(int64_t n, const double *a, const double *b) {
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
static double dot_omp(int64_t n, const double *a, const double *b) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* "Exact" dot product (threaded_blas1 == 2): double-double accumulation -- exact products by fma (TwoProd), error-free
 * additions (TwoSum) -- rounded to fp64 once at the end.  The value no longer depends on the summation order (up to a
 * ~1e-24 relative chance per call of sitting on a rounding boundary), which is what lets a CPU run and a GPU run --
 * whose sums are formed in completely different orders -- produce the same CG scalars, hence the same iterates and the
 * same iteration counts.  Not what Julia's BLAS dot does (no implementation can reproduce that order); it is the
 * order-free reference both sides can meet.  Product side: tb_cg_set_exact_dot. */
typedef struct { double hi, lo; } orc_dd;
static inline void dd_add_pair(orc_dd *a, double h, double l) {
    double s = a->hi + h;
    double bb = s - a->hi;
    double e = (a->hi - (s - bb)) + (h - bb);
    e += a->lo + l;
    double hi = s + e;
    a->lo = e - (hi - s);
    a->hi = hi;
}
static double dot_exact(int64_t n, const double *a, const double *b) {
    orc_dd tot = {0.0, 0.0};
#pragma omp parallel
    {
        orc_dd acc = {0.0, 0.0};
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < n; i++) {
            double p = a[i] * b[i];
            dd_add_pair(&acc, p, fma(a[i], b[i], -p));
        }
#pragma omp critical
        dd_add_pair(&tot, acc.hi, acc.lo);
    }
    return tot.hi + tot.lo;
}
static double dot_mode(int mode, int64_t n, const double *a, const double *b) {
    return mode == 2 ? dot_exact(n, a, b) : mode == 1 ? dot_omp(n, a, b) : dot_serial(n, a, b);
}

int64_t orc_cg(int64_t n, const int64_t *rowptr, const int64_t *colidx, const double *vals, const double *b, double *x,
               double atol, double rtol, int64_t itmax, int threaded_blas1, double *work, double *rnorm,
               int32_t *converged) {
    double *r = work, *p = work + n, *Ap = work + 2 * n;
    for (int64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; }
    double gamma = dot_mode(threaded_blas1, n, r, r);
    double rn = sqrt(gamma);
    double eps = atol + rtol * rn;
    int solved = rn <= eps;
    int64_t iter = 0;
    int tired = iter >= itmax;
    while (!(solved || tired)) {
        orc_spmv(n, rowptr, colidx, vals, p, Ap);
        double pAp = dot_mode(threaded_blas1, n, p, Ap);
        double alpha = gamma / pAp;
        if (threaded_blas1) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
        } else {
            for (int64_t i = 0; i < n; i++) x[i] += alpha * p[i];
            for (int64_t i = 0; i < n; i++) r[i] -= alpha * Ap[i];
        }
        double gnext = dot_mode(threaded_blas1, n, r, r);
        rn = sqrt(gnext);
        solved = rn <= eps;
        if (!solved) {
            double beta = gnext / gamma;
            gamma = gnext;
            if (threaded_blas1) {
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
            } else {
                for (int64_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
            }
        }
        iter++;
        tired = iter >= itmax;
    }
    *rnorm = rn;
    *converged = solved;
    return iter;
}

/* Preconditioned CG as LinearSolve.KrylovJL_CG(precs = ..., ldiv = false) runs it (Krylov.jl 0.10.9 cg!, the usage
 * shown in bak/examples-gpu/spiral-wave.jl:95-105 and recommended at ep01_spiral-wave.jl:129-131): z = M r with M the
 * stored INVERSE of the preconditioner, gamma = r.z, the stopping test is on sqrt(r.z) (the M-norm of the residual),
 * p = z + beta p.  Here M = diag(A)^-1 (Jacobi = KrylovPreconditioners' BlockJacobi with blocks of one row).
 * dinv: n doubles, 1/a_ii.  work: 4n doubles. */
int64_t orc_pcg_jacobi(int64_t n, const int64_t *rowptr, const int64_t *colidx, const double *vals, const double *b,
                       double *x, double atol, double rtol, int64_t itmax, double *dinv, double *work, double *rnorm,
                       int32_t *converged, int dmode) {
    double *r = work, *p = work + n, *Ap = work + 2 * n, *z = work + 3 * n;
    for (int64_t i = 0; i < n; i++) {
        dinv[i] = 0.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++)
            if (colidx[k] == i) dinv[i] = 1.0 / vals[k];
    }
    for (int64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; z[i] = dinv[i] * r[i]; p[i] = z[i]; }
    double gamma = dot_mode(dmode, n, r, z);
    double rn = sqrt(gamma);
    double eps = atol + rtol * rn;
    int solved = rn <= eps;
    int64_t iter = 0;
    int tired = iter >= itmax;
    while (!(solved || tired)) {
        orc_spmv(n, rowptr, colidx, vals, p, Ap);
        double pAp = dot_mode(dmode, n, p, Ap);
        double alpha = gamma / pAp;
        for (int64_t i = 0; i < n; i++) x[i] += alpha * p[i];
        for (int64_t i = 0; i < n; i++) r[i] -= alpha * Ap[i];
        for (int64_t i = 0; i < n; i++) z[i] = dinv[i] * r[i];
        double gnext = dot_mode(dmode, n, r, z);
        rn = sqrt(gnext);
        solved = rn <= eps;
        if (!solved) {
            double beta = gnext / gamma;
            gamma = gnext;
            for (int64_t i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        }
        iter++;
        tired = iter >= itmax;
    }
    *rnorm = rn;
    *converged = solved;
    return iter;
}

/* ------------------------------------------------------------------------------------------
 * Cell models.  prm arrays hold the struct fields in declaration order.
 * FHN      src/modeling/cells/fhn.jl:6-34       prm = a,b,c,d,e,f
 * PCG2019  src/modeling/cells/pcg2019.jl:4-133  prm = the 36 fields g_Na .. E_Ca
 * ---------------------------------------------------------------------------------------- */
#define ORC_FHN 0
#define ORC_PCG2019 1
#define ORC_ALIEV_PANFILOV 2 /* src/modeling/cells/aliev-panfilov.jl:1-34; states (s, phi_m) */
#define ORC_TIMEPROBE 99 /* test/test_time_integrator.jl:275-280: du = 1 + sin(t) */

int orc_cell_nstates(int model) { return model == ORC_FHN || model == ORC_ALIEV_PANFILOV ? 2 : model == ORC_PCG2019 ? 7 : 1; }

void orc_fhn_default_params(double *p) {
    p[0] = 0.1; p[1] = 0.5; p[2] = 1.0; p[3] = 0.0; p[4] = 0.01; p[5] = 1.0;
}
void orc_aliev_panfilov_default_params(double *p) { /* aliev-panfilov.jl:2-7 */
    p[0] = 1.0 / 12.9; p[1] = 8.0; p[2] = 0.05; p[3] = 0.002; p[4] = 0.2; p[5] = 0.3;
}
void orc_pcg2019_default_params(double *p) {
    static const double d[36] = {12.0, -52.244, 6.5472, 0.12, -78.7, 5.93, 0.799163, 6.80738, /* I_Na */
                                 0.73893, -91.9655, 12.4997,                                   /* I_K1 */
                                 0.1688, 14.3116, 11.462, -47.9286, 4.9314, 9.90669,           /* I_to */
                                 0.11503, 0.7, 4.3, -15.7, 4.6, 30.0,                          /* I_CaL */
                                 0.056, -26.6, 6.5, 334.0, -49.6, 23.5,                        /* I_Kr */
                                 0.008, 24.6, 12.1, 628.0,                                     /* I_Ks */
                                 65.0, -85.0, 50.0};
    memcpy(p, d, sizeof(d));
}

static inline double sigmoid(double phi, double E, double k, double sign) { return 1.0 / (1.0 + exp(sign * (phi - E) / k)); }

enum { g_Na, E_m, k_m, tau_m, E_h, k_h, delta_h, tau_h0, g_K1, E_z, k_z, g_to, E_r, k_r, E_s, k_s, tau_s, g_CaL, E_d, k_d,
       E_f, k_f, tau_f, g_Kr, E_xr, k_xr, tau_xr, E_y, k_y, g_Ks, E_xs, k_xs, tau_xs, E_Na, E_K, E_Ca };

void orc_pcg2019_default_state(const double *p, double *u0) {
    u0[0] = p[E_K];
    u0[1] = sigmoid(u0[0], p[E_h], p[k_h], 1.0);
    u0[2] = sigmoid(u0[0], p[E_m], p[k_m], -1.0);
    u0[3] = sigmoid(u0[0], p[E_f], p[k_f], 1.0);
    u0[4] = sigmoid(u0[0], p[E_s], p[k_s], 1.0);
    u0[5] = sigmoid(u0[0], p[E_xs], p[k_xs], -1.0);
    u0[6] = sigmoid(u0[0], p[E_xr], p[k_xr], -1.0);
}

void orc_cell_rhs(int model, const double *p, const double *u, double t, double *du) {
    if (model == ORC_FHN) {
        double phi = u[0], s = u[1];
        du[0] = p[5] * (phi * (1 - phi) * (phi - p[0]) - s);
        du[1] = p[4] * (p[1] * phi - p[2] * s - p[3]);
    } else if (model == ORC_ALIEV_PANFILOV) { /* aliev-panfilov.jl:15-34: u = (s, phi) */
        double ct = p[0], k = p[1], a = p[2], e0 = p[3], mu1 = p[4], mu2 = p[5];
        double phi = u[1], s = u[0];
        double eps = e0 + s * mu1 / (phi + mu2);
        du[1] = ct * (k * phi * (phi - 1.0) * (phi - a) - phi * s);
        du[0] = ct * eps * (-s - k * phi * (phi - a - 1.0));
    } else if (model == ORC_PCG2019) {
        const double C_m = 1.0;
        double phi = u[0], h = u[1], m = u[2], f = u[3], s = u[4], xs = u[5], xr = u[6];
        double r_inf = sigmoid(phi, p[E_r], p[k_r], -1.0);
        double d_inf = sigmoid(phi, p[E_d], p[k_d], -1.0);
        double z_inf = sigmoid(phi, p[E_z], p[k_z], 1.0);
        double y_inf = sigmoid(phi, p[E_y], p[k_y], 1.0);
        double I_Na = p[g_Na] * m * m * m * h * h * (phi - p[E_Na]);
        double I_K1 = p[g_K1] * z_inf * (phi - p[E_K]);
        double I_to = p[g_to] * r_inf * s * (phi - p[E_K]);
        double I_CaL = p[g_CaL] * d_inf * f * (phi - p[E_Ca]);
        double I_Kr = p[g_Kr] * xr * y_inf * (phi - p[E_K]);
        double I_Ks = p[g_Ks] * xs * (phi - p[E_K]);
        double I_total = I_Na + I_K1 + I_to + I_CaL + I_Kr + I_Ks;
        du[0] = -I_total / C_m;
        double tau_h = (2.0 * p[tau_h0] * exp(p[delta_h] * (phi - p[E_h]) / p[k_h])) / (1.0 + exp((phi - p[E_h]) / p[k_h]));
        double h_inf = sigmoid(phi, p[E_h], p[k_h], 1.0);
        du[1] = (h_inf - h) / tau_h;
        double m_inf = sigmoid(phi, p[E_m], p[k_m], -1.0);
        du[2] = (m_inf - m) / p[tau_m];
        double f_inf = sigmoid(phi, p[E_f], p[k_f], 1.0);
        du[3] = (f_inf - f) / p[tau_f];
        double s_inf = sigmoid(phi, p[E_s], p[k_s], 1.0);
        du[4] = (s_inf - s) / p[tau_s];
        double xs_inf = sigmoid(phi, p[E_xs], p[k_xs], -1.0);
        du[5] = (xs_inf - xs) / p[tau_xs];
        double xr_inf = sigmoid(phi, p[E_xr], p[k_xr], -1.0);
        du[6] = (xr_inf - xr) / p[tau_xr];
    } else {
        du[0] = 1.0 + sin(t);
    }
}

/* Cell sweep (src/solver/time/partitioned_solver.jl:38-52 outer, :80-99 forward Euler,
 * :196-234 adaptive substepper).  u, du are SoA: state s of node i at [s*ld + i]
 * (src/modeling/solution_variables.jl:60-63).  du is written to memory exactly as the reference
 * does.  substeps <= 1 selects plain forward Euler; phi_idx 0-based. */
void orc_cell_step(int model, const double *p, double *u, double *du, int64_t n, int64_t ld, double t, double dt,
                   int substeps, double threshold, int phi_idx) {
    int ns = orc_cell_nstates(model);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double ul[8], dl[8];
        for (int s = 0; s < ns; s++) ul[s] = u[s * ld + i];
        orc_cell_rhs(model, p, ul, t, dl);
        if (substeps <= 1 || fabs(dl[phi_idx]) < threshold) {
            for (int s = 0; s < ns; s++) ul[s] += dt * dl[s];
        } else {
            double dts = dt / substeps;
            for (int s = 0; s < ns; s++) ul[s] += dts * dl[s];
            for (int k = 2; k <= substeps; k++) {
                double ts = t + (k - 1) * dts;
                orc_cell_rhs(model, p, ul, ts, dl);
                for (int s = 0; s < ns; s++) ul[s] += dts * dl[s];
            }
        }
        for (int s = 0; s < ns; s++) { u[s * ld + i] = ul[s]; du[s * ld + i] = dl[s]; }
    }
}

/* ------------------------------------------------------------------------------------------
 * One LieTrotterGodunov step over (BackwardEulerSolver, cell solver): heat first, then cells
 * (OrdinaryDiffEqOperatorSplitting 0.4.0 via src/solver/time/integrator/operatorsplitting-interface.jl:23-232;
 * src/solver/time/euler.jl:71-101).  A must already hold M - dt*K.  bS may be NULL
 * (LinearNullOperator) else it is ADDED as is (euler.jl:88-91: add! is unconditional).
 * work: 5n doubles.  Returns CG iterations; *converged as orc_cg.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_ltg_step(int64_t n, const int64_t *rowptr, const int64_t *colidx, const double *Avals, const double *Mvals,
                     const double *bS, int model, const double *prm, double *u, double *du, int64_t ld, int phi_idx,
                     double t, double dt, int substeps, double threshold, double atol, double rtol, int64_t itmax,
                     int threaded_blas1, double *work, double *rnorm, int32_t *converged) {
    double *b = work, *x = work + n, *cgw = work + 2 * n;
    double *phi = u + (int64_t)phi_idx * ld;
    orc_spmv(n, rowptr, colidx, Mvals, phi, b); /* b = M u_{n-1} */
    if (bS)
        for (int64_t i = 0; i < n; i++) b[i] += bS[i];
    int64_t it = orc_cg(n, rowptr, colidx, Avals, b, x, atol, rtol, itmax, threaded_blas1, cgw, rnorm, converged);
    memcpy(phi, x, sizeof(double) * (size_t)n);
    orc_cell_step(model, prm, u, du, n, ld, t, dt, substeps, threshold, phi_idx);
    return it;
}

/* ------------------------------------------------------------------------------------------
 * Closed-form operator of a UNIFORM hexahedral grid with a diagonal diffusion tensor -- a second, matrix-free
 * restatement used where the assembled CSR image does not fit the host (BASELINE config 5: 101 M dofs, 2.7 G
 * nonzeros).  For Lagrange-1 on a uniform grid the 2x2x2 Gauss rule integrates mass and stiffness exactly, so
 *     M = Mx (x) My (x) Mz,   K_ref = -( kx Kx(x)My(x)Mz + ky Mx(x)Ky(x)Mz + kz Mx(x)My(x)Kz )      (negative, a-7)
 * with the 1D matrices  M1 = h*[1/6 2/3 1/6] (ends h*[1/3 1/6]),  K1 = (1/h)*[-1 2 -1] (ends (1/h)*[1 -1]).
 * Vectors are in GRID ORDER (node a + (nx+1)*(b + (ny+1)*c)), not in dof order: this oracle never numbers dofs; the
 * harness matches nodes by coordinates.  tests/test_oracle_independent.py pins the closed form against the assembled
 * oracle matrices (orc_assemble_bilinear) on small grids.
 *   y = cm * M x + ck * K_ref x
 * ---------------------------------------------------------------------------------------- */
static void stencil_1d(int64_t n, double h, int64_t i, double *m3, double *k3) {
    /* row i of the (n+1)x(n+1) 1D mass / stiffness matrices, offsets -1, 0, +1 */
    const int lo = i > 0, hi = i < n;
    m3[0] = lo ? h / 6.0 : 0.0;
    m3[2] = hi ? h / 6.0 : 0.0;
    m3[1] = (lo ? h / 3.0 : 0.0) + (hi ? h / 3.0 : 0.0);
    k3[0] = lo ? -1.0 / h : 0.0;
    k3[2] = hi ? -1.0 / h : 0.0;
    k3[1] = (lo ? 1.0 / h : 0.0) + (hi ? 1.0 / h : 0.0);
}

void orc_stencil_apply(const int64_t *nel, const double *h, const double *kappa, double cm, double ck, const double *x,
                       double *y) {
    const int64_t nx = nel[0], ny = nel[1], nz = nel[2];
    const int64_t sx = nx + 1, sy = ny + 1, sz = nz + 1;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < sz; c++) {
        double mz[3], kz[3];
        stencil_1d(nz, h[2], c, mz, kz);
        for (int64_t b = 0; b < sy; b++) {
            double my[3], ky[3];
            stencil_1d(ny, h[1], b, my, ky);
            for (int64_t a = 0; a < sx; a++) {
                double mx[3], kx[3];
                stencil_1d(nx, h[0], a, mx, kx);
                double acc = 0.0;
                for (int dc = -1; dc <= 1; dc++) {
                    if (c + dc < 0 || c + dc >= sz) continue;
                    for (int db = -1; db <= 1; db++) {
                        if (b + db < 0 || b + db >= sy) continue;
                        for (int da = -1; da <= 1; da++) {
                            if (a + da < 0 || a + da >= sx) continue;
                            const double m = mx[da + 1] * my[db + 1] * mz[dc + 1];
                            const double k = -(kappa[0] * (kx[da + 1] * my[db + 1] * mz[dc + 1]) +
                                               kappa[1] * (mx[da + 1] * ky[db + 1] * mz[dc + 1]) +
                                               kappa[2] * (mx[da + 1] * my[db + 1] * kz[dc + 1]));
                            acc += (cm * m + ck * k) * x[(a + da) + sx * ((b + db) + sy * (c + dc))];
                        }
                    }
                }
                y[a + sx * (b + sy * c)] = acc;
            }
        }
    }
}

/* one stored entry of the closed-form M (which = 0) or K_ref (which = 1) between grid nodes (a,b,c) and (a+da,b+db,c+dc) */
double orc_stencil_entry(const int64_t *nel, const double *h, const double *kappa, int which, const int64_t *node,
                         const int *off) {
    double m[3][3], k[3][3];
    for (int d = 0; d < 3; d++) stencil_1d(nel[d], h[d], node[d], m[d], k[d]);
    for (int d = 0; d < 3; d++)
        if (off[d] < -1 || off[d] > 1 || node[d] + off[d] < 0 || node[d] + off[d] > nel[d]) return 0.0;
    const double mx = m[0][off[0] + 1], my = m[1][off[1] + 1], mz = m[2][off[2] + 1];
    const double kx = k[0][off[0] + 1], ky = k[1][off[1] + 1], kz = k[2][off[2] + 1];
    if (which == 0) return mx * my * mz;
    return -(kappa[0] * (kx * my * mz) + kappa[1] * (mx * ky * mz) + kappa[2] * (mx * my * kz));
}

/* orc_ltg_step with the closed-form operator: b = M phi; CG (same recurrence and stopping rule as orc_cg, threaded
 * dot products) on A = M - dt*K_ref; phi <- x; cell sweep.  u, du: SoA in grid order; work: 5n doubles. */
int64_t orc_stencil_ltg_step(const int64_t *nel, const double *h, const double *kappa, int model, const double *prm,
                             double *u, double *du, int phi_idx, double t, double dt, int substeps, double threshold,
                             double atol, double rtol, int64_t itmax, double *work, double *rnorm, int32_t *converged) {
    const int64_t n = (nel[0] + 1) * (nel[1] + 1) * (nel[2] + 1);
    double *b = work, *x = work + n, *r = work + 2 * n, *p = work + 3 * n, *Ap = work + 4 * n;
    double *phi = u + (int64_t)phi_idx * n;
    orc_stencil_apply(nel, h, kappa, 1.0, 0.0, phi, b);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; }
    double gamma = dot_omp(n, r, r);
    double rn = sqrt(gamma);
    const double eps = atol + rtol * rn;
    int solved = rn <= eps;
    int64_t iter = 0;
    int tired = iter >= itmax;
    while (!(solved || tired)) {
        orc_stencil_apply(nel, h, kappa, 1.0, -dt, p, Ap);
        const double pAp = dot_omp(n, p, Ap);
        const double alpha = gamma / pAp;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
        const double gnext = dot_omp(n, r, r);
        rn = sqrt(gnext);
        solved = rn <= eps;
        if (!solved) {
            const double beta = gnext / gamma;
            gamma = gnext;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
        }
        iter++;
        tired = iter >= itmax;
    }
    *rnorm = rn;
    *converged = solved;
    memcpy(phi, x, sizeof(double) * (size_t)n);
    orc_cell_step(model, prm, u, du, n, n, t, dt, substeps, threshold, phi_idx);
    return iter;
}

/* ------------------------------------------------------------------------------------------
 * Multi-subdomain splits (src/discretization/fem.jl:434-542).
 * orc_cell_step_strided: the cell solvers over one StateBlock with an arbitrary layout (solution_variables.jl:60-68):
 * state s of point k at u[k*pstride + s*sstride] (PointBlockedLayout: pstride = nstates, sstride = 1).  du uses the
 * same indexing.  Same arithmetic as orc_cell_step.
 * ---------------------------------------------------------------------------------------- */
void orc_cell_step_strided(int model, const double *p, double *u, double *du, int64_t n, int64_t pstride, int64_t sstride,
                           double t, double dt, int substeps, double threshold, int phi_idx) {
    int ns = orc_cell_nstates(model);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double ul[8], dl[8];
        for (int s = 0; s < ns; s++) ul[s] = u[i * pstride + s * sstride];
        orc_cell_rhs(model, p, ul, t, dl);
        if (substeps <= 1 || fabs(dl[phi_idx]) < threshold) {
            for (int s = 0; s < ns; s++) ul[s] += dt * dl[s];
        } else {
            double dts = dt / substeps;
            for (int s = 0; s < ns; s++) ul[s] += dts * dl[s];
            for (int k = 2; k <= substeps; k++) {
                double ts = t + (k - 1) * dts;
                orc_cell_rhs(model, p, ul, ts, dl);
                for (int s = 0; s < ns; s++) ul[s] += dts * dl[s];
            }
        }
        for (int s = 0; s < ns; s++) { u[i * pstride + s * sstride] = ul[s]; du[i * pstride + s * sstride] = dl[s]; }
    }
}

/* BilinearInterfaceDiffusionElementCache / assemble_element! (src/modeling/core/diffusion.jl:81-140): an interface cell is a
 * pair of coincident facets; basis functions 0..k-1 live on the "here" side, k..2k-1 on the "there" side;
 * shape_value_jump = value_there - value_here, i.e. -N_i on the here side and +N_i on the there side (⚠ mem:
 * FerriteInterfaceElements' convention; the element matrix is a product of two jumps, so the sign cancels);
 * dOmega = getdetJdV_average = (detJ_here + detJ_there)/2 * w.  Facets: line (2 nodes) in 2D, quadrilateral (4 nodes) in 3D;
 * Gauss-Legendre `qorder` points per direction, first coordinate fastest.  Ke: (2k) x (2k), row-major. */
void orc_interface_diffusion_element(int k, int sdim, int qorder, const double *Xh, const double *Xt, double D, double *Ke) {
    double gp[8], gw[8];
    int ng = gauss_1d(qorder, gp, gw);
    int fd = sdim - 1, nd = 2 * k;
    int nq = fd == 1 ? ng : ng * ng;
    for (int i = 0; i < nd * nd; i++) Ke[i] = 0;
    static const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
    for (int q = 0; q < nq; q++) {
        double N[4], dN[4][2], w;
        if (fd == 1) {
            N[0] = 0.5 * (1.0 - gp[q]); N[1] = 0.5 * (1.0 + gp[q]);
            dN[0][0] = -0.5; dN[1][0] = 0.5;
            w = gw[q];
        } else {
            int a = q % ng, b = q / ng;
            for (int v = 0; v < 4; v++) {
                N[v] = 0.25 * (1.0 + sx[v] * gp[a]) * (1.0 + sy[v] * gp[b]);
                dN[v][0] = 0.25 * sx[v] * (1.0 + sy[v] * gp[b]);
                dN[v][1] = 0.25 * (1.0 + sx[v] * gp[a]) * sy[v];
            }
            w = gw[a] * gw[b];
        }
        double dO = 0;
        for (int side = 0; side < 2; side++) {
            const double *X = side ? Xt : Xh;
            double t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0};
            for (int a = 0; a < k; a++)
                for (int d = 0; d < sdim; d++) {
                    t0[d] += X[a * sdim + d] * dN[a][0];
                    if (fd == 2) t1[d] += X[a * sdim + d] * dN[a][1];
                }
            double dj;
            if (fd == 1) dj = sqrt(t0[0] * t0[0] + t0[1] * t0[1]);
            else {
                double n0 = t0[1] * t1[2] - t0[2] * t1[1], n1 = t0[2] * t1[0] - t0[0] * t1[2], n2 = t0[0] * t1[1] - t0[1] * t1[0];
                dj = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
            }
            dO += dj * w;
        }
        dO = dO / 2.0;
        for (int i = 0; i < nd; i++) {
            double ji = i < k ? -N[i] : N[i - k];
            for (int j = 0; j < nd; j++) {
                double jj = j < k ? -N[j] : N[j - k];
                Ke[i * nd + j] -= (ji * D * jj) * dO;
            }
        }
    }
}

/* sequential interface-cell loop scattering into a CSR pattern that already holds the interface couplings; vals is NOT
 * zeroed (the caller adds the interface operator on top of the bulk one, as BilinearMultiIntegrator assembles them into
 * one matrix, subdomain after subdomain) */
int orc_assemble_interface_diffusion(int k, int sdim, int qorder, int64_t nif, const int64_t *dofs, const double *Xh,
                                     const double *Xt, double D, const int64_t *rowptr, const int64_t *colidx, double *vals) {
    int nd = 2 * k, miss = 0;
    double Ke[64];
    for (int64_t c = 0; c < nif; c++) {
        orc_interface_diffusion_element(k, sdim, qorder, Xh + c * k * sdim, Xt + c * k * sdim, D, Ke);
        for (int i = 0; i < nd; i++) {
            int64_t r = dofs[c * nd + i];
            for (int j = 0; j < nd; j++) {
                int64_t p = find_col(colidx, rowptr[r], rowptr[r + 1], dofs[c * nd + j]);
                if (p < rowptr[r + 1] && colidx[p] == dofs[c * nd + j]) vals[p] += Ke[i * nd + j];
                else miss = 1;
            }
        }
    }
    return miss;
}
