// Test-only host build of the product's host/device inline arithmetic (tb_cells.cuh, tb_elements.cuh)
// so that `-m "not gpu"` tests can compare it with the oracle without a GPU.  Not shipped, not a fallback.
#include <cstdint>
#include <cstring>
#include "../../thunderbolt.jl_b200/csrc/tb_cells.cuh"
#include "../../thunderbolt.jl_b200/csrc/tb_elements.cuh"
#include "../../thunderbolt.jl_b200/csrc/tb_grid.cuh"

template <int NV> static void unpack(const double *acc, double *full) {
    for (int i = 0; i < NV; i++)
        for (int j = 0; j < NV; j++) full[i * NV + j] = i <= j ? acc[tb_sym<NV>(i, j)] : acc[tb_sym<NV>(j, i)];
}

extern "C" {

// closed-form first-touch dof id of every node of a structured Quadrilateral (dim 2) / Hexahedron (dim 3) grid
void hm_grid_dofs(int dim, const int64_t *nel, int64_t *node2dof) {
    GridDesc g;
    g.celltype = dim == 2 ? 0 : 1;
    g.dim = dim;
    for (int d = 0; d < 3; d++) {
        g.nel[d] = d < dim ? nel[d] : 1;
        g.nn[d] = d < dim ? nel[d] + 1 : 1;
        g.left[d] = 0.0;
        g.right[d] = 1.0;
    }
    for (int64_t c = 0; c < g.nn[2]; c++)
        for (int64_t b = 0; b < g.nn[1]; b++)
            for (int64_t a = 0; a < g.nn[0]; a++) node2dof[(c * g.nn[1] + b) * g.nn[0] + a] = tb_grid_dof(g, a, b, c);
}

double hm_cell_node_step(int model, int adaptive, const double *prm, double *u, double t, double dt, int substeps, double thr) {
    tb_cell_params P;
    std::memset(&P, 0, sizeof(P));
    for (int i = 0; i < (model == 1 ? 36 : 6); i++) P.p[i] = prm[i];
    if (model == 2) return adaptive ? tb_cell_node_step<2, true>(P, u, t, dt, substeps, thr) : tb_cell_node_step<2, false>(P, u, t, dt, substeps, thr);
    if (model == 0) return adaptive ? tb_cell_node_step<0, true>(P, u, t, dt, substeps, thr) : tb_cell_node_step<0, false>(P, u, t, dt, substeps, thr);
    return adaptive ? tb_cell_node_step<1, true>(P, u, t, dt, substeps, thr) : tb_cell_node_step<1, false>(P, u, t, dt, substeps, thr);
}

int hm_tables(int celltype, int qorder, int *nq, double *xi, double *w, double *N, double *dN) {
    tb_elem_tables T;
    if (tb_build_tables(celltype, qorder, &T)) return 1;
    *nq = T.nq;
    std::memcpy(xi, T.xi, sizeof(double) * T.nq * T.dim);
    std::memcpy(w, T.w, sizeof(double) * T.nq);
    std::memcpy(N, T.N, sizeof(double) * T.nq * T.nv);
    std::memcpy(dN, T.dN, sizeof(double) * T.nq * T.nv * T.dim);
    return 0;
}

// op 0 mass, 1 diffusion; out is nv x nv full matrix
int hm_element_matrix(int celltype, int qorder, int op, const double *X, double rho, int kind, const double *data,
                      double cmchi, int64_t cell, double *out) {
    tb_elem_tables T;
    if (tb_build_tables(celltype, qorder, &T)) return 1;
    double acc[36];
#define RUN(NV, DIM)                                                                              \
    {                                                                                             \
        if (op == 0) tb_element_mass<NV, DIM, 1>(tb_view_of(&T), X, rho, acc);                                \
        else tb_element_diffusion<NV, DIM, 1>(tb_view_of(&T), X, kind, data, cmchi, cell, acc);               \
        unpack<NV>(acc, out);                                                                     \
    }
    switch (celltype) {
    case 0: RUN(4, 2) break;
    case 1: RUN(8, 3) break;
    case 2: RUN(3, 2) break;
    default: RUN(4, 3) break;
    }
#undef RUN
    return 0;
}

// full (non-symmetrised) diffusion matrix of the gather assembly: must be BITWISE the oracle's
int hm_element_diffusion_full(int celltype, int qorder, const double *X, int kind, const double *data, double cmchi,
                              int64_t cell, double *out) {
    tb_elem_tables T;
    if (tb_build_tables(celltype, qorder, &T)) return 1;
    switch (celltype) {
    case 0: tb_element_diffusion_full<4, 2, 1>(tb_view_of(&T), X, kind, data, cmchi, cell, out); break;
    case 1: tb_element_diffusion_full<8, 3, 1>(tb_view_of(&T), X, kind, data, cmchi, cell, out); break;
    case 2: tb_element_diffusion_full<3, 2, 1>(tb_view_of(&T), X, kind, data, cmchi, cell, out); break;
    default: tb_element_diffusion_full<4, 3, 1>(tb_view_of(&T), X, kind, data, cmchi, cell, out); break;
    }
    return 0;
}

int hm_element_source(int celltype, int qorder, const double *X, int kind, const double *prm, double t, const double *fq,
                      double *be) {
    tb_elem_tables T;
    if (tb_build_tables(celltype, qorder, &T)) return 1;
    switch (celltype) {
    case 0: tb_element_source<4, 2, 1>(tb_view_of(&T), X, kind, prm, t, fq, be); break;
    case 1: tb_element_source<8, 3, 1>(tb_view_of(&T), X, kind, prm, t, fq, be); break;
    case 2: tb_element_source<3, 2, 1>(tb_view_of(&T), X, kind, prm, t, fq, be); break;
    default: tb_element_source<4, 3, 1>(tb_view_of(&T), X, kind, prm, t, fq, be); break;
    }
    return 0;
}

// traced source program (tb_program_build + tb_program_eval) at npts points; returns tb_program_build's code
int hm_program_eval(int dim, const int32_t *code, int ncode, const double *consts, int nconsts, const double *x, int npts,
                    double t, double *out) {
    tb_src_program P;
    const int rc = tb_program_build(code, ncode, consts, nconsts, dim, &P);
    if (rc) return rc;
    for (int i = 0; i < npts; i++)
        out[i] = dim == 2 ? tb_program_eval<2>(P, x + 2 * i, t) : tb_program_eval<3>(P, x + 3 * i, t);
    return 0;
}
}
