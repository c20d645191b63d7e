// Test-only host build of the product's host/device inline arithmetic (tb_cells.cuh, tb_elements.cuh)
// so that `-m "not gpu"` tests can compare it with the oracle without a GPU.  Not shipped, not a fallback.
#include <cstdint>
#include <cstring>
#include "../../thunderbolt.jl_b200/csrc/tb_cells.cuh"
#include "../../thunderbolt.jl_b200/csrc/tb_elements.cuh"

template <int NV> static void unpack(const double *acc, double *full) {
    for (int i = 0; i < NV; i++)
        for (int j = 0; j < NV; j++) full[i * NV + j] = i <= j ? acc[tb_sym<NV>(i, j)] : acc[tb_sym<NV>(j, i)];
}

extern "C" {

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
}
