// Structured-grid description and the closed form of Ferrite's first-touch dof numbering on Quadrilateral / Hexahedron grids
// (also compiled for the host by tests/hostmath to check it against the oracle's close! restatement without a GPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TB_GRID_HD __host__ __device__ __forceinline__
#else
#define TB_GRID_HD inline
#endif

struct GridDesc {
    int celltype, dim;
    int64_t nel[3];
    int64_t nn[3];
    double left[3], right[3];
};

TB_GRID_HD int64_t tb_grid_dof(const GridDesc &g, int64_t a, int64_t b, int64_t c) {
    const bool d3 = g.dim == 3;
    const int64_t i = a > 0 ? a - 1 : 0, j = b > 0 ? b - 1 : 0, k = d3 ? (c > 0 ? c - 1 : 0) : 0;
    const int64_t Ck = (d3 && k == 0) ? 2 : 1, Bj = j == 0 ? 2 : 1;
    int64_t n = 0;
    if (d3 && k >= 1) n += g.nn[0] * g.nn[1] * (k + 1);
    if (j >= 1) n += g.nn[0] * (j + 1) * Ck;
    if (i >= 1) n += (i + 1) * Bj * Ck;
    const int da = (int)(a - i), db = (int)(b - j), dc = d3 ? (int)(c - k) : 0;
    // rank among the new vertices of cell (i, j, k) in local vertex order (--,+-,++,-+ bottom, then top)
    const int oa[8] = {0, 1, 1, 0, 0, 1, 1, 0}, ob[8] = {0, 0, 1, 1, 0, 0, 1, 1}, oc[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    const int nvl = d3 ? 8 : 4;
    int pos = 0;
    for (int v = 0; v < nvl; v++) {
        if (oa[v] == da && ob[v] == db && oc[v] == dc) break;
        pos += ((oa[v] == 1 || i == 0) && (ob[v] == 1 || j == 0) && (!d3 || oc[v] == 1 || k == 0)) ? 1 : 0;
    }
    return n + pos;
}

