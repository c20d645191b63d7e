"""Host-side input producers for BASELINE config 4: idealized LV geometry + rule-based fibres.

`generate_ideal_lv_mesh` restates src/mesh/generators.jl:521-677 (+ `_ellipsoid_point`, :738-757): nodes ring by
ring, circumferential index fastest; hexahedra for every ring section; a fan of wedges around the singular apex
edge.  The reference mesh is Hex + Wedge.  **The tetrahedral split below is this project's addition** (config 4
asks for a tetrahedral LV; the reference has no such generator): every hexahedron is cut with Ferrite's own
6-tet pattern (generate_grid(Tetrahedron)) and every wedge into 3 tets with face diagonals chosen to
match, so the result is conforming.

`odb25lt_fibres` restates `compute_local_microstructure(::ODB25LTMicrostructureParameters, …)`
(src/modeling/microstructure.jl:192-244; rotate_around / orthogonalize: src/utils.jl:95-112) with an ANALYTIC
transmural coordinate and local axes from the ellipsoid parametrisation instead of the Laplace solves of
src/modeling/core/coordinate_systems.jl (setup-time geometry, out of scope): per element, per local node f, s, n
-- exactly the `FieldCoefficient` layout `create_microstructure_model` fills (microstructure.jl:280-333).
"""
from __future__ import annotations

import numpy as np

HEX_TO_TETS = ((0, 1, 3, 7), (0, 4, 1, 7), (1, 2, 3, 7), (1, 6, 2, 7), (1, 4, 5, 7), (1, 5, 6, 7))
# wedge (s_j, a, b, s_j+1, c, d): diagonals s_j-c, s_j-d on the faces through the singular edge, b-c on the outer quad
WEDGE_TO_TETS = ((0, 1, 2, 4), (0, 2, 5, 4), (0, 3, 4, 5))


def ellipsoid_point(theta, phi, rp, inner_radius=0.7, outer_radius=1.0, apex_inner=1.3, apex_outer=1.5):
    """_ellipsoid_point with septum_flatness = 0, axis_ratio = 1, eccentricity = 0 (the fan variant)."""
    radius = inner_radius * (1.0 - rp) + outer_radius * rp
    z = np.where(theta < np.pi / 2, (apex_inner * (1.0 - rp) + apex_outer * rp) * np.cos(theta), apex_outer * np.cos(theta))
    x = radius * (np.cos(phi) * np.sin(theta))
    y = radius * np.sin(phi) * np.sin(theta)
    return np.stack([x, y, z], axis=-1)


def generate_ideal_lv_mesh(nc: int, nr: int, nl: int, inner_radius=0.7, outer_radius=1.0, longitudinal_upper=0.2,
                           apex_inner=1.3, apex_outer=1.5):
    """Returns (nodes [n,3], hexes [nh,8], wedges [nw,6], params [n,3] = (theta, phi, rp) per node); 0-based ids."""
    phi = np.linspace(0.0, 2 * np.pi, nc + 1)[:-1]
    rps = np.linspace(0.0, 1.0, nr + 1)
    thetas = np.linspace(0.0, (1.0 + longitudinal_upper) * np.pi / 2, nl + 2)[1:]
    T, R, P = np.meshgrid(thetas, rps, phi, indexing="ij")          # ring (slowest), radial, circumferential (fastest)
    prm = np.stack([T.ravel(), P.ravel(), R.ravel()], axis=1)
    kw = dict(inner_radius=inner_radius, outer_radius=outer_radius, apex_inner=apex_inner, apex_outer=apex_outer)
    nodes = ellipsoid_point(prm[:, 0], prm[:, 1], prm[:, 2], **kw)
    n_ring = nc * (nr + 1) * (nl + 1)
    na = lambda i, j, k: (k * (nr + 1) + j) * nc + i                  # node_array[i,j,k], 0-based
    hexes = []
    for k in range(nl):
        for j in range(nr):
            for i in range(nc):
                i2 = (i + 1) % nc
                hexes.append((na(i, j, k), na(i2, j, k), na(i2, j + 1, k), na(i, j + 1, k),
                              na(i, j, k + 1), na(i2, j, k + 1), na(i2, j + 1, k + 1), na(i, j + 1, k + 1)))
    apex_prm = np.stack([np.zeros(nr + 1), np.zeros(nr + 1), rps], axis=1)
    apex_nodes = ellipsoid_point(apex_prm[:, 0], apex_prm[:, 1], apex_prm[:, 2], **kw)
    nodes = np.concatenate([nodes, apex_nodes])
    prm = np.concatenate([prm, apex_prm])
    wedges = []
    for j in range(nr):
        for i in range(nc):
            i2 = (i + 1) % nc
            s = n_ring + j
            wedges.append((s, na(i, j, 0), na(i2, j, 0), s + 1, na(i, j + 1, 0), na(i2, j + 1, 0)))
    return nodes, np.array(hexes, dtype=np.int64), np.array(wedges, dtype=np.int64), prm


def tetrahedralize(nodes, hexes, wedges):
    """Conforming tet mesh (this project's addition): 6 tets per hexahedron, 3 per wedge, positively oriented."""
    tets = [hexes[:, list(t)] for t in HEX_TO_TETS]
    tets = np.stack(tets, axis=1).reshape(-1, 4) if len(hexes) else np.empty((0, 4), dtype=np.int64)
    wt = [wedges[:, list(t)] for t in WEDGE_TO_TETS]
    wt = np.stack(wt, axis=1).reshape(-1, 4) if len(wedges) else np.empty((0, 4), dtype=np.int64)
    tets = np.concatenate([tets, wt])
    X = nodes[tets]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])
    flip = vol < 0
    tets[flip] = tets[flip][:, [0, 2, 1, 3]]
    if np.any(np.abs(vol) < 1e-14):
        raise ValueError("degenerate tetrahedron in the LV split")
    return tets


def _normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _rotate_around(v, a, theta):
    """src/utils.jl: v cos(t) + (a x v) sin(t) + a (a.v)(1 - cos(t))"""
    c, s = np.cos(theta)[..., None], np.sin(theta)[..., None]
    return v * c + np.cross(a, v) * s + a * np.sum(a * v, axis=-1, keepdims=True) * (1.0 - c)


def odb25lt_fibres(prm, cells, alpha_endo=np.deg2rad(60.0), alpha_epi=np.deg2rad(-60.0), beta_endo=0.0, beta_epi=0.0,
                   gamma_endo=0.0, gamma_epi=0.0, **geo):
    """f, s, n per cell and local node: array [ncells, nv, 3, 3] (last-but-one axis: f, s, n)."""
    theta, phi, rp = prm[:, 0], prm[:, 1], prm[:, 2]
    eps = 1e-6
    th = np.maximum(theta, 1e-3)                                     # the azimuth is undefined on the apex edge
    x = lambda t, p, r: ellipsoid_point(t, p, r, **geo)
    transmural = _normalize(x(th, phi, np.minimum(rp + eps, 1.0)) - x(th, phi, np.maximum(rp - eps, 0.0)))
    circumferential = _normalize(x(th, phi + eps, rp) - x(th, phi - eps, rp))
    apicobasal = _normalize(x(th + eps, phi, rp) - x(th - eps, phi, rp))
    a = (1 - rp) * alpha_endo + rp * alpha_epi
    b = (1 - rp) * beta_endo + rp * beta_epi
    g = (1 - rp) * gamma_endo + rp * gamma_epi
    f0 = _normalize(_rotate_around(circumferential, transmural, a))
    f0 = _normalize(_rotate_around(f0, apicobasal, -b))
    s0 = _normalize(_rotate_around(circumferential, transmural, a + np.pi / 2.0))
    s0 = _normalize(s0 - np.sum(s0 * f0, axis=-1, keepdims=True) * f0)
    s0 = _normalize(_rotate_around(s0, f0, -g))
    n0 = _normalize(np.cross(f0, s0))
    fsn_nodes = np.stack([f0, s0, n0], axis=1)                       # [nnodes, 3, 3]
    return fsn_nodes[cells]                                          # [ncells, nv, 3, 3]
