"""Output staging: the reference's `ParaViewWriter` / `store_timestep!` / `store_timestep_field!` (src/ferrite-addons/io.jl:18-93)
over the C ABI's snapshot staging (tb_vec_stage_col / tb_stage_wait).

The reference evaluates the field at the grid nodes and hands it to WriteVTK inside the time loop (blocking).  Here a snapshot
of the state column leaves the device on its own copy stream into a pinned ring while stepping continues; the file is written
on the host when the slot is recycled or at `finalize!`.  Files: one `<name>/<t>.vtu` (VTK XML UnstructuredGrid, ASCII) per
stored time step and `<name>.pvd` as the collection, like WriteVTK's `paraview_collection`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import _lib as L
from .core import B200Vector, DeviceMesh

_VTK_TYPE = {L.QUAD4: 9, L.HEX8: 12, L.TRI3: 5, L.TET4: 10}


def _write_vtu(path: Path, points: np.ndarray, cells: np.ndarray, celltype: int, point_data: dict):
    npts, (ncells, nv) = points.shape[0], cells.shape
    p3 = np.zeros((npts, 3))
    p3[:, :points.shape[1]] = points
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian">\n')
        f.write(f'<UnstructuredGrid><Piece NumberOfPoints="{npts}" NumberOfCells="{ncells}">\n')
        f.write('<Points><DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        np.savetxt(f, p3, fmt="%.17g")
        f.write('</DataArray></Points>\n<Cells>\n<DataArray type="Int64" Name="connectivity" format="ascii">\n')
        np.savetxt(f, cells, fmt="%d")
        f.write('</DataArray>\n<DataArray type="Int64" Name="offsets" format="ascii">\n')
        np.savetxt(f, (np.arange(1, ncells + 1) * nv)[None, :], fmt="%d")
        f.write('</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
        np.savetxt(f, np.full((1, ncells), _VTK_TYPE[celltype]), fmt="%d")
        f.write('</DataArray>\n</Cells>\n<PointData>\n')
        for name, data in point_data.items():
            f.write(f'<DataArray type="Float64" Name="{name}" format="ascii">\n')
            np.savetxt(f, np.asarray(data, dtype=np.float64)[None, :], fmt="%.17g")
            f.write('</DataArray>\n')
        f.write('</PointData>\n</Piece></UnstructuredGrid>\n</VTKFile>\n')


class ParaViewWriter:
    """ParaViewWriter(filename): io.jl:3-16.  `ring` = number of snapshots that may be in flight (staged but not yet written)."""

    def __init__(self, filename: str, ring: int = 2):
        self.filename = str(filename)
        self.entries = []                       # (t, relative file) of the collection
        self.current = None                     # (t, grid arrays, point data dict) of the open time step
        self.ring = max(1, int(ring))
        self._slots = []                        # pinned buffers: [ptr, nbytes, pending job or None]
        self._mesh_cache = {}
        self._dev = None

    # ---- pinned ring ------------------------------------------------------------------------------------------------
    def _slot(self, dev, n):
        for s in self._slots:
            if s[2] is None and s[1] >= n * 8:
                return s
        if len(self._slots) >= self.ring:        # recycle the oldest pending slot: its copy has to land, its file gets written
            self._flush(dev, self._slots[0])
            s = self._slots.pop(0)
            self._slots.append(s)
            if s[1] >= n * 8:
                return s
            L.call("tb_host_free", s[0])
            self._slots.remove(s)
        p = C.c_void_p()
        L.call("tb_host_alloc", int(n * 8), C.byref(p))
        s = [p, n * 8, None]
        self._slots.append(s)
        return s

    def _flush(self, dev, slot):
        job = slot[2]
        if job is None:
            return
        if not isinstance(job, tuple):
            return
        L.call("tb_stage_wait", dev.h)
        t, grid, fields = job
        nodes, cells, celltype, node2dof = grid
        data = {}
        for name, (n, sl) in fields.items():
            u = np.ctypeslib.as_array(C.cast(sl[0], C.POINTER(C.c_double)), shape=(n,))
            data[name] = u[node2dof]            # Lagrange-1: the value at a grid node is its dof's value (_evaluate_at_grid_nodes)
        self._write(t, nodes, cells, celltype, data)
        for name, (n, sl) in fields.items():
            sl[2] = None

    def _write(self, t, nodes, cells, celltype, data):
        Path(self.filename).mkdir(parents=True, exist_ok=True)
        rel = f"{Path(self.filename).name}/{t}.vtu"
        _write_vtu(Path(self.filename) / f"{t}.vtu", nodes, cells, celltype, data)
        self.entries.append((t, rel))
        with open(self.filename + ".pvd", "w") as f:                      # "this updates the PVD file", io.jl:84-86
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">\n<Collection>\n')
            for tt, r in self.entries:
                f.write(f'<DataSet timestep="{tt}" part="0" file="{r}"/>\n')
            f.write('</Collection>\n</VTKFile>\n')

    def _grid_of(self, dh: DeviceMesh):
        key = id(dh)
        if key not in self._mesh_cache:
            conn, coords, celldofs = dh.download()
            node2dof = np.zeros(coords.shape[0], dtype=np.int64)
            node2dof[conn.ravel()] = celldofs.ravel()
            self._mesh_cache[key] = (coords, conn, dh.celltype, node2dof)
        return self._mesh_cache[key]


def store_timestep_(io: ParaViewWriter, t, grid: DeviceMesh):
    """store_timestep!(io, t, grid): open the time step (io.jl:18-26)"""
    if io.current is None:
        io._dev = grid.dev
        io.current = (t, io._grid_of(grid), {}, grid.dev)


def store_timestep_field_(io: ParaViewWriter, t, dh: DeviceMesh, u, sym: str, name: str | None = None, col: int = 0):
    """store_timestep_field!(io, t, dh, u, sym, name) (io.jl:34-58).  u: B200Vector (state column `col`; staged asynchronously) or a
    host array of ndofs values (written as is)."""
    assert io.current is not None
    name = name or sym
    tcur, grid, fields, dev = io.current
    if isinstance(u, B200Vector):
        slot = io._slot(dev, u.n)
        L.call("tb_vec_stage_col", u.h, int(col), slot[0])
        slot[2] = "staged"
        fields[name] = (u.n, slot)
    else:
        fields[name] = np.ascontiguousarray(u, dtype=np.float64)


def finalize_timestep_(io: ParaViewWriter, t):
    """finalize_timestep!(io, t) (io.jl:79-86): host arrays are written now; staged device snapshots when their slot is recycled
    or at finalize_ -- the time loop does not wait for the copy"""
    tcur, grid, fields, dev = io.current
    staged = {k: v for k, v in fields.items() if isinstance(v, tuple)}
    if not staged:
        nodes, cells, celltype, node2dof = grid
        io._write(tcur, nodes, cells, celltype, {k: v[node2dof] for k, v in fields.items()})
    else:
        if len(staged) != len(fields):
            raise ValueError("mixing staged device fields and host fields in one time step is not supported")
        job = (tcur, grid, staged)
        for n, sl in staged.values():
            sl[2] = job
    io.current = None


def store_timestep(io: ParaViewWriter, t, grid: DeviceMesh, fn):
    """store_timestep!(f, io, t, grid) (io.jl:28-32)"""
    store_timestep_(io, t, grid)
    fn(io)
    finalize_timestep_(io, t)


def finalize_(io: ParaViewWriter):
    """finalize!(io) (io.jl:88-90): lands and writes every snapshot still in flight, releases the pinned ring"""
    for s in list(io._slots):
        if isinstance(s[2], tuple):
            io._flush(io._dev, s)
    for s in io._slots:
        L.call("tb_host_free", s[0])
    io._slots = []
