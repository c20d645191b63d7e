"""Output staging (src/ferrite-addons/io.jl:18-93): ParaViewWriter / store_timestep! / store_timestep_field! -- the VTU / PVD
files (CPU) and the asynchronous device snapshots behind them (GPU)."""
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _read_vtu(path):
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    pts = np.array(piece.find("Points/DataArray").text.split(), dtype=float).reshape(-1, 3)
    arr = {d.get("Name"): np.array(d.text.split(), dtype=float) for d in piece.findall("Cells/DataArray")}
    pd = {d.get("Name"): np.array(d.text.split(), dtype=float) for d in piece.findall("PointData/DataArray")}
    return int(piece.get("NumberOfPoints")), int(piece.get("NumberOfCells")), pts, arr, pd


def test_vtu_and_pvd_files(tmp_path, tb, oracle):
    O = oracle
    m = O.generate_grid(O.HEX8, (3, 2, 2), (0, 0, 0), (3, 2, 2))
    data = np.arange(m.nnodes, dtype=float) * 0.5
    tb.io._write_vtu(tmp_path / "a.vtu", m.coords, m.conn, tb.Hexahedron, {"φₘ": data})
    npts, ncells, pts, arr, pd = _read_vtu(tmp_path / "a.vtu")
    assert (npts, ncells) == (m.nnodes, m.ncells) and np.array_equal(pts, m.coords)
    assert np.array_equal(arr["connectivity"].reshape(-1, 8), m.conn) and np.all(arr["types"] == 12)
    assert np.array_equal(arr["offsets"], 8 * np.arange(1, m.ncells + 1)) and np.array_equal(pd["φₘ"], data)


@pytest.mark.gpu
def test_staged_snapshots_match_the_state_at_their_time(tmp_path, tb, dev, oracle):
    """store_timestep! inside the time loop: the snapshot is staged asynchronously, stepping continues (and overwrites the
    state), the file written later holds the state AS OF the store call; ring of 2 with 5 stored steps forces recycling"""
    md = tb.generate_mesh(tb.Hexahedron, (10, 8, 6), (0, 0, 0), (2.5, 2.0, 1.5), device=dev)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, np.diag([0.03, 0.013, 0.013]), 1.0)
    ion = tb.FHNModel()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    x = md.dof_coords()
    N = md.ndofs
    u = tb.B200Vector.from_host(dev, np.concatenate([np.where(x[:, 0] < 0.8, 1.0, 0.0), np.zeros(N)]), 2)
    io = tb.io.ParaViewWriter(str(tmp_path / "wave"), ring=2)
    expected = {}
    for step in range(10):
        st.step(u, float(step), 1.0)
        if step % 2 == 1:
            t = float(step + 1)
            tb.io.store_timestep(io, t, md, lambda w: tb.io.store_timestep_field_(w, t, md, u, "φₘ"))
            expected[t] = None                  # filled below from an independent, synchronous download
            expected[t] = u.column(0)
    tb.io.finalize_(io)
    conn, coords, celldofs = md.download()
    node2dof = np.zeros(coords.shape[0], dtype=np.int64)
    node2dof[conn.ravel()] = celldofs.ravel()
    pvd = ET.parse(str(tmp_path / "wave.pvd")).getroot()
    entries = pvd.findall("Collection/DataSet")
    assert [float(e.get("timestep")) for e in entries] == sorted(expected)
    for e in entries:
        npts, ncells, pts, arr, pd = _read_vtu(tmp_path / e.get("file"))
        assert npts == md.nnodes and ncells == md.ncells
        assert np.array_equal(pd["φₘ"], expected[float(e.get("timestep"))][node2dof])
