"""MESHER database -> module variables: the native reader
axisem_b200/hostcxx/meshdb.cpp against databases written record for record as
MESHER/pdb.f90:2205-2382 writes them (axisem_b200/host/meshdb_io.py) for the synthetic
meshes, whose arrays are known.  Index maps must come back bit-exact, and the quantities the
SOLVER derives in def_grid (axis flags, glob2el) must equal the host builder's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from axisem_b200.capi import fortran_matrix
from axisem_b200.host import prem_mesh_spec
from axisem_b200.host.mesh import build_rank
from axisem_b200.host.meshdb_io import read_axbprob, write_meshdb
from axisem_b200.host.spectral import SpectralBasis

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_meshdb2axbp")


def _exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["bash", os.path.join(ROOT, "axisem_b200", "hostcxx", "build.sh")])
    return EXE


def _convert(mesh, tmp_path, rank):
    db = str(tmp_path / f"meshdb.dat{rank:04d}")
    out = str(tmp_path / f"mesh{rank}.axbp")
    write_meshdb(mesh, db, period=50.0, courant=0.6, dt=0.25)
    r = subprocess.run([_exe(), db, str(rank), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the converter checks the database's GLL/GLJ arrays against the native spectral basis
    dev = float(r.stdout.split("max deviation")[1].split()[0])
    assert dev < 1e-6, r.stdout
    return read_axbprob(out), db


@pytest.mark.parametrize("nranks", [1, 2, 4])
def test_meshdb_round_trip(nranks, tmp_path):
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    basis = SpectralBasis(4)
    for rank in range(nranks):
        mesh = build_rank(spec, rank, nranks, basis)
        v, _ = _convert(mesh, tmp_path, rank)
        assert int(v["data_proc%nproc"]) == nranks and int(v["data_proc%mynum"]) == rank
        assert int(v["data_mesh%npol"]) == 4
        for k in ("nel_solid", "nel_fluid", "nglob_solid", "nglob_fluid", "nel_bdry"):
            assert int(v["data_mesh%" + k]) == int(getattr(mesh, k)), k
        for k in ("igloc_solid", "igloc_fluid", "ax_el_solid", "ax_el_fluid", "axis_solid", "axis_fluid"):
            assert np.array_equal(v["data_mesh%" + k], np.asarray(getattr(mesh, k)).reshape(-1)), k
        if mesh.nel_bdry:
            for k in ("bdry_solid_el", "bdry_fluid_el", "bdry_jpol_solid", "bdry_jpol_fluid"):
                assert np.array_equal(v["data_mesh%" + k], getattr(mesh, k)), k
        assert np.array_equal(v["data_spec%G0"], basis.G0)
        for k in ("G1", "G1T", "G2", "G2T"):
            assert np.array_equal(v["data_spec%" + k], fortran_matrix(getattr(basis, k))), k
        assert np.array_equal(v["data_spec%eta"], basis.eta) and np.array_equal(v["data_spec%wt"], basis.wt)
        assert float(v["data_time%deltat"]) == 0.25 and float(v["data_time%period"]) == 50.0
        assert float(v["data_mesh%router"]) == float(spec.router)
        assert bytes(v["data_mesh%bkgrdmodel"].astype(np.uint8)).decode() == "prem_iso"
        assert np.all(v["data_mesh%eltype"] == 0)
        for dom, hs in (("solid", mesh.halo_solid), ("fluid", mesh.halo_fluid)):
            assert int(v[f"data_comm%sizerecv_{dom}"]) == hs.nmsg
            if hs.nmsg:
                assert np.array_equal(v[f"data_comm%listrecv_{dom}"], hs.list_peer)
                assert np.array_equal(v[f"data_comm%sizemsgrecv_{dom}"], hs.sizemsg)
                assert np.array_equal(v[f"data_comm%glocal_index_msg_recv_{dom}"], hs.glocal_index_msg)
                # def_grid.f90:95-180 restated in the reader == the host builder's list
                assert int(v[f"data_comm%num_comm_gll_{dom}"]) == hs.num_comm_gll
                assert np.array_equal(v[f"data_comm%glob2el_{dom}"], hs.glob2el.T)


def test_meshdb_reader_rejects_damaged_files(tmp_path):
    spec = prem_mesh_spec(ntheta=8, nr_target=10)
    mesh = build_rank(spec, 0, 1, SpectralBasis(4))
    _, db = _convert(mesh, tmp_path, 0)
    raw = open(db, "rb").read()
    # truncated
    (tmp_path / "short").write_bytes(raw[: len(raw) // 2])
    r = subprocess.run([_exe(), str(tmp_path / "short"), "0", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR" in r.stderr
    # npoint inconsistent with nelem (4th record)
    b = bytearray(raw)
    off = 3 * 12 + 4
    b[off:off + 4] = struct.pack("<i", 7)
    (tmp_path / "bad").write_bytes(bytes(b))
    r = subprocess.run([_exe(), str(tmp_path / "bad"), "0", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "inconsistent" in r.stderr
    # mismatching record markers
    b = bytearray(raw)
    b[8:12] = struct.pack("<i", 5)
    (tmp_path / "mark").write_bytes(bytes(b))
    r = subprocess.run([_exe(), str(tmp_path / "mark"), "0", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "markers" in r.stderr
