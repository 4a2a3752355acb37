"""The host mirror's ExodusII ingest (NetCDF classic CDF-1/CDF-2 parsed directly,
xyst_b200/host/exodus.cpp; cf. src/IO/ExodusIIMeshReader.cpp) and diagnostics writer
(src/IO/DiagWriter.cpp) -- no GPU. The reader must deliver exactly what the committed fixtures
hold (those were flattened from the same files with scipy's NetCDF reader by
tests/golden/make_mesh_fixtures.py), and a solver created from the file must build the same
setup as one created from the arrays."""
import glob
import os
import numpy as np
import pytest
import oraclelib as O
from xyst_b200 import hostapi as H, capi
from host_common import fixture_to_host_mesh

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference/tests/regression/inciter"
FILES = {"riecg_sod": os.path.join(GOLDEN, "riecg_sod.exo"),
         "riecg_sedov": REF + "/RieCG/Sedov/sedov_coarse.exo",
         "riecg_taylor_green": REF + "/RieCG/TaylorGreen/unitcube_1k.exo",
         "laxcg_bump": REF + "/LaxCG/Bump/bump.exo",
         "chocg_poiseuille": REF + "/ChoCG/Poiseuille/poiseuille1tetz.exo",
         "chocg_pidiv4": REF + "/ChoCG/Poisson/unitcube_0pidiv4_1k.exo",
         "sphere2_5k": REF + "/ChoCG/Sphere/sphere2_5K.exo",
         "unitsquare_3_6k": REF + "/ZalCG/SlotCyl/unitsquare_01_3.6k.exo",
         "riecg_canyon": REF + "/RieCG/Canyon/canyon.exo"}


def sets_of(m):
    out = {}
    for i, s in enumerate(m["set_id"]):
        a, b = int(m["set_off"][i]), int(m["set_off"][i + 1])
        out[int(s)] = np.asarray(m["set_tri"][a:b], np.int64).tolist()
    return out


@pytest.mark.parametrize("case", list(FILES))
def test_reader_matches_fixture(case):
    path = FILES[case]
    if not os.path.exists(path):
        pytest.skip("mesh file not available here")
    m = H.exo_read(path)
    hm = fixture_to_host_mesh(O.load_mesh(case))
    assert np.array_equal(m["coord"], hm["coord"])
    assert np.array_equal(m["tets"], np.asarray(hm["tets"], np.uint64))
    assert sets_of(m) == sets_of(hm)


def test_solver_from_file_equals_solver_from_arrays():
    kw = O.CASES["riecg_sod"]
    a = H.Solver.exo(H.make_cfg(**kw), FILES["riecg_sod"])
    hm = fixture_to_host_mesh(O.load_mesh("riecg_sod"))
    b = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    for s in (a, b):
        s.prepare(); s.host_setup()
    for n in ("gid", "inpoel", "x", "vol", "dsupedge0", "dsupint0", "dsupedge1", "dsupint1", "triinpoel",
              "symbcnodes", "symbcnorms", "u0"):
        assert np.array_equal(a.get(n), b.get(n)), n


def test_reader_rejects_what_it_cannot_read(tmp_path):
    p = tmp_path / "x.exo"
    p.write_bytes(b"not a mesh")
    with pytest.raises(capi.XystError, match="Not a NetCDF classic"):
        H.exo_read(str(p))
    with pytest.raises(capi.XystError, match="Cannot open"):
        H.exo_read(str(tmp_path / "missing.exo"))
    p.write_bytes(b"CDF\x05" + b"\0" * 64)
    with pytest.raises(capi.XystError, match="Unsupported NetCDF"):
        H.exo_read(str(p))
    data = open(FILES["riecg_sod"], "rb").read()
    p.write_bytes(data[:4000])                      # truncated
    with pytest.raises(capi.XystError):
        H.exo_read(str(p))
