"""The partitioner, pinned to the reference's own: oracle/_ref/libzoltan_ref.so is the reference's vendored Zoltan
3.901 (src/zoltan, compiled where it lies against a one-rank MPI stub, oracle/stub/mpi) behind a restatement of the
reference's call into it (inciter::geomPartMesh, src/Partition/ZoltanGeom.cpp:139-244; oracle/zoltan_geom.c). The
host mirror's rcb() and rib() (xyst_b200/host/mesh.cpp: Zoltan's serial_rcb / serial_rib / inertial3d / find_median /
average-cut restated) must give every element the part Zoltan gives it -- on every regression mesh and on box meshes, for 2..8 parts, powers of
two or not. CPU only; skipped where the reference tree (and with it the library) is absent."""
import ctypes as C
import os
import numpy as np
import pytest
import oraclelib as O
from xyst_b200 import hostapi as H
from host_common import fixture_to_host_mesh

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libzoltan_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libzoltan_ref.so not built (needs /root/reference)")

MESHES = ["riecg_sod", "riecg_sedov", "riecg_taylor_green", "laxcg_bump", "chocg_unitcube", "chocg_pidiv4",
          "chocg_poiseuille", "sphere2_5k", "unitsquare_3_6k", "riecg_canyon"]


def zoltan(alg, coord, tets, n):
    L = C.CDLL(LIB)
    L.orc_zoltan_geom.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p]
    t = tets.astype(np.int64)
    cen = [np.ascontiguousarray((c[t[:, 0]] + c[t[:, 1]] + c[t[:, 2]] + c[t[:, 3]]) / 4.0) for c in coord]   # ZoltanGeom.cpp:133-135
    out = np.full(len(t), -1, np.int32)
    rc = L.orc_zoltan_geom(alg.encode(), len(t), cen[0].ctypes.data, cen[1].ctypes.data, cen[2].ctypes.data, n,
                           out.ctypes.data)
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("name", MESHES)
def test_own_rcb_is_zoltans_rcb_on_the_regression_meshes(name):
    hm = fixture_to_host_mesh(O.load_mesh(name))
    for n in (2, 3, 4, 5, 7, 8):
        z = zoltan("RCB", hm["coord"], hm["tets"], n)
        m = H.rcb(hm["coord"], hm["tets"], n)
        assert z.min() == 0 and z.max() == n - 1
        assert np.array_equal(z, m), (name, n, int((z != m).sum()))


@pytest.mark.parametrize("name", MESHES)
def test_own_rib_is_zoltans_rib_on_the_regression_meshes(name):
    """part = "rib" (ZalCG/Bump/*.q): recursive inertial bisection -- centre of mass, inertia tensor, eigenvector of
    the largest eigenvalue (cubic roots + pivoted elimination, rcb/inertial3d.c), projections, the same median."""
    hm = fixture_to_host_mesh(O.load_mesh(name))
    for n in (2, 3, 4, 5, 7, 8):
        z = zoltan("RIB", hm["coord"], hm["tets"], n)
        assert np.array_equal(z, H.rib(hm["coord"], hm["tets"], n)), (name, n)


@pytest.mark.parametrize("dims", [(8, 8, 8), (12, 6, 4), (5, 7, 9)])
def test_own_rcb_is_zoltans_rcb_on_box_meshes(dims):
    """Structured meshes: many centroids share a coordinate, so the cuts go through Zoltan's tie handling (dots on
    the median are moved one by one in list order until the target weight is met, par_median.c:373-412)."""
    m = H.box_mesh(*dims)
    for n in (2, 3, 4, 6, 8):
        z = zoltan("RCB", m["coord"], m["tets"], n)
        assert np.array_equal(z, H.rcb(m["coord"], m["tets"], n)), (dims, n)
        assert np.array_equal(zoltan("RIB", m["coord"], m["tets"], n), H.rib(m["coord"], m["tets"], n)), (dims, n)


def test_chare_count_of_the_references_over_decomposition():
    """tk::linearLoadDistributor (Base/LoadDistributor.cpp:24-97; the -u command-line argument): the assertions of the
    reference's unit test (tests/unit/Base/TestLoadDistributor.cpp:57-160) and the values behind its -u goldens."""
    n, chunk, rem = H.chare_count(0.5, 1234, 2)
    assert n < 1235 and chunk < 1235 and rem < chunk and n * chunk + rem == 1234
    n0, chunk0, rem0 = H.chare_count(0.0, 1234, 2)
    assert (n0, chunk0, rem0) == (2, 617, 0)                       # no virtualization: one chare per PE
    n1, chunk1, rem1 = H.chare_count(1.0, 1234, 2)
    assert (n1, chunk1, rem1) == (1234, 1, 0)                      # full virtualization: one chare per element
    from xyst_b200 import capi
    for bad in ((-0.5, 1234, 2), (1.5, 1234, 2), (0.5, 1234, 0)):
        with pytest.raises(capi.XystError):
            H.chare_count(*bad)
    assert H.chare_count(0.5, 730, 4)[0] == 8                      # VorticalFlow -u 0.5 on 4 PEs (the HLLC goldens)
