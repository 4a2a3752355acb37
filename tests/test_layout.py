"""CPU check of the device data layout (xyst_b200/csrc/layout.hpp): internal node order, owner
slots, incidence and incoming-edge lists, replayed on the host (tests/layout_check.cpp) for box
meshes and the regression fixtures; and how well the owner kernels' neighbour gathers coalesce."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import oraclelib as O
from xyst_b200 import hostapi as H
from host_common import fixture_to_host_mesh

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liblayout_check.so")


def _lib():
    src = os.path.join(HERE, "layout_check.cpp")
    dep = [src] + [os.path.join(HERE, "..", "xyst_b200", "csrc", f) for f in ("layout.hpp", "locality.hpp")]
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in dep):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fopenmp", "-Wl,-Bsymbolic", "-o", SO, src], check=True)
    return C.CDLL(SO)


def _check(s, reorder, tile=256):
    s.prepare(); s.host_setup()
    co = np.array([s.get("x"), s.get("y"), s.get("z")])
    npoin = co.shape[1]
    se = [np.ascontiguousarray(s.get("dsupedge%d" % k), np.uint64) for k in range(3)]
    si = [np.ascontiguousarray(s.get("dsupint%d" % k), np.float64) for k in range(3)]
    nsup = (C.c_size_t * 3)(len(se[0]) // 4, len(se[1]) // 3, len(se[2]) // 2)
    pe = (C.c_void_p * 3)(*[a.ctypes.data for a in se]); pi = (C.c_void_p * 3)(*[a.ctypes.data for a in si])
    stats = (C.c_size_t * 6)(); msg = C.create_string_buffer(256)
    x, y, z = (np.ascontiguousarray(co[i]) for i in range(3))
    rc = _lib().layout_check(C.c_size_t(npoin), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                             z.ctypes.data_as(C.c_void_p), nsup, pe, pi, C.c_size_t(3), reorder,
                             C.c_size_t(tile), stats, msg, C.c_size_t(256))
    assert rc == 0, msg.value.decode()
    return dict(zip(("ne", "nslot", "nent", "maxdeg", "sectors_x1000", "lines_x1000"), list(stats)))


@pytest.mark.parametrize("n,reorder,tile", [(6, 0, 256), (6, 1, 256), (17, 1, 256), (17, 1, 128), (24, 0, 256), (24, 1, 256)])
def test_box_layout(n, reorder, tile):
    cfg = H.make_cfg(problem="sedov", gamma=5.0 / 3.0, p0=1.0, cfl=0.5, sym=(1, 3, 5))
    st = _check(H.Solver.box(cfg, n, n, n), reorder, tile)
    assert st["ne"] == 7 * n ** 3 + 9 * n ** 2 + 3 * n
    assert st["maxdeg"] == 14
    # a warp-wide 16-byte gather of the other ends touches close to the ideal 16 sectors
    assert st["sectors_x1000"] < 22000
    print(n, reorder, tile, st)


@pytest.mark.parametrize("case", ["riecg_sod", "riecg_sedov"])
def test_unstructured_layout(case):
    kw = O.CASES[case]
    m = fixture_to_host_mesh(O.load_mesh(case))
    cfg = H.make_cfg(**kw)
    s = H.Solver.mesh(cfg, m["coord"], m["tets"], m["set_id"], m["set_off"], m["set_tri"])
    for reorder in (0, 1):
        st = _check(s, reorder)
        print(case, reorder, st)
