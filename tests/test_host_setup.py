"""Host mirror (C++, xyst_b200/host) vs oracle: what RieCG's setup builds (no GPU).

Bit-exact: local renumbering (gid, inpoel), coordinates, nodal volumes, the set of edges
with their integrals, the tetrahedron superedges (ids AND integrals, in order), the
boundary faces per side set with orientation, symmetry-BC node lists and normals, ICs.
Triangle superedges are compared in order too: the host mirror reproduces the reference's
hash-set walk (xyst_b200/host/siphash.hpp explains why that is part of the algorithm).
Allowed to differ (documented in DESIGN.md): the ORDER of the single-edge list, of the side
sets inside triinpoel and of the Dirichlet node list -- none of which the device depends on.
"""
import numpy as np
import pytest
import oraclelib as O
from xyst_b200 import hostapi as H
from host_common import fixture_to_host_mesh, host_mesh_to_oracle, edge_dict, face_multiset

EXACT = ["gid", "inpoel", "x", "y", "z", "vol", "v", "dsupedge0", "dsupint0", "dsupedge1", "dsupint1"]


def compare(o, s, kw):
    for n in EXACT:
        assert np.array_equal(o.get(n), s.get(n)), n
    eo, es = edge_dict(o.get), edge_dict(s.get)
    assert eo.keys() == es.keys()
    assert all(eo[k] == es[k] for k in eo)                       # integrals bitwise
    # single edges: same set with the same orientation (order is hash order in the reference)
    so = sorted(map(tuple, o.get("dsupedge2").reshape(-1, 2).tolist()))
    ss = sorted(map(tuple, s.get("dsupedge2").reshape(-1, 2).tolist()))
    assert so == ss
    assert face_multiset(o.get("triinpoel"), o.get("bface")) == face_multiset(s.get("triinpoel"), s.get("bface"))
    assert np.array_equal(o.get("symbcnodes"), s.get("symbcnodes"))
    assert np.array_equal(o.get("symbcnorms"), s.get("symbcnorms"))
    # besym is per face node: compare as multiset keyed by node
    def bes(g):
        return sorted(zip(g("triinpoel").tolist(), g("besym").tolist()))
    assert bes(o.get) == bes(s.get)
    dm_o = o.get("dirbcmasks").reshape(-1, 6); dm_s = s.get("dirbcmasks").reshape(-1, 6)
    assert np.array_equal(dm_o[np.argsort(dm_o[:, 0])], dm_s[np.argsort(dm_s[:, 0])])
    assert np.array_equal(o.get("u"), s.get("u0"))
    assert s.scalar("meshvol") == o.scalar("meshvol")


@pytest.mark.parametrize("case", list(O.CASES))
def test_regression_meshes(case):
    kw = O.CASES[case]
    mesh = O.load_mesh(case)
    o = O.Oracle(mesh, O.make_cfg(**kw), "port")
    hm = fixture_to_host_mesh(mesh)
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.host_setup()
    compare(o, s, kw)


@pytest.mark.parametrize("n", [(3, 3, 3), (6, 4, 5)])
def test_box_mesh(n):
    kw = dict(O.CASES["riecg_sedov"], sym=(1, 3, 5))
    m = H.box_mesh(*n, 1.2, 1.2, 1.2)
    npn = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
    assert m["coord"].shape == (3, npn) and len(m["tets"]) == 6 * n[0] * n[1] * n[2]
    o = O.Oracle(host_mesh_to_oracle(m), O.make_cfg(**kw), "port")     # also checks J > 0
    s = H.Solver.box(H.make_cfg(**kw), *n, 1.2, 1.2, 1.2)
    s.prepare(); s.host_setup()
    compare(o, s, kw)
    nx, ny, nz = n
    assert s.scalar("nedge") == 7 * nx * ny * nz + 3 * (nx * ny + ny * nz + nx * nz) + nx + ny + nz


def test_no_bc_means_no_boundary_faces():
    m = H.box_mesh(3, 3, 3)
    kw = dict(problem="sod", gamma=1.4, cfl=0.5)
    s = H.Solver.mesh(H.make_cfg(**kw), m["coord"], m["tets"], m["set_id"], m["set_off"], m["set_tri"])
    s.prepare(); s.host_setup()
    assert s.scalar("ntri") == 0
    o = O.Oracle(host_mesh_to_oracle(m), O.make_cfg(**kw), "port")
    assert len(o.get("triinpoel")) == 0


def test_rcb_matches_box_part_ranges():
    n = 4
    m = H.box_mesh(n, n, n)
    for nparts in (2, 4, 8):
        part = H.rcb(m["coord"], m["tets"], nparts)
        cen = m["coord"][:, m["tets"].astype(np.int64)].mean(axis=2) * n      # hex units
        for p in range(nparts):
            r = H.box_part_range(n, n, n, nparts, p).astype(float)
            inside = np.all([(cen[d] > r[2 * d]) & (cen[d] < r[2 * d + 1]) for d in range(3)], axis=0)
            assert np.array_equal(inside, part == p), (nparts, p)


def test_face_set_iteration_order_equals_libstdcxx_unordered_set():
    """RefOrderFaceSet (flat arrays) must walk surviving faces in exactly the order of the
    container the reference uses, std::unordered_set<Face, SipHash, Eq> -- across rehashes,
    duplicate inserts in other node orders, and erasures."""
    import ctypes as C
    L = H.lib()
    L.xyst_test_faceset_order.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
    rng = np.random.default_rng(7)
    for n, maxid in [(5, 10), (40, 12), (1000, 60), (50000, 4000), (300000, 200000)]:
        f = rng.integers(0, maxid, size=(n, 3)).astype(np.uint64)
        f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
        e = f[rng.integers(0, len(f), size=len(f) // 3)][:, ::-1].copy()      # erase by permuted ids
        a = np.zeros((len(f), 3), np.uint64); b = np.zeros((len(f), 3), np.uint64)
        k = C.c_size_t()
        assert L.xyst_test_faceset_order(len(f), f.ctypes.data, len(e), e.ctypes.data,
                                         a.ctypes.data, b.ctypes.data, C.byref(k)) == 0
        assert k.value > 0
        assert np.array_equal(a[:k.value], b[:k.value]), (n, maxid)


@pytest.mark.parametrize("case", list(O.ZCASES))
def test_zalcg_setup_matches_oracle(case):
    """ZalCG variant of the setup: no renumbering, 4 integrals per edge (normal + J/120)."""
    kw = O.ZCASES[case]
    mesh = O.load_mesh(kw["mesh"])
    o = O.Oracle(mesh, O.make_cfg(**kw), "port")
    hm = fixture_to_host_mesh(mesh)
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.host_setup()
    for n in ["gid", "inpoel", "x", "vol", "dsupedge0", "dsupint0", "dsupedge1", "dsupint1", "symbcnodes", "symbcnorms"]:
        assert np.array_equal(o.get(n), s.get(n)), n
    assert np.array_equal(np.sort(o.get("dsupint2").reshape(-1, 4), axis=0), np.sort(s.get("dsupint2").reshape(-1, 4), axis=0))
    assert np.array_equal(o.get("gid"), np.arange(len(o.get("gid")), dtype=np.uint64))      # not renumbered


@pytest.mark.parametrize("case", ["chocg_poiseuille_damp4", "chocg_ldc", "chocg_poisson_neumann",
                                  "lohcg_poiseuille_damp4", "lohcg_ldc"])
def test_chocg_setup(case):
    """ChoCG: stride-5 edge integrals (normal, J/120, Laplacian term; ChoCG.cpp:399-446), Dirichlet
    masks/values of velocity and pressure (:210-300), no-slip nodes (:655-682), the pressure
    Poisson matrix in tk::CSR form (:146-188) -- all bitwise equal to the oracle's. LohCG: the same
    with stride-4 integrals (normal, Laplacian term; LohCG.cpp:407-453) and four unknowns."""
    kw = {**O.CCASES, **O.HCASES}[case]
    st, nc = (4, 4) if kw["solver"] == "lohcg" else (5, 3)
    mesh = O.load_mesh(kw["mesh"])
    o = O.Oracle(mesh, O.make_cfg(**kw), "port")
    hm = fixture_to_host_mesh(mesh)
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.host_setup()
    for n in EXACT + ["plhs_ia", "plhs_ja", "plhs_a", "noslipbcnodes", "symbcnodes", "symbcnorms"]:
        assert np.array_equal(o.get(n), s.get(n)), n
    # single edges: same (edge, 5 integrals) set; order is hash order in the reference
    def singles(g):
        e = g("dsupedge2").reshape(-1, 2); d = g("dsupint2").reshape(-1, st)
        return sorted((tuple(a), tuple(b)) for a, b in zip(e.tolist(), d.tolist()))
    assert singles(o.get) == singles(s.get)
    assert face_multiset(o.get("triinpoel"), o.get("bface")) == face_multiset(s.get("triinpoel"), s.get("bface"))
    for masks, vals, w in (("dirbcmasks", "dirbcval", nc + 1), ("dirbcmaskp", "dirbcvalp", 2)):
        mo, ms = o.get(masks).reshape(-1, w), s.get(masks).reshape(-1, w)
        assert np.array_equal(mo[np.argsort(mo[:, 0])], ms[np.argsort(ms[:, 0])]), masks
        vo, vs = o.get(vals).reshape(-1, w), s.get(vals).reshape(-1, w)
        assert np.array_equal(vo[np.argsort(vo[:, 0])], vs[np.argsort(vs[:, 0])]), vals
    assert s.scalar("meshvol") == o.scalar("meshvol")     # (the oracle's u0 already carries BC(t0): compared on the GPU)
