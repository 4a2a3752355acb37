"""Transported scalars on the device (ncomp = 5 + ns): scalar gradients, scalar MUSCL
(Riemann.cpp:145-209), scalar Riemann fluxes upwinded with the reconstructed flow states (:462-470,
:636-643), boundary flux, source, Dirichlet BCs and diagnostics of the scalar columns -- against the
oracle on the reference's RieCG/SlotCyl regression case (problems::slot_cyl: a cone, a hump and a
slotted cylinder carried by a rotating flow), kernel by kernel through the C ABI and as a full run
through the host mirror against the oracle's diagnostics and the reference's golden file."""
import numpy as np
import pytest
import oraclelib as O
import xyst_b200
from xyst_b200 import hostapi as H
from gpu_common import relerr
from host_common import fixture_to_host_mesh

pytestmark = pytest.mark.gpu
TOL = 1.0e-12


def _ctx(o, kw, exact):
    g = o.get
    ctx = xyst_b200.Context(device=0, flux=kw.get("flux", "rusanov"), gamma=kw["gamma"], exact_muscl=exact,
                            ncomp=kw["ncomp"])
    ctx.mesh_upload(g("x"), g("y"), g("z"), [g("dsupedge0"), g("dsupedge1"), g("dsupedge2")],
                    [g("dsupint0"), g("dsupint1"), g("dsupint2")], g("triinpoel"), g("besym"), g("vol"), g("v"))
    U0 = g("u")
    dm = g("dirbcmasks")
    dv = U0[dm.reshape(-1, kw["ncomp"] + 1)[:, 0].astype(np.int64)]
    ctx.bc_upload(dirbcmasks=dm, dirvals=dv, symbcnodes=g("symbcnodes"), symbcnorms=g("symbcnorms"))
    # slot_cyl::src (Problems.cpp:636-670): s = (0, -rho v, rho u, 0, 0, 0) of the prescribed (steady) rotation
    S = np.zeros_like(U0); S[:, 1] = -U0[:, 2]; S[:, 2] = U0[:, 1]
    ctx.src_upload(S)
    ctx.state_set(U0)
    return ctx


@pytest.mark.parametrize("case", ["riecg_slot_cyl", "riecg_slot_cyl_hllc"])
@pytest.mark.parametrize("exact", [True, False])
def test_scalar_grad_and_rhs_match_oracle(case, exact):
    kw = O.SCASES[case]
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    ctx = _ctx(o, kw, exact)
    for rep in range(2):                       # the initial state, then a developed one
        ctx.state_set(o.get("u"))
        ctx.grad(); ctx.rhs()
        G = ctx.grad_get(); R = ctx.rhs_get()
        o.kernel("grad"); o.kernel("rhs", 0, o.scalar("t"))
        Go, Ro = o.get("grad"), o.get("rhs")
        assert G.shape == Go.shape == (len(G), 18) and R.shape == Ro.shape == (len(R), 6)
        assert relerr(G[:, :15], Go[:, :15]) < TOL and relerr(G[:, 15:], Go[:, 15:]) < TOL
        # per component, on the scale of the momentum equations (the flow is two-dimensional: the
        # z-momentum rhs is rounding noise around 1e-18 in both)
        scale = np.abs(Ro[:, :5]).max()
        for c in range(5):
            assert np.abs(R[:, c] - Ro[:, c]).max() <= TOL * scale, c
        assert relerr(R[:, 5], Ro[:, 5]) < TOL
        o.step(5)


@pytest.mark.parametrize("case", ["riecg_rayleigh_taylor_st"])
def test_stationary_rayleigh_taylor_run_matches_oracle(case):
    """RieCG/RayleighTaylor with kappa = 0 (stationary variant; source term and Dirichlet values from the
    analytic solution): one more of the goldens that pinned the oracle only."""
    kw = O.OCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw.get("mesh", case)))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw.get("mesh", case)), O.make_cfg(**kw), "port")
    n = kw["nstep"]
    rows = s.step(n); o.step(n); d = o.diag()
    assert rows.shape == d.shape
    for c in range(1, d.shape[1]):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max() + 1e-15, c


@pytest.mark.parametrize("case", ["riecg_slot_cyl", "riecg_slot_cyl_hllc"])
def test_scalar_transport_run_matches_oracle_and_golden(case):
    kw = O.SCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    n = kw["nstep"]
    rows = s.step(n); o.step(n); d = o.diag()
    assert rows.shape == d.shape
    for c in range(1, d.shape[1]):          # (columns of the z-momentum are identically zero in the oracle)
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max() + 1e-15, c
    U, Uo = s.get("u"), o.get("u")
    for c in range(6):
        assert np.abs(U[:, c] - Uo[:, c]).max() <= 1e-11 * np.abs(Uo[:, c]).max() + 1e-15, c
    if case == "riecg_slot_cyl":               # tests/regression/inciter/RieCG/SlotCyl/diag.std, 12 printed digits
        gold = O.load_golden_diag(case)
        assert gold.shape == rows.shape
        assert (np.abs(rows - gold) <= 1e-10 * np.abs(gold) + 1e-15).all()


@pytest.mark.parametrize("far", [False, True])
def test_point_source_run_matches_oracle_and_golden(far):
    """RieCG/Canyon (problems::point_src: the scalar is set to 1 inside a sphere after every stage,
    RieCG.cpp:1023-1025; pressure BCs at inlet/outlet, symmetry walls) through the host mirror: against
    the oracle at 1e-11 and the reference's golden (printed with 6 digits). far=True: the variant with
    far-field BCs the oracle is pinned to as well."""
    kw = dict(O.OCASES["riecg_canyon_farfield"]) if far else dict(O.CANYON)
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    n = kw["nstep"]
    rows = s.step(n); o.step(n); d = o.diag()
    assert rows.shape == d.shape
    cols = [c for c in range(1, d.shape[1]) if c != 14]      # 14: increment norm of the pinned scalar, rounding noise
    for c in cols:
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max() + 1e-15, c
    # pointwise after 50 steps of a pressure-driven flow: on the scale of the momentum (the cross-flow
    # component is a small difference of large terms and carries the amplified rounding of both runs)
    U, Uo = s.get("u"), o.get("u")
    mom = np.abs(Uo[:, 1:4]).max()
    for c in range(6):
        scale = mom if c in (1, 2, 3) else np.abs(Uo[:, c]).max()
        assert np.abs(U[:, c] - Uo[:, c]).max() <= 1e-10 * scale + 1e-15, c
    if not far:
        gold = O.load_golden_diag("riecg_canyon")
        m = min(len(gold), len(rows))
        assert (np.abs(rows[:m, cols] - gold[:m, cols]) <= 6e-7 * np.abs(gold[:m, cols]) + 1e-15).all()


def test_scalar_configurations_that_are_not_implemented_fail_loudly():
    with pytest.raises(xyst_b200.XystError):
        xyst_b200.Context(device=0, ncomp=4)
    with pytest.raises(xyst_b200.XystError):
        xyst_b200.Context(device=0, ncomp=14)


# ---- ChoCG with a transported scalar (chorin::vgrad / adv_damp2 / adv_damp4 / boundary integral for the
# ---- scalar rows, time-dependent Dirichlet values, problems::point_src) ----------------------------------
PCASES = {"chocg_slot_cyl": O.SCASES["chocg_slot_cyl"], "chocg_slot_cyl_damp4": O.SCASES["chocg_slot_cyl_damp4"],
          "chocg_slot_cyl_damp4_freeze": O.SCASES["chocg_slot_cyl_damp4_freeze"], "chocg_sphere_point_src": O.SPHERE_SRC}


@pytest.mark.parametrize("case", list(PCASES))
def test_chocg_with_a_transported_scalar_matches_oracle_and_golden(case):
    """{ChoCG/SlotCyl/slot_cyl.q, slot_cyl_damp4.q, slot_cyl_damp4_freeze.q (frozen flow: after t = 0.1 the time
    step doubles and the velocity of time level n comes back after every stage), ChoCG/Sphere/sphere_point_src.q}:
    3 velocities + 1 scalar
    through the C++ host mirror and the device: scalar rows of the advection flux with the flow flux's normal
    velocities and stabilisation (damp2; damp4 with the limited reconstruction on the scalar's own nodal
    gradient), boundary integral, RK update, Dirichlet values of the rotating scalar field re-evaluated at every
    BC time (t + rk dt in the stages, t + dt after the projection), analytic-solution error norms of all
    four components, the point source pinned after every stage. Free-running against the oracle: same
    iteration count of every pressure solve, diagnostics and fields 1e-9 (CG bound of the ChoCG tests),
    then the reference's golden rows."""
    from xyst_b200 import hostapi as H
    from host_common import fixture_to_host_mesh
    kw = PCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    gold = O.load_golden_diag(case)
    n = int(gold[-1, 0]) + (1 if "point_src" in case else 0)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    assert s.get("u").shape == o.get("u").shape == (hm["coord"].shape[1], 4)
    assert rel(s.get("u"), o.get("u")) < 1e-9, "after the start-up projection"
    rows = []
    for it in range(n):
        r = s.step(1); o.step(1)
        if len(r):
            rows.append(r[0])
        # (slot_cyl: the rotation is divergence-free, the Poisson right-hand side is rounding noise of 1e-16 and
        # so is the iteration at which its residual falls below p_tol times its norm: 154 vs 155)
        assert "slot_cyl" in case or int(s.scalar("pit")) == int(o.scalar("pit")), it
        U, Uo = s.get("u"), o.get("u")
        assert rel(U[:, :3], Uo[:, :3]) < 1e-9, ("velocity", it)
        assert np.abs(U[:, 3] - Uo[:, 3]).max() < 1e-9 * max(np.abs(Uo[:, 3]).max(), 1e-3), ("scalar", it)
    rows = np.asarray(rows); ro = o.diag()
    assert rows.shape == ro.shape == gold.shape
    assert (np.abs(rows[:, :3] - ro[:, :3]) <= TOL * np.abs(ro[:, :3])).all()
    vs = np.abs(ro[:, 3:]).max(axis=1, keepdims=True)
    err = np.abs(rows - ro) / (np.abs(ro) + 1e-11 * vs)
    if "slot_cyl" in case:
        # Every node of this one-cell-thick mesh is a velocity Dirichlet node, so the velocity is prescribed
        # (it equals the oracle's to the last bit) and nothing depends on the pressure -- which the case solves
        # with p_tol = 1e-2 and a single pinned node. That pressure is ill-determined in the reference
        # itself: summing the oracle's dot products from the last node to the first (ORACLE_DOT_REVERSE=1,
        # test_oracle_cg_sensitivity.py) moves the norm of its pressure by up to 1.3e-3 (step 19), the increment norm
        # by 9e-2 and the iteration count of 12 of the 20 solves by one. Columns 3 and 8 (norms of p and of its
        # increment) get those bounds.
        rest = [c for c in range(rows.shape[1]) if c not in (3, 8)]
        assert (err[:, rest] <= 1e-9).all()
        assert (err[:, 3] <= 5e-3).all() and (err[:, 8] <= 0.3).all()
        assert rel(s.get("pr"), o.get("pr")) < 5e-3
        # the reference's own acceptance test of these goldens (SlotCyl/diag.ndiff.cfg)
        assert O.numdiff_ok(rows[:, 1:8], gold[:, 1:8], 1.0e-5, 2.0e-3).all()
        assert O.numdiff_ok(rows[:, 8:13], gold[:, 8:13], 3.0e-3, 1.0e-6).all()
        if case != "chocg_slot_cyl":         # (the damp2 golden's scalar columns differ from the reference's own objects)
            assert (np.abs(rows[:, rest] - gold[:, rest]) <= 2e-8 * np.abs(gold[:, rest]) + 1e-11 * vs).all()
    else:
        assert (err <= 1e-9).all()
        assert rel(s.get("pr"), o.get("pr")) < 1e-8
        assert (np.abs(rows - gold) <= 2e-8 * np.abs(gold) + 1e-11 * vs).all()
    print(case, "max rel diag diff vs oracle", err.max())


@pytest.mark.parametrize("case", ["lohcg_slot_cyl", "lohcg_slot_cyl_damp4"])
def test_lohcg_with_a_transported_scalar_matches_oracle_and_golden(case):
    """LohCG/SlotCyl/slot_cyl.q and slot_cyl_damp4.q: (p,u,v,w) + 1 scalar -- the scalar rows of lohner::grad,
    adv_damp2 / adv_damp4 (with LohCG's stab2 speed |vn| + s |n|), the boundary integral, the momentum source
    (lohner::src), time-dependent Dirichlet values (t + dt at start-up and after the projection, t + rk dt in
    the stages), error norms of all components. No linear solve in the time loop: fields and diagnostics are
    held to 1e-11 free-running; then the golden rows to their 12 printed digits."""
    kw = O.SCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    gold = O.load_golden_diag(case)
    n = int(gold[-1, 0])
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    assert s.get("u").shape == o.get("u").shape == (hm["coord"].shape[1], 5)
    # (the start-up pressure, u[:, 0], is the solution of the loosely converged, ill-determined Poisson problem of
    # the ChoCG slot_cyl test above; everything else does not depend on it: all nodes are velocity Dirichlet nodes)
    assert rel(s.get("u")[:, 1:], o.get("u")[:, 1:]) < 1e-11, "after the start-up projection"
    worst = 0.0; rows = []
    for it in range(n):
        r = s.step(1); o.step(1)
        if len(r):
            rows.append(r[0])
        U, Uo = s.get("u"), o.get("u")
        worst = max(worst, rel(U[:, 1:4], Uo[:, 1:4]), float(np.abs(U[:, 4] - Uo[:, 4]).max() / 0.6))
        assert worst < 1e-11, it
    print(case, "velocity / scalar max rel diff vs oracle over the run", worst, "pressure", rel(s.get("u")[:, 0], o.get("u")[:, 0]))
    rows = np.asarray(rows); ro = o.diag()
    assert rows.shape == ro.shape == gold.shape
    vs = np.abs(ro[:, 3:]).max(axis=1, keepdims=True)
    err = np.abs(rows - ro) / (np.abs(ro) + 1e-11 * vs)
    # columns: it t dt | L2 of p,u,v,w,s | L2 of their increments | L2 errors u,v,w,s | L1 errors
    rest = [c for c in range(rows.shape[1]) if c not in (3, 8)]
    print(case, "diag max rel diff vs oracle", err[:, rest].max(), "pressure columns", err[:, 3].max(), err[:, 8].max())
    assert (err[:, rest] <= 1e-10).all()
    assert (err[:, 3] <= 5e-3).all() and (err[:, 8] <= 0.3).all()
    assert (np.abs(rows[:, rest] - gold[:, rest]) <= 2e-8 * np.abs(gold[:, rest]) + 1e-11 * vs).all()


@pytest.mark.parametrize("base,kwx", [("chocg_ldc", dict(freezeflow=2.0, freezetime=0.0)),
                                      ("chocg_poiseuille_rk3", dict(freezeflow=1.5, freezetime=0.0)),
                                      ("chocg_poiseuille_rk4", dict(freezeflow=2.0, freezetime=0.05))])
def test_chocg_frozen_flow_matches_oracle(base, kwx):
    """tag::freezeflow on meshes WITH interior velocity nodes (on the SlotCyl mesh every node is a velocity
    Dirichlet node and the restored velocity equals the BC value): the velocity of time level n comes back
    after the BCs and the velocity gradient of every stage and, at the last stage, after the divergence
    (ChoCG::solve :1550-1570 in a serial run; oracle/chocg_port.hpp), dt is multiplied from the first step
    that starts after freezetime. No golden exists for these; oracle only."""
    kw = dict(O.CCASES[base], **kwx)
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o1 = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**O.CCASES[base]), "port")      # the same case, flow not frozen
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    rows = []
    for it in range(10):
        r = s.step(1); o.step(1); o1.step(1)
        rows.append(r[0])
        assert int(s.scalar("pit")) == int(o.scalar("pit")), it
        assert rel(s.get("u"), o.get("u")) < 1e-9 and rel(s.get("pr"), o.get("pr")) < 1e-8, it
    rows = np.asarray(rows); ro = o.diag()
    vs = np.abs(ro[:, 3:]).max(axis=1, keepdims=True)
    assert (np.abs(rows[:, :3] - ro[:, :3]) <= TOL * np.abs(ro[:, :3])).all()
    assert (np.abs(rows - ro) <= 1e-9 * np.abs(ro) + 1e-11 * vs).all()
    assert rel(o1.get("u"), o.get("u")) > 1e-6           # freezing does change this flow
    print(base, "frozen vs free-running flow differ by", rel(o1.get("u"), o.get("u")))


def test_kozcg_with_a_transported_scalar_and_frozen_flow_matches_oracle_and_golden():
    """KozCG/SlotCyl/slot_cyl.q: element-based Taylor-Galerkin + FCT of one transported scalar next to the flow
    (kozak::rhs scalar rows Kozak.cpp:84-86,133-135 with the element's half-step flow state; antidiffusive element
    contributions, allowed bounds and limit coefficients of the scalar as for a flow component), momentum source
    at nodes and centroids, time-dependent Dirichlet values, and freezeflow = 3: from the second step on
    (t > freezetime = 0) dt is tripled and only the scalar advances (KozCG::dt :669-674, solve :1140-1176).
    Every step against the oracle at 1e-12, then the golden rows."""
    case = "kozcg_slot_cyl"
    kw = O.SCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    gold = O.load_golden_diag(case)
    n = int(gold[-1, 0])
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    rows = []; worst = 0.0
    for it in range(n):
        r = s.step(1); o.step(1)
        rows.append(r[0])
        U, Uo = s.get("u"), o.get("u")
        worst = max(worst, rel(U[:, :5], Uo[:, :5]), float(np.abs(U[:, 5] - Uo[:, 5]).max() / 0.6))
        assert worst < TOL, it
    rows = np.asarray(rows); ro = o.diag()
    assert rows.shape == ro.shape == gold.shape
    assert rows[1, 2] > 2.9 * rows[0, 2]                 # dt tripled from the second step on
    sc = np.abs(ro).max(axis=0)
    # (+ 1e-15: norms of the z-momentum, which is zero up to rounding in this planar rotation)
    assert (np.abs(rows - ro) <= TOL * np.maximum(np.abs(ro), 1e-3 * sc) + 1e-15).all()
    assert (np.abs(rows - gold) <= 2e-11 * np.maximum(np.abs(gold), 1e-3 * sc) + 1e-15).all()
    print(case, "fields max rel diff vs oracle over the run", worst)


def test_zalcg_with_a_transported_scalar_source_and_frozen_flow_matches_oracle_and_golden():
    """ZalCG/SlotCyl/slot_cyl.q: Taylor-Galerkin edge flux + FCT of one transported scalar next to the flow
    (zalesak::rhs scalar rows Zalesak.cpp:113-116,146-149, boundary term, the three FCT node passes per scalar),
    the source term of zalesak::rhs (momentum source at the end nodes into the half step and at the edge
    midpoints to both nodes, :118-128,152-163, evaluated by the host mirror on the device's edge list),
    time-dependent Dirichlet values, and freezeflow = 3 from the second step on (ZalCG::dt :948-952, solve
    :1549,1577-1584). Every step against the oracle at 1e-12, then the golden rows."""
    case = "zalcg_slot_cyl"
    kw = O.SCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    gold = O.load_golden_diag(case)
    n = int(gold[-1, 0])
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    rows = []; worst = 0.0
    for it in range(n):
        r = s.step(1); o.step(1)
        rows.append(r[0])
        U, Uo = s.get("u"), o.get("u")
        worst = max(worst, rel(U[:, :5], Uo[:, :5]), float(np.abs(U[:, 5] - Uo[:, 5]).max() / 0.6))
        assert worst < TOL, it
    rows = np.asarray(rows); ro = o.diag()
    assert rows.shape == ro.shape == gold.shape
    assert rows[1, 2] > 2.9 * rows[0, 2]                 # dt tripled from the second step on
    sc = np.abs(ro).max(axis=0)
    assert (np.abs(rows - ro) <= TOL * np.maximum(np.abs(ro), 1e-3 * sc) + 1e-15).all()
    assert (np.abs(rows - gold) <= 2e-11 * np.maximum(np.abs(gold), 1e-3 * sc) + 1e-15).all()
    print(case, "fields max rel diff vs oracle over the run", worst)
