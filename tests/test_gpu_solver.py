"""GPU tests of the full drop-in path: the C++ host mirror of the reference's RieCG solver
class (setup pipeline + time loop) driving the CUDA kernels through the C ABI, compared
with the oracle, with the reference's golden diagnostics, and -- at BASELINE.json's full
size -- through size-independent properties."""
import numpy as np
import pytest
import oraclelib as O
from xyst_b200 import hostapi as H
from host_common import fixture_to_host_mesh, host_mesh_to_oracle

pytestmark = pytest.mark.gpu
TOL = 1.0e-12            # north_star: norms/diagnostics within 1e-12 relative, fp64


ALLCASES = {**O.CASES, **O.VCASES, **O.TCASES, **O.PCASES}


def solver_for(case, **over):
    kw = dict(ALLCASES[case], **over)
    hm = fixture_to_host_mesh(O.load_mesh(kw.get("mesh", case)))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    return s, kw


@pytest.mark.parametrize("case", list(O.CASES))
def test_every_step_matches_oracle_in_lockstep(case):
    """The time-step operator (dt reduction, 3 x (grad, flux, gather+update, BC), diagnostics)
    maps the SAME input state to the same output as the oracle within 1e-12, for every step
    along the oracle's trajectory. (Free-running trajectories are compared below with looser
    bounds: MUSCL's first-order fallback, Riemann.cpp:129-130, is discontinuous, so rounding
    differences from the different but fixed summation order get amplified on shocks.)"""
    gold = O.load_golden_diag(case)
    nsteps = min(int(gold[-1, 0]), 12)
    s, kw = solver_for(case)
    o = O.Oracle(O.load_mesh(case), O.make_cfg(**kw), "port")
    for it in range(nsteps):
        s.set_u(o.get("u"))
        row = s.step(1)
        o.step(1)
        d = o.diag()
        U, Uo = s.get("u"), o.get("u")
        for c in range(5):
            scale = max(np.abs(Uo[:, c]).max(), 1e-3 * np.abs(Uo).max())
            assert np.abs(U[:, c] - Uo[:, c]).max() <= TOL * scale, (it, c)
        if len(row) and len(d) and d[-1, 0] == row[0, 0]:
            dd = d[-1]
            for c in list(range(2, 8)) + [13]:          # dt, L2 norms, total energy
                assert abs(row[0, c] - dd[c]) <= TOL * abs(dd[c]), (it, c)


@pytest.mark.parametrize("case", list(O.CASES))
def test_diag_rows_match_oracle_and_golden(case):
    gold = O.load_golden_diag(case)
    nsteps = int(gold[-1, 0])
    s, kw = solver_for(case)
    rows = s.step(nsteps)
    o = O.Oracle(O.load_mesh(case), O.make_cfg(**kw), "port")
    o.step(nsteps)
    d = o.diag()
    assert rows.shape == d.shape == gold.shape
    assert np.array_equal(rows[:, 0], d[:, 0])
    # free-running for the whole regression run (10 / 10 / 68 steps): the dominant norms
    # (density, energy, total energy) stay within 1e-11, everything else -- including the
    # tiny transverse momenta of the 1D Sod problem, relative to their own magnitude --
    # within 1e-8
    big = [3, 7, 13]
    for c in big:
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max(), c
    for c in range(1, d.shape[1]):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-8 * np.abs(d[:, c]).max(), c
    # and the reference's own acceptance test against its golden file
    assert O.numdiff_ok(rows[:, 1:8], gold[:, 1:8], 2.0e-4, 1.0e-5).all()
    assert O.numdiff_ok(rows[:, 8:13], gold[:, 8:13], 3.0e-4, 1.0e-7).all()
    assert (np.abs(rows - gold) / np.maximum(np.abs(gold), 1e-300)).max() < 1e-7


@pytest.mark.parametrize("case", list(O.VCASES))
def test_vortical_flow_matches_oracle_and_golden(case):
    """RieCG vortical flow (manufactured solution with a source term; Rusanov / HLLC, stab2, steady-state
    local time stepping with the residual stop test) through the host mirror: the whole regression run
    against the oracle's serial run (1e-12 on the dominant norms), and the serially recorded goldens to
    their printed digits (the two HLLC goldens were recorded on partitioned runs, see
    test_oracle_golden.py: for them the reference's own ndiff acceptance only)."""
    gold = O.load_golden_diag(case)
    nsteps = int(gold[-1, 0])
    s, kw = solver_for(case)
    rows = s.step(nsteps)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o.step(nsteps)
    d = o.diag()
    assert rows.shape == d.shape == gold.shape
    for c in (1, 2, 3, 7, 13):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max(), c
    for c in range(1, d.shape[1]):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-8 * np.abs(d[:, c]).max(), c
    U, Uo = s.get("u"), o.get("u")
    assert np.abs(U - Uo).max() <= 1e-10 * np.abs(Uo).max()
    if "hllc" not in case:
        assert (np.abs(rows[:, 1:8] - gold[:, 1:8]) <= 2e-8 * np.abs(gold[:, 1:8])).all()
    else:
        assert (np.abs(rows[:, 1:8] - gold[:, 1:8]) <= 1e-3 * np.abs(gold[:, 1:8])).all()


@pytest.mark.parametrize("case", list(O.TCASES))
def test_time_dependent_problems_match_oracle_and_golden(case):
    """Nonlinear energy growth and Rayleigh-Taylor: Dirichlet values refreshed per stage
    (xyst_dirbc_values at t + rk dt) and the source term per step (xyst_src_upload at t) by the host
    mirror, stage-wise device calls; whole regression run vs the oracle and the serial golden, incl.
    the L2/L1 errors against the time-dependent analytic solution."""
    gold = O.load_golden_diag(case)
    nsteps = int(gold[-1, 0])
    s, kw = solver_for(case)
    rows = s.step(nsteps)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o.step(nsteps)
    d = o.diag()
    assert rows.shape == d.shape == gold.shape
    assert (np.abs(rows - d) <= 1e-10 * np.abs(d) + 1e-13).all()
    for c in (1, 2, 3, 7, 13):
        assert np.abs(rows[:, c] - d[:, c]).max() <= TOL * np.abs(d[:, c]).max(), c
    U, Uo = s.get("u"), o.get("u")
    assert np.abs(U - Uo).max() <= 1e-11 * np.abs(Uo).max()
    assert (np.abs(rows - gold) <= 1e-8 * np.abs(gold) + 1e-15).all()


def test_pipe_with_pressure_bc_matches_oracle_and_golden():
    """Pipe flow: user-defined IC, symmetry walls, pressure BCs (physics::prebc) at inlet/outlet through
    the host mirror; 20 steps vs the oracle and the reference's 12-digit golden."""
    case = "riecg_pipe"
    gold = O.load_golden_diag(case)
    s, kw = solver_for(case)
    rows = s.step(20)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o.step(20)
    d = o.diag()
    assert rows.shape == d.shape == gold.shape
    for c in (1, 2, 3, 4, 7, 13):
        assert np.abs(rows[:, c] - d[:, c]).max() <= TOL * np.abs(d[:, c]).max(), c
    U, Uo = s.get("u"), o.get("u")
    for c in range(5):       # (momenta are ~1e-2 next to an energy of 2.5e5: same scale rule as the lock-step test)
        scale = max(np.abs(Uo[:, c]).max(), 1e-3 * np.abs(Uo).max())
        assert np.abs(U[:, c] - Uo[:, c]).max() <= 1e-11 * scale, c
    assert np.abs(U[:, 1] - Uo[:, 1]).max() <= 1e-9 * np.abs(Uo[:, 1]).max()
    for c in (1, 2, 3, 4, 7, 13):
        assert (np.abs(rows[:, c] - gold[:, c]) <= 5e-12 * np.abs(gold[:, c])).all(), c


@pytest.mark.parametrize("case", ["riecg_sod", "riecg_taylor_green"])
def test_reference_shaped_members_equal_fused_step(case):
    """dt/advance/grad/rhs/solve (materialised R, RieCG.cpp:871-1057) == fused stage kernels."""
    a, _ = solver_for(case); b, _ = solver_for(case)
    a.step(4, want_diag=False); b.step_unfused(4)
    assert np.array_equal(a.get("u"), b.get("u"))
    assert a.scalar("t") == b.scalar("t")


def test_exact_and_fast_limiter_agree():
    a, _ = solver_for("riecg_sedov", exact_muscl=True); b, _ = solver_for("riecg_sedov", exact_muscl=False)
    ra = a.step(10); rb = b.step(10)
    assert np.abs(ra[:, 3:8] - rb[:, 3:8]).max() <= TOL * np.abs(ra[:, 3:8]).max()


def test_box_vs_oracle_small():
    n = 8
    kw = dict(O.CASES["riecg_sedov"], sym=(1, 3, 5), p0=4.13e-2 / ((1.2 / n) ** 3 / 4))
    m = H.box_mesh(n, n, n, 1.2, 1.2, 1.2)
    o = O.Oracle(host_mesh_to_oracle(m), O.make_cfg(**kw), "port")
    s = H.Solver.box(H.make_cfg(**kw), n, n, n, 1.2, 1.2, 1.2)
    s.prepare(); s.attach(0); s.setup()
    rows = s.step(10); o.step(10); d = o.diag()
    for c in list(range(1, 8)) + [13]:
        assert np.abs(rows[:, c] - d[:, c]).max() <= TOL * np.abs(d[:, c]).max(), c
    U, Uo = s.get("u"), o.get("u")
    assert np.abs(U - Uo).max() <= 1e-11 * np.abs(Uo).max()


def _box_pair(n, reforder):
    import bench
    L = 1.2 * n / 150.0; h = L / n
    kw = dict(problem="sedov", gamma=bench.GAMMA, p0=4.13e-2 / (h ** 3 / 4.0), cfl=0.5, sym=(1, 3, 5))
    m = H.box_mesh(n, n, n, L, L, L)
    o = O.Oracle(host_mesh_to_oracle(m), O.make_cfg(**kw), "port")
    s = H.Solver.box(H.make_cfg(reforder=reforder, **kw), n, n, n, L, L, L)
    s.prepare(); s.attach(0); s.setup()
    return s, o


def _same_supedges(s, o):
    """The mirror reproduces the reference's grouping bitwise (single edges: as a set, their order is immaterial)."""
    for k in range(2):
        assert np.array_equal(s.get("dsupedge%d" % k), o.get("dsupedge%d" % k)), k
    e2 = lambda g: np.unique(np.asarray(g("dsupedge2")).reshape(-1, 2), axis=0)
    assert np.array_equal(e2(s.get), e2(o.get))


def _steps_match(s, o, steps):
    """Nodal state pointwise and all diagnostics columns at 1e-12. The reference accumulates its
    diagnostics sums serially over the nodes (NodeDiagnostics.cpp:85-118); beyond ~10^5 nodes that
    sum's own rounding reaches 1e-12 relative (3.6e-12 at 10^6 nodes, the same at every step), so
    the L2 columns of the LAST row are checked against the exactly summed (math.fsum) oracle state,
    and every row against the oracle's own serial sums at a tolerance that grows with the node count."""
    import math
    rows = s.step(steps); o.step(steps); d = o.diag()
    U, Uo = s.get("u"), o.get("u")
    for c in range(5):
        assert np.abs(U[:, c] - Uo[:, c]).max() <= TOL * np.abs(Uo[:, c]).max(), c
    v = o.get("v"); meshvol = o.scalar("meshvol")
    for c in range(5):
        exact = math.sqrt(math.fsum(Uo[:, c] ** 2 * v) / meshvol)
        assert abs(rows[-1, 3 + c] - exact) <= TOL * exact, c
    exact = math.fsum(Uo[:, 4] * v)
    assert abs(rows[-1, 13] - exact) <= TOL * abs(exact)
    tol_serial = max(TOL, 3.0e-17 * len(v))
    for c in list(range(1, 8)) + [13]:
        assert np.abs(rows[:, c] - d[:, c]).max() <= tol_serial * np.abs(d[:, c]).max(), c


@pytest.mark.parametrize("reforder", [1, 0])
def test_box_vs_oracle_n40(reforder):
    """The benchmark's code path against the oracle on a 40^3 box (384k tets), 10 steps, all diag
    columns 1e-12, U pointwise. reforder=1: triangle superedges in the reference's hash order,
    oracle untouched. reforder=0: the host mirror's element-order walk; the oracle then runs the
    reference's kernels on THESE superedges (orc_set_supedge): kernel against kernel, same inputs."""
    s, o = _box_pair(40, reforder)
    if reforder:
        _same_supedges(s, o)
    else:
        o.set_supedges(s.get)
    _steps_match(s, o, 10)


def test_box_vs_oracle_above_4m_tets():
    """n = 90: 4.37M tets, above the size at which round 1 silently switched the triangle walk.
    Reference order: same superedges as the oracle, 3 steps at 1e-12. Then the element-order
    superedges of a second solver are given to the SAME oracle and one gradient + right-hand side
    evaluation on the stepped state is compared (the oracle's set-up dominates the run time)."""
    n = 90
    s, o = _box_pair(n, 1)
    _same_supedges(s, o)
    _steps_match(s, o, 3)
    Uo = o.get("u")
    s.close()
    import bench
    L = 1.2 * n / 150.0
    kw = dict(problem="sedov", gamma=bench.GAMMA, p0=4.13e-2 / ((L / n) ** 3 / 4.0), cfl=0.5, sym=(1, 3, 5))
    s0 = H.Solver.box(H.make_cfg(reforder=0, **kw), n, n, n, L, L, L)
    s0.prepare(); s0.attach(0); s0.setup()
    o.set_supedges(s0.get)
    ctx = s0.ctx()
    ctx.state_set(Uo)
    ctx.grad(); ctx.rhs()
    o.kernel("grad"); o.kernel("rhs", 0, o.scalar("t"))
    G, R = ctx.grad_get(), ctx.rhs_get()
    Go, Ro = o.get("grad"), o.get("rhs")
    assert np.abs(G - Go).max() <= TOL * np.abs(Go).max()
    assert np.abs(R - Ro).max() <= TOL * np.abs(Ro).max()


def test_full_size_box_properties():
    """BASELINE.json configs[1]: 20.25M-tet box (n=150). Properties that need no oracle:
    conservation of total energy with closed (symmetry) boundaries, finiteness, positivity
    of the time step, run-to-run bit-reproducibility."""
    import bench
    n = 150; h = 1.2 / n
    cfg = bench.sedov_cfg(H.make_cfg, h)
    cfg.diag_iter = 1
    res = []
    for rep in range(2):
        s = H.Solver.box(cfg, n, n, n, 1.2, 1.2, 1.2)
        s.prepare(); s.attach(0); s.setup()
        assert s.scalar("nedge") == 23827950 and s.scalar("npoin") == 3442951
        rows = s.step(5 if rep == 0 else 2)
        res.append(rows)
        if rep == 0:
            assert np.isfinite(rows).all()
            mE = rows[:, 13]
            assert np.abs(mE - mE[0]).max() <= 1e-12 * abs(mE[0])          # energy conserved
            assert np.abs(rows[:, 3] - 1.0).max() < 1e-6                    # L2(rho) ~ 1
            assert (rows[:, 2] > 0).all()
    assert np.array_equal(res[0][:2], res[1][:2])                          # deterministic


@pytest.mark.parametrize("solver", ["zalcg", "kozcg"])
def test_full_size_fct_box_properties(solver):
    """BASELINE.json configs[2]: flux-corrected transport on the 50M-tet box (n=203, 50.19M tets,
    8.49M nodes). Properties that need no oracle: conservation of mass and total energy with
    closed (symmetry) boundaries -- the FCT anti-diffusive edge/element contributions cancel
    pairwise -- finiteness, a positive time step, density bounded below by its initial minimum
    (the limiter's purpose)."""
    import bench
    n = 203; h = 1.2 / n
    cfg = bench.sedov_cfg(H.make_cfg, h, solver=solver, fctsys=(1, 2, 3, 4, 5))
    cfg.diag_iter = 1
    s = H.Solver.box(cfg, n, n, n, 1.2, 1.2, 1.2)
    s.prepare(); s.attach(0); s.setup()
    assert s.scalar("ntet") == 6 * n ** 3 == 50192562 and s.scalar("npoin") == 8489664
    rows = s.step(3)
    assert np.isfinite(rows).all() and (rows[:, 2] > 0).all()
    mE = rows[:, 13]
    assert np.abs(mE - mE[0]).max() <= 1e-11 * abs(mE[0])
    assert np.abs(rows[:, 3] - 1.0).max() < 1e-5                           # L2(rho) ~ 1
    u = s.get("u")
    assert u[:, 0].min() > 0.0 and u[:, 4].min() > 0.0


def test_full_size_chocg_box_properties():
    """BASELINE.json configs[3]: ChoCG on the 50M-tet box -- lid-driven cavity set-up of the
    reference's regression case (tests/regression/inciter/ChoCG/Lid) on the 203^3 Kuhn box,
    pressure Poisson matrix with 8.49M rows. Properties: the assembled Laplacian has zero row sums
    (constants are in its null space) and is symmetric, the CG solve reduces the residual to its
    tolerance or uses all its iterations, the state stays finite, no-slip and lid values hold
    exactly after every step, and the run is bit-reproducible."""
    n = 203
    kw = dict(solver="chocg", ncomp=3, cfl=0.9, flux="damp4", mu=0.01, p_iter=40, p_tol=1.0e-3, p_pc="jacobi",
              p_hydrostat=0, problem="userdef", noslip=(1, 2, 3, 5, 6), dir_=((4, 2, 2, 2),),
              dirval=((4, 1.0, 0.0, 0.0),), nstep=2)
    res = []
    for rep in range(2):
        s = H.Solver.box(H.make_cfg(**kw), n, n, n)
        s.prepare(); s.host_setup()
        if rep == 0:
            ia = s.get("plhs_ia").astype(np.int64) - 1; a = s.get("plhs_a")
            assert len(ia) - 1 == 8489664
            rs = np.add.reduceat(a, ia[:-1])
            assert np.abs(rs).max() <= 1e-12 * np.abs(a).max()
        s.attach(0); s.setup()
        rows = s.step(2)
        assert rows.shape[0] == 2 and np.isfinite(rows).all()
        assert 1 <= int(s.scalar("pit")) <= 40
        u = s.get("u")
        x = s.get("x"); y = s.get("y"); z = s.get("z")
        wall = np.isclose(x, 0) | np.isclose(x, 1) | np.isclose(y, 0) | np.isclose(z, 0) | np.isclose(z, 1)
        lid = np.isclose(y, 1.0) & ~wall                     # no-slip is applied last: it wins on the lid's rim
        assert lid.sum() == (n - 1) ** 2
        assert np.array_equal(u[lid], np.tile([1.0, 0.0, 0.0], (lid.sum(), 1)))
        assert not u[wall].any()
        res.append(rows)
        s.close()
    assert np.array_equal(res[0], res[1])


def test_mesh_file_to_diag_file_end_to_end(tmp_path):
    """The reference's own workflow for a regression case, through the product alone: ExodusII mesh
    file in (tests/golden/riecg_sod.exo = RieCG/Sod/rectangle_01_1.5k.exo), control-file equivalent,
    10 steps on the GPU, `diag` text file out in the reference's format (src/IO/DiagWriter.cpp) --
    compared with the golden diag.std by the reference's own acceptance rule (diag.ndiff.cfg) and,
    tighter, to the digits the default precision prints."""
    import os
    kw = O.CASES["riecg_sod"]
    exo = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "riecg_sod.exo")
    s = H.Solver.exo(H.make_cfg(**kw), exo)
    s.prepare(); s.attach(0); s.setup()
    out = str(tmp_path / "diag")
    s.diag_file(out)
    s.step(10)
    s.close()
    lines = open(out).read().splitlines()
    gl = open(os.path.join(os.path.dirname(exo), "riecg_sod.diag.std")).read().splitlines()
    assert lines[0].split() == gl[0].split()                       # same header: "# 1:it 2:t 3:dt 4:L2(r) ..."
    rows = np.asarray([[float(x) for x in l.split()] for l in lines[1:]])
    gold = O.load_golden_diag("riecg_sod")
    assert rows.shape == gold.shape
    assert np.array_equal(rows[:, 0], gold[:, 0])
    assert O.numdiff_ok(rows[:, 1:8], gold[:, 1:8], 2.0e-4, 1.0e-5).all()
    assert (np.abs(rows[:, 1:] - gold[:, 1:]) <= 2e-8 * np.abs(gold[:, 1:]) + 1e-14).all()
