"""GPU parity of the LaxCG variant (time-derivative preconditioning, src/Physics/Lax.cpp,
src/Inciter/LaxCG.cpp) and of steady-state local time stepping through the C ABI, against the
oracle on the reference's LaxCG Bump regression case. Tolerance 1e-12 relative (fp64) for
kernels, 1e-10 for the 20-step diagnostics."""
import numpy as np
import pytest
import oraclelib as O
from gpu_common import context_from_oracle, drive_steps, relerr

pytestmark = pytest.mark.gpu
TOL = 1.0e-12


@pytest.mark.parametrize("case", list(O.LCASES))
def test_lax_grad_rhs_match_oracle(case):
    kw = O.LCASES[case]
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    ctx = context_from_oracle(o, kw, exact_muscl=True)
    dt = ctx.dt_min(kw["cfl"])
    o.kernel("mindt")
    assert abs(dt - o.scalar("dt")) <= TOL * dt
    ctx.grad(); ctx.rhs()
    o.kernel("lgrad"); o.kernel("lrhs", 0, 0.0)
    G = ctx.grad_get(); Go = o.get("grad")
    for c in range(5):
        assert relerr(G[:, 3*c:3*c+3], Go[:, 3*c:3*c+3]) < TOL, c
    R = ctx.rhs_get(); Ro = o.get("rhs")
    for c in range(5):
        assert relerr(R[:, c], Ro[:, c]) < 1e-11, c


@pytest.mark.parametrize("case", list(O.LCASES))
@pytest.mark.parametrize("exact", [True, False])
def test_laxcg_time_stepping_matches_oracle(case, exact):
    kw = O.LCASES[case]
    nsteps = 20
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    ctx = context_from_oracle(o, kw, exact_muscl=exact)
    t, dts = drive_steps([ctx], kw, nsteps)
    o.step(nsteps)
    d = o.diag()
    assert abs(t - o.scalar("t")) <= 1e-11 * abs(t)
    U = ctx.state_get(); Uo = o.get("u")
    for c in range(5):
        scale = max(np.abs(Uo[:, c]).max(), 1e-3 * np.abs(Uo).max())
        assert np.abs(U[:, c] - Uo[:, c]).max() <= 1e-10 * scale, c
    s = ctx.diag()
    meshvol = o.scalar("meshvol")
    l2 = np.sqrt(s[0:5] / meshvol); l2res = np.sqrt(s[5:10] / meshvol)
    assert np.abs(l2 - d[-1, 3:8]).max() <= 1e-11 * np.abs(d[-1, 3:8]).max()
    assert (np.abs(l2res - d[-1, 8:13]) <= 1e-8 * np.abs(d[-1, 8:13])).all()
    assert abs(s[10] - d[-1, 13]) <= TOL * abs(d[-1, 13])
    if case == "laxcg_bump":      # serial golden of the reference at its printed precision
        gold = O.load_golden_diag(case)
        assert np.abs(l2 - gold[-1, 3:8]).max() <= 1e-9 * np.abs(gold[-1, 3:8]).max()
        assert (np.abs(l2res - gold[-1, 8:13]) <= 1e-7 * np.abs(gold[-1, 8:13])).all()


def test_riecg_steady_local_time_stepping_matches_oracle():
    """steady = true with the RieCG solver (RieCG.cpp:812-825,1013) on the Sod mesh."""
    kw = dict(O.CASES["riecg_sod"], steady=True, mesh="riecg_sod")
    o = O.Oracle(O.load_mesh("riecg_sod"), O.make_cfg(**kw), "port")
    ctx = context_from_oracle(o, kw, exact_muscl=True)
    drive_steps([ctx], kw, 5)
    o.step(5)
    assert relerr(ctx.state_get(), o.get("u")) < 1e-11


@pytest.mark.parametrize("case", list(O.LCASES))
def test_laxcg_host_mirror_diag_rows(case):
    """Full drop-in path (C++ host mirror of LaxCG's setup + time loop, exact limiter divisions)
    vs oracle and, for the serial Rusanov case, the reference's golden file."""
    from xyst_b200 import hostapi as H
    from host_common import fixture_to_host_mesh
    kw = O.LCASES[case]
    gold = O.load_golden_diag(case)
    nsteps = int(gold[-1, 0])
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(exact_muscl=True, **kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    rows = s.step(nsteps)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o.step(nsteps); d = o.diag()
    assert rows.shape == d.shape == gold.shape
    for c in range(1, 8):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-10 * np.abs(d[:, c]).max(), c
    for c in range(8, 13):
        assert (np.abs(rows[:, c] - d[:, c]) <= 1e-8 * np.abs(d[:, c])).all(), c
    if case == "laxcg_bump":
        assert O.numdiff_ok(rows[:, 1:13], gold[:, 1:13], 1.0e-5, 1.0e-5).all()      # diag.ndiff.cfg
        assert (np.abs(rows[:, 3:8] - gold[:, 3:8]) <= 1e-9 * np.abs(gold[:, 3:8])).all()
