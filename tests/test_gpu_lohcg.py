"""GPU parity of the LohCG path (artificial-compressibility solver, unknowns p,u,v,w: Lohner edge
operators lohner::div/grad/vgrad/flux/rhs, explicit RK stages on all four unknowns, BCs incl. the
pressure Dirichlet BC, initial projection through the conjugate-gradient pressure solve) through the
drop-in path -- the C++ host mirror driving the device over the C ABI -- against the oracle's serial
restatement of LohCG.cpp and the reference's golden diagnostics
(tests/regression/inciter/LohCG/{Poiseuille,Lid}/diag_*.std)."""
import numpy as np
import pytest
import oraclelib as O
from gpu_common import relerr

pytestmark = pytest.mark.gpu

TOL = 1.0e-12          # north_star: fields and diagnostics within 1e-12 relative in fp64


def host_solver_for(case):
    from xyst_b200 import hostapi as H
    from host_common import fixture_to_host_mesh
    kw = O.HCASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    return s, kw


@pytest.mark.parametrize("case", list(O.HCASES))
def test_every_lohcg_step_matches_oracle_in_lockstep(case):
    """The time-step operator (dt reduction, rk x (gradients of all unknowns for damp4, lohner::rhs,
    update, dirbc/dirbcp/symbc/noslipbc), diagnostics) maps the SAME input state to the same output as
    the oracle within 1e-12, for every step along the oracle's trajectory."""
    s, kw = host_solver_for(case)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    # the start-up (two pressure solves) leaves the same state up to the CG iterate's rounding
    assert int(s.scalar("pit")) == int(o.scalar("pit"))
    U0, O0 = s.get("u"), o.get("u").reshape(-1, 4)
    assert relerr(U0[:, 1:], O0[:, 1:]) < 1e-9
    assert relerr(U0[:, 0], O0[:, 0]) < 1e-8
    for it in range(12):
        s.set_u(o.get("u"))
        row = s.step(1)
        o.step(1)
        U, Uo = s.get("u"), o.get("u").reshape(-1, 4)
        for c in range(4):
            scale = max(np.abs(Uo[:, c]).max(), 1e-3 * np.abs(Uo[:, 1:]).max())
            assert np.abs(U[:, c] - Uo[:, c]).max() <= TOL * scale, (it, c)
        if not len(row):                                 # no diagnostics this step (diag_iter)
            continue
        d = o.diag()[-1]
        assert row.shape[1] == len(d) and row[0, 0] == d[0]
        for c in range(0, 7):                            # it, t, dt, L2 norms of p,u,v,w
            assert abs(row[0, c] - d[c]) <= TOL * abs(d[c]) + 1e-300, (it, c)


@pytest.mark.parametrize("case", list(O.HCASES))
def test_host_mirror_lohcg_matches_oracle_and_golden(case):
    """Free-running: setup (stride-4 integrals, BC lists, Poisson matrix, initial projection) and the
    reference's number of steps; diagnostics against the oracle and the reference's golden rows."""
    s, kw = host_solver_for(case)
    gold = O.load_golden_diag(case)
    n = int(gold[-1, 0])
    rows = s.step(n)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o.step(n)
    ro = o.diag()
    assert rows.shape == ro.shape == gold.shape
    assert (np.abs(rows[:, :3] - ro[:, :3]) <= TOL * np.abs(ro[:, :3])).all()
    # columns: L2 norms of (p,u,v,w), then of their increments; a norm that is zero up to rounding
    # (e.g. w in the channel) is compared on the scale of the velocity norm
    vs = np.abs(ro[:, 4:7]).max(axis=1, keepdims=True)
    assert (np.abs(rows - ro) <= 1e-9 * np.abs(ro) + 1e-11 * vs).all()
    assert relerr(s.get("u"), o.get("u")) < 1e-9
    assert (np.abs(rows - gold) <= 2e-8 * np.abs(gold) + 1e-11 * vs).all()
    print(case, "max rel diag diff vs oracle", (np.abs(rows - ro) / np.maximum(np.abs(ro), 1e-300)).max())


def test_lohcg_and_chocg_entries_reject_the_other_mesh():
    s, kw = host_solver_for("lohcg_ldc")
    ctx = s.ctx()
    from xyst_b200 import capi
    with pytest.raises(capi.XystError, match="LohCG mesh"):
        ctx.chocg_rhs()
    with pytest.raises(capi.XystError, match="LohCG mesh"):
        ctx.chocg_stage(0, 1.0, 1e-3)
