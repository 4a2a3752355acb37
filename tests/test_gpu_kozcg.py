"""GPU parity of the KozCG variant (element-based Taylor-Galerkin + flux-corrected transport,
src/Physics/Kozak.cpp:29-180, src/Inciter/KozCG.cpp:691-1197) through the C ABI against the oracle
on the reference's KozCG regression cases. Tolerance 1e-12 relative (fp64)."""
import numpy as np
import pytest
import oraclelib as O
from gpu_common import context_from_oracle, drive_steps, relerr

pytestmark = pytest.mark.gpu
TOL = 1.0e-12


@pytest.mark.parametrize("case", list(O.KCASES))
def test_kozak_rhs_matches_oracle(case):
    kw = O.KCASES[case]
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    ctx = context_from_oracle(o, kw)
    dt = ctx.dt_min(kw["cfl"])
    ctx.kozcg_rhs(dt)
    o.kernel("krhs", 0, 0.0, dt)
    assert relerr(ctx.rhs_get(), o.get("rhs")) < TOL


@pytest.mark.parametrize("case", list(O.KCASES))
def test_kozcg_time_stepping_matches_oracle(case):
    kw = O.KCASES[case]
    gold = O.load_golden_diag(case)
    nsteps = int(gold[-1, 0])
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    ctx = context_from_oracle(o, kw)
    t, dts = drive_steps([ctx], kw, nsteps)
    o.step(nsteps)
    d = o.diag()
    assert abs(t - o.scalar("t")) <= TOL * abs(t)
    U = ctx.state_get(); Uo = o.get("u")
    for c in range(5):
        scale = max(np.abs(Uo[:, c]).max(), 1e-3 * np.abs(Uo).max())
        assert np.abs(U[:, c] - Uo[:, c]).max() <= 1e-11 * scale, c
    s = ctx.diag()
    meshvol = o.scalar("meshvol")
    l2 = np.sqrt(s[0:5] / meshvol)
    assert np.abs(l2 - d[-1, 3:8]).max() <= TOL * np.abs(d[-1, 3:8]).max()
    assert abs(s[10] - d[-1, 13]) <= TOL * abs(d[-1, 13])
    # and against the reference's golden file at its printed precision
    assert np.abs(l2 - gold[-1, 3:8]).max() <= 1e-8 * np.abs(gold[-1, 3:8]).max()


def test_kozcg_fct_variants():
    """fct without clipping and without system limiting; larger diffusion coefficient."""
    for extra in (dict(fctclip=False, fctsys=()), dict(fctdif=0.5, fctsys=(1, 2, 3, 4, 5))):
        kw = dict(O.KCASES["kozcg_sod"], **extra)
        o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
        ctx = context_from_oracle(o, kw)
        drive_steps([ctx], kw, 5)
        o.step(5)
        assert relerr(ctx.state_get(), o.get("u")) < 1e-11


@pytest.mark.parametrize("case", list(O.KCASES) + list(O.KTCASES))
def test_kozcg_host_mirror_diag_rows(case):
    """Full drop-in path (C++ host mirror of KozCG's setup + time loop) vs oracle and golden; the
    time-dependent problems refresh nodal sources (t), centroid sources (t + dt/2) and Dirichlet
    values (t + dt) every step."""
    from xyst_b200 import hostapi as H
    from host_common import fixture_to_host_mesh
    kw = {**O.KCASES, **O.KTCASES}[case]
    gold = O.load_golden_diag(case)
    nsteps = int(gold[-1, 0])
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.attach(0); s.setup()
    rows = s.step(nsteps)
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    o.step(nsteps); d = o.diag()
    assert rows.shape == d.shape == gold.shape
    for c in range(1, 8):
        assert np.abs(rows[:, c] - d[:, c]).max() <= 1e-11 * np.abs(d[:, c]).max(), c
    assert O.numdiff_ok(rows[:, 1:8], gold[:, 1:8], 1.0e-8, 1.0e-7).all()     # reference's own tolerance
    assert O.numdiff_ok(rows[:, 8:13], gold[:, 8:13], 1.0e-8, 1.0e-6).all()
    if rows.shape[1] > 14:      # L2 and L1 errors against the analytic solution (manufactured problems)
        assert (np.abs(rows[:, 14:] - d[:, 14:]) <= 1e-9 * np.abs(d[:, 14:]) + 1e-14).all()
