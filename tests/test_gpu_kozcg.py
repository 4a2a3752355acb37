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
