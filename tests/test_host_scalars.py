"""Host mirror on the transported-scalar cases of every solver (CPU only): the problem functions behind them
(problems::slot_cyl for the compressible solvers, ChoCG and LohCG unknowns; problems::point_src), the Dirichlet
mask lists with problem_ncomp + 1 entries per node and the configuration switches (freezeflow / freezetime),
against the oracle's chare after its setup. The device side of these cases is tests/test_gpu_scalars.py."""
import numpy as np
import pytest
import oraclelib as O
from xyst_b200 import hostapi as H
from host_common import fixture_to_host_mesh

CASES = {**O.SCASES, "chocg_sphere_point_src": O.SPHERE_SRC, "riecg_canyon": O.CANYON}


@pytest.mark.parametrize("case", list(CASES))
def test_scalar_case_host_setup_matches_oracle(case):
    kw = CASES[case]
    hm = fixture_to_host_mesh(O.load_mesh(kw["mesh"]))
    s = H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
    s.prepare(); s.host_setup()
    o = O.Oracle(O.load_mesh(kw["mesh"]), O.make_cfg(**kw), "port")
    nc = kw["ncomp"]
    u0, uo = s.get("u0"), o.get("u")
    assert u0.shape == uo.shape == (hm["coord"].shape[1], nc)
    solver = kw.get("solver", "riecg")
    nflow = {"chocg": 3, "lohcg": 4}.get(solver, 5)
    # the scalar columns are untouched by the oracle's start-up (projection, BCs re-impose the same IC values)
    assert np.array_equal(u0[:, nflow:], uo[:, nflow:])
    if solver not in ("chocg", "lohcg") and not kw.get("pre"):      # no start-up projection, no pressure BC: the whole initial field
        assert np.array_equal(u0, uo)
    if "slot_cyl" in case:
        assert 0.59 < u0[:, nflow].max() <= 0.6 and u0[:, nflow].min() == 0.0
    dm = s.get("dirbcmasks").reshape(-1, nc + 1); do = o.get("dirbcmasks").reshape(-1, nc + 1)
    assert np.array_equal(dm[np.argsort(dm[:, 0])], do[np.argsort(do[:, 0])])


def test_freezeflow_is_refused_where_it_is_not_implemented():
    kw = dict(O.CASES["riecg_sod"], freezeflow=2.0)
    with pytest.raises(Exception, match="freezeflow"):
        hm = fixture_to_host_mesh(O.load_mesh(kw.get("mesh", "riecg_sod")))
        H.Solver.mesh(H.make_cfg(**kw), hm["coord"], hm["tets"], hm["set_id"], hm["set_off"], hm["set_tri"])
