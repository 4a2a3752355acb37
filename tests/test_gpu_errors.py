"""Error behaviour at the drop-in boundary (the reference throws tk::Exception from Assert/ErrChk/
Throw, src/Base/Exception.hpp:40-52; the C ABI returns non-zero and keeps the message in
xyst_last_error()): misuse must fail loudly, never compute on garbage."""
import numpy as np
import pytest
import oraclelib as O
import xyst_b200
from gpu_common import context_from_oracle

pytestmark = pytest.mark.gpu


def test_compute_before_upload_fails():
    ctx = xyst_b200.Context()
    for call in (ctx.grad, ctx.rhs, ctx.apply_bc, lambda: ctx.dt_min(0.5), lambda: ctx.step(1e-3)):
        with pytest.raises(xyst_b200.XystError, match="no mesh uploaded"):
            call()
    with pytest.raises(xyst_b200.XystError, match="no matrix uploaded"):
        ctx.cg_solve(10, 1e-3)
    with pytest.raises(xyst_b200.XystError, match="no mesh uploaded"):
        ctx.npoin = 1; ctx.chocg_get("u")


def test_bad_arguments_fail():
    with pytest.raises(xyst_b200.XystError, match="ncomp"):
        xyst_b200.Context(ncomp=4)
    with pytest.raises(xyst_b200.XystError, match="invalid device"):
        xyst_b200.Context(device=99)
    kw = O.CASES["riecg_sod"]
    o = O.Oracle(O.load_mesh("riecg_sod"), O.make_cfg(**kw), "port")
    g = o.get
    ctx = xyst_b200.Context(gamma=kw["gamma"])
    se = [g("dsupedge0").copy(), g("dsupedge1"), g("dsupedge2")]
    se[0][3] = len(g("x")) + 7                                     # node id beyond npoin
    with pytest.raises(xyst_b200.XystError, match="node id out of range"):
        ctx.mesh_upload(g("x"), g("y"), g("z"), se, [g("dsupint0"), g("dsupint1"), g("dsupint2")],
                        g("triinpoel"), g("besym"), g("vol"), g("v"))
    ctx = context_from_oracle(o, kw)
    with pytest.raises(xyst_b200.XystError, match="stage"):
        ctx.stage(3, 1e-3)
    with pytest.raises(xyst_b200.XystError, match="stride-4"):
        ctx.zalcg_step(1e-3)                                       # RieCG upload, ZalCG call
    with pytest.raises(xyst_b200.XystError, match="xyst_chocg_mesh_upload"):
        ctx.chocg_rhs()                                            # RieCG upload, ChoCG call
    with pytest.raises(xyst_b200.XystError, match="out of range"):
        ctx.bc_upload(symbcnodes=[10 ** 9], symbcnorms=[1.0, 0.0, 0.0])


def test_host_mirror_rejects_unknown_configuration():
    from xyst_b200 import hostapi as H
    m = H.box_mesh(2, 2, 2)
    for bad, msg in ((dict(problem="sod", solver="nosuch"), "Unknown solver"),
                     (dict(problem="nosuch", sym=(1,)), "not hooked up"),
                     (dict(problem="sod", flux="nosuch", sym=(1,)), "Flux not configured")):
        with pytest.raises(xyst_b200.XystError, match=msg):
            s = H.Solver.mesh(H.make_cfg(gamma=1.4, cfl=0.5, **bad), m["coord"], m["tets"], m["set_id"],
                              m["set_off"], m["set_tri"])
            s.prepare(); s.attach(0); s.setup()
