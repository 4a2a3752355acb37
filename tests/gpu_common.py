"""Helpers shared by the GPU parity tests: drive the C ABI (include/xyst_b200.h) with the
arrays the oracle's serial RieCG restatement built, the way the reference's solver class
would (RieCG::dt/advance/grad/rhs/solve, src/Inciter/RieCG.cpp:787-1057)."""
import numpy as np
import oraclelib as O
import xyst_b200

RK = (1.0 / 3.0, 0.5, 1.0)


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / (s if s > 0 else 1.0))


def tg_source(x, y):
    s = np.zeros((len(x), 5))
    s[:, 4] = 3.0 * np.pi / 8.0 * (np.cos(3.0 * np.pi * x) * np.cos(np.pi * y)
                                   - np.cos(3.0 * np.pi * y) * np.cos(np.pi * x))
    return s


def context_from_oracle(o, kw, chare=0, exact_muscl=False, device=0):
    """Upload one oracle chare's mesh/BC/state arrays into a device context."""
    ctx = xyst_b200.Context(device=device, flux=kw.get("flux", "rusanov"), gamma=kw["gamma"],
                            stab2=kw.get("stab2", False), stab2coef=kw.get("stab2coef", 0.2),
                            exact_muscl=exact_muscl)
    g = lambda n: o.get(n, chare)
    x, y, z = g("x"), g("y"), g("z")
    zal = kw.get("solver") == "zalcg"
    if kw.get("solver") == "kozcg":
        return _kozcg_context(ctx, o, kw, chare)
    ctx.mesh_upload(x, y, z, [g("dsupedge0"), g("dsupedge1"), g("dsupedge2")],
                    [g("dsupint0"), g("dsupint1"), g("dsupint2")], g("triinpoel"), g("besym"),
                    g("vol"), g("v"), stride=4 if zal else 3)
    if zal:
        ctx.zalcg_config(kw.get("fct", True), kw.get("fctclip", False), kw.get("fctsys", ()), kw.get("fctdif", 1.0))
    if kw.get("solver") == "laxcg":
        ctx.laxcg_config(kw.get("rgas", 287.052874), kw.get("turkel", 0.5), kw.get("velinf", (1.0, 1.0, 1.0)))
    if kw.get("steady"):
        ctx.steady(True)
    U0 = g("u")
    dm = g("dirbcmasks")
    dv = U0[dm.reshape(-1, 6)[:, 0].astype(np.int64)] if len(dm) else None
    ctx.bc_upload(dirbcmasks=dm, dirvals=dv, symbcnodes=g("symbcnodes"), symbcnorms=g("symbcnorms"),
                  farbcnodes=g("farbcnodes"), farbcnorms=g("farbcnorms"),
                  far=(kw.get("far_density", 0.0), kw.get("far_pressure", 0.0), kw.get("far_velocity", (0.0, 0.0, 0.0))))
    if kw["problem"] == "taylor_green":
        ctx.src_upload(tg_source(x, y))
    ctx.state_set(U0)
    return ctx


def _kozcg_context(ctx, o, kw, chare):
    g = lambda n: o.get(n, chare)
    x, y, z = g("x"), g("y"), g("z")
    inpoel = g("inpoel").reshape(-1, 4).astype(np.int64)
    Sn = Sc = None
    if kw["problem"] == "taylor_green":
        Sn = tg_source(x, y)
        Sc = tg_source(x[inpoel].sum(axis=1) / 4.0, y[inpoel].sum(axis=1) / 4.0)
    ctx.kozcg_mesh_upload(x, y, z, inpoel, g("vol"), g("v"), Sn, Sc)
    ctx.zalcg_config(kw.get("fct", True), kw.get("fctclip", False), kw.get("fctsys", ()), kw.get("fctdif", 1.0))
    U0 = g("u")
    dm = g("dirbcmasks")
    dv = U0[dm.reshape(-1, 6)[:, 0].astype(np.int64)] if len(dm) else None
    ctx.bc_upload(dirbcmasks=dm, dirvals=dv, symbcnodes=g("symbcnodes"), symbcnorms=g("symbcnorms"))
    ctx.state_set(U0)
    return ctx


def analytic_prim(kw, o, chare=0):
    """Analytic solution in primitive form for the diagnostics (time-independent ICs)."""
    if kw["problem"] != "taylor_green":
        return None
    # the oracle's initial state IS the analytic solution for taylor_green
    return None


def drive_steps(ctxs, kw, nsteps, t0=0.0, fused=True, allreduce_min=None):
    """RieCG time loop over the C ABI (single or several partitions in one process,
    no halo): returns (t, list of dt)."""
    t = t0; dts = []
    for _ in range(nsteps):
        if abs(kw.get("dt", 0.0)) > np.finfo(float).eps:
            dt = kw["dt"]
        else:
            dt = min(c.dt_min(kw["cfl"]) for c in ctxs)
        if t + dt > kw.get("term", 1e300):
            dt = kw["term"] - t
        for c in ctxs:
            if kw.get("solver") == "zalcg":
                c.zalcg_step(dt)
            elif kw.get("solver") == "kozcg":
                c.kozcg_step(dt)
            elif fused:
                c.step(dt)
            else:
                for s in range(3):
                    c.grad(); c.rhs(); c.rk_update(s, dt); c.apply_bc()
        t += dt; dts.append(dt)
    return t, dts
