"""Helpers shared by the GPU parity tests: drive the C ABI (include/xyst_b200.h) with the
arrays the oracle's serial RieCG restatement built, the way the reference's solver class
would (RieCG::dt/advance/grad/rhs/solve, src/Inciter/RieCG.cpp:787-1057)."""
import numpy as np
import oraclelib as O
import xyst_b200

RK = (1.0 / 3.0, 0.5, 1.0)


def relerr(a, b):
    a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / (s if s > 0 else 1.0))


def tg_source(x, y):
    s = np.zeros((len(x), 5))
    s[:, 4] = 3.0 * np.pi / 8.0 * (np.cos(3.0 * np.pi * x) * np.cos(np.pi * y)
                                   - np.cos(3.0 * np.pi * y) * np.cos(np.pi * x))
    return s


def source(kw, x, y, z):
    """Nodal values of problems::SRC() for the problems that have one (time-independent ones)."""
    if kw["problem"] == "taylor_green":
        return tg_source(x, y)
    if kw["problem"] == "vortical_flow":                 # Problems.cpp:480-507
        a, k, g = kw["alpha"], kw["kappa"], kw["gamma"]
        ru = a * x - k * y; rv = k * x + a * y
        s = np.zeros((len(x), 5))
        s[:, 1] = a * ru - k * rv
        s[:, 2] = k * ru + a * rv
        s[:, 4] = s[:, 1] * ru + s[:, 2] * rv + 8.0 * a * a * a * z * z / (g - 1.0)
        return s
    return None


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
    S = source(kw, x, y, z)
    if S is not None:
        ctx.src_upload(S)
    ctx.state_set(U0)
    return ctx


def _kozcg_context(ctx, o, kw, chare):
    g = lambda n: o.get(n, chare)
    x, y, z = g("x"), g("y"), g("z")
    inpoel = g("inpoel").reshape(-1, 4).astype(np.int64)
    Sn = source(kw, x, y, z)
    Sc = None if Sn is None else source(kw, x[inpoel].sum(axis=1) / 4.0, y[inpoel].sum(axis=1) / 4.0,
                                        z[inpoel].sum(axis=1) / 4.0)
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


# ---- ChoCG ------------------------------------------------------------------------------------
CHO_RK = {1: (1.0,), 2: (0.5, 1.0), 3: (1.0 / 3.0, 0.5, 1.0), 4: (0.25, 1.0 / 3.0, 0.5, 1.0)}   # ChoCG.cpp:43-48


class ChoDriver:
    """The ChoCG solver's control flow (src/Inciter/ChoCG.cpp + chocg.ci, one partition) over the
    C ABI: every nodal/edge loop runs on the device, this class only sequences the calls the
    way the chare's SDAG code does and evaluates nothing itself. Mesh-derived arrays, BC node
    lists and the values of the problem functions come from the oracle's setup."""

    def __init__(self, o, kw, device=0):
        g = lambda n: o.get(n, 0)
        self.kw = kw
        self.rk = CHO_RK[kw.get("rk", 1)]
        self.ctx = ctx = xyst_b200.Context(device=device)
        self.damp4 = kw.get("flux", "damp2") == "damp4"
        ctx.chocg_mesh_upload(g("x"), g("y"), g("z"), [g("dsupedge0"), g("dsupedge1"), g("dsupedge2")],
                              [g("dsupint0"), g("dsupint1"), g("dsupint2")], g("triinpoel"), g("vol"), g("v"),
                              flux=kw.get("flux", "damp2"), stab=kw.get("stab", True), stab2=kw.get("stab2", False),
                              stab2coef=kw.get("stab2coef", 0.2), mu=kw.get("mu", 0.0))
        u0 = g("u0").reshape(-1, 3)
        dm = g("dirbcmasks").reshape(-1, 4).astype(np.int64)
        dv = g("dirbcval").reshape(-1, 4)
        val = np.zeros((len(dm), 3))
        for i in range(len(dm)):                     # physics::dirbc: mask 1 = IC value, 2 = given value
            for c in range(3):
                if dm[i, 1 + c] == 1:
                    val[i, c] = u0[dm[i, 0], c]
                elif dm[i, 1 + c] == 2 and len(dv):
                    val[i, c] = dv[i, 1 + c]
        mask = dm[:, 1:].copy()
        if not len(dv):
            mask[mask == 2] = 0
        ctx.chocg_bc_upload(dm[:, 0], mask, val, g("symbcnodes"), g("symbcnorms"), g("noslipbcnodes"))
        ctx.csr_upload(g("plhs_ia"), g("plhs_ja"), g("plhs_a"), 1)
        # pressure BCs of ChoCG::pinit :1047-1115
        pm = g("dirbcmaskp").reshape(-1, 2).astype(np.int64)
        pv = g("dirbcvalp").reshape(-1, 2)
        pic = g("p_ic")
        self.pbc = {}
        for i in range(len(pm)):
            if pm[i, 1] == 1:
                self.pbc[int(pm[i, 0])] = pic[pm[i, 0]]
            elif pm[i, 1] == 2 and len(pv):
                self.pbc[int(pm[i, 0])] = pv[i, 1]
        h = g("hydrostat")
        if len(h) and int(h[0]) not in self.pbc:
            self.pbc[int(h[0])] = pic[int(h[0])]
        self.neubc = g("neubc") if len(g("neubc")) else None
        self.prhs = g("p_rhs") if len(g("p_rhs")) else None
        self.psol = g("p_sol") if len(g("p_sol")) else None
        self.meshvol = o.scalar("meshvol")
        self.t = 0.0; self.dt = kw.get("dt", 0.0); self.it = 0
        self.np = 0; self.initial = True; self.finished = False
        self.pit = 0; self.rows = []
        ctx.chocg_set_u(u0)
        # ChoCG::merge onwards: initial projection and pressure
        self.div_u()
        self.pinit(); self.psolve()
        self.sgrad(); self.psolved()

    def div_u(self):
        self.ctx.chocg_div(0, self.dt, self.np > 1)

    def pinit(self):
        nodes = np.asarray(sorted(self.pbc), np.uint64)
        vals = np.asarray([0.0 if self.np > 1 else self.pbc[int(n)] for n in nodes])
        self.ctx.chocg_pinit(self.dt if self.np > 1 else 1.0, nodes, vals, self.neubc, self.prhs,
                             self.kw.get("p_pc", "none"))

    def psolve(self):
        _, self.pit = self.ctx.cg_solve(self.kw["p_iter"], self.kw["p_tol"])

    def sgrad(self):
        self.ctx.chocg_grad(0)

    def pgrad(self):
        self.ctx.chocg_grad(1)

    def psolved(self):
        c = self.ctx
        if self.np != 1:
            c.chocg_project(self.dt if self.np > 1 else 1.0)
        if self.initial:
            if self.kw.get("nstep") == 1:
                c.chocg_pressure_update(False)
                self.diag(); self.finished = True
            else:
                self.np += 1
                if self.np < 2:
                    c.chocg_vgrad(); c.chocg_flux(); c.chocg_div(1, self.dt, self.np > 1)
                    self.pinit(); self.psolve()
                    self.psolved()
                else:
                    c.chocg_pressure_update(False)
                    self.pgrad()
                    self.initial = False
        else:
            c.chocg_pressure_update(True)
            self.pgrad()
            self.diag()

    def step(self):
        if self.finished:
            return False
        kw = self.kw; c = self.ctx
        eps = np.finfo(float).eps
        mindt = kw["dt"] if abs(kw.get("dt", 0.0)) > eps else c.chocg_dt_min(kw["cfl"], kw.get("dif", 0.0))
        if mindt < eps:
            self.finished = True
        self.dt = mindt
        if self.t + self.dt > kw.get("term", 1e300):
            self.dt = kw["term"] - self.t
        for s, rk in enumerate(self.rk):
            c.chocg_stage(s, rk, self.dt)
        self.div_u()
        self.pinit(); self.psolve()
        self.sgrad(); self.psolved()
        if self.it >= kw.get("nstep", 1 << 62) or abs(self.t - kw.get("term", 1e300)) < eps:
            self.finished = True
        return not self.finished

    def diag(self):
        self.it += 1; self.t += self.dt
        if (self.it + 1) % self.kw.get("diag_iter", 1):
            return
        d = self.ctx.chocg_diag(self.psol, None)
        ncomp = 0 if self.psol is not None else 3
        row = [float(self.it), self.t, self.dt]
        row += [np.sqrt(d[i] / self.meshvol) for i in range(ncomp + 1)]
        row += [np.sqrt(d[4 + i] / self.meshvol) for i in range(ncomp + 1)]
        if self.psol is not None:
            row += [np.sqrt(d[8] / self.meshvol), d[9] / self.meshvol]
        self.rows.append(row)
