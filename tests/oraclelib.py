"""ctypes front end of the CPU oracle (oracle/): TEST INFRASTRUCTURE ONLY.

Loads oracle/_build/liboracle.so (the self-contained restatement, "port") or
oracle/_ref/liboracle_ref.so (the same serial driver running the reference's own
unmodified Physics/Mesh objects, "reference"). Builds them on demand with
oracle/Makefile; the "reference" flavour can only be (re)built where
/root/reference exists, elsewhere the prebuilt file is used if present.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Cfg(C.Structure):
    _fields_ = [
        ("problem", C.c_char * 32), ("flux", C.c_char * 16),
        ("ncomp", C.c_int32), ("stab2", C.c_int32), ("steady", C.c_int32),
        ("nsym", C.c_int32), ("sym", C.c_int32 * 16),
        ("ndir", C.c_int32), ("dir", (C.c_int32 * 12) * 16),
        ("nfar", C.c_int32), ("far_sets", C.c_int32 * 16),
        ("npre", C.c_int32), ("pre_sets", C.c_int32 * 16),
        ("nfieldout", C.c_int32), ("fieldout_sets", C.c_int32 * 16),
        ("solver", C.c_char * 16), ("fct", C.c_int32), ("fctclip", C.c_int32), ("nfctsys", C.c_int32),
        ("fctsys", C.c_int32 * 8), ("fctdif", C.c_double),
        ("nstep", C.c_uint64), ("diag_iter", C.c_uint64),
        ("gamma", C.c_double), ("p0", C.c_double), ("cfl", C.c_double), ("dt", C.c_double),
        ("t0", C.c_double), ("term", C.c_double), ("stab2coef", C.c_double),
        ("far_density", C.c_double), ("far_pressure", C.c_double), ("far_velocity", C.c_double * 3),
        ("pre_density", C.c_double * 16), ("pre_pressure", C.c_double * 16),
        ("rgas", C.c_double), ("turkel", C.c_double), ("velinf", C.c_double * 3), ("residual", C.c_double),
        ("rescomp", C.c_uint64),
        ("ic_density", C.c_double), ("ic_pressure", C.c_double), ("ic_velocity", C.c_double * 3),
        ("mu", C.c_double), ("dif", C.c_double), ("stab", C.c_int32), ("rk", C.c_uint64),
        ("nnoslip", C.c_int32), ("noslip", C.c_int32 * 16),
        ("ndirval", C.c_int32), ("dirval", (C.c_double * 12) * 16),
        ("p_iter", C.c_uint64), ("p_tol", C.c_double), ("p_pc", C.c_char * 16),
        ("np_dir", C.c_int32), ("p_dir", (C.c_int32 * 2) * 16),
        ("np_dirval", C.c_int32), ("p_dirval", (C.c_double * 2) * 16),
        ("np_sym", C.c_int32), ("p_sym", C.c_int32 * 16),
        ("p_hydrostat_set", C.c_int32), ("p_hydrostat", C.c_uint64),
        ("alpha", C.c_double), ("kappa", C.c_double),
        ("r0", C.c_double), ("ce", C.c_double), ("beta", C.c_double * 3),
        ("soundspeed", C.c_double),
        ("src_location", C.c_double * 3), ("src_radius", C.c_double), ("src_release_time", C.c_double),
        ("freezeflow", C.c_double), ("freezetime", C.c_double),
        ("theta", C.c_double), ("mom_iter", C.c_uint64), ("mom_tol", C.c_double), ("mom_pc", C.c_char * 16),
        ("fctfreeze", C.c_double),
    ]


def make_cfg(problem, mesh=None, flux="rusanov", gamma=1.4, p0=0.0, cfl=0.0, dt=0.0, t0=0.0, term=1e300,
             nstep=2**63, sym=(), dir_=(), stab2=False, stab2coef=0.2, diag_iter=1, ncomp=5,
             fieldout=(), solver="riecg", fct=True, fctclip=False, fctsys=(), fctdif=1.0,
             steady=False, residual=0.0, rescomp=1, rgas=287.052874, turkel=0.5, velinf=(1.0, 1.0, 1.0),
             far=(), far_density=0.0, far_pressure=0.0, far_velocity=(0.0, 0.0, 0.0),
             ic_density=0.0, ic_pressure=0.0, ic_velocity=(0.0, 0.0, 0.0),
             mu=0.0, dif=0.0, stab=True, rk=1, noslip=(), dirval=(), p_iter=10, p_tol=1.0e-3, p_pc="none",
             p_dir=(), p_dirval=(), p_sym=(), p_hydrostat=None, alpha=0.0, kappa=0.0, r0=0.0, ce=0.0, beta=(0.0, 0.0, 0.0), pre=(), soundspeed=1.0,
             theta=0.0, mom_iter=10, mom_tol=1.0e-3, mom_pc="none", freezeflow=1.0, freezetime=0.0, fctfreeze=0.0,
             src_location=(0.0, 0.0, 0.0), src_radius=0.0, src_release_time=0.0, cls=Cfg):
    """Control-file equivalent; defaults are the reference's (InciterConfig.cpp:1707-1757)."""
    c = cls()
    c.problem = problem.encode(); c.flux = flux.encode(); c.ncomp = ncomp
    c.gamma = gamma; c.p0 = p0; c.cfl = cfl; c.dt = dt; c.t0 = t0; c.term = term; c.alpha = alpha; c.kappa = kappa
    c.r0 = r0; c.ce = ce; c.soundspeed = soundspeed
    c.freezeflow = freezeflow; c.freezetime = freezetime; c.fctfreeze = fctfreeze
    c.src_radius = src_radius; c.src_release_time = src_release_time
    for i in range(3):
        c.src_location[i] = src_location[i]
    c.theta = theta; c.mom_iter = mom_iter; c.mom_tol = mom_tol; c.mom_pc = mom_pc.encode()
    for i in range(3):
        c.beta[i] = beta[i]
    c.npre = len(pre)
    for i, (sid, dens, pres) in enumerate(pre):          # bc_pre = { name = { sideset = { sid }, density, pressure } }
        c.pre_sets[i] = sid; c.pre_density[i] = dens; c.pre_pressure[i] = pres
    c.nstep = nstep; c.stab2 = int(stab2); c.stab2coef = stab2coef; c.steady = int(steady)
    c.diag_iter = diag_iter
    c.residual = residual; c.rescomp = rescomp; c.rgas = rgas; c.turkel = turkel
    c.nfar = len(far); c.far_density = far_density; c.far_pressure = far_pressure
    for i, s_ in enumerate(far):
        c.far_sets[i] = s_
    c.ic_density = ic_density; c.ic_pressure = ic_pressure
    c.mu = mu; c.dif = dif; c.stab = int(stab); c.rk = rk
    c.nnoslip = len(noslip)
    for i, s_ in enumerate(noslip):
        c.noslip[i] = s_
    c.ndirval = len(dirval)
    for i, m in enumerate(dirval):
        for j, v in enumerate(m):
            c.dirval[i][j] = v
    c.p_iter = p_iter; c.p_tol = p_tol; c.p_pc = p_pc.encode()
    c.np_dir = len(p_dir)
    for i, m in enumerate(p_dir):
        c.p_dir[i][0], c.p_dir[i][1] = m
    c.np_dirval = len(p_dirval)
    for i, m in enumerate(p_dirval):
        c.p_dirval[i][0], c.p_dirval[i][1] = m
    c.np_sym = len(p_sym)
    for i, s_ in enumerate(p_sym):
        c.p_sym[i] = s_
    c.p_hydrostat_set = int(p_hydrostat is not None); c.p_hydrostat = p_hydrostat or 0
    for i in range(3):
        c.velinf[i] = velinf[i]; c.far_velocity[i] = far_velocity[i]; c.ic_velocity[i] = ic_velocity[i]
    c.solver = solver.encode(); c.fct = int(fct); c.fctclip = int(fctclip); c.fctdif = fctdif
    c.nfctsys = len(fctsys)
    for i, s_ in enumerate(fctsys):
        c.fctsys[i] = s_
    c.nsym = len(sym)
    for i, s in enumerate(sym):
        c.sym[i] = s
    c.nfieldout = len(fieldout)
    for i, s in enumerate(fieldout):
        c.fieldout_sets[i] = s
    c.ndir = len(dir_)
    for i, m in enumerate(dir_):
        for j, v in enumerate(m):
            c.dir[i][j] = v
    return c


# LaxCG regression case (tests/regression/inciter/LaxCG/Bump/{bump.q,bump_hllc.q}): steady-state
# local time stepping, free stream rho=1.225, p=101325, Mach 0.675
_C_INF = (1.4 * 101325.0 / 1.225) ** 0.5
_BUMP = dict(solver="laxcg", problem="userdef", gamma=1.4, cfl=0.7, nstep=20, steady=True, residual=1.0e-14,
             rescomp=1, sym=(3,), far=(4,), far_density=1.225, far_pressure=101325.0,
             far_velocity=(_C_INF * 0.675, 0.0, 0.0), ic_density=1.225, ic_pressure=101325.0,
             ic_velocity=(_C_INF * 0.675, 0.0, 0.0), velinf=(_C_INF * 0.675, 0.0, 0.0), mesh="laxcg_bump")
LCASES = {
    "laxcg_bump": dict(_BUMP),
    "laxcg_bump_hllc": dict(_BUMP, flux="hllc"),
}

# ChoCG regression cases (tests/regression/inciter/ChoCG/{Poisson,Poiseuille,Lid}/*.q)
_PDIR6 = tuple((s, 1) for s in range(1, 7))
_POIS = dict(solver="chocg", ncomp=3, cfl=0.5, nstep=20, mu=0.01, p_iter=500, p_tol=1.0e-3, p_pc="jacobi",
             p_dir=((1, 2), (2, 2)), p_dirval=((1, 2.4), (2, 0.0)), problem="userdef", noslip=(3, 4),
             dir_=((1, 0, 1, 1), (5, 0, 1, 1)), mesh="chocg_poiseuille")
CCASES = {
    "chocg_poisson_const": dict(solver="chocg", ncomp=3, nstep=1, dt=0.5, problem="poisson_const", p_iter=100,
                                p_tol=1.0e-6, p_dir=_PDIR6, mesh="chocg_unitcube"),
    "chocg_poisson_sine": dict(solver="chocg", ncomp=3, nstep=1, dt=0.5, problem="poisson_sine", p_iter=100,
                               p_tol=1.0e-6, p_dir=_PDIR6, mesh="chocg_unitcube"),
    "chocg_poisson_sine3": dict(solver="chocg", ncomp=3, nstep=1, dt=0.5, problem="poisson_sine3", p_iter=100,
                                p_tol=1.0e-6, p_dir=_PDIR6, mesh="chocg_unitcube"),
    "chocg_poisson_neumann": dict(solver="chocg", ncomp=3, nstep=1, dt=0.5, problem="poisson_neumann", p_iter=100,
                                  p_tol=1.0e-6, p_dir=((1, 1), (2, 1), (3, 1)), p_sym=(4, 5, 6),
                                  mesh="chocg_pidiv4"),
    "chocg_poiseuille_damp2": dict(_POIS, flux="damp2", cfl=0.05),
    "chocg_poiseuille_damp4": dict(_POIS, flux="damp4", cfl=0.05),
    "chocg_poiseuille_rk2": dict(_POIS, flux="damp2", rk=2),
    "chocg_poiseuille_rk3": dict(_POIS, flux="damp2", rk=3),
    "chocg_poiseuille_rk4": dict(_POIS, flux="damp4", rk=4, cfl=1.0),
    # ChoCG/Sphere/inviscid_sphere.q: potential flow around a sphere, symmetry + inflow Dirichlet, pressure
    # Dirichlet at the outflow from the pressure IC
    "chocg_inviscid_sphere": dict(solver="chocg", ncomp=3, nstep=10, cfl=0.5, flux="damp2", p_iter=300, p_tol=1.0e-3,
                                  p_pc="jacobi", p_dir=((3, 1),), problem="userdef", ic_velocity=(1.0, 0.0, 0.0),
                                  dir_=((2, 1, 0, 0),), sym=(1, 4), mesh="sphere2_5k"),
    # ChoCG/Sphere/sphere_chocg_viscous_test.q: Re = 40, damp4, rk 4, no-slip sphere
    "chocg_viscous_sphere": dict(solver="chocg", ncomp=3, nstep=20, cfl=0.3, flux="damp4", rk=4, p_iter=300, p_tol=1.0e-3,
                                 p_pc="jacobi", p_dir=((3, 1),), mu=1.0 / 40.0, problem="userdef",
                                 ic_velocity=(1.0, 0.0, 0.0), dir_=((2, 1, 0, 0), (4, 1, 1, 1)), noslip=(1,),
                                 mesh="sphere2_5k"),
    "chocg_ldc": dict(solver="chocg", ncomp=3, nstep=10, cfl=0.9, flux="damp4", mu=0.01, p_iter=500, p_tol=1.0e-3,
                      p_pc="jacobi", p_hydrostat=0, problem="userdef", noslip=(1, 2, 3, 5, 6),
                      dir_=((4, 2, 2, 2),), dirval=((4, 1.0, 0.0, 0.0),), mesh="riecg_taylor_green"),
}

# ChoCG with the semi-implicit momentum solve (tests/regression/inciter/ChoCG/Poiseuille/poiseuille_theta.q;
# golden recorded on 2 PEs)
ICASES = {
    "chocg_poiseuille_theta": dict(solver="chocg", ncomp=3, nstep=20, cfl=0.5, flux="damp2", theta=0.5, mu=0.01,
                                   mom_iter=50, mom_tol=1.0e-3, mom_pc="jacobi", p_iter=500, p_tol=1.0e-3, p_pc="jacobi",
                                   p_dir=((1, 2), (2, 2)), p_dirval=((1, 2.4), (2, 0.0)), problem="userdef", noslip=(3, 4),
                                   dir_=((1, 0, 1, 1), (5, 0, 1, 1)), mesh="chocg_poiseuille"),
}

# LohCG regression cases (tests/regression/inciter/LohCG/{Poiseuille,Lid}/*.q): artificial compressibility,
# unknowns (p,u,v,w); serial goldens printed with 12 digits
_LPOIS = dict(solver="lohcg", ncomp=4, nstep=20, soundspeed=10.0, mu=0.01, p_iter=500, p_tol=1.0e-3, p_pc="jacobi",
              p_dir=((1, 2), (2, 2)), p_dirval=((1, 2.4), (2, 0.0)), problem="userdef", noslip=(3, 4),
              dir_=((5, 0, 0, 1, 1),), mesh="chocg_poiseuille")
HCASES = {
    "lohcg_poiseuille_damp2": dict(_LPOIS, flux="damp2", cfl=0.5, rk=3),
    "lohcg_poiseuille_damp4": dict(_LPOIS, flux="damp4", cfl=0.3, rk=4),
    # LohCG/Sphere/sphere_lohcg_viscous_test.q: viscous flow around a sphere, Re = 100, damp4 + stab2, rk 4,
    # no-slip sphere, diagnostics every 10th step
    "lohcg_viscous_sphere": dict(solver="lohcg", ncomp=4, nstep=20, cfl=0.1, flux="damp4", stab2=True, stab2coef=0.05,
                                 soundspeed=10.0, rk=4, p_iter=300, p_tol=1.0e-3, p_pc="jacobi", p_dir=((3, 1),),
                                 mu=1.0 / 100.0, problem="userdef", ic_velocity=(1.0, 0.0, 0.0),
                                 dir_=((2, 0, 1, 1, 1), (3, 1, 0, 0, 0), (4, 0, 1, 0, 0)), noslip=(1,), diag_iter=10,
                                 mesh="sphere2_5k"),
    "lohcg_ldc": dict(solver="lohcg", ncomp=4, nstep=20, cfl=0.1, flux="damp2", soundspeed=10.0, rk=2, mu=0.01,
                      p_iter=500, p_tol=1.0e-3, p_pc="jacobi", p_hydrostat=0, problem="userdef",
                      noslip=(1, 2, 3, 5, 6), dir_=((4, 0, 2, 2, 2),), dirval=((4, 0.0, 1.0, 0.0, 0.0),),
                      mesh="riecg_taylor_green"),
}

# KozCG regression cases (tests/regression/inciter/KozCG/{Sod/sod.q,TaylorGreen/taylor_green.q})
KCASES = {
    "kozcg_sod": dict(solver="kozcg", problem="sod", gamma=1.4, cfl=0.5, nstep=10,
                      sym=(2, 4, 5, 6), dir_=((1, 1, 1, 1, 1, 1), (3, 1, 1, 1, 1, 1)),
                      fctclip=True, fctsys=(1, 2, 5), mesh="riecg_sod"),
    "kozcg_taylor_green": dict(solver="kozcg", problem="taylor_green", gamma=5.0 / 3.0, cfl=0.8, term=1.0,
                               fct=False, diag_iter=2, dir_=tuple((s, 1, 1, 1, 1, 1) for s in range(1, 7)),
                               mesh="riecg_taylor_green"),
    # KozCG/VorticalFlow/vortical_flow.q: no FCT, manufactured solution with nodal + centroid source terms
    "kozcg_vortical_flow": dict(solver="kozcg", problem="vortical_flow", alpha=0.1, kappa=1.0, p0=10.0,
                                gamma=5.0 / 3.0, cfl=0.8, term=1.0, fct=False,
                                dir_=tuple((s, 1, 1, 1, 1, 1) for s in range(1, 7)), mesh="riecg_taylor_green"),
}

# ZalCG regression cases (tests/regression/inciter/ZalCG/{Sod/sod.q,Sedov/sedov.q})
ZCASES = {
    "zalcg_sod": dict(solver="zalcg", problem="sod", gamma=1.4, cfl=0.5, nstep=10, term=0.2,
                      sym=(2, 4, 5, 6), dir_=((1, 1, 1, 1, 1, 1), (3, 1, 1, 1, 1, 1)),
                      fctclip=True, fctsys=(1, 2, 5), mesh="riecg_sod"),
    "zalcg_sedov": dict(solver="zalcg", problem="sedov", gamma=5.0 / 3.0, p0=4.86e3, cfl=0.5, nstep=20,
                        term=1.0, sym=(1, 2, 3), fctsys=(1, 2, 3, 4, 5), mesh="riecg_sedov"),
}

# The three RieCG regression cases pinned by golden diag.std files (control files:
# tests/regression/inciter/RieCG/{Sod/sod.q,Sedov/sedov.q,TaylorGreen/taylor_green.q})
CASES = {
    "riecg_sod": dict(problem="sod", gamma=1.4, cfl=0.5, nstep=10, term=0.2, sym=(2, 4, 5, 6)),
    "riecg_sedov": dict(problem="sedov", gamma=5.0 / 3.0, p0=4.86e3, cfl=0.5, nstep=10, term=1.0,
                        sym=(1, 2, 3)),
    "riecg_taylor_green": dict(problem="taylor_green", gamma=5.0 / 3.0, cfl=0.8, term=1.0,
                               diag_iter=2, dir_=tuple((s, 1, 1, 1, 1, 1) for s in range(1, 7))),
}


# Vortical flow (tests/regression/inciter/RieCG/VorticalFlow/vortical_flow{,_hllc,_stab2,_hllc_stab2,_steady}.q):
# a steady manufactured solution with a source term on the unit cube (the mesh of the Taylor-Green case),
# Dirichlet BCs on all sides -- golden diagnostics for both Riemann solvers, stab2 and the steady-state path
_VF = dict(problem="vortical_flow", alpha=0.1, kappa=1.0, p0=10.0, gamma=5.0 / 3.0, cfl=0.8, term=1.0,
           dir_=tuple((s, 1, 1, 1, 1, 1) for s in range(1, 7)), mesh="riecg_taylor_green")
# Time-dependent manufactured solutions on the same mesh: Dirichlet values and source term change in time
# (RieCG/NonlinearEnergyGrowth/nleg.q, RieCG/RayleighTaylor/rayleigh_taylor.q), goldens recorded serially
_DIR6 = tuple((s, 1, 1, 1, 1, 1) for s in range(1, 7))
TCASES = {
    "riecg_nleg": dict(problem="nonlinear_energy_growth", alpha=0.25, r0=2.0, ce=-1.0, kappa=0.8,
                       beta=(1.0, 0.75, 0.5), gamma=5.0 / 3.0, cfl=0.8, term=1.0, dir_=_DIR6,
                       mesh="riecg_taylor_green"),
    "riecg_rayleigh_taylor": dict(problem="rayleigh_taylor", alpha=1.0, beta=(1.0, 1.0, 1.0), p0=1.0, r0=1.0,
                                  kappa=1.0, gamma=5.0 / 3.0, cfl=0.5, nstep=50, dir_=_DIR6,
                                  mesh="riecg_taylor_green"),
}
# Scalar transport (slotted cylinder, cone and hump in a rotating flow, problems::slot_cyl): one transported
# scalar next to the flow variables -- {RieCG,ZalCG,KozCG,ChoCG,LohCG}/SlotCyl/*.q on unitsquare_01_3.6k.exo.
# The RieCG ones also run on the device (tests/test_gpu_scalars.py); the ZalCG/KozCG/ChoCG/LohCG ones pin the
# ORACLE (port and reference objects) for the scalar rows of SURVEY section 8 (a5).
_SC6 = dict(problem="slot_cyl", ncomp=6, gamma=5.0 / 3.0, nstep=20, dir_=((1, 1, 1, 1, 1, 1, 1), (2, 1, 1, 1, 1, 1, 0)),
            mesh="unitsquare_3_6k")
_SCCHO = dict(solver="chocg", problem="slot_cyl", ncomp=4, gamma=5.0 / 3.0, cfl=0.9, nstep=20, flux="damp2", rk=3,
              p_iter=300, p_tol=1.0e-2, p_pc="jacobi", p_hydrostat=0, dir_=((1, 1, 1, 1, 1), (2, 1, 1, 1, 0)),
              mesh="unitsquare_3_6k")
_SCLOH = dict(solver="lohcg", problem="slot_cyl", ncomp=5, gamma=5.0 / 3.0, cfl=0.9, nstep=20, flux="damp2", rk=3,
              stab2=True, stab2coef=0.1, p_iter=300, p_tol=1.0e-2, p_pc="jacobi", p_hydrostat=0,
              dir_=((1, 0, 1, 1, 1, 1), (2, 0, 1, 1, 1, 0)), mesh="unitsquare_3_6k")
# RieCG/Canyon/canyon.q: dispersion from a point source in a street canyon (problems::point_src: the scalar is
# set to 1 inside a sphere every stage), pressure BCs at inlet/outlet, symmetry walls; golden printed with 6 digits
CANYON = dict(problem="point_src", ncomp=6, gamma=1.4, cfl=0.5, nstep=50, sym=(1, 2, 3, 4, 5),
              pre=((6, 1.225, 1.0e5), (7, 1.225, 0.9e5)), ic_density=1.225, ic_pressure=1.0e5,
              ic_velocity=(0.0, 0.0, 0.0), src_location=(3.0, 0.01, 0.0), src_radius=0.2, src_release_time=0.0,
              diag_iter=10, mesh="riecg_canyon")
# ChoCG/Sphere/sphere_point_src.q: the same point source in the projection solver (3 velocities + scalar)
SPHERE_SRC = dict(solver="chocg", problem="point_src", ncomp=4, nstep=20, cfl=0.5, flux="damp2", p_iter=300, p_tol=1.0e-3,
                  p_pc="jacobi", p_dir=((3, 1),), ic_velocity=(1.0, 0.0, 0.0), dir_=((2, 1, 0, 0, 0),), sym=(1, 4),
                  src_location=(-4.95, 0.0, 0.0), src_radius=2.0, src_release_time=0.0, diag_iter=5, mesh="sphere2_5k")
SCASES = {
    "riecg_slot_cyl": dict(_SC6, cfl=0.9),
    "riecg_slot_cyl_hllc": dict(_SC6, cfl=0.9, flux="hllc"),      # (no golden of its own: port vs reference objects only)
    "zalcg_slot_cyl": dict(_SC6, solver="zalcg", cfl=0.5, freezeflow=3.0, freezetime=0.0),
    "kozcg_slot_cyl": dict(_SC6, solver="kozcg", cfl=0.5, freezeflow=3.0, freezetime=0.0),
    "chocg_slot_cyl": dict(_SCCHO),
    "chocg_slot_cyl_damp4": dict(_SCCHO, flux="damp4", rk=4),
    # ChoCG/SlotCyl/slot_cyl_damp4_freeze.q: after t = 0.1 the flow is frozen and the scalar advances with 2 dt
    "chocg_slot_cyl_damp4_freeze": dict(_SCCHO, flux="damp4", rk=4, freezeflow=2.0, freezetime=1.0e-1),
    "lohcg_slot_cyl": dict(_SCLOH),
    "lohcg_slot_cyl_damp4": dict(_SCLOH, flux="damp4", rk=4),
}
# ZalCG/Bump/bump.q: steady-state local time stepping (edge dt = mean of the end nodes' dt), stab2,
# far-field BC, FCT with defaults; serial golden printed with 12 digits
ZSCASES = {
    "zalcg_bump": dict(solver="zalcg", problem="userdef", gamma=1.4, cfl=0.7, nstep=20, steady=True, residual=1.0e-9,
                       rescomp=1, stab2=True, stab2coef=0.05, sym=(3,), far=(4,), far_density=1.0, far_pressure=1.0,
                       far_velocity=(0.7987, 0.0, 0.0), ic_density=1.0, ic_pressure=1.0,
                       ic_velocity=(0.7987, 0.0, 0.0), mesh="laxcg_bump"),
}
# KozCG/{NonlinearEnergyGrowth/nleg.q,RayleighTaylor/rayleigh_taylor.q}: the time-dependent manufactured
# solutions through the element-based solver (no FCT): nodal sources at t, centroid sources at t + dt/2
KTCASES = {
    "kozcg_nleg": dict(TCASES["riecg_nleg"], solver="kozcg", fct=False),
    "kozcg_rayleigh_taylor": dict(TCASES["riecg_rayleigh_taylor"], solver="kozcg", fct=False),
}
# further goldens pinned on the oracle only (same code paths as cases above; not in the GPU lists yet):
# stationary Rayleigh-Taylor (kappa = 0) through RieCG and KozCG, Canyon with far-field instead of pressure BCs
OCASES = {
    "riecg_rayleigh_taylor_st": dict(TCASES["riecg_rayleigh_taylor"], kappa=0.0, nstep=10),
    "kozcg_rayleigh_taylor_st": dict(KTCASES["kozcg_rayleigh_taylor"], kappa=0.0, nstep=10),
    "riecg_canyon_farfield": dict({k: v for k, v in CANYON.items() if k != "pre"}, far=(6, 7), far_density=1.225,
                                  far_pressure=1.0e5, far_velocity=(10.0, 0.0, 0.0)),
}
# RieCG/Pipe/pipe.q: user-defined quiescent IC, symmetry walls, pressure BCs at inlet and outlet
# (physics::prebc, BC.cpp:222-241); serial golden printed with 12 digits
PCASES = {
    "riecg_pipe": dict(problem="userdef", gamma=1.4, cfl=0.5, nstep=20, sym=(2, 4, 5, 6),
                       pre=((1, 1.0, 100000.0), (3, 1.0, 99900.0)), ic_density=1.0, ic_pressure=99950.0,
                       ic_velocity=(0.0, 0.0, 0.0), mesh="riecg_sod"),
}
VCASES = {
    "riecg_vortical_flow": dict(_VF),
    "riecg_vortical_flow_hllc": dict(_VF, flux="hllc"),
    "riecg_vortical_flow_stab2": dict(_VF, stab2=True),
    "riecg_vortical_flow_hllc_stab2": dict(_VF, flux="hllc", stab2=True),
    "riecg_vortical_flow_steady": dict(_VF, term=1e300, nstep=10, steady=True, residual=1.0e-8, rescomp=1),
}


def load_mesh(name):
    m = np.load(os.path.join(GOLDEN, name + ".mesh.npz"))
    return {k: m[k] for k in m.files}


def load_golden_diag(name):
    rows = []
    with open(os.path.join(GOLDEN, name + ".diag.std")) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            rows.append([float(x) for x in line.split()])
    return np.asarray(rows)


def build(flavour):
    target = {"port": "port", "reference": "ref"}[flavour]
    so = os.path.join(ORC, "_build", "liboracle.so") if flavour == "port" else \
        os.path.join(ORC, "_ref", "liboracle_ref.so")
    if flavour == "reference" and not os.path.isdir("/root/reference/src"):
        return so if os.path.exists(so) else None
    subprocess.run(["make", "-s", "-C", ORC, target], check=True)
    return so


_libs = {}


def lib(flavour="port"):
    if flavour in _libs:
        return _libs[flavour]
    so = build(flavour)
    if so is None or not os.path.exists(so):
        _libs[flavour] = None
        return None
    L = C.CDLL(so)
    L.orc_backend.restype = C.c_char_p
    L.orc_last_error.restype = C.c_char_p
    L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_step.argtypes = [C.c_void_p, C.c_int]
    L.orc_ndiag.argtypes = [C.c_void_p]; L.orc_ndiag.restype = C.c_size_t
    L.orc_diagrow.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.orc_diagrow.restype = C.c_size_t
    L.orc_scalar.argtypes = [C.c_void_p, C.c_char_p]; L.orc_scalar.restype = C.c_double
    L.orc_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_size_t]
    L.orc_get.restype = C.c_size_t
    L.orc_set_u.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_kernel.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_double]
    L.orc_set_supedge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_siphash_ids.argtypes = [C.c_void_p, C.c_int]; L.orc_siphash_ids.restype = C.c_uint64
    _libs[flavour] = L
    return L


_DT = {"gid": np.uint64, "inpoel": np.uint64, "triinpoel": np.uint64, "besym": np.uint8,
       "dsupedge0": np.uint64, "dsupedge1": np.uint64, "dsupedge2": np.uint64,
       "dirbcmasks": np.uint64, "symbcnodes": np.uint64, "farbcnodes": np.uint64,
       "prebcnodes": np.uint64, "bface": np.uint64, "commmap": np.uint64,
       "plhs_ia": np.uint64, "plhs_ja": np.uint64, "dirbcmaskp": np.uint64, "noslipbcnodes": np.uint64}


class Oracle:
    """One oracle run (mesh + configuration), `nchare` partitions stepped serially."""

    def __init__(self, mesh, cfg, flavour="port", nchare=1, target=None):
        self.L = lib(flavour)
        if self.L is None:
            raise RuntimeError("oracle flavour %s unavailable" % flavour)
        self.cfg = cfg
        co = np.ascontiguousarray(mesh["coord"], dtype=np.float64)
        tets = np.ascontiguousarray(mesh["tets"], dtype=np.uint64)
        tris = np.ascontiguousarray(mesh["tris"], dtype=np.uint64)
        bt = np.ascontiguousarray(mesh["block_type"], dtype=np.int32)
        bn = np.ascontiguousarray(mesh["block_n"], dtype=np.uint64)
        sid = np.ascontiguousarray(mesh["set_id"], dtype=np.int32)
        soff = np.ascontiguousarray(mesh["set_off"], dtype=np.uint64)
        sel = np.ascontiguousarray(mesh["set_elem"], dtype=np.uint64)
        ssd = np.ascontiguousarray(mesh["set_side"], dtype=np.uint64)
        tg = None if target is None else np.ascontiguousarray(target, dtype=np.uint64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.L.orc_create.argtypes = [C.c_size_t] + [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_void_p]
        self.h = self.L.orc_create(co.shape[1], p(co[0]), p(co[1]), p(co[2]), len(tets), p(tets),
                                   len(tris), p(tris) if len(tris) else None, len(bt), p(bt), p(bn),
                                   len(sid), p(sid), p(soff), p(sel), p(ssd), C.byref(cfg), nchare,
                                   None if tg is None else p(tg))
        if not self.h:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        self.nchare = nchare
        self.ncomp = cfg.ncomp

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def step(self, n=1):
        r = self.L.orc_step(self.h, n)
        if r < 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        return r

    def scalar(self, name):
        return self.L.orc_scalar(self.h, name.encode())

    def diag(self):
        rows = []
        for i in range(self.L.orc_ndiag(self.h)):
            n = self.L.orc_diagrow(self.h, i, None, 0)
            a = np.zeros(n)
            self.L.orc_diagrow(self.h, i, a.ctypes.data_as(C.c_void_p), n)
            rows.append(a)
        return np.asarray(rows)

    def get(self, name, chare=0):
        nb = self.L.orc_get(self.h, chare, name.encode(), None, 0)
        if nb == C.c_size_t(-1).value:
            raise KeyError(name)
        dt = _DT.get(name, np.float64)
        a = np.zeros(nb // np.dtype(dt).itemsize, dtype=dt)
        if nb:
            self.L.orc_get(self.h, chare, name.encode(), a.ctypes.data_as(C.c_void_p), nb)
        if name in ("u", "un", "rhs", "a"):
            a = a.reshape(-1, self.ncomp)
        elif name == "grad":
            a = a.reshape(-1, 3 * self.ncomp)
        elif name in ("p", "q"):
            a = a.reshape(-1, 2 * self.ncomp)
        return a

    def set_u(self, u, chare=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_set_u(self.h, chare, u.ctypes.data_as(C.c_void_p))

    def set_supedges(self, get, chare=0):
        """Take the three superedge groups from `get(name)` (e.g. a host-mirror Solver.get)."""
        for k in range(3):
            ids = np.ascontiguousarray(get("dsupedge%d" % k), np.uint64)
            ints = np.ascontiguousarray(get("dsupint%d" % k), np.float64)
            if self.L.orc_set_supedge(self.h, chare, k, len(ids), ids.ctypes.data_as(C.c_void_p), len(ints),
                                      ints.ctypes.data_as(C.c_void_p)) != 0:
                raise RuntimeError(self.L.orc_last_error().decode())

    def kernel(self, what, stage=0, t=0.0, dt=0.0, chare=0):
        if self.L.orc_kernel(self.h, chare, what.encode(), stage, t, dt) != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())


def numdiff_ok(res, gold, abs_tol, rel_tol):
    """numdiff's 'any' constraint: pass if abs OR rel error is within tolerance
    (tests/regression/inciter/RieCG/Sod/diag.ndiff.cfg)."""
    res = np.asarray(res); gold = np.asarray(gold)
    ae = np.abs(res - gold)
    re = ae / np.maximum(np.minimum(np.abs(res), np.abs(gold)), 1e-300)
    return (ae <= abs_tol) | (re <= rel_tol)


# ---- linear solver oracle (oracle/cg_port.hpp) ------------------------------------------------
def psup_of(inpoel, npoin):
    """Points surrounding points in the reference's linked-list form (psup1 with a leading 0,
    psup2 offsets), neighbour ids ascending per point (tk::genPsup)."""
    nb = [set() for _ in range(npoin)]
    for t in np.asarray(inpoel, dtype=np.int64).reshape(-1, 4):
        for a in t:
            for b in t:
                if a != b:
                    nb[a].add(int(b))
    p1 = [0]; p2 = [0]
    for s in nb:
        p1.extend(sorted(s)); p2.append(len(p1) - 1)
    return np.asarray(p1, np.uint64), np.asarray(p2, np.uint64)


class CGOracle:
    """Serial multi-partition conjugate gradients (restatement of ConjugateGradients.cpp)."""

    def __init__(self, flavour="port", pc="none"):
        self.L = lib(flavour)
        L = self.L
        L.orc_cg_create.restype = C.c_void_p; L.orc_cg_create.argtypes = [C.c_char_p]
        L.orc_cg_destroy.argtypes = [C.c_void_p]
        L.orc_cg_backend.restype = C.c_char_p
        L.orc_cg_add.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t] + [C.c_void_p] * 3 + \
            [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cg_laplacian.argtypes = [C.c_void_p, C.c_int, C.c_size_t] + [C.c_void_p] * 4
        L.orc_cg_dirichlet.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_double, C.c_size_t]
        L.orc_cg_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cg_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_size_t]
        L.orc_cg_get.restype = C.c_size_t
        L.orc_cg_mult.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cg_setup.argtypes = [C.c_void_p]; L.orc_cg_setup.restype = C.c_double
        L.orc_cg_solve.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.POINTER(C.c_uint64)]
        L.orc_cg_solve.restype = C.c_double
        self.h = L.orc_cg_create(pc.encode())
        self.ncomp = []; self.np = []

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_cg_destroy(self.h); self.h = None

    def add(self, inpoel, npoin, ncomp=1, gid=None, comm=None):
        p1, p2 = psup_of(inpoel, npoin)
        gid = np.arange(npoin, dtype=np.uint64) if gid is None else np.ascontiguousarray(gid, np.uint64)
        comm = comm or {}
        ranks = np.asarray(sorted(comm), np.int32)
        off = [0]; g = []
        for r in ranks:
            g.extend(comm[int(r)]); off.append(len(g))
        off = np.asarray(off, np.uint64); g = np.asarray(g, np.uint64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        k = self.L.orc_cg_add(self.h, npoin, ncomp, len(p1), p(p1), p(p2), p(gid), len(ranks),
                              p(ranks) if len(ranks) else None, p(off), p(g) if len(g) else None)
        assert k >= 0, self.L.orc_last_error()
        self.ncomp.append(ncomp); self.np.append(npoin)
        return k

    def laplacian(self, part, inpoel, coord):
        t = np.ascontiguousarray(inpoel, np.uint64); co = np.ascontiguousarray(coord, np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert self.L.orc_cg_laplacian(self.h, part, len(t), p(t), p(co[0]), p(co[1]), p(co[2])) == 0

    def dirichlet(self, part, node, val=0.0, pos=0):
        assert self.L.orc_cg_dirichlet(self.h, part, node, val, pos) == 0

    def set(self, part, x=None, b=None):
        x = None if x is None else np.ascontiguousarray(x, np.float64)
        b = None if b is None else np.ascontiguousarray(b, np.float64)
        self.L.orc_cg_set(self.h, part, None if x is None else x.ctypes.data, None if b is None else b.ctypes.data)

    def get(self, part, name):
        nb = self.L.orc_cg_get(self.h, part, name.encode(), None, 0)
        dt = np.uint64 if name in ("ia", "ja") else np.float64
        a = np.zeros(nb // 8, dt)
        self.L.orc_cg_get(self.h, part, name.encode(), a.ctypes.data, nb)
        return a

    def mult(self, part, x):
        x = np.ascontiguousarray(x, np.float64); r = np.zeros_like(x)
        self.L.orc_cg_mult(self.h, part, x.ctypes.data, r.ctypes.data)
        return r

    def setup(self):
        return self.L.orc_cg_setup(self.h)

    def solve(self, maxit, tol):
        it = C.c_uint64()
        r = self.L.orc_cg_solve(self.h, maxit, tol, C.byref(it))
        return r, int(it.value)
