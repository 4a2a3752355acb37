"""ctypes bindings of include/xyst_host.h (the C++ host mirror: Discretization + RieCG)."""
import ctypes as C
import os
import numpy as np
from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libxyst_host.so")


class HostCfg(C.Structure):
    _fields_ = [
        ("problem", C.c_char * 32), ("flux", C.c_char * 16),
        ("ncomp", C.c_int32), ("stab2", C.c_int32), ("exact_muscl", C.c_int32), ("reforder", C.c_int32),
        ("nsym", C.c_int32), ("sym", C.c_int32 * 16),
        ("ndir", C.c_int32), ("dir", (C.c_int32 * 12) * 16),
        ("nfar", C.c_int32), ("far_sets", C.c_int32 * 16),
        ("npre", C.c_int32), ("pre_sets", C.c_int32 * 16),
        ("nstep", C.c_uint64), ("diag_iter", C.c_uint64),
        ("gamma", C.c_double), ("p0", C.c_double), ("cfl", C.c_double), ("dt", C.c_double),
        ("t0", C.c_double), ("term", C.c_double), ("stab2coef", C.c_double),
        ("far_density", C.c_double), ("far_pressure", C.c_double), ("far_velocity", C.c_double * 3),
        ("pre_density", C.c_double * 16), ("pre_pressure", C.c_double * 16),
        ("solver", C.c_char * 16), ("fct", C.c_int32), ("fctclip", C.c_int32), ("nfctsys", C.c_int32),
        ("fctsys", C.c_int32 * 8), ("fctdif", C.c_double),
        ("steady", C.c_int32), ("rescomp", C.c_uint64), ("residual", C.c_double),
        ("rgas", C.c_double), ("turkel", C.c_double), ("velinf", C.c_double * 3),
        ("ic_density", C.c_double), ("ic_pressure", C.c_double), ("ic_velocity", C.c_double * 3),
        ("mu", C.c_double), ("dif", C.c_double), ("stab", C.c_int32),
        ("nnoslip", C.c_int32), ("noslip", C.c_int32 * 16), ("rk", C.c_uint64),
        ("ndirval", C.c_int32), ("dirval", (C.c_double * 12) * 16),
        ("p_iter", C.c_uint64), ("p_tol", C.c_double), ("p_pc", C.c_char * 16),
        ("np_dir", C.c_int32), ("p_dir", (C.c_int32 * 2) * 16),
        ("np_dirval", C.c_int32), ("p_dirval", (C.c_double * 2) * 16),
        ("np_sym", C.c_int32), ("p_sym", C.c_int32 * 16),
        ("p_hydrostat_set", C.c_int32), ("p_hydrostat", C.c_uint64),
        ("alpha", C.c_double), ("kappa", C.c_double),
        ("r0", C.c_double), ("ce", C.c_double), ("beta", C.c_double * 3),
        ("theta", C.c_double), ("mom_iter", C.c_uint64), ("mom_tol", C.c_double), ("mom_pc", C.c_char * 16),
        ("soundspeed", C.c_double),
        ("src_location", C.c_double * 3), ("src_radius", C.c_double), ("src_release_time", C.c_double),
        ("freezeflow", C.c_double), ("freezetime", C.c_double),
    ]


COMM_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_double))


def make_cfg(problem, flux="rusanov", gamma=1.4, p0=0.0, cfl=0.0, dt=0.0, t0=0.0, term=1e300,
             nstep=2**63, sym=(), dir_=(), stab2=False, stab2coef=0.2, diag_iter=1, ncomp=5,
             exact_muscl=False, reforder=-1, solver="riecg", fct=True, fctclip=False, fctsys=(),
             fctdif=1.0, steady=False, residual=0.0, rescomp=1, rgas=287.052874, turkel=0.5,
             velinf=(1.0, 1.0, 1.0), far=(), far_density=0.0, far_pressure=0.0, far_velocity=(0.0, 0.0, 0.0),
             ic_density=0.0, ic_pressure=0.0, ic_velocity=(0.0, 0.0, 0.0), mu=0.0, dif=0.0, stab=True, rk=1,
             noslip=(), dirval=(), p_iter=10, p_tol=1.0e-3, p_pc="none", p_dir=(), p_dirval=(), p_sym=(),
             p_hydrostat=None, alpha=0.0, kappa=0.0, r0=0.0, ce=0.0, beta=(0.0, 0.0, 0.0), pre=(),
             soundspeed=1.0, theta=0.0, mom_iter=10, mom_tol=1.0e-3, mom_pc="none", src_location=(0.0, 0.0, 0.0),
             src_radius=-1.0, src_release_time=0.0, freezeflow=1.0, freezetime=0.0, **_ignored):
    c = HostCfg()
    c.problem = problem.encode(); c.flux = flux.encode(); c.ncomp = ncomp
    c.gamma = gamma; c.p0 = p0; c.cfl = cfl; c.dt = dt; c.t0 = t0; c.term = term; c.alpha = alpha; c.kappa = kappa
    c.r0 = r0; c.ce = ce; c.soundspeed = soundspeed
    c.src_radius = src_radius; c.src_release_time = src_release_time
    c.freezeflow = freezeflow; c.freezetime = freezetime
    for i in range(3):
        c.src_location[i] = src_location[i]
    c.theta = theta; c.mom_iter = mom_iter; c.mom_tol = mom_tol; c.mom_pc = mom_pc.encode()
    for i in range(3):
        c.beta[i] = beta[i]
    c.npre = len(pre)
    for i, (sid, dens, pres) in enumerate(pre):
        c.pre_sets[i] = sid; c.pre_density[i] = dens; c.pre_pressure[i] = pres
    c.nstep = nstep; c.stab2 = int(stab2); c.stab2coef = stab2coef
    c.exact_muscl = int(exact_muscl); c.diag_iter = diag_iter; c.reforder = reforder
    c.solver = solver.encode(); c.fct = int(fct); c.fctclip = int(fctclip); c.fctdif = fctdif
    c.nfctsys = len(fctsys)
    for i, s_ in enumerate(fctsys):
        c.fctsys[i] = s_
    c.steady = int(steady); c.residual = residual; c.rescomp = rescomp; c.rgas = rgas; c.turkel = turkel
    c.nfar = len(far); c.far_density = far_density; c.far_pressure = far_pressure
    for i, s_ in enumerate(far):
        c.far_sets[i] = s_
    c.ic_density = ic_density; c.ic_pressure = ic_pressure
    for i in range(3):
        c.velinf[i] = velinf[i]; c.far_velocity[i] = far_velocity[i]; c.ic_velocity[i] = ic_velocity[i]
    c.nsym = len(sym)
    for i, s in enumerate(sym):
        c.sym[i] = s
    c.ndir = len(dir_)
    for i, m in enumerate(dir_):
        for j, v in enumerate(m):
            c.dir[i][j] = v
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
    c.p_hydrostat_set = int(p_hydrostat is not None); c.p_hydrostat = 0 if p_hydrostat is None else p_hydrostat
    return c


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    capi.lib()                      # libxyst_b200.so first (RTLD_GLOBAL), no fallback
    if not os.path.exists(SO):
        raise capi.XystError("libxyst_host.so is missing: run `python -m xyst_b200.build`")
    L = C.CDLL(SO, mode=C.RTLD_GLOBAL)
    L.xyst_host_last_error.restype = C.c_char_p
    vp = C.c_void_p
    L.xyst_solver_create_box.argtypes = [C.POINTER(HostCfg), C.c_size_t, C.c_size_t, C.c_size_t,
                                         C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                         C.POINTER(vp)]
    L.xyst_solver_create_mesh.argtypes = [C.POINTER(HostCfg), C.c_size_t, vp, vp, vp, C.c_size_t, vp,
                                          C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.xyst_solver_create_exo.argtypes = [C.POINTER(HostCfg), C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.xyst_exo_read.argtypes = [C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int),
                                C.POINTER(C.c_size_t)] + [vp] * 7
    L.xyst_solver_diag_file.argtypes = [vp, C.c_char_p, C.c_int]
    for f in ("destroy", "prepare", "host_setup", "setup"):
        getattr(L, "xyst_solver_" + f).argtypes = [vp]
    L.xyst_solver_attach.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.xyst_solver_set_comm.argtypes = [vp, COMM_FN, vp, C.c_int, C.c_int]
    L.xyst_solver_set_u0.argtypes = [vp, vp]
    L.xyst_solver_set_u.argtypes = [vp, vp]
    L.xyst_solver_step.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_size_t)]
    L.xyst_solver_step_unfused.argtypes = [vp, C.c_int]
    L.xyst_solver_scalar.argtypes = [vp, C.c_char_p]; L.xyst_solver_scalar.restype = C.c_double
    L.xyst_solver_get.argtypes = [vp, C.c_char_p, vp, C.c_size_t]; L.xyst_solver_get.restype = C.c_size_t
    L.xyst_solver_ctx.argtypes = [vp]; L.xyst_solver_ctx.restype = vp
    L.xyst_box_counts.argtypes = [C.c_size_t] * 3 + [C.POINTER(C.c_size_t)] * 3
    L.xyst_box_mesh.argtypes = [C.c_size_t] * 3 + [C.c_double] * 3 + [vp] * 7
    L.xyst_rcb.argtypes = [C.c_size_t, vp, vp, vp, C.c_size_t, vp, C.c_int, vp]
    L.xyst_chare_count.argtypes = [C.c_double, C.c_uint64, C.c_int, vp, vp, vp]
    L.xyst_rib.argtypes = [C.c_size_t, vp, vp, vp, C.c_size_t, vp, C.c_int, vp]
    L.xyst_box_part_range.argtypes = [C.c_size_t] * 3 + [C.c_int, C.c_int, vp]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise capi.XystError(lib().xyst_host_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_DT = {"gid": np.uint64, "inpoel": np.uint64, "triinpoel": np.uint64, "besym": np.uint8,
       "dsupedge0": np.uint64, "dsupedge1": np.uint64, "dsupedge2": np.uint64,
       "dirbcmasks": np.uint64, "symbcnodes": np.uint64, "bface": np.uint64, "commmap": np.uint64,
       "shared": np.uint64, "plhs_ia": np.uint64, "plhs_ja": np.uint64, "dirbcmaskp": np.uint64,
       "noslipbcnodes": np.uint64, "mbcrows": np.uint64}


def box_mesh(nx, ny, nz, Lx=1.0, Ly=1.0, Lz=1.0):
    """Full structured box mesh as arrays (coord 3xN, tets ntetx4, set ids, offsets, triangles)."""
    L = lib()
    npn, nt, ntri = C.c_size_t(), C.c_size_t(), C.c_size_t()
    L.xyst_box_counts(nx, ny, nz, C.byref(npn), C.byref(nt), C.byref(ntri))
    co = np.zeros((3, npn.value)); tets = np.zeros((nt.value, 4), np.uint64)
    sid = np.zeros(6, np.int32); soff = np.zeros(7, np.uint64); stri = np.zeros((ntri.value, 3), np.uint64)
    _ck(L.xyst_box_mesh(nx, ny, nz, Lx, Ly, Lz, _p(co[0]), _p(co[1]), _p(co[2]), _p(tets), _p(sid),
                        _p(soff), _p(stri)))
    return dict(coord=co, tets=tets, set_id=sid, set_off=soff, set_tri=stri)


def exo_read(path):
    """An ExodusII mesh file as arrays (coord 3xN, tets ntetx4, side-set ids, offsets, triangles)."""
    L = lib()
    npn, nt, ns, ntri = C.c_size_t(), C.c_size_t(), C.c_int(), C.c_size_t()
    _ck(L.xyst_exo_read(path.encode(), C.byref(npn), C.byref(nt), C.byref(ns), C.byref(ntri), *([None] * 7)))
    co = np.zeros((3, npn.value)); tets = np.zeros((nt.value, 4), np.uint64)
    sid = np.zeros(ns.value, np.int32); soff = np.zeros(ns.value + 1, np.uint64); stri = np.zeros((ntri.value, 3), np.uint64)
    _ck(L.xyst_exo_read(path.encode(), None, None, None, None, _p(co[0]), _p(co[1]), _p(co[2]), _p(tets), _p(sid),
                        _p(soff), _p(stri)))
    return dict(coord=co, tets=tets, set_id=sid, set_off=soff, set_tri=stri)


def rcb(coord, tets, nparts):
    L = lib()
    co = np.ascontiguousarray(coord, np.float64); t = np.ascontiguousarray(tets, np.uint64)
    part = np.zeros(len(t), np.int32)
    _ck(L.xyst_rcb(co.shape[1], _p(co[0]), _p(co[1]), _p(co[2]), len(t), _p(t), nparts, _p(part)))
    return part


def rib(coord, tets, nparts):
    L = lib()
    co = np.ascontiguousarray(coord, np.float64); t = np.ascontiguousarray(tets, np.uint64)
    part = np.zeros(len(t), np.int32)
    _ck(L.xyst_rib(co.shape[1], _p(co[0]), _p(co[1]), _p(co[2]), len(t), _p(t), nparts, _p(part)))
    return part


def chare_count(virtualization, load, npe):
    """(nchare, chunksize, remainder) of tk::linearLoadDistributor (the reference's -u over-decomposition)."""
    v = (C.c_uint64 * 3)()
    a = C.addressof(v)
    _ck(lib().xyst_chare_count(virtualization, load, npe, a, a + 8, a + 16))
    return int(v[2]), int(v[0]), int(v[1])


def box_part_range(nx, ny, nz, nparts, part):
    r = np.zeros(6, np.uint64)
    _ck(lib().xyst_box_part_range(nx, ny, nz, nparts, part, _p(r)))
    return r


class Solver:
    """The host-mirror RieCG solver of one mesh partition (thin wrapper, no logic)."""

    def __init__(self, handle, cfg, ncomp=5):
        self.L = lib(); self.h = handle; self.cfg = cfg; self.ncomp = ncomp; self._cb = None

    @classmethod
    def box(cls, cfg, nx, ny, nz, Lx=1.0, Ly=1.0, Lz=1.0, nparts=1, part=0):
        h = C.c_void_p()
        _ck(lib().xyst_solver_create_box(C.byref(cfg), nx, ny, nz, Lx, Ly, Lz, nparts, part, C.byref(h)))
        return cls(h, cfg, cfg.ncomp)

    @classmethod
    def mesh(cls, cfg, coord, tets, set_id, set_off, set_tri, nparts=1, part=0, tetpart=None):
        co = np.ascontiguousarray(coord, np.float64); t = np.ascontiguousarray(tets, np.uint64)
        sid = np.ascontiguousarray(set_id, np.int32); so = np.ascontiguousarray(set_off, np.uint64)
        st = np.ascontiguousarray(set_tri, np.uint64)
        tp = None if tetpart is None else np.ascontiguousarray(tetpart, np.int32)
        h = C.c_void_p()
        _ck(lib().xyst_solver_create_mesh(C.byref(cfg), co.shape[1], _p(co[0]), _p(co[1]), _p(co[2]),
                                          len(t), _p(t), len(sid), _p(sid), _p(so), _p(st),
                                          nparts, part, _p(tp), C.byref(h)))
        return cls(h, cfg, cfg.ncomp)

    @classmethod
    def exo(cls, cfg, path, nparts=1, part=0):
        h = C.c_void_p()
        _ck(lib().xyst_solver_create_exo(C.byref(cfg), path.encode(), nparts, part, C.byref(h)))
        return cls(h, cfg, cfg.ncomp)

    def diag_file(self, path, precision=8):
        _ck(self.L.xyst_solver_diag_file(self.h, path.encode(), precision))

    def close(self):
        if getattr(self, "h", None):
            self.L.xyst_solver_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def prepare(self):
        _ck(self.L.xyst_solver_prepare(self.h))

    def attach(self, device=0, nranks=1, rank=0, ncclid=None):
        _ck(self.L.xyst_solver_attach(self.h, device, nranks, rank, ncclid))

    def set_comm(self, fn, nranks, rank):
        """fn(op, w, n, array): op 0 halo sum on unique shared nodes, 1 allreduce sum, 2 min."""
        def cb(user, op, w, n, vals):
            a = np.ctypeslib.as_array(vals, shape=(n * w,))
            fn(op, w, n, a)
        self._cb = COMM_FN(cb)
        _ck(self.L.xyst_solver_set_comm(self.h, self._cb, None, nranks, rank))

    def set_u0(self, u0):
        u0 = np.ascontiguousarray(u0, np.float64)
        _ck(self.L.xyst_solver_set_u0(self.h, _p(u0)))

    def set_u(self, u):
        u = np.ascontiguousarray(u, np.float64)
        _ck(self.L.xyst_solver_set_u(self.h, _p(u)))

    def host_setup(self):
        _ck(self.L.xyst_solver_host_setup(self.h))

    def setup(self):
        _ck(self.L.xyst_solver_setup(self.h))

    def step(self, nsteps=1, want_diag=True):
        rows = np.zeros((max(nsteps, 1), 32)) if want_diag else None
        nr, nc = C.c_size_t(), C.c_size_t()
        _ck(self.L.xyst_solver_step(self.h, nsteps, _p(rows), 0 if rows is None else rows.size,
                                    C.byref(nr), C.byref(nc)))
        if rows is None or nr.value == 0:
            return np.zeros((0, 0))
        return rows.reshape(-1)[: nr.value * nc.value].reshape(nr.value, nc.value).copy()

    def step_unfused(self, nsteps=1):
        _ck(self.L.xyst_solver_step_unfused(self.h, nsteps))

    def scalar(self, name):
        return self.L.xyst_solver_scalar(self.h, name.encode())

    def get(self, name):
        nb = self.L.xyst_solver_get(self.h, name.encode(), None, 0)
        if nb == C.c_size_t(-1).value:
            raise capi.XystError(self.L.xyst_host_last_error().decode())
        dt = _DT.get(name, np.float64)
        a = np.zeros(nb // np.dtype(dt).itemsize, dtype=dt)
        if nb:
            self.L.xyst_solver_get(self.h, name.encode(), _p(a), nb)
        if name in ("u", "u0"):
            a = a.reshape(-1, self.ncomp)
        return a

    def ctx(self):
        """A capi.Context view of the solver's device context (not owning)."""
        c = capi.Context.__new__(capi.Context)
        c.L = capi.lib(); c.h = C.c_void_p(self.L.xyst_solver_ctx(self.h)); c.ncomp = self.ncomp
        c.npoin = int(self.scalar("npoin")); c.close = lambda: None
        return c
