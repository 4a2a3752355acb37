"""ctypes bindings of include/xyst_b200.h (the device-side C ABI)."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libxyst_b200.so")


class XystError(RuntimeError):
    pass


class LaxParams(C.Structure):
    _fields_ = [("rgas", C.c_double), ("turkel", C.c_double), ("velinf", C.c_double * 3)]


class ZalParams(C.Structure):
    _fields_ = [("fct", C.c_int32), ("fctclip", C.c_int32), ("fctsys_mask", C.c_int32), ("pad_", C.c_int32),
                ("fctdif", C.c_double)]


class ChoParams(C.Structure):
    _fields_ = [("flux", C.c_int32), ("stab", C.c_int32), ("stab2", C.c_int32), ("pad_", C.c_int32),
                ("stab2coef", C.c_double), ("mu", C.c_double)]


class LohParams(C.Structure):
    _fields_ = [("flux", C.c_int32), ("stab", C.c_int32), ("stab2", C.c_int32), ("pad_", C.c_int32),
                ("stab2coef", C.c_double), ("mu", C.c_double), ("soundspeed", C.c_double)]


class Params(C.Structure):
    _fields_ = [("ncomp", C.c_int32), ("flux", C.c_int32), ("stab2", C.c_int32),
                ("exact_muscl", C.c_int32), ("gamma", C.c_double), ("stab2coef", C.c_double)]


FLUX = {"rusanov": 0, "hllc": 1}
_lib = None

SYMBOLS = [
    "xyst_last_error", "xyst_device_count", "xyst_ctx_create", "xyst_ctx_set_stream",
    "xyst_ctx_destroy", "xyst_sync", "xyst_mesh_upload", "xyst_bc_upload", "xyst_dirbc_values",
    "xyst_src_upload", "xyst_state_set", "xyst_state_get", "xyst_riecg_grad", "xyst_grad_get",
    "xyst_riecg_rhs", "xyst_rhs_get", "xyst_rk_update", "xyst_apply_bc", "xyst_dt_min", "xyst_dt_min_all", "xyst_grad_set", "xyst_besym_upload", "xyst_v_upload", "xyst_scalar_pin",
    "xyst_riecg_stage", "xyst_riecg_step", "xyst_diag", "xyst_comm_unique_id", "xyst_comm_init",
    "xyst_halo_upload", "xyst_halo_sum", "xyst_allreduce_min", "xyst_allreduce_sum", "xyst_launch_count",
    "xyst_nedge", "xyst_kernel_time", "xyst_csr_upload", "xyst_csr_mult", "xyst_cg_setup",
    "xyst_cg_solve", "xyst_cg_get_x", "xyst_zalcg_config", "xyst_zalcg_mesh_upload", "xyst_zalcg_rhs",
    "xyst_zalcg_step", "xyst_laxcg_config", "xyst_steady", "xyst_kozcg_mesh_upload", "xyst_kozcg_rhs", "xyst_kozcg_step",
    "xyst_chocg_mesh_upload", "xyst_chocg_bc_upload", "xyst_chocg_set_u", "xyst_chocg_get_u", "xyst_chocg_set_p",
    "xyst_chocg_get", "xyst_chocg_apply_bc", "xyst_chocg_div", "xyst_chocg_vgrad", "xyst_chocg_flux",
    "xyst_chocg_grad", "xyst_chocg_src", "xyst_chocg_rhs", "xyst_chocg_stage", "xyst_chocg_pinit",
    "xyst_chocg_project", "xyst_chocg_pressure_update", "xyst_chocg_dt_min", "xyst_chocg_diag",
    "xyst_kozcg_src", "xyst_chocg_minit", "xyst_chocg_mupdate", "xyst_cg_select", "xyst_csr_update",
    "xyst_lohcg_mesh_upload", "xyst_lohcg_bc_upload", "xyst_lohcg_set_u", "xyst_lohcg_get_u", "xyst_lohcg_get_rhs",
    "xyst_lohcg_apply_bc", "xyst_lohcg_rhs", "xyst_lohcg_stage", "xyst_lohcg_project", "xyst_lohcg_pressure_set",
    "xyst_lohcg_dt_min", "xyst_lohcg_diag", "xyst_chocg_scalars", "xyst_chocg_dirbc_values", "xyst_chocg_pin",
    "xyst_lohcg_scalars", "xyst_lohcg_src", "xyst_chocg_restore_velocity", "xyst_kozcg_freeze",
    "xyst_zalcg_freeze", "xyst_nslot", "xyst_edge_list", "xyst_zalcg_src",
]


def lib():
    """Load libxyst_b200.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    so = os.environ.get("XYST_B200_LIB", SO)      # alternative CUDA builds (e.g. -fmad=false)
    if not os.path.exists(so):
        raise XystError("libxyst_b200.so is missing: run `python -m xyst_b200.build` "
                        "(the CUDA path has no CPU fallback)")
    L = C.CDLL(so, mode=C.RTLD_GLOBAL)
    L.xyst_last_error.restype = C.c_char_p
    L.xyst_ctx_create.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
    L.xyst_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_ctx_destroy.argtypes = [C.c_void_p]
    L.xyst_sync.argtypes = [C.c_void_p]
    L.xyst_mesh_upload.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 3 + \
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_bc_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                 C.c_size_t, C.c_void_p, C.c_void_p]
    L.xyst_dirbc_values.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_src_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_state_set.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_state_get.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_riecg_grad.argtypes = [C.c_void_p]
    L.xyst_grad_get.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_riecg_rhs.argtypes = [C.c_void_p]
    L.xyst_rhs_get.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_rk_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.xyst_apply_bc.argtypes = [C.c_void_p]
    L.xyst_dt_min.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    L.xyst_riecg_stage.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.xyst_riecg_step.argtypes = [C.c_void_p, C.c_double]
    L.xyst_diag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_comm_unique_id.argtypes = [C.c_void_p]
    L.xyst_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.xyst_halo_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_halo_sum.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.xyst_allreduce_min.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.xyst_allreduce_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.xyst_launch_count.argtypes = [C.c_void_p]; L.xyst_launch_count.restype = C.c_uint64
    L.xyst_nedge.argtypes = [C.c_void_p]; L.xyst_nedge.restype = C.c_uint64
    L.xyst_kernel_time.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                   C.POINTER(C.c_uint64)]
    L.xyst_zalcg_config.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_zalcg_mesh_upload.argtypes = L.xyst_mesh_upload.argtypes
    L.xyst_zalcg_rhs.argtypes = [C.c_void_p, C.c_double]
    L.xyst_zalcg_step.argtypes = [C.c_void_p, C.c_double]
    L.xyst_laxcg_config.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_steady.argtypes = [C.c_void_p, C.c_int]
    L.xyst_kozcg_mesh_upload.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 3 + [C.c_size_t] + [C.c_void_p] * 5
    L.xyst_kozcg_rhs.argtypes = [C.c_void_p, C.c_double]
    L.xyst_kozcg_step.argtypes = [C.c_void_p, C.c_double]
    L.xyst_csr_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_csr_mult.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_cg_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_double)]
    L.xyst_cg_solve.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.POINTER(C.c_size_t),
                                C.POINTER(C.c_double)]
    L.xyst_cg_get_x.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_chocg_mesh_upload.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 3 + \
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_chocg_bc_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    for f in (L.xyst_chocg_set_u, L.xyst_chocg_get_u, L.xyst_chocg_set_p, L.xyst_chocg_src):
        f.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_kozcg_freeze.argtypes = [C.c_void_p, C.c_int]
    L.xyst_zalcg_freeze.argtypes = [C.c_void_p, C.c_int]
    L.xyst_nslot.argtypes = [C.c_void_p]; L.xyst_nslot.restype = C.c_size_t
    L.xyst_edge_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_zalcg_src.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_chocg_scalars.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.xyst_chocg_dirbc_values.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_lohcg_scalars.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.xyst_lohcg_src.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_chocg_pin.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double]
    L.xyst_chocg_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    for f in (L.xyst_chocg_apply_bc, L.xyst_chocg_vgrad, L.xyst_chocg_flux, L.xyst_chocg_rhs, L.xyst_chocg_restore_velocity):
        f.argtypes = [C.c_void_p]
    L.xyst_chocg_div.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
    L.xyst_chocg_grad.argtypes = [C.c_void_p, C.c_int]
    L.xyst_chocg_stage.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    L.xyst_chocg_pinit.argtypes = [C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int]
    L.xyst_chocg_project.argtypes = [C.c_void_p, C.c_double]
    L.xyst_chocg_pressure_update.argtypes = [C.c_void_p, C.c_int]
    L.xyst_chocg_dt_min.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.xyst_chocg_diag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_kozcg_src.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_chocg_minit.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    L.xyst_chocg_mupdate.argtypes = [C.c_void_p, C.c_int]
    L.xyst_cg_select.argtypes = [C.c_void_p, C.c_int]
    L.xyst_csr_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.xyst_lohcg_mesh_upload.argtypes = L.xyst_chocg_mesh_upload.argtypes
    L.xyst_lohcg_bc_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]
    for f in (L.xyst_lohcg_set_u, L.xyst_lohcg_get_u, L.xyst_lohcg_get_rhs):
        f.argtypes = [C.c_void_p, C.c_void_p]
    L.xyst_lohcg_apply_bc.argtypes = [C.c_void_p, C.c_int]
    for f in (L.xyst_lohcg_rhs, L.xyst_lohcg_project, L.xyst_lohcg_pressure_set):
        f.argtypes = [C.c_void_p]
    L.xyst_lohcg_stage.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    L.xyst_lohcg_dt_min.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.xyst_lohcg_diag.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One device context = one mesh partition on one GPU (thin wrapper, no logic)."""

    def __init__(self, device=0, flux="rusanov", gamma=1.4, stab2=False, stab2coef=0.2,
                 exact_muscl=False, ncomp=5):
        self.L = lib()
        self.ncomp = ncomp
        prm = Params(ncomp, FLUX[flux], int(stab2), int(exact_muscl), gamma, stab2coef)
        h = C.c_void_p()
        self._ck(self.L.xyst_ctx_create(device, C.byref(prm), C.byref(h)))
        self.h = h
        self.npoin = 0

    def _ck(self, rc):
        if rc != 0:
            raise XystError(self.L.xyst_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.xyst_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.xyst_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._ck(self.L.xyst_sync(self.h))

    def zalcg_config(self, fct=True, fctclip=False, fctsys=(), fctdif=1.0):
        mask = 0
        for c_ in fctsys:
            mask |= 1 << (c_ - 1)
        zp = ZalParams(int(fct), int(fctclip), mask, 0, fctdif)
        self._ck(self.L.xyst_zalcg_config(self.h, C.byref(zp)))

    def laxcg_config(self, rgas=287.052874, turkel=0.5, velinf=(1.0, 1.0, 1.0)):
        p = LaxParams(rgas, turkel, (C.c_double * 3)(*velinf))
        self._ck(self.L.xyst_laxcg_config(self.h, C.byref(p)))

    def steady(self, on=True):
        self._ck(self.L.xyst_steady(self.h, int(on)))

    def kozcg_mesh_upload(self, x, y, z, inpoel, vol, v, Sn=None, Sc=None):
        x, y, z, vol, v = map(_f64, (x, y, z, vol, v))
        t = _u64(inpoel)
        Sn = None if Sn is None else _f64(Sn); Sc = None if Sc is None else _f64(Sc)
        self.npoin = len(x)
        self._ck(self.L.xyst_kozcg_mesh_upload(self.h, len(x), _p(x), _p(y), _p(z), t.size // 4, _p(t),
                                               _p(vol), _p(v), _p(Sn), _p(Sc)))

    def kozcg_rhs(self, dt):
        self._ck(self.L.xyst_kozcg_rhs(self.h, dt))

    def kozcg_step(self, dt):
        self._ck(self.L.xyst_kozcg_step(self.h, dt))

    def zalcg_rhs(self, dt):
        self._ck(self.L.xyst_zalcg_rhs(self.h, dt))

    def zalcg_step(self, dt):
        self._ck(self.L.xyst_zalcg_step(self.h, dt))

    def mesh_upload(self, x, y, z, dsupedge, dsupint, triinpoel, besym, vol, v, stride=3):
        x, y, z, vol, v = map(_f64, (x, y, z, vol, v))
        se = [_u64(a) for a in dsupedge]
        si = [_f64(a) for a in dsupint]
        nsup = (C.c_size_t * 3)(len(se[0]) // 4, len(se[1]) // 3, len(se[2]) // 2)
        pe = (C.c_void_p * 3)(*[a.ctypes.data for a in se])
        pi = (C.c_void_p * 3)(*[a.ctypes.data for a in si])
        tri = _u64(triinpoel)
        bs = np.ascontiguousarray(besym, dtype=np.uint8)
        self.npoin = len(x)
        self._keep = (x, y, z, vol, v, se, si, tri, bs)
        fn = self.L.xyst_mesh_upload if stride == 3 else self.L.xyst_zalcg_mesh_upload
        self._ck(fn(self.h, len(x), _p(x), _p(y), _p(z), nsup, pe, pi,
                    len(tri) // 3, _p(tri), _p(bs), _p(vol), _p(v)))

    def bc_upload(self, dirbcmasks=(), dirvals=None, symbcnodes=(), symbcnorms=(), farbcnodes=(),
                  farbcnorms=(), far=(0.0, 0.0, (0.0, 0.0, 0.0)), prebcnodes=(), prebcvals=()):
        dm = _u64(dirbcmasks); ndir = len(dm) // (self.ncomp + 1)
        dv = _f64(dirvals) if dirvals is not None and ndir else None
        sn = _u64(symbcnodes); snn = _f64(symbcnorms)
        fn = _u64(farbcnodes); fnn = _f64(farbcnorms)
        pn = _u64(prebcnodes); pv = _f64(prebcvals)
        fu = (C.c_double * 3)(*far[2])
        self._ck(self.L.xyst_bc_upload(self.h, ndir, _p(dm), _p(dv), len(sn), _p(sn), _p(snn),
                                       len(fn), _p(fn), _p(fnn), far[0], far[1], fu,
                                       len(pn), _p(pn), _p(pv)))

    def src_upload(self, S):
        S = None if S is None else _f64(S)
        self._ck(self.L.xyst_src_upload(self.h, _p(S)))

    def state_set(self, U):
        U = _f64(U)
        assert U.size == self.npoin * self.ncomp
        self._ck(self.L.xyst_state_set(self.h, _p(U)))

    def state_get(self):
        U = np.empty((self.npoin, self.ncomp))
        self._ck(self.L.xyst_state_get(self.h, _p(U)))
        return U

    def grad(self):
        self._ck(self.L.xyst_riecg_grad(self.h))

    def grad_get(self):
        G = np.empty((self.npoin, 3 * self.ncomp))
        self._ck(self.L.xyst_grad_get(self.h, _p(G)))
        return G

    def rhs(self):
        self._ck(self.L.xyst_riecg_rhs(self.h))

    def rhs_get(self):
        R = np.empty((self.npoin, self.ncomp))
        self._ck(self.L.xyst_rhs_get(self.h, _p(R)))
        return R

    def rk_update(self, stage, dt):
        self._ck(self.L.xyst_rk_update(self.h, stage, dt))

    def apply_bc(self):
        self._ck(self.L.xyst_apply_bc(self.h))

    def dt_min(self, cfl):
        dt = C.c_double()
        self._ck(self.L.xyst_dt_min(self.h, cfl, C.byref(dt)))
        return dt.value

    def stage(self, stage, dt):
        self._ck(self.L.xyst_riecg_stage(self.h, stage, dt))

    def step(self, dt):
        self._ck(self.L.xyst_riecg_step(self.h, dt))

    def diag(self, an=None):
        out = np.zeros(4 * self.ncomp + 1)
        an = None if an is None else _f64(an)
        self._ck(self.L.xyst_diag(self.h, _p(an), _p(out)))
        return out

    # ---- linear solver (tk::CSR + ConjugateGradients) ----
    def csr_upload(self, ia, ja, a, ncomp=1):
        ia = _u64(ia); ja = _u64(ja); a = _f64(a)
        self.cg_nrow = len(ia) - 1
        self._ck(self.L.xyst_csr_upload(self.h, self.cg_nrow, ncomp, _p(ia), _p(ja), _p(a)))

    def csr_mult(self, x):
        x = _f64(x); r = np.empty_like(x)
        self._ck(self.L.xyst_csr_mult(self.h, _p(x), _p(r)))
        return r

    def cg_setup(self, x, b, pc="none", slave=None, count=None):
        x = _f64(x); b = _f64(b)
        sl = None if slave is None else np.ascontiguousarray(slave, np.uint8)
        ct = None if count is None else _f64(count)
        nb = C.c_double()
        self._ck(self.L.xyst_cg_setup(self.h, _p(x), _p(b), {"none": 0, "jacobi": 1}[pc], _p(sl), _p(ct),
                                      C.byref(nb)))
        return nb.value

    def cg_solve(self, maxit, tol):
        it = C.c_size_t(); nr = C.c_double()
        self._ck(self.L.xyst_cg_solve(self.h, maxit, tol, C.byref(it), C.byref(nr)))
        return nr.value, int(it.value)

    def cg_x(self):
        x = np.empty(self.cg_nrow)
        self._ck(self.L.xyst_cg_get_x(self.h, _p(x)))
        return x

    # ---- ChoCG (projection method; Chorin edge operators) ----
    CHO_W = {"pr": 1, "div": 1, "dp": 1, "sgrad": 3, "pgrad": 3, "flux": 3, "rhs": 3, "un": 3, "u": 3, "vgrad": 9}

    def chocg_mesh_upload(self, x, y, z, dsupedge, dsupint, triinpoel, vol, v, flux="damp2", stab=True,
                          stab2=False, stab2coef=0.1, mu=0.0):
        x, y, z, vol, v = map(_f64, (x, y, z, vol, v))
        se = [_u64(a) for a in dsupedge]
        si = [_f64(a) for a in dsupint]
        nsup = (C.c_size_t * 3)(len(se[0]) // 4, len(se[1]) // 3, len(se[2]) // 2)
        pe = (C.c_void_p * 3)(*[a.ctypes.data for a in se])
        pi = (C.c_void_p * 3)(*[a.ctypes.data for a in si])
        tri = _u64(triinpoel)
        self.npoin = len(x)
        prm = ChoParams({"damp2": 0, "damp4": 1}[flux], int(stab), int(stab2), 0, stab2coef, mu)
        self._ck(self.L.xyst_chocg_mesh_upload(self.h, len(x), _p(x), _p(y), _p(z), nsup, pe, pi,
                                               len(tri) // 3, _p(tri), _p(vol), _p(v), C.byref(prm)))

    def chocg_bc_upload(self, dirnodes=(), dirmask=(), dirval=None, symbcnodes=(), symbcnorms=(), noslipbcnodes=()):
        dn = _u64(dirnodes); dm = np.ascontiguousarray(dirmask, np.int32)
        dv = None if dirval is None else _f64(dirval)
        sn = _u64(symbcnodes); snn = _f64(symbcnorms); nn = _u64(noslipbcnodes)
        self._ck(self.L.xyst_chocg_bc_upload(self.h, len(dn), _p(dn), _p(dm), _p(dv), len(sn), _p(sn), _p(snn),
                                             len(nn), _p(nn)))

    def chocg_set_u(self, u):
        u = _f64(u); self._ck(self.L.xyst_chocg_set_u(self.h, _p(u)))

    def chocg_set_p(self, p):
        p = _f64(p); self._ck(self.L.xyst_chocg_set_p(self.h, _p(p)))

    def chocg_get(self, what):
        w = self.CHO_W[what]
        out = np.empty((self.npoin, w)) if w > 1 else np.empty(self.npoin)
        self._ck(self.L.xyst_chocg_get(self.h, what.encode(), _p(out)))
        return out

    def chocg_apply_bc(self):
        self._ck(self.L.xyst_chocg_apply_bc(self.h))

    def chocg_div(self, which=0, dt=0.0, stab=False):
        self._ck(self.L.xyst_chocg_div(self.h, which, dt, int(stab)))

    def chocg_vgrad(self):
        self._ck(self.L.xyst_chocg_vgrad(self.h))

    def chocg_flux(self):
        self._ck(self.L.xyst_chocg_flux(self.h))

    def chocg_grad(self, which):
        self._ck(self.L.xyst_chocg_grad(self.h, which))

    def chocg_src(self, S):
        S = None if S is None else _f64(S)
        self._ck(self.L.xyst_chocg_src(self.h, _p(S)))

    def chocg_rhs(self):
        self._ck(self.L.xyst_chocg_rhs(self.h))

    def chocg_stage(self, stage, rkcoef, dt):
        self._ck(self.L.xyst_chocg_stage(self.h, stage, rkcoef, dt))

    def chocg_pinit(self, divisor=1.0, bcnodes=(), bcvals=None, neubc=None, rhs0=None, pc="none"):
        bn = _u64(bcnodes); bv = None if bcvals is None else _f64(bcvals)
        ne = None if neubc is None else _f64(neubc); r0 = None if rhs0 is None else _f64(rhs0)
        self.cg_nrow = self.npoin
        self._ck(self.L.xyst_chocg_pinit(self.h, divisor, len(bn), _p(bn), _p(bv), _p(ne), _p(r0),
                                         {"none": 0, "jacobi": 1}[pc]))

    def chocg_project(self, pdt):
        self._ck(self.L.xyst_chocg_project(self.h, pdt))

    def chocg_pressure_update(self, increment):
        self._ck(self.L.xyst_chocg_pressure_update(self.h, int(increment)))

    def chocg_dt_min(self, cfl, dif=0.0):
        dt = C.c_double()
        self._ck(self.L.xyst_chocg_dt_min(self.h, cfl, dif, C.byref(dt)))
        return dt.value

    def chocg_diag(self, an_p=None, an_u=None):
        out = np.zeros(16)
        ap = None if an_p is None else _f64(an_p); au = None if an_u is None else _f64(an_u)
        self._ck(self.L.xyst_chocg_diag(self.h, _p(ap), _p(au), _p(out)))
        return out

    def chocg_minit(self, bcrows=(), pc="none"):
        br = _u64(bcrows)
        self._ck(self.L.xyst_chocg_minit(self.h, len(br), _p(br), {"none": 0, "jacobi": 1}[pc]))

    def chocg_mupdate(self, stage):
        self._ck(self.L.xyst_chocg_mupdate(self.h, stage))

    def cg_select(self, which):
        self._ck(self.L.xyst_cg_select(self.h, which))

    def csr_update(self, ia, a):
        ia = _u64(ia); a = _f64(a)
        self._ck(self.L.xyst_csr_update(self.h, _p(ia), _p(a)))

    # ---- LohCG (artificial compressibility; Lohner edge operators). div/vgrad/flux/grad(0)/pinit/get of
    # the chocg_* methods act on the velocity part of the same context ----
    def lohcg_mesh_upload(self, x, y, z, dsupedge, dsupint, triinpoel, vol, v, flux="damp2", stab=True,
                          stab2=False, stab2coef=0.1, mu=0.0, soundspeed=1.0):
        x, y, z, vol, v = map(_f64, (x, y, z, vol, v))
        se = [_u64(a) for a in dsupedge]
        si = [_f64(a) for a in dsupint]
        nsup = (C.c_size_t * 3)(len(se[0]) // 4, len(se[1]) // 3, len(se[2]) // 2)
        pe = (C.c_void_p * 3)(*[a.ctypes.data for a in se])
        pi = (C.c_void_p * 3)(*[a.ctypes.data for a in si])
        tri = _u64(triinpoel)
        self.npoin = len(x)
        prm = LohParams({"damp2": 0, "damp4": 1}[flux], int(stab), int(stab2), 0, stab2coef, mu, soundspeed)
        self._ck(self.L.xyst_lohcg_mesh_upload(self.h, len(x), _p(x), _p(y), _p(z), nsup, pe, pi,
                                               len(tri) // 3, _p(tri), _p(vol), _p(v), C.byref(prm)))

    def lohcg_bc_upload(self, dirnodes=(), dirmask=(), dirval=None, pdirnodes=(), pdirval=(), symbcnodes=(),
                        symbcnorms=(), noslipbcnodes=()):
        dn = _u64(dirnodes); dm = np.ascontiguousarray(dirmask, np.int32)
        dv = None if dirval is None else _f64(dirval)
        pn = _u64(pdirnodes); pv = _f64(pdirval)
        sn = _u64(symbcnodes); snn = _f64(symbcnorms); nn = _u64(noslipbcnodes)
        self._ck(self.L.xyst_lohcg_bc_upload(self.h, len(dn), _p(dn), _p(dm), _p(dv), len(pn), _p(pn), _p(pv),
                                             len(sn), _p(sn), _p(snn), len(nn), _p(nn)))

    def lohcg_set_u(self, u):
        u = _f64(u); self._ck(self.L.xyst_lohcg_set_u(self.h, _p(u)))

    def lohcg_get_u(self):
        out = np.empty((self.npoin, 4)); self._ck(self.L.xyst_lohcg_get_u(self.h, _p(out))); return out

    def lohcg_get_rhs(self):
        out = np.empty((self.npoin, 4)); self._ck(self.L.xyst_lohcg_get_rhs(self.h, _p(out))); return out

    def lohcg_apply_bc(self, pressure=True):
        self._ck(self.L.xyst_lohcg_apply_bc(self.h, int(pressure)))

    def lohcg_rhs(self):
        self._ck(self.L.xyst_lohcg_rhs(self.h))

    def lohcg_stage(self, stage, rkcoef, dt):
        self._ck(self.L.xyst_lohcg_stage(self.h, stage, rkcoef, dt))

    def lohcg_project(self):
        self._ck(self.L.xyst_lohcg_project(self.h))

    def lohcg_pressure_set(self):
        self._ck(self.L.xyst_lohcg_pressure_set(self.h))

    def lohcg_dt_min(self, cfl, dif=0.0):
        dt = C.c_double(0.0)
        self._ck(self.L.xyst_lohcg_dt_min(self.h, cfl, dif, C.byref(dt)))
        return dt.value

    def lohcg_diag(self, an=None):
        out = np.zeros(16)
        a = None if an is None else _f64(an)
        self._ck(self.L.xyst_lohcg_diag(self.h, _p(a), _p(out)))
        return out

    def launch_count(self):
        return int(self.L.xyst_launch_count(self.h))

    def nedge(self):
        return int(self.L.xyst_nedge(self.h))

    def kernel_time(self, name, reset=False):
        ms = C.c_double(); n = C.c_uint64()
        self._ck(self.L.xyst_kernel_time(self.h, name.encode(), int(reset), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)
