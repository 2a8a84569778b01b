"""ctypes binding of oracle/libhot_oracle.so — TEST INFRASTRUCTURE (the checker), never the product path."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# HOT_ORACLE_LIB: another build of the same oracle (oracle/_ref/libhot_oracle_lbfgsref.so = the oracle + the reference's own LBFGS loop)
_lib = C.CDLL(os.environ.get("HOT_ORACLE_LIB") or os.path.join(ROOT, "oracle", "libhot_oracle.so"))
_lib.orc_create.restype = C.c_void_p
_lib.orc_create.argtypes = [C.c_double, C.c_double, C.c_double]
_lib.orc_last_error.restype = C.c_char_p
for _n in ["orc_num_particles", "orc_num_groups", "orc_num_pages", "orc_activate"]:
    getattr(_lib, _n).restype = C.c_long
lib = _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _vp(h):
    return C.c_void_p(h)


def mask_info(fp32=False):
    out = (C.c_int * 6)()
    masks = (C.c_ulonglong * 3)()
    _lib.orc_mask_info(int(fp32), out, masks)
    return list(out), [int(m) for m in masks]


def linear_offset(ijk, fp32=False):
    ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
    out = np.empty(len(ijk), dtype=np.uint64)
    _lib.orc_linear_offset(int(fp32), C.c_long(len(ijk)), _p(ijk), _p(out))
    return out


def linear_to_coord(off, fp32=False):
    off = np.ascontiguousarray(off, dtype=np.uint64)
    out = np.empty((len(off), 3), dtype=np.int32)
    _lib.orc_linear_to_coord(int(fp32), C.c_long(len(off)), _p(off), _p(out))
    return out


def packed_add(a, b, fp32=False):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty(len(a), dtype=np.uint64)
    _lib.orc_packed_add(int(fp32), C.c_long(len(a)), _p(a), _p(b), _p(out))
    return out


def activate(group_offsets, fp32=False):
    g = np.ascontiguousarray(group_offsets, dtype=np.uint64)
    out = np.empty(9 * len(g) + 8, dtype=np.uint64)
    n = _lib.orc_activate(int(fp32), C.c_long(len(g)), _p(g), _p(out), C.c_long(len(out)))
    return out[:n].copy()


class OracleSim:
    """Same method names as hot_b200.MpmSimulationB200 so one harness drives both."""
    ELEMENTS_PER_BLOCK = 32

    def __init__(self, dx, apic_rpic_ratio=1.0, cfl=0.6):
        self._h = _lib.orc_create(float(dx), float(apic_rpic_ratio), float(cfl))
        self.dx = dx

    def close(self):
        if self._h:
            _lib.orc_destroy(_vp(self._h))
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(_lib.orc_last_error(_vp(self._h)).decode())

    def set_particles(self, X, V, mass, C_, F, vol, mu, lam):
        n = len(mass)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = [f(X), f(V), f(mass), f(C_), f(F), f(vol), f(mu), f(lam)]
        self._check(_lib.orc_set_particles(_vp(self._h), C.c_long(n), *[_p(a) for a in arrs]))
        self.N = n

    def get_particles(self, gradV=True):
        n = self.N
        X = np.empty((n, 3)); V = np.empty((n, 3)); Cm = np.empty((n, 9)); F = np.empty((n, 9)); G = np.empty((n, 9))
        self._check(_lib.orc_get_particles(_vp(self._h), _p(X), _p(V), _p(Cm), _p(F), _p(G)))
        return dict(X=X, V=V, C=Cm, F=F, gradV=G)

    def sortParticlesAndPolluteGrid(self):
        self._check(_lib.orc_sort_and_activate(_vp(self._h)))

    @property
    def num_groups(self):
        return int(_lib.orc_num_groups(_vp(self._h)))

    @property
    def num_pages(self):
        return int(_lib.orc_num_pages(_vp(self._h)))

    @property
    def num_nodes(self):
        return int(_lib.orc_num_nodes(_vp(self._h)))

    def get_sort(self):
        n = self.N
        sorter = np.empty(n, dtype=np.uint64); order = np.empty(n, dtype=np.int32); base = np.empty(n, dtype=np.uint64)
        self._check(_lib.orc_get_sort(_vp(self._h), _p(sorter), _p(order), _p(base)))
        return sorter, order, base

    def get_groups(self):
        g = self.num_groups
        first = np.empty(g, dtype=np.int32); last = np.empty(g, dtype=np.int32); blk = np.empty(g, dtype=np.uint64)
        self._check(_lib.orc_get_groups(_vp(self._h), _p(first), _p(last), _p(blk)))
        return first, last, blk

    def get_pages(self):
        out = np.empty(self.num_pages, dtype=np.uint64)
        self._check(_lib.orc_get_pages(_vp(self._h), _p(out)))
        return out

    def particlesToGrid(self):
        n = C.c_int(0)
        self._check(_lib.orc_p2g(_vp(self._h), C.byref(n)))
        return n.value

    def get_grid(self):
        gn = self.num_pages * self.ELEMENTS_PER_BLOCK
        idx = np.empty(gn, dtype=np.int64); m = np.empty(gn); v = np.empty((gn, 3))
        self._check(_lib.orc_get_grid(_vp(self._h), _p(idx), _p(m), _p(v)))
        return idx, m, v

    def get_id2coord(self):
        out = np.empty((self.num_nodes, 3), dtype=np.int32)
        self._check(_lib.orc_get_id2coord(_vp(self._h), _p(out)))
        return out

    def buildMassMatrix(self):
        out = np.empty(self.num_nodes)
        self._check(_lib.orc_get_mass_matrix(_vp(self._h), _p(out)))
        return out

    def set_dv(self, dv):
        dv = np.ascontiguousarray(dv, dtype=np.float64)
        self._check(_lib.orc_set_dv(_vp(self._h), _p(dv)))

    def gridToParticles(self, dt, want_flags=True):
        flags = (C.c_int * 2)(0, 0)
        self._check(_lib.orc_g2p(_vp(self._h), C.c_double(dt), flags))
        return (flags[0], flags[1])


# ---- force model (oracle_force.inl) ------------------------------------------------------------------------
def constitutive(F, mu, lam, project=True, dF=None):
    """single-particle CorotatedIsotropic evaluation: psi, P, dPdF (9x9, index i+3j), dP(dF), U, sigma, V"""
    F = np.ascontiguousarray(np.asarray(F, dtype=np.float64).reshape(3, 3).T)  # -> column-major buffer
    psi = C.c_double(0)
    P = np.empty(9); H = np.empty(81); dP = np.empty(9); U = np.empty(9); sg = np.empty(3); V = np.empty(9)
    dFb = None if dF is None else np.ascontiguousarray(np.asarray(dF, dtype=np.float64).reshape(3, 3).T)
    _lib.orc_constitutive(_p(F), C.c_double(mu), C.c_double(lam), int(project), C.byref(psi), _p(P), _p(H), _p(dFb),
                          _p(dP) if dF is not None else None, _p(U), _p(sg), _p(V))
    cm = lambda a: a.reshape(3, 3).T.copy()
    return dict(psi=psi.value, P=cm(P), dPdF=H.reshape(9, 9).T.copy(), dP=cm(dP) if dF is not None else None, U=cm(U), sigma=sg, V=cm(V))


def set_constitutive_model_global(model):
    """the oracle's constitutive model switch is process-wide (0 CorotatedIsotropic, 1 neo-Hookean extension)"""
    _lib.orc_set_constitutive_model(None, int(model))


def _add_force_methods(cls):
    def set_dt_gravity(self, dt, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        self._check(_lib.orc_set_dt_gravity(_vp(self._h), C.c_double(dt), _p(g)))
        self.dt = dt

    def set_project(self, project):
        self._check(_lib.orc_set_project(_vp(self._h), int(project)))

    def set_bc(self, node_id, P=None, R=None, Rinv=None, slip=None, dv_bc=None, mode=0):
        node_id = np.ascontiguousarray(node_id, dtype=np.int32)
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        P, R, Rinv, dv_bc = f(P), f(R), f(Rinv), f(dv_bc)
        slip = None if slip is None else np.ascontiguousarray(slip, dtype=np.int32)
        self._check(_lib.orc_set_bc(_vp(self._h), int(mode), len(node_id), _p(node_id), _p(P), _p(R), _p(Rinv), _p(slip), _p(dv_bc)))

    def get_dv(self):
        out = np.empty((self.num_nodes, 3))
        self._check(_lib.orc_get_dv(_vp(self._h), _p(out)))
        return out

    def backupStrain(self):
        self._check(_lib.orc_backup_strain(_vp(self._h)))

    def restoreStrain(self):
        self._check(_lib.orc_restore_strain(_vp(self._h)))

    def updateState(self, dv=None):
        e = C.c_double(0)
        dvb = None if dv is None else np.ascontiguousarray(dv, dtype=np.float64)
        self._check(_lib.orc_update_state(_vp(self._h), _p(dvb), C.byref(e)))
        return e.value

    def get_stress(self):
        S = np.empty((self.N, 9)); F = np.empty((self.N, 9))
        self._check(_lib.orc_get_stress(_vp(self._h), _p(S), _p(F)))
        return S, F

    def computeResidual(self):
        r = np.empty((self.num_nodes, 3))
        self._check(_lib.orc_compute_residual(_vp(self._h), _p(r)))
        return r

    def project(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64).copy()
        self._check(_lib.orc_project(_vp(self._h), _p(v)))
        return v

    def multiply(self, x):
        """matrix-free Hessian apply (ImplicitSolverObjective::multiply with --matfree)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        b = np.empty_like(x)
        self._check(_lib.orc_hessian_apply_mf(_vp(self._h), _p(x), _p(b)))
        return b

    def set_constitutive_model(self, model):
        m = {"corotated": 0, "fixed_corotated": 0, "neo_hookean": 1}.get(model, model)
        self._check(_lib.orc_set_constitutive_model(_vp(self._h), int(m)))

    def set_plasticity(self, model, params=()):
        m = {"none": 0, "von_mises": 1, "snow": 2, "drucker_prager": 3}.get(model, model)
        p = np.ascontiguousarray(list(params) + [0.0] * (5 - len(params)), dtype=np.float64)
        self._check(_lib.orc_set_plasticity(_vp(self._h), int(m), _p(p)))

    def applyPlasticity(self):
        self._check(_lib.orc_apply_plasticity(_vp(self._h)))

    def get_plastic_state(self):
        Jp = np.empty(self.N); mu = np.empty(self.N); lam = np.empty(self.N)
        self._check(_lib.orc_get_plastic_state(_vp(self._h), _p(Jp), _p(mu), _p(lam)))
        return Jp, mu, lam

    def evaluatePerNodeCNTolerance(self, eps, dt):
        tol = np.empty(self.num_nodes)
        self._check(_lib.orc_eval_cn_tolerance(_vp(self._h), C.c_double(eps), C.c_double(dt), _p(tol)))
        return tol

    for k, v in list(locals().items()):
        if callable(v) and k != "cls":
            setattr(cls, k, v)


_add_force_methods(OracleSim)


# ---- assembled matrix / multigrid (oracle_matrix.inl) --------------------------------------------------------
def _add_matrix_methods(cls):
    _ip = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)

    def buildMatrix(self, bcproject=True):
        self._check(_lib.orc_build_matrix(_vp(self._h), int(bcproject)))

    def buildDiagonal(self, Ainv=1):
        out = np.empty((self.num_nodes, 9))
        self._check(_lib.orc_build_diagonal(_vp(self._h), int(Ainv), _p(out)))
        return out

    def get_matrix(self):
        n = self.num_nodes
        col = np.empty((n, 125), dtype=np.int32); val = np.empty((n, 125, 9))
        self._check(_lib.orc_get_matrix(_vp(self._h), _ip(col), _p(val)))
        return col, val

    def buildMultigrid(self, levels=3, smoother=5, coarseSolver=2, Ainv=1, times=1, levelscale=0, topomega=0.1):
        self._check(_lib.orc_build_mg(_vp(self._h), levels, smoother, coarseSolver, Ainv, times, levelscale, C.c_double(topomega)))

    def estimate2norm(self, level):
        out = (C.c_double * 2)()
        self._check(_lib.orc_estimate_2norm(_vp(self._h), level, out))
        return float(out[0]), float(out[1])

    def estimate2norm_from(self, level, start):
        """the power iteration of estimate2norm from a given +-1 start vector; leaves lMax / lMin of the level set (the Chebyshev smoother reads them)"""
        start = np.ascontiguousarray(start, dtype=np.float64)
        out = (C.c_double * 2)()
        self._check(_lib.orc_estimate_2norm_from(_vp(self._h), level, _p(start), out))
        return float(out[0]), float(out[1])

    def level_dofs(self):
        L = _lib.orc_mg_levels(_vp(self._h))
        out = (C.c_int * L)()
        self._check(_lib.orc_get_level_dofs(_vp(self._h), out))
        return list(out)

    def level_coords(self, level):
        out = np.empty((self.level_dofs()[level], 3), dtype=np.int32)
        self._check(_lib.orc_get_level_coords(_vp(self._h), level, _ip(out)))
        return out

    def level_matrix(self, level, kind=0):
        cs = C.c_int(0)
        self._check(_lib.orc_get_level_matrix(_vp(self._h), level, kind, C.byref(cs), None, None))
        d = self.level_dofs()
        rows = d[level + 1] if kind == 2 else d[level]
        col = np.empty((rows, cs.value), dtype=np.int32); val = np.empty((rows, cs.value, 9))
        self._check(_lib.orc_get_level_matrix(_vp(self._h), level, kind, C.byref(cs), _ip(col), _p(val)))
        return col, val

    def level_diagonal(self, level):
        n = self.level_dofs()[level]
        D = np.empty((n, 9)); Di = np.empty((n, 9))
        self._check(_lib.orc_get_level_diagonal(_vp(self._h), level, _p(D), _p(Di)))
        return D, Di

    def color_order(self, level):
        out = np.empty((self.level_dofs()[level], 3), dtype=np.int32)
        self._check(_lib.orc_get_color_order(_vp(self._h), level, _ip(out)))
        return out

    def spmv(self, level, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        b = np.empty_like(x)
        self._check(_lib.orc_spmv(_vp(self._h), level, _p(x), _p(b)))
        return b

    def restrict(self, level, fine):
        fine = np.ascontiguousarray(fine, dtype=np.float64)
        out = np.empty((self.level_dofs()[level + 1], 3))
        self._check(_lib.orc_restrict(_vp(self._h), level, _p(fine), _p(out)))
        return out

    def prolong(self, level, coarse):
        coarse = np.ascontiguousarray(coarse, dtype=np.float64)
        out = np.empty((self.level_dofs()[level], 3))
        self._check(_lib.orc_prolong(_vp(self._h), level, _p(coarse), _p(out)))
        return out

    def smooth(self, level, kind, u, r, iterations, tolerance=0.0, initial_residual=None):
        u = np.ascontiguousarray(u, dtype=np.float64).copy(); r = np.ascontiguousarray(r, dtype=np.float64).copy()
        ir = None if initial_residual is None else np.ascontiguousarray(initial_residual, dtype=np.float64)
        self._check(_lib.orc_smooth(_vp(self._h), level, kind, _p(u), _p(r), iterations, C.c_double(tolerance), _p(ir)))
        return u, r

    def vcycle(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.empty_like(r)
        self._check(_lib.orc_vcycle(_vp(self._h), _p(r), _p(out)))
        return out

    def vcycle_timing(self):
        t = np.zeros((10, 4)); it = C.c_int(0)
        self._check(_lib.orc_vcycle_timing(_vp(self._h), _p(t), C.byref(it)))
        return t, it.value

    for k, v in list(locals().items()):
        if callable(v) and not k.startswith("_") and k != "cls":
            setattr(cls, k, v)


_add_matrix_methods(OracleSim)


# ---- solvers (oracle_solver.inl) -----------------------------------------------------------------------------
def _add_solver_methods(cls):
    from hot_b200._lib import SolverOptions, SolveLog   # plain ctypes mirrors of the public C structs

    def default_options(self, **kw):
        o = SolverOptions()
        _lib.orc_default_options(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        return o

    def pcg(self, b, x0=None, tolerance=1.0, max_iterations=10000, matfree=False, preconditioner=1):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        it = C.c_int(0)
        self._check(_lib.orc_pcg(_vp(self._h), _p(b), _p(x), C.c_double(tolerance), int(max_iterations), int(matfree),
                                 int(preconditioner), C.byref(it)))
        return x, it.value

    def obj_multiply(self, x, matfree=False):
        x = np.ascontiguousarray(x, dtype=np.float64); b = np.empty_like(x)
        self._check(_lib.orc_obj_multiply(_vp(self._h), int(matfree), _p(x), _p(b)))
        return b

    def obj_precondition(self, r, matfree=False, preconditioner=1):
        r = np.ascontiguousarray(r, dtype=np.float64); z = np.empty_like(r)
        self._check(_lib.orc_obj_precondition(_vp(self._h), int(matfree), int(preconditioner), _p(r), _p(z)))
        return z

    def minres(self, b, x0=None, relative_tolerance=1.0, tolerance=1.0, max_iterations=10000, matfree=False, preconditioner=1):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        it = C.c_int(0)
        self._check(_lib.orc_minres(_vp(self._h), _p(b), _p(x), C.c_double(relative_tolerance), C.c_double(tolerance), int(max_iterations),
                                    int(matfree), int(preconditioner), C.byref(it)))
        return x, it.value

    def backwardEulerStep(self, options=None, **kw):
        o = options if options is not None else self.default_options(**kw)
        log = SolveLog()
        self._check(_lib.orc_backward_euler_step(_vp(self._h), C.byref(o), C.byref(log)))
        return log.as_dict()

    def backwardEulerStepReferenceLBFGS(self, options=None, **kw):
        """orc_backward_euler_step with the L-BFGS loop replaced by the reference's own ZIRAN::LBFGS::solve (oracle/lbfgs_ref_shim.cpp);
        only in the library built from the reference (HOT_ORACLE_LIB=oracle/_ref/libhot_oracle_lbfgsref.so)"""
        o = options if options is not None else self.default_options(**kw)
        log = SolveLog()
        self._check(_lib.zr_lbfgs_backward_euler_step(_vp(self._h), C.byref(o), C.byref(log)))
        return log.as_dict()

    def get_dv0(self):
        out = np.empty((self.num_nodes, 3))
        self._check(_lib.orc_get_dv0(_vp(self._h), _p(out)))
        return out

    for k, v in list(locals().items()):
        if callable(v) and not k.startswith("_") and k not in ("cls", "SolverOptions", "SolveLog"):
            setattr(cls, k, v)


_add_solver_methods(OracleSim)
