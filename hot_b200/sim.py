"""Host-side mirror of the reference's MpmSimulationBase operator surface for the hot path.

Method names follow Lib/MPM/MpmSimulationBase.h:138-220 (sortParticlesAndPolluteGrid, particlesToGrid,
gridToParticles ...) so that the parity tests read like calls into the reference.  All compute happens in
libhot_b200.so on the GPU; numpy arrays are only the caller-owned host buffers of the C ABI.
"""
import ctypes as C
import numpy as np

from ._lib import load_library, SolverOptions, SolveLog, Transport, T_ALL_REDUCE, T_ALL_GATHER, T_NEIGHBOR_EXCHANGE


class HotError(RuntimeError):
    """Raised where the reference throws std::runtime_error from ZIRAN_ASSERT (Lib/Ziran/CS/Util/Debug.h:19-42)."""


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class MpmSimulationB200:
    ELEMENTS_PER_BLOCK = 32  # GridState<double,3>: 2x4x4 nodes per 4 KB page

    def __init__(self, dx, apic_rpic_ratio=1.0, cfl=0.6, device=-1, stream=None):
        self._lib = load_library()
        self._h = self._lib.hot_create(float(dx), float(apic_rpic_ratio), float(cfl), int(device))
        if not self._h:
            raise HotError("hot_create failed: no usable CUDA device (hot_b200 has no CPU fallback)")
        self.dx = float(dx)
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hot_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise HotError(self._lib.hot_last_error(self._h).decode())

    # ---- plumbing
    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.hot_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    def synchronize(self):
        self._check(self._lib.hot_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._lib.hot_launch_count(self._h))

    def timing(self, enable=1):
        self._check(self._lib.hot_timing(self._h, int(enable)))

    def get_timings(self):
        """{kernel class: (total ms, launches)} measured with CUDA events on the handle's stream"""
        ms = (C.c_double * 64)(); cnt = (C.c_longlong * 64)()
        k = self._lib.hot_get_timings(self._h, 64, ms, cnt)
        return {self._lib.hot_timing_name(i).decode(): (ms[i], int(cnt[i])) for i in range(k) if cnt[i]}

    # ---- SPGrid addressing
    def linear_offset(self, ijk, fp32=False):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        out = np.empty(len(ijk), dtype=np.uint64)
        self._check((self._lib.hot_linear_offset_f32 if fp32 else self._lib.hot_linear_offset)(self._h, len(ijk), ijk.ctypes.data_as(C.POINTER(C.c_int)),
                                                out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return out

    def linear_to_coord(self, off, fp32=False):
        off = np.ascontiguousarray(off, dtype=np.uint64)
        out = np.empty((len(off), 3), dtype=np.int32)
        self._check((self._lib.hot_linear_to_coord_f32 if fp32 else self._lib.hot_linear_to_coord)(self._h, len(off), off.ctypes.data_as(C.POINTER(C.c_ulonglong)),
                                                  out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def packed_add(self, a, b, fp32=False):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.empty(len(a), dtype=np.uint64)
        p = C.POINTER(C.c_ulonglong)
        self._check((self._lib.hot_packed_add_f32 if fp32 else self._lib.hot_packed_add)(self._h, len(a), a.ctypes.data_as(p), b.ctypes.data_as(p),
                                                                                          out.ctypes.data_as(p)))
        return out

    # ---- particles
    def set_particles(self, X, V, mass, C_, F, vol, mu, lam):
        n = len(mass)
        arrs = [_f64(X, (n, 3)), _f64(V, (n, 3)), _f64(mass, (n,)), _f64(C_, (n, 9)), _f64(F, (n, 9)), _f64(vol, (n,)),
                _f64(mu, (n,)), _f64(lam, (n,))]
        self._check(self._lib.hot_set_particles(self._h, n, *[_ptr(a) for a in arrs]))
        self.N = n

    def set_particles_ptr(self, n, ptrs):
        """Raw-pointer variant for pinned host buffers (bench.py e2e leg)."""
        self._check(self._lib.hot_set_particles(self._h, n, *[C.c_void_p(int(p)) for p in ptrs]))
        self.N = n

    # pipelined state exchange (hot_upload_state_async ... hot_wait_download): raw pointers of pinned host buffers X, V, C, F
    def upload_state_async(self, ptrs):
        self._check(self._lib.hot_upload_state_async(self._h, *[C.c_void_p(int(p)) for p in ptrs]))

    def commit_state(self):
        self._check(self._lib.hot_commit_state(self._h))

    def download_state_async(self, ptrs):
        self._check(self._lib.hot_download_state_async(self._h, *[C.c_void_p(int(p)) for p in ptrs]))

    def wait_download(self):
        self._check(self._lib.hot_wait_download(self._h))

    def get_particles_ptr(self, ptrs):
        self._check(self._lib.hot_get_particles(self._h, *[None if p is None else C.c_void_p(int(p)) for p in ptrs]))

    def get_particles(self, gradV=True):
        n = self.N
        X = np.empty((n, 3)); V = np.empty((n, 3)); Cm = np.empty((n, 9)); F = np.empty((n, 9))
        G = np.empty((n, 9)) if gradV else None
        self._check(self._lib.hot_get_particles(self._h, _ptr(X), _ptr(V), _ptr(Cm), _ptr(F), _ptr(G)))
        return dict(X=X, V=V, C=Cm, F=F, gradV=G)

    # ---- a5
    def sortParticlesAndPolluteGrid(self):
        self._check(self._lib.hot_sort_and_activate(self._h))

    @property
    def num_groups(self):
        return int(self._lib.hot_num_groups(self._h))

    @property
    def num_pages(self):
        return int(self._lib.hot_num_pages(self._h))

    def get_sort(self):
        n = self.N
        sorter = np.empty(n, dtype=np.uint64); order = np.empty(n, dtype=np.int32); base = np.empty(n, dtype=np.uint64)
        self._check(self._lib.hot_get_sort(self._h, _ptr(sorter), _ptr(order), _ptr(base)))
        return sorter, order, base

    def get_groups(self):
        g = self.num_groups
        first = np.empty(g, dtype=np.int32); last = np.empty(g, dtype=np.int32); blk = np.empty(g, dtype=np.uint64)
        self._check(self._lib.hot_get_groups(self._h, _ptr(first), _ptr(last), _ptr(blk)))
        return first, last, blk

    def get_pages(self):
        out = np.empty(self.num_pages, dtype=np.uint64)
        self._check(self._lib.hot_get_pages(self._h, _ptr(out)))
        return out

    # ---- a6 / a7
    def particlesToGrid(self):
        n = C.c_int(0)
        self._check(self._lib.hot_p2g(self._h, C.byref(n)))
        return n.value

    @property
    def num_nodes(self):
        return int(self._lib.hot_num_nodes(self._h))

    def get_grid(self):
        gn = self.num_pages * self.ELEMENTS_PER_BLOCK
        idx = np.empty(gn, dtype=np.int64); m = np.empty(gn); v = np.empty((gn, 3))
        self._check(self._lib.hot_get_grid(self._h, _ptr(idx), _ptr(m), _ptr(v)))
        return idx, m, v

    def get_id2coord(self):
        out = np.empty((self.num_nodes, 3), dtype=np.int32)
        self._check(self._lib.hot_get_id2coord(self._h, _ptr(out)))
        return out

    def buildMassMatrix(self):
        out = np.empty(self.num_nodes)
        self._check(self._lib.hot_get_mass_matrix(self._h, _ptr(out)))
        return out

    # ---- a23
    def set_dv(self, dv):
        dv = _f64(dv, (self.num_nodes, 3))
        self._check(self._lib.hot_set_dv(self._h, _ptr(dv)))

    def gridToParticles(self, dt, want_flags=True):
        flags = (C.c_int * 2)(0, 0)
        self._check(self._lib.hot_g2p(self._h, float(dt), flags if want_flags else None))
        return (flags[0], flags[1])

    # ---- force model / objective (ImplicitSolverObjective surface, Projects/multigrid/ImplicitSolver.h)
    def set_dt_gravity(self, dt, g):
        g = _f64(g, (3,))
        self._check(self._lib.hot_set_dt_gravity(self._h, float(dt), _ptr(g)))
        self.dt = float(dt)

    def set_project(self, project):
        self._check(self._lib.hot_set_project(self._h, int(project)))

    def set_bc(self, node_id, P=None, R=None, Rinv=None, slip=None, dv_bc=None, mode=0):
        node_id = np.ascontiguousarray(node_id, dtype=np.int32)
        f = lambda a: None if a is None else _f64(a)
        P, R, Rinv, dv_bc = f(P), f(R), f(Rinv), f(dv_bc)
        slip = None if slip is None else np.ascontiguousarray(slip, dtype=np.int32)
        self._check(self._lib.hot_set_bc(self._h, int(mode), len(node_id), _ptr(node_id), _ptr(P), _ptr(R), _ptr(Rinv),
                                         _ptr(slip), _ptr(dv_bc)))

    def get_dv(self):
        out = np.empty((self.num_nodes, 3))
        self._check(self._lib.hot_get_dv(self._h, _ptr(out)))
        return out

    def corotated_eval(self, F, mu, lam, project=True, dF=None, single=False):
        """CorotatedIsotropic<T,3> on deformation gradients F (n x 3 x 3, numpy row-major): psi, P, dP(dF), dense dPdF, U, sigma, V"""
        F = np.ascontiguousarray(np.asarray(F, dtype=np.float64).reshape(-1, 3, 3).transpose(0, 2, 1))   # column-major per item
        n = len(F)
        dFb = None if dF is None else np.ascontiguousarray(np.asarray(dF, dtype=np.float64).reshape(-1, 3, 3).transpose(0, 2, 1))
        psi = np.empty(n); P = np.empty((n, 9)); dP = np.empty((n, 9)); H = np.empty((n, 81)); U = np.empty((n, 9)); sg = np.empty((n, 3)); V = np.empty((n, 9))
        self._check(self._lib.hot_corotated_eval(self._h, n, _ptr(F), float(mu), float(lam), int(project), _ptr(dFb), _ptr(psi), _ptr(P),
                                                 _ptr(dP) if dF is not None else None, _ptr(H), _ptr(U), _ptr(sg), _ptr(V)))
        t = lambda a: a.reshape(n, 3, 3).transpose(0, 2, 1).copy()
        out = dict(psi=psi, P=t(P), dP=t(dP) if dF is not None else None, dPdF=H.reshape(n, 9, 9).transpose(0, 2, 1).copy(), U=t(U), sigma=sg, V=t(V))
        if single:
            out = {k: (v[0] if v is not None else None) for k, v in out.items()}
        return out

    def backupStrain(self):
        self._check(self._lib.hot_backup_strain(self._h))

    def restoreStrain(self):
        self._check(self._lib.hot_restore_strain(self._h))

    def updateState(self, dv=None, want_energy=True):
        e = C.c_double(0)
        dvb = None if dv is None else _f64(dv, (self.num_nodes, 3))
        self._check(self._lib.hot_update_state(self._h, _ptr(dvb), C.byref(e) if want_energy else None))
        return e.value

    def get_stress(self):
        S = np.empty((self.N, 9)); F = np.empty((self.N, 9))
        self._check(self._lib.hot_get_stress(self._h, _ptr(S), _ptr(F)))
        return S, F

    def computeResidual(self):
        r = np.empty((self.num_nodes, 3))
        self._check(self._lib.hot_compute_residual(self._h, _ptr(r)))
        return r

    def project(self, v):
        v = _f64(v, (self.num_nodes, 3)).copy()
        self._check(self._lib.hot_project(self._h, _ptr(v)))
        return v

    def multiply(self, x):
        """matrix-free Hessian apply (ImplicitSolverObjective::multiply with --matfree)"""
        x = _f64(x, (self.num_nodes, 3))
        b = np.empty_like(x)
        self._check(self._lib.hot_hessian_apply_mf(self._h, _ptr(x), _ptr(b)))
        return b

    def evaluatePerNodeCNTolerance(self, eps, dt):
        tol = np.empty(self.num_nodes)
        self._check(self._lib.hot_eval_cn_tolerance(self._h, float(eps), float(dt), _ptr(tol)))
        return tol

    # ---- assembled system / Galerkin multigrid (ImplicitSolver.h:470-603, MultigridPreconditioner.h)
    def buildMatrix(self, bcproject=True):
        self._check(self._lib.hot_build_matrix(self._h, int(bcproject)))

    def get_matrix(self):
        n = self.num_nodes
        col = np.empty((n, 125), dtype=np.int32); val = np.empty((n, 125, 9))
        self._check(self._lib.hot_get_matrix(self._h, _ptr(col), _ptr(val)))
        return col, val

    def buildDiagonal(self, Ainv=1):
        out = np.empty((self.num_nodes, 9))
        self._check(self._lib.hot_build_diagonal(self._h, int(Ainv), _ptr(out)))
        return out

    def buildMultigrid(self, levels=3, smoother=5, coarseSolver=2, Ainv=1, times=1, levelscale=0, topomega=0.1):
        self._check(self._lib.hot_build_mg(self._h, levels, smoother, coarseSolver, Ainv, times, levelscale, float(topomega)))

    def estimate2norm(self, level):
        """(lMax, lMin) of A_level: SquareMatrix::estimate2norm"""
        out = (C.c_double * 2)()
        self._check(self._lib.hot_estimate_2norm(self._h, level, out))
        return float(out[0]), float(out[1])

    def level_dofs(self):
        L = self._lib.hot_mg_levels(self._h)
        out = (C.c_int * max(L, 1))()
        self._check(self._lib.hot_get_level_dofs(self._h, out))
        return list(out)[:L]

    def level_coords(self, level):
        out = np.empty((self.level_dofs()[level], 3), dtype=np.int32)
        self._check(self._lib.hot_get_level_coords(self._h, level, _ptr(out)))
        return out

    def level_matrix(self, level, kind=0):
        """kind 0: (col, val[n,125,9]); kind 1 / 2: (col, scalar weights) of P_l / R_l"""
        cs = C.c_int(0)
        self._check(self._lib.hot_get_level_matrix(self._h, level, kind, C.byref(cs), None, None))
        d = self.level_dofs()
        rows = d[level + 1] if kind == 2 else d[level]
        col = np.empty((rows, cs.value), dtype=np.int32)
        val = np.empty((rows, cs.value, 9)) if kind == 0 else np.empty((rows, cs.value))
        self._check(self._lib.hot_get_level_matrix(self._h, level, kind, C.byref(cs), _ptr(col), _ptr(val)))
        return col, val

    def level_diagonal(self, level):
        n = self.level_dofs()[level]
        D = np.empty((n, 9)); Di = np.empty((n, 9))
        self._check(self._lib.hot_get_level_diagonal(self._h, level, _ptr(D), _ptr(Di)))
        return D, Di

    def gs_schedule(self, level):
        nb = C.c_int(0); cfb = (C.c_int * 9)()
        self._check(self._lib.hot_get_gs_schedule(self._h, level, C.byref(nb), cfb, None, None))
        seq = np.empty(self.level_dofs()[level], dtype=np.int32); start = np.empty(nb.value + 1, dtype=np.int32)
        self._check(self._lib.hot_get_gs_schedule(self._h, level, C.byref(nb), cfb, _ptr(seq), _ptr(start)))
        return seq, start, list(cfb)

    def color_order(self, level):
        """(colour, block id within the colour, 1-based position in the block) per node = SquareMatrix::colorOrder"""
        seq, start, cfb = self.gs_schedule(level)
        out = np.empty((len(seq), 3), dtype=np.int32)
        for c in range(8):
            for b in range(cfb[c], cfb[c + 1]):
                nodes = seq[start[b]:start[b + 1]]
                out[nodes, 0] = c; out[nodes, 1] = b - cfb[c]; out[nodes, 2] = np.arange(1, len(nodes) + 1)
        return out

    def _dofvec(self, level, a):
        return _f64(a, (self.level_dofs()[level] if self._lib.hot_mg_levels(self._h) else self.num_nodes, 3))

    def spmv(self, level, x):
        x = self._dofvec(level, x)
        b = np.empty_like(x)
        self._check(self._lib.hot_spmv(self._h, level, _ptr(x), _ptr(b)))
        return b

    def restrict(self, level, fine):
        fine = self._dofvec(level, fine)
        out = np.empty((self.level_dofs()[level + 1], 3))
        self._check(self._lib.hot_restrict(self._h, level, _ptr(fine), _ptr(out)))
        return out

    def prolong(self, level, coarse):
        coarse = self._dofvec(level + 1, coarse)
        out = np.empty((self.level_dofs()[level], 3))
        self._check(self._lib.hot_prolong(self._h, level, _ptr(coarse), _ptr(out)))
        return out

    def smooth(self, level, kind, u, r, iterations, tolerance=0.0, initial_residual=None):
        u = self._dofvec(level, u).copy(); r = self._dofvec(level, r).copy()
        ir = None if initial_residual is None else self._dofvec(level, initial_residual)
        self._check(self._lib.hot_smooth(self._h, level, kind, _ptr(u), _ptr(r), iterations, float(tolerance), _ptr(ir)))
        return u, r

    def vcycle(self, r):
        r = self._dofvec(0, r)
        out = np.empty_like(r)
        self._check(self._lib.hot_vcycle(self._h, _ptr(r), _ptr(out)))
        return out

    def vcycle_timing(self):
        t = np.zeros((10, 4)); it = C.c_int(0)
        self._check(self._lib.hot_vcycle_timing(self._h, _ptr(t), C.byref(it)))
        return t, it.value

    def vcycle_bench(self, reps):
        ms = C.c_double(0)
        self._check(self._lib.hot_vcycle_bench(self._h, int(reps), C.byref(ms)))
        return ms.value / reps

    # ---- solvers (InexactConjugateGradient.h, LBFGS.h, ExtendedNewtonsMethod.h, MultigridSimulation.h:188-233)
    def default_options(self, **kw):
        o = SolverOptions()
        self._lib.hot_default_options(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        return o

    def pcg(self, b, x0=None, tolerance=1.0, max_iterations=10000, matfree=False, preconditioner=1):
        b = _f64(b, (self.num_nodes, 3))
        x = np.zeros_like(b) if x0 is None else _f64(x0, (self.num_nodes, 3)).copy()
        it = C.c_int(0)
        self._check(self._lib.hot_pcg(self._h, _ptr(b), _ptr(x), float(tolerance), int(max_iterations), int(matfree),
                                      int(preconditioner), C.byref(it)))
        return x, it.value

    def backwardEulerStep(self, options=None, **kw):
        o = options if options is not None else self.default_options(**kw)
        log = SolveLog()
        self._check(self._lib.hot_backward_euler_step(self._h, C.byref(o), C.byref(log)))
        return log.as_dict()

    def get_dv0(self):
        out = np.empty((self.num_nodes, 3))
        self._check(self._lib.hot_get_dv0(self._h, _ptr(out)))
        return out

    OPS = {"hessian_apply": 0, "spmv": 1, "update_state": 2, "residual": 3, "smooth": 4, "build_matrix": 5, "build_mg": 6, "coarse_solve": 7}

    def op_bench(self, op, reps=10, level=0):
        """milliseconds per application of one device-resident operator (CUDA events)"""
        ms = C.c_double(0)
        self._check(self._lib.hot_op_bench(self._h, self.OPS[op], int(level), int(reps), C.byref(ms)))
        return ms.value / reps

    def level_nnz_blocks(self, level):
        n = C.c_longlong(0)
        self._check(self._lib.hot_level_nnz_blocks(self._h, int(level), C.byref(n)))
        return n.value

    def addScaledForces(self, scale, f):
        f = _f64(f, (self.num_nodes, 3)).copy()
        self._check(self._lib.hot_add_scaled_forces(self._h, float(scale), _ptr(f)))
        return f

    def addScaledForceDifferentials(self, scale, x, f):
        x = _f64(x, (self.num_nodes, 3)); f = _f64(f, (self.num_nodes, 3)).copy()
        self._check(self._lib.hot_add_scaled_force_differentials(self._h, float(scale), _ptr(x), _ptr(f)))
        return f

    # ---- one object over several GPUs (include/hot_b200.h "one object over the GPUs of a box")
    def init_nccl(self, rank, world, unique_id):
        """NCCL inside the library: `unique_id` = the 128 bytes of hot_b200.comm_unique_id() from rank 0, distributed by the caller"""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.hot_comm_init_nccl(self._h, int(rank), int(world), buf))

    def set_partition(self, rank, world, all_reduce=None, all_gather=None, neighbor_exchange=None):
        """the caller's collectives, on HOST numpy arrays (this wrapper stages the device buffers with hot_memcpy_d2h / h2d):
        all_reduce(array f64, op 0 sum / 1 max) -> array; all_gather(array u8) -> array of world x len;
        neighbor_exchange(peers, [array f64 per peer]) -> [array f64 per peer] (same lengths)."""
        lib, h = self._lib, self._h

        def d2h(ptr, nbytes, dtype):
            a = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
            if nbytes:
                self._check(lib.hot_memcpy_d2h(h, a.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), nbytes))
            return a

        def h2d(ptr, a):
            a = np.ascontiguousarray(a)
            if a.nbytes:
                self._check(lib.hot_memcpy_h2d(h, C.c_void_p(ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))

        def guard(f):
            def g(*args):
                try:
                    f(*args)
                    return 0
                except Exception:  # pragma: no cover - surfaces as HotError through the return code
                    import traceback; traceback.print_exc()
                    return 1
            return g

        @guard
        def cb_all_reduce(user, dev, count, op):
            h2d(dev, np.asarray(all_reduce(d2h(dev, 8 * count, np.float64), int(op)), dtype=np.float64))

        @guard
        def cb_all_gather(user, send, recv, nbytes):
            h2d(recv, np.asarray(all_gather(d2h(send, nbytes, np.uint8)), dtype=np.uint8).reshape(-1))

        @guard
        def cb_exchange(user, n_peers, peers, send, recv, count):
            ps = [int(peers[j]) for j in range(n_peers)]
            out = neighbor_exchange(ps, [d2h(send[j], 8 * count[j], np.float64) for j in range(n_peers)])
            for j in range(n_peers):
                h2d(recv[j], np.asarray(out[j], dtype=np.float64))

        if world > 1:
            self._transport = Transport(None, T_ALL_REDUCE(cb_all_reduce), T_ALL_GATHER(cb_all_gather), T_NEIGHBOR_EXCHANGE(cb_exchange))
            self._check(lib.hot_set_partition(h, int(rank), int(world), C.byref(self._transport)))
        else:
            self._check(lib.hot_set_partition(h, int(rank), int(world), None))

    def get_partition(self):
        out = (C.c_long * 8)()
        self._check(self._lib.hot_get_partition(self._h, out))
        return dict(zip(("rank", "world", "neighbors", "shared_pages", "exchange_pages", "owned_nodes", "global_nodes", "particles"), [int(v) for v in out]))

    def set_constitutive_model(self, model):
        m = {"corotated": 0, "fixed_corotated": 0, "neo_hookean": 1}.get(model, model)
        self._check(self._lib.hot_set_constitutive_model(self._h, int(m)))

    def set_ghost_ring(self, on=True):
        """partitioned objects: hold the 27-neighbourhood of the shared pages too (needed by buildMatrix / buildMultigrid / vcycle and
        the solvers with an assembled matrix when world > 1); takes effect with the next sortParticlesAndPolluteGrid"""
        self._check(self._lib.hot_set_ghost_ring(self._h, 1 if on else 0))

    def get_transport(self):
        return {0: "single rank", 1: "grouped ncclSend/ncclRecv", 2: "caller callbacks", 3: "peer memory (NVLink P2P stores + flags)"}[int(self._lib.hot_get_transport(self._h))]

    # ---- plasticity (PlasticityApplier.cpp): applied by gridToParticles after evolveStrain
    def set_plasticity(self, model, params=()):
        m = {"none": 0, "von_mises": 1, "snow": 2, "drucker_prager": 3}.get(model, model)
        p = _f64(list(params) + [0.0] * (5 - len(params)), (5,))
        self._check(self._lib.hot_set_plasticity(self._h, int(m), _ptr(p)))

    def applyPlasticity(self):
        self._check(self._lib.hot_apply_plasticity(self._h))

    def get_plastic_state(self):
        n = self.N
        Jp = np.empty(n); mu = np.empty(n); lam = np.empty(n)
        self._check(self._lib.hot_get_plastic_state(self._h, _ptr(Jp), _ptr(mu), _ptr(lam)))
        return Jp, mu, lam
