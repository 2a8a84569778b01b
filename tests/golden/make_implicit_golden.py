"""Generates tests/golden/implicit_ref.npz: the REFERENCE'S OWN implicit-solver objective, ZIRAN::ImplicitSolverObjective<Simulation>
(Projects/multigrid/ImplicitSolver.h: updateState / totalEnergy, computeResidual, buildMatrix<projectSystem>, buildDiagonal,
evaluatePerNodeCNTolerance, multiply, shouldExitByCN), compiled where it lies and instantiated on a stand-in simulation whose grid is the
reference's MpmGrid / SPGrid code and whose constitutive model is the reference's CorotatedIsotropic (oracle/implicit_ref_shim.cpp ->
oracle/_ref/libimplicit_ref.so; the particle loops of MpmForceBase / FBasedMpmForceHelper / MassLumpedInertia that sit between them are written
out in that file).  tests/test_oracle_implicit_ref.py compares the oracle's restatement of rows a9-a15 (oracle_force.inl, oracle_matrix.inl) and
the CUDA path with these results.  Inputs are stored with the outputs.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_implicit_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libimplicit_ref.so")
OUT = os.path.join(ROOT, "tests", "golden", "implicit_ref.npz")
GRAVITY = (0.0, -9.8, 0.0)
CN_EPS = 1e-3
# case -> scene arguments, dt, BC mode (0: P projection, all sticky; 1: BC-projected system with slip nodes), makePD on / off
CASES = {
    "sticky": (dict(cells=(5, 4, 4), dx=1.0 / 32, ppc=4, seed=3, E=2e5), 6e-3, 0, True),
    "slip": (dict(cells=(4, 3, 3), dx=1.0 / 32, ppc=5, seed=4, E=5e4), 4e-3, 1, True),
    "slip_noproject": (dict(cells=(4, 5, 3), dx=1.0 / 64, ppc=6, seed=7, E=1e5, origin_cells=(29, 14, 60)), 2e-3, 1, False),
}


# whole implicit solves (MultigridSimulation::backwardEulerStep): scene arguments of tests/golden/make_lbfgs_golden.scene, solver options.
# The first three are the substeps of lbfgs_ref.npz (there: the reference's L-BFGS loop on the ORACLE'S objective; here: on the reference's own).
HOT = dict(lsolver=3, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, bcproject=1, usecn=1)   # tog.sh:38
STEPS = {
    "hot_default": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT)),
    "hot_no_linesearch": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, linesearch=0)),
    "hot_stiff_long": (dict(cells=(7, 9, 6), E=2e6, dt=8e-3, seed=1), dict(HOT, cneps=1e-9)),
    "pn_mgpcg": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=2, max_newton_iterations=10)),
    "pn_pcg": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=2, mg_level=1, max_newton_iterations=10)),
    "pn_pcg_no_cn": (dict(cells=(6, 5, 6), E=3e5, dt=5e-3, seed=2), dict(HOT, lsolver=2, mg_level=1, usecn=0, cneps=1e-6, linesearch=0, max_newton_iterations=10)),
    "pn_pcg_matfree": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=2, matfree=1, mg_level=1, bcproject=0, max_newton_iterations=10)),
    "pn_pcg_entry_diagonal": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=2, mg_level=1, Ainv=0, max_newton_iterations=10)),
    "pn_mgpcg_jacobi_levels": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=2, mg_level=2, smoother=1, coarse_solver=0, max_newton_iterations=10)),
    "hot_jacobi_smoother": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, smoother=0)),
    "hot_adaptive_hessian": (dict(cells=(7, 9, 6), E=2e6, dt=8e-3, seed=1), dict(HOT, cneps=1e-9, adaptive_h=1)),
    "hot_times3_jacobi": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, mg_times=3, smoother=0)),
    "hot_levelscale2_times2": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, mg_scale=2, mg_times=2)),
    "hot_optimal_jacobi_omega": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, smoother=1, coarse_solver=1, topomega=0.5)),
    "hot_pcg_smoother": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, smoother=2)),
    "lbfgs_h": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, bcproject=0, mg_level=1, mg_times=10000, smoother=2, coarse_solver=2)),
    "hot_four_levels_no_pd": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, mg_level=4, project=0)),
    "hot_two_materials": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0, stiff_layer=(25.0, 0.4)), dict(HOT)),     # tolerance from the stiffest dP/dF
    "pn_mgpcg_two_materials": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0, stiff_layer=(25.0, 0.4)), dict(HOT, lsolver=2, max_newton_iterations=10)),
    # -lsolver 1 (f4): Newton + MINRES; its absolute tolerance is maxcntol with --usecn (MultigridSimulation.h:206), the constructor's 1 without
    "pn_mgminres": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=1, max_newton_iterations=10)),
    "pn_minres": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=1, mg_level=1, max_newton_iterations=10)),
    "pn_minres_matfree": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=1, mg_level=1, matfree=1, bcproject=0, max_newton_iterations=10)),
    "pn_minres_no_cn": (dict(cells=(6, 6, 6), E=4e5, dt=5e-3, seed=0), dict(HOT, lsolver=1, mg_level=1, usecn=0, cneps=1e-6, linesearch=0, max_newton_iterations=10)),
}
FULL_MATRIX = ("slip",)        # cases whose block rows are stored entry by entry (the others: columns, block row sums, action on a vector)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def rotation(normal):
    """rotation that takes `normal` to the x axis (the frame slip nodes are solved in: first component = normal component)"""
    n = np.asarray(normal, dtype=np.float64); n = n / np.linalg.norm(n)
    a = np.array([0.0, 0.0, 1.0]) if abs(n[2]) < 0.9 else np.array([0.0, 1.0, 0.0])
    t1 = np.cross(n, a); t1 /= np.linalg.norm(t1)
    t2 = np.cross(n, t1)
    return np.stack([n, t1, t2])          # rows: R @ n = e_x


def make_inputs(name):
    from hot_b200 import scenes
    kw, dt, mode, project = CASES[name]
    sc = scenes.block(**kw)
    rng = np.random.default_rng(17)
    inp = {k: np.ascontiguousarray(sc[k]) for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")}
    inp["F"] = inp["F"] + 0.08 * (rng.random(inp["F"].shape) - 0.5)     # strained start of the step (some particles get a clamped Hessian)
    if name == "slip_noproject":                                       # two materials: every third particle 20 x stiffer
        inp["mu"] = inp["mu"].copy(); inp["lam"] = inp["lam"].copy()
        inp["mu"][::3] *= 20.0; inp["lam"][::3] *= 20.0
    return inp, sc["dx"], dt, mode, project


def make_bc(coord, mode):
    """floor nodes sticky; in mode 1 the nodes of one side wall slip along a tilted plane"""
    ymin, xmax = coord[:, 1].min(), coord[:, 0].max()
    sticky = np.nonzero(coord[:, 1] <= ymin + 1)[0]
    node = sticky; slip = np.zeros(len(sticky), dtype=np.int32)
    n = len(node)
    I = np.tile(np.eye(3).reshape(1, 9), (n, 1))
    P = np.zeros((n, 9)); R = I.copy(); Rinv = I.copy()
    if mode == 1:
        wall = np.nonzero((coord[:, 0] >= xmax - 1) & (coord[:, 1] > ymin + 1))[0]
        Rm = rotation((1.0, 0.2, -0.1))
        node = np.concatenate([sticky, wall]); slip = np.concatenate([slip, np.ones(len(wall), dtype=np.int32)])
        Rw = np.tile(Rm.T.reshape(1, 9), (len(wall), 1))          # column-major buffers
        Rwi = np.tile(Rm.reshape(1, 9), (len(wall), 1))           # R^-1 = R^T
        nrm = Rm[0]
        Pw = np.tile((np.eye(3) - np.outer(nrm, nrm)).T.reshape(1, 9), (len(wall), 1))
        P = np.concatenate([P, Pw]); R = np.concatenate([R, Rw]); Rinv = np.concatenate([Rinv, Rwi])
    return node.astype(np.int32), P, R, Rinv, slip.astype(np.int32)


class Reference:
    """ImplicitSolverObjective<stand-in simulation> (oracle/implicit_ref_shim.cpp)"""
    def __init__(self, dx, dt, gravity=GRAVITY):
        self.lib = C.CDLL(REF_LIB)
        self.lib.implicit_ref_create.restype = C.c_void_p
        self.lib.implicit_ref_create.argtypes = [C.c_double, C.c_double, C.c_void_p]
        self.lib.implicit_ref_update_state.restype = C.c_double
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        self.h = C.c_void_p(self.lib.implicit_ref_create(float(dx), float(dt), _p(g)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.implicit_ref_destroy(self.h)
            self.h = None

    def setup(self, inp, project):
        a = [np.ascontiguousarray(inp[k], dtype=np.float64) for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")]
        self.N = len(a[2])
        self.num_nodes = int(self.lib.implicit_ref_setup(self.h, C.c_long(self.N), *[_p(x) for x in a], int(project)))
        return self.num_nodes

    def set_bc(self, mode, node, P, R, Rinv, slip):
        self.lib.implicit_ref_set_bc(self.h, int(mode), len(node), _p(node), _p(P), _p(R), _p(Rinv), _p(slip))

    def set_dv(self, dv):
        dv = np.ascontiguousarray(dv, dtype=np.float64)
        self.lib.implicit_ref_set_dv(self.h, _p(dv))

    def updateState(self, dv=None):
        d = None if dv is None else np.ascontiguousarray(dv, dtype=np.float64)
        return float(self.lib.implicit_ref_update_state(self.h, _p(d), 1))

    def get_F(self):
        F = np.empty((self.N, 9)); self.lib.implicit_ref_get_F(self.h, _p(F)); return F

    def computeResidual(self):
        r = np.empty((self.num_nodes, 3)); self.lib.implicit_ref_compute_residual(self.h, _p(r)); return r

    def buildMatrix(self, bcproject):
        col = np.empty((self.num_nodes, 125), dtype=np.int32); val = np.empty((self.num_nodes, 125, 9))
        self.lib.implicit_ref_build_matrix(self.h, int(bcproject), _p(col), _p(val))
        return col, val

    def buildDiagonal(self, Ainv):
        out = np.empty((self.num_nodes, 9)); self.lib.implicit_ref_build_diagonal(self.h, int(Ainv), _p(out)); return out

    def evaluatePerNodeCNTolerance(self, eps, dt):
        tol = np.empty(self.num_nodes); self.lib.implicit_ref_cn_tolerance(self.h, C.c_double(eps), C.c_double(dt), _p(tol)); return tol

    def multiply(self, x, matfree):
        x = np.ascontiguousarray(x, dtype=np.float64); b = np.empty_like(x)
        self.lib.implicit_ref_multiply(self.h, int(matfree), _p(x), _p(b)); return b

    def backwardEulerStep(self, lsolver=3, mg_level=3, smoother=5, coarse_solver=2, Ainv=1, linesearch=1, usecn=1, cneps=1e-7, max_iterations=10000,
                          adaptive_h=0, matfree=0, bcproject=1, max_linear_iterations=10000, mg_times=1, mg_scale=0, topomega=0.1):
        """the whole implicit solve in the reference's solver templates; returns dict(iterations, converged, tolerance, dv)"""
        dv = np.empty((self.num_nodes, 3)); out = np.zeros(3)
        self.lib.implicit_ref_backward_euler_step(self.h, int(lsolver), int(mg_level), int(smoother), int(coarse_solver), int(Ainv), int(linesearch), int(usecn),
                                                  C.c_double(cneps), int(max_iterations), int(adaptive_h), int(matfree), int(bcproject), int(max_linear_iterations), int(mg_times), int(mg_scale), C.c_double(topomega),
                                                  _p(dv), _p(out))
        return dict(iterations=int(out[0]), converged=int(out[1]), tolerance=float(out[2]), dv=dv)

    def begin_step(self):
        self.num_nodes = int(self.lib.implicit_ref_begin_step(self.h))
        return self.num_nodes

    def get_id2coord(self):
        out = np.empty((self.num_nodes, 3), dtype=np.int32); self.lib.implicit_ref_get_id2coord(self.h, _p(out)); return out

    def end_step(self, dt, plastic_model=0, params=()):
        q = np.ascontiguousarray(list(params) + [0.0] * (5 - len(params)), dtype=np.float64)
        flags = (C.c_int * 2)(0, 0)
        self.lib.implicit_ref_end_step(self.h, C.c_double(dt), int(plastic_model), _p(q), flags)
        return (flags[0], flags[1])

    def get_state(self):
        n = self.N
        out = dict(X=np.empty((n, 3)), V=np.empty((n, 3)), C=np.empty((n, 9)), F=np.empty((n, 9)), Jp=np.empty(n), mu=np.empty(n), lam=np.empty(n))
        self.lib.implicit_ref_get_state(self.h, *[_p(out[k]) for k in ("X", "V", "C", "F", "Jp", "mu", "lam")])
        return out

    def shouldExitByCN(self, r, useCN, cneps):
        r = np.ascontiguousarray(r, dtype=np.float64)
        return int(self.lib.implicit_ref_should_exit_by_cn(self.h, _p(r), int(useCN), C.c_double(cneps)))


def _sibling(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", "golden", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def step_scene(make_sim, cells, E, dt, seed, stiff_layer=None):
    """the substep scene of make_lbfgs_golden.scene (seeded block, perturbed F, sticky floor), returned with what the reference side needs;
    stiff_layer = (factor, fraction): the particles of the upper `fraction` of the block get `factor` times the Lame parameters (two materials)"""
    from hot_b200 import scenes
    sc = scenes.block(cells, 1.0 / 32, ppc=6, seed=seed, E=E)
    sc["F"] = sc["F"] + 0.08 * (np.random.default_rng(seed).random(sc["F"].shape) - 0.5)
    if stiff_layer is not None:
        factor, fraction = stiff_layer
        y = sc["X"][:, 1]
        top = y > y.max() - fraction * (y.max() - y.min())
        sc["mu"] = np.where(top, factor * sc["mu"], sc["mu"]); sc["lam"] = np.where(top, factor * sc["lam"], sc["lam"])
    s = make_sim(sc["dx"])
    s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    s.set_dt_gravity(dt, GRAVITY)
    s.sortParticlesAndPolluteGrid(); s.particlesToGrid()
    coord = s.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
    s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    return s, sc, bc


def reference_step(sc, bc, dv0, dt, opts):
    ref = Reference(sc["dx"], dt)
    ref.setup({k: sc[k] for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")}, opts.get("project", 1))
    ref.set_bc(0, bc, np.zeros((len(bc), 9)), None, None, np.zeros(len(bc), dtype=np.int32))
    ref.set_dv(dv0)
    lsolver = opts.get("lsolver", 3)
    return ref.backwardEulerStep(lsolver=lsolver, mg_level=opts.get("mg_level", 3), smoother=opts.get("smoother", 5), coarse_solver=opts.get("coarse_solver", 2),
                                 Ainv=opts.get("Ainv", 1), linesearch=opts.get("linesearch", 1), usecn=opts.get("usecn", 1), cneps=opts.get("cneps", 1e-7),
                                 adaptive_h=opts.get("adaptive_h", 0), matfree=opts.get("matfree", 0), bcproject=opts.get("bcproject", 1), max_linear_iterations=opts.get("max_cg_iterations", 10000), mg_times=opts.get("mg_times", 1), mg_scale=opts.get("mg_scale", 0),
                                 topomega=opts.get("topomega", 0.1),
                                 max_iterations=opts.get("max_lbfgs_iterations", 10000) if lsolver == 3 else opts.get("max_newton_iterations", 3))


# whole time steps (MultigridSimulation::advanceOneTimeStep, MultigridSimulation.h:235-297): sort, P2G, implicit solve, G2P, evolveStrain, plasticity
RUNS = {
    "elastic_hot": (dict(cells=(5, 6, 5), E=2e5, dt=4e-3, seed=3), dict(HOT), ("none", [])),
    "snow_hot": (dict(cells=(5, 6, 5), E=2e5, dt=4e-3, seed=4), dict(HOT), ("snow", [10.0, 2e-2, 7.5e-3, 0.6, 20.0])),      # hardening: mu / lambda change every step
    "von_mises_pn_mgpcg": (dict(cells=(5, 6, 5), E=2e5, dt=4e-3, seed=5), dict(HOT, lsolver=2, max_newton_iterations=10), ("von_mises", [300.0])),
}
RUN_STEPS = 3
PLASTIC_CODE = {"none": 0, "von_mises": 1, "snow": 2}


def floor_bc(coord):
    return np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)


def run_scene(cells, E, dt, seed):
    from hot_b200 import scenes
    sc = scenes.block(cells, 1.0 / 32, ppc=6, seed=seed, E=E)
    sc["F"] = sc["F"] + 0.05 * (np.random.default_rng(seed).random(sc["F"].shape) - 0.5)
    sc["V"] = sc["V"] + np.array([0.0, -1.5, 0.0])                 # the block is pushed into its sticky floor
    return sc


def run_sim(make_sim, sc_args, opts, plastic, steps=RUN_STEPS):
    """`steps` whole time steps through the interface the oracle and the CUDA object share; the state after every step"""
    sc = run_scene(**sc_args)
    dt = sc_args["dt"]
    s = make_sim(sc["dx"])
    s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    s.set_dt_gravity(dt, GRAVITY)
    s.set_plasticity(*plastic)
    states, logs = [], []
    for _ in range(steps):
        s.sortParticlesAndPolluteGrid(); s.particlesToGrid()
        bc = floor_bc(s.get_id2coord())
        s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
        logs.append(s.backwardEulerStep(**opts))
        s.gridToParticles(dt)
        p = s.get_particles()
        Jp, mu, lam = s.get_plastic_state()
        states.append(dict(X=p["X"], V=p["V"], C=p["C"], F=p["F"], Jp=Jp, mu=mu, lam=lam))
    return states, [l["iterations"] for l in logs]


def run_reference(sc_args, opts, plastic, steps=RUN_STEPS):
    """the same time steps in the reference's code (oracle/implicit_ref_shim.cpp): one create() = one process, state carried inside"""
    sc = run_scene(**sc_args)
    dt = sc_args["dt"]
    ref = Reference(sc["dx"], dt)
    ref.setup({k: sc[k] for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")}, opts.get("project", 1))
    states, its = [], []
    lsolver = opts.get("lsolver", 3)
    for k in range(steps):
        n = ref.num_nodes if k == 0 else ref.begin_step()
        bc = floor_bc(ref.get_id2coord())
        ref.set_bc(0, bc, np.zeros((len(bc), 9)), None, None, np.zeros(len(bc), dtype=np.int32))
        dv0 = np.tile(dt * np.asarray(GRAVITY), (n, 1)); dv0[bc] = 0.0      # buildInitialDvAndVnForNewton with a sticky floor at rest
        ref.set_dv(dv0)
        r = ref.backwardEulerStep(lsolver=lsolver, mg_level=opts.get("mg_level", 3), smoother=opts.get("smoother", 5), coarse_solver=opts.get("coarse_solver", 2),
                                  linesearch=opts.get("linesearch", 1), usecn=opts.get("usecn", 1), cneps=opts.get("cneps", 1e-7),
                                  max_iterations=opts.get("max_lbfgs_iterations", 10000) if lsolver == 3 else opts.get("max_newton_iterations", 3))
        its.append(r["iterations"])
        ref.end_step(dt, PLASTIC_CODE[plastic[0]], plastic[1])
        states.append(ref.get_state())
    return states, its


def vectors(n, seed):
    rng = np.random.default_rng(seed)
    return rng.random((n, 3)) - 0.5


if __name__ == "__main__":
    gold = {}
    for name in CASES:
        inp, dx, dt, mode, project = make_inputs(name)
        ref = Reference(dx, dt)
        n = ref.setup(inp, project)
        # node coordinates for the BC sets come from the pinned transfer library's twin in this one: id2coord of buildMatrix
        col0, _ = ref.buildMatrix(False)
        mg = _sibling("make_mpmgrid_golden")
        g = mg.Reference(dx); g.set_particles(inp["X"], inp["V"], inp["mass"], inp["C"]); g.sortParticlesAndPolluteGrid(); g.particlesToGrid()
        coord = g.get_id2coord()
        node, P, R, Rinv, slip = make_bc(coord, mode)
        ref.set_bc(mode, node, P, R, Rinv, slip)
        dv = 0.05 * vectors(n, 21) + dt * np.asarray(GRAVITY)
        dv[node[slip == 0]] = 0.0
        out = dict(num_nodes=np.int64(n), bc_node=node, bc_P=P, bc_R=R, bc_Rinv=Rinv, bc_slip=slip, dv=dv)
        out["energy"] = np.float64(ref.updateState(dv))
        out["F"] = ref.get_F()
        out["residual"] = ref.computeResidual()
        for bc in (0, 1):
            c, v = ref.buildMatrix(bc)
            out[f"col{bc}"] = c
            if name in FULL_MATRIX:
                out[f"val{bc}"] = v
            out[f"valsum{bc}"] = v.sum(1)          # block row sums (every entry enters once)
            x = vectors(n, 30 + bc)
            out[f"x{bc}"] = x
            out[f"Ax{bc}"] = ref.multiply(x, False)
        x = vectors(n, 40)
        out["x_mf"] = x; out["Ax_mf"] = ref.multiply(x, True)
        for a in (0, 1):
            out[f"diag{a}"] = ref.buildDiagonal(a)
        out["cntol"] = ref.evaluatePerNodeCNTolerance(CN_EPS, dt)
        r = out["residual"]
        scaled = (r ** 2).sum(1) / out["cntol"] ** 2
        f = np.sqrt(n / scaled.sum())            # residual scaled to sit just below / above the CN exit threshold
        out["exit"] = np.array([ref.shouldExitByCN(0.99 * f * r, 1, 0.0), ref.shouldExitByCN(1.01 * f * r, 1, 0.0),
                                ref.shouldExitByCN(r, 0, 1.01 * np.sqrt((r ** 2).sum())), ref.shouldExitByCN(r, 0, 0.99 * np.sqrt((r ** 2).sum()))], dtype=np.int32)
        out["exit_scale"] = np.float64(f)
        gold[name + "/dx"] = np.float64(dx); gold[name + "/dt"] = np.float64(dt); gold[name + "/mode"] = np.int64(mode)
        gold[name + "/project"] = np.int64(project)
        for k, v in inp.items():
            gold[f"{name}/in_{k}"] = v
        for k, v in out.items():
            gold[f"{name}/{k}"] = v
        print(name, "particles", len(inp["mass"]), "nodes", n, "bc", len(node), "slip", int(slip.sum()), "energy", out["energy"], "exit", out["exit"])
    import oracle_binding
    for name, (sc_args, opts) in STEPS.items():
        o, sc, bc = step_scene(oracle_binding.OracleSim, **sc_args)      # (the oracle only supplies the start value buildInitialDvAndVnForNewton leaves in dv)
        r = reference_step(sc, bc, o.get_dv(), sc_args["dt"], opts)
        gold[f"step/{name}/iterations"] = np.int64(r["iterations"]); gold[f"step/{name}/converged"] = np.int64(r["converged"])
        gold[f"step/{name}/tolerance"] = np.float64(r["tolerance"]); gold[f"step/{name}/dv"] = r["dv"]
        print("step", name, "iterations", r["iterations"], "converged", r["converged"], "tolerance", r["tolerance"])
    for name, (sc_args, opts, plastic) in RUNS.items():
        states, its = run_reference(sc_args, opts, plastic)
        gold[f"run/{name}/iterations"] = np.array(its, dtype=np.int64)
        for k, st in enumerate(states):
            for key, v in st.items():
                if k == len(states) - 1 or key in ("Jp", "mu"):           # full state after the last step, the plastic history after every step
                    gold[f"run/{name}/{k}/{key}"] = v
        print("run", name, "iterations per step", its, "Jp range", states[-1]["Jp"].min(), states[-1]["Jp"].max(),
              "mu range", states[-1]["mu"].min() / states[0]["mu"].max())
    np.savez_compressed(OUT, **gold)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
