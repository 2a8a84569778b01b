"""Generates tests/golden/krylov_ref.npz: the REFERENCE'S OWN Krylov solver classes - ZIRAN::InexactConjugateGradient
(Lib/Ziran/Math/Linear/InexactConjugateGradient.h:49-103) and ZIRAN::Minres (Lib/Ziran/Math/Linear/Minres.h:71-176), compiled where they lie into
oracle/_ref/libziran_ref.so (oracle/ziran_krylov_shim.cpp) - run on the operator of a small seeded MPM system that the oracle
provides through callbacks (multiply / project / precondition).  tests/test_oracle_krylov_ref.py compares the oracle's restatement of the
two iterations (oracle_solver.inl: inexact_pcg, minres_solve) with these iteration counts and solutions.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_krylov_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

APPLY = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long)
PROJECT = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_long)

# (name, solver, matfree, preconditioner of orc_pcg / orc_minres, solver arguments)
# `scale`: the right-hand side is scaled so that its preconditioned norm |r|_M-1 is `scale` - the inexact CG's forcing sequence
# min(0.5, sqrt(max(|r|, tolerance))) then asks for a reduction by sqrt(scale) instead of 0.5 (more iterations, the other branch)
CASES = [
    ("pcg_mf_none", "pcg", 1, 0, None, dict(tolerance=1e-30, max_iterations=12)),
    ("pcg_mf_jacobi", "pcg", 1, 1, None, dict(tolerance=1e-30, max_iterations=200)),
    ("pcg_mat_jacobi_half", "pcg", 0, 1, None, dict(tolerance=1e-30, max_iterations=200)),
    ("pcg_mf_jacobi_1e-6", "pcg", 1, 1, 1e-6, dict(tolerance=1e-30, max_iterations=200)),
    ("pcg_mat_jacobi_1e-8", "pcg", 0, 1, 1e-8, dict(tolerance=1e-30, max_iterations=300)),
    ("pcg_mat_vcycle_1e-8", "pcg", 0, 2, 1e-8, dict(tolerance=1e-30, max_iterations=200)),
    ("pcg_tolerance_floor", "pcg", 0, 1, 1e-12, dict(tolerance=1e-6, max_iterations=200)),   # max(|r|, tolerance) takes the tolerance
    ("pcg_max_iterations", "pcg", 0, 1, 1e-10, dict(tolerance=1e-30, max_iterations=3)),
    ("minres_mf_jacobi", "minres", 1, 1, None, dict(relative_tolerance=1e-3, tolerance=1e300, max_iterations=200)),
    ("minres_mat_vcycle", "minres", 0, 2, None, dict(relative_tolerance=1e-6, tolerance=1e300, max_iterations=200)),
    ("minres_abs_tol", "minres", 0, 1, 1.0, dict(relative_tolerance=1.0, tolerance=1e-5, max_iterations=200)),
    ("minres_max_iterations", "minres", 0, 1, None, dict(relative_tolerance=1e-12, tolerance=1e300, max_iterations=4)),
]


def scene(orc):
    """the system of tests/test_oracle_solver.py::_scene with the matrix and the hierarchy built"""
    from hot_b200 import scenes
    sc = scenes.block((6, 5, 6), 1.0 / 32, ppc=6, seed=3, E=3e6)
    rng = np.random.default_rng(3)
    sc["F"] = sc["F"] + 0.05 * (rng.random(sc["F"].shape) - 0.5)
    o = orc.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.set_dt_gravity(1e-2, (0.0, -9.8, 0.0))
    o.sortParticlesAndPolluteGrid(); o.particlesToGrid()
    coord = o.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
    o.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    o.backupStrain(); o.updateState()
    o.buildMatrix(True); o.buildMultigrid(levels=3)
    b = o.project(np.random.default_rng(0).random((o.num_nodes, 3)) - 0.5)
    return o, b


def scaled_rhs(o, b, matfree, precond, scale):
    if scale is None:
        return b
    z = o.obj_precondition(b, matfree=bool(matfree), preconditioner=precond)
    return b * (scale / np.sqrt(float((z * b).sum())))


def reference_solve(lib, o, b, solver, matfree, precond, **kw):
    n = b.size
    shape = b.shape

    def mul(_, pin, pout, m):
        x = np.ctypeslib.as_array(pin, (m,)).reshape(shape)
        np.ctypeslib.as_array(pout, (m,))[:] = o.obj_multiply(x, matfree=bool(matfree)).reshape(-1)

    def proj(_, pv, m):
        v = np.ctypeslib.as_array(pv, (m,))
        v[:] = o.project(v.reshape(shape)).reshape(-1)

    def prec(_, pin, pout, m):
        r = np.ctypeslib.as_array(pin, (m,)).reshape(shape)
        np.ctypeslib.as_array(pout, (m,))[:] = o.obj_precondition(r, matfree=bool(matfree), preconditioner=precond).reshape(-1)
    cb = (APPLY(mul), PROJECT(proj), APPLY(prec))
    x = np.zeros(n)
    px, pb = x.ctypes.data_as(C.POINTER(C.c_double)), np.ascontiguousarray(b.reshape(-1)).ctypes.data_as(C.POINTER(C.c_double))
    if solver == "pcg":
        lib.zr_inexact_cg.restype = C.c_int
        it = lib.zr_inexact_cg(None, cb[0], cb[1], cb[2], C.c_long(n), px, pb, int(kw["max_iterations"]), C.c_double(kw["tolerance"]))
    else:
        lib.zr_minres.restype = C.c_int
        it = lib.zr_minres(None, cb[0], cb[1], cb[2], C.c_long(n), px, pb, int(kw["max_iterations"]), C.c_double(kw["tolerance"]),
                           C.c_double(kw["relative_tolerance"]))
    return x.reshape(shape), int(it)


def main():
    import oracle_binding as orc
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libziran_ref.so"))
    o, b = scene(orc)
    out = {"b": b}
    for name, solver, matfree, precond, scale, kw in CASES:
        bs = scaled_rhs(o, b, matfree, precond, scale)
        x, it = reference_solve(lib, o, bs, solver, matfree, precond, **kw)
        out[name + "_x"] = x
        out[name + "_it"] = np.int64(it)
        print(name, "iterations", it, "|x|", float(np.abs(x).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "krylov_ref.npz"), **out)


if __name__ == "__main__":
    main()
