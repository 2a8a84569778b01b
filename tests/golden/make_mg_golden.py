"""Generates tests/golden/mg_ref.npz: the REFERENCE'S OWN Galerkin-multigrid code - ZIRAN::MultigridBuilder::build
(Projects/multigrid/MultigridPreconditioner.h:553-703), SquareMatrix::{buildDiagonal, buildTransposeMatrix, buildCoarseMatrix, multiply}
(SquareMatrix.h), the smoothers and MultigridOperator::operator() (MultigridPreconditioner.h:160-318,362-421), compiled where they lie into
oracle/_ref/libziran_ref.so (oracle/mg_ref_shim.cpp) - run on the level-0 system of a small seeded MPM scene, which the oracle assembles
and hands over as arrays (id2coord, entryCol, entryVal, mass: the builder's own interface).  tests/test_oracle_mg_ref.py compares the
oracle's restatement of rows a16-a20 (oracle_matrix.inl) and the CUDA path with these results.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_mg_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libziran_ref.so")
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)
LEVELS = 3
# (smoother, coarseSolver, times): the hierarchies that are built; smoother calls checked per level: kind -> iterations
CONFIGS = [(5, 2, 1), (0, 2, 2), (1, 0, 2), (5, 5, 1)]
SMOOTHERS = {5: 2, 0: 3, 1: 3, 2: 6}


def _d(a):
    return a.ctypes.data_as(DP)


def _i(a):
    return a.ctypes.data_as(IP)


def scene(make_sim):
    """a seeded block on a sticky floor with the level-0 matrix assembled (BC-projected, like HOT's --bcproject)"""
    from hot_b200 import scenes
    sc = scenes.block((7, 6, 5), 1.0 / 32, ppc=6, seed=5, E=2e5)
    rng = np.random.default_rng(5)
    sc["F"] = sc["F"] + 0.06 * (rng.random(sc["F"].shape) - 0.5)
    s = make_sim(sc["dx"])
    s.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    s.set_dt_gravity(6e-3, (0.0, -9.8, 0.0))
    s.sortParticlesAndPolluteGrid(); s.particlesToGrid()
    coord = s.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)
    s.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    s.backupStrain(); s.updateState()
    s.buildMatrix(True)
    return s


def vectors(n, seed):
    rng = np.random.default_rng(seed)
    return rng.random((n, 3)) - 0.5, 0.01 * (rng.random((n, 3)) - 0.5)


class Reference:
    """the reference's multigrid on arrays (oracle/mg_ref_shim.cpp)"""
    def __init__(self):
        self.lib = C.CDLL(REF_LIB)
        self.lib.zr_mg_level_entries.restype = C.c_long

    def build(self, o, smoother, coarse, times, Ainv=1, levelscale=0, topomega=0.1, cneps=1e-7):
        col, val = o.get_matrix()
        coord = np.ascontiguousarray(o.get_id2coord(), dtype=np.int32)
        mass = np.ascontiguousarray(o.buildMassMatrix(), dtype=np.float64)
        col = np.ascontiguousarray(col, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
        rc = self.lib.zr_mg_build(int(o.num_nodes), _i(coord), _i(col), _d(val), _d(mass), LEVELS, int(smoother), int(coarse), int(Ainv), int(times),
                                  int(levelscale), C.c_double(topomega), C.c_double(cneps))
        assert rc == 0
        self.dofs = [self.lib.zr_mg_level_dofs(l) for l in range(LEVELS)]

    def spmv(self, level, x):
        x = np.ascontiguousarray(x, dtype=np.float64); b = np.empty_like(x)
        self.lib.zr_mg_spmv(level, _d(x), _d(b))
        return b

    def diagonal(self, level):
        n = self.dofs[level]
        D = np.empty((n, 9)); Di = np.empty((n, 9))
        self.lib.zr_mg_level_diagonal(level, _d(D), _d(Di))
        return D, Di

    def color_order(self, level):
        out = np.empty((self.dofs[level], 3), dtype=np.int32)
        self.lib.zr_mg_color_order(level, _i(out))
        return out

    def transfer(self, level, kind):
        cs = C.c_int(0)
        ne = self.lib.zr_mg_level_entries(level, kind)
        col = np.empty(ne, dtype=np.int32); val = np.empty((ne, 9))
        self.lib.zr_mg_level_matrix(level, kind, C.byref(cs), _i(col), _d(val))
        return col.reshape(-1, cs.value), val.reshape(-1, cs.value, 9)

    def smooth(self, level, kind, u, r, iterations, tolerance=0.0, initial_residual=None):
        u = np.ascontiguousarray(u, dtype=np.float64).copy(); r = np.ascontiguousarray(r, dtype=np.float64).copy()
        ir = None if initial_residual is None else _d(np.ascontiguousarray(initial_residual, dtype=np.float64))
        rc = self.lib.zr_mg_smooth(level, kind, _d(u), _d(r), int(iterations), C.c_double(tolerance), ir)
        assert rc == 0
        return u, r

    def vcycle(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64); out = np.empty_like(r)
        self.lib.zr_mg_vcycle(_d(r), _d(out))
        return out

    def estimate2norm(self, level):
        """SquareMatrix::estimate2norm as the reference runs it (clock-seeded start): returns the start vector it used, lMax, lMin"""
        start = np.empty((self.dofs[level], 3)); lmax = C.c_double(0); lmin = C.c_double(0)
        rc = self.lib.zr_mg_estimate_two_norm(level, _d(start), C.byref(lmax), C.byref(lmin))
        assert rc == 0
        return start, lmax.value, lmin.value


def main():
    import oracle_binding as orc
    o = scene(orc.OracleSim)
    ref = Reference()
    out = {}
    for smoother, coarse, times in CONFIGS:
        tag = f"s{smoother}c{coarse}t{times}"
        ref.build(o, smoother, coarse, times)
        out[tag + "_dofs"] = np.asarray(ref.dofs, dtype=np.int64)
        b, _ = vectors(ref.dofs[0], 1)
        out[tag + "_vcycle"] = ref.vcycle(b)
        print(tag, "dofs", ref.dofs, "|V-cycle|", float(np.abs(out[tag + "_vcycle"]).max()))
        if (smoother, coarse) == (5, 2):
            for l in range(LEVELS):
                x, u0 = vectors(ref.dofs[l], 10 + l)
                out[f"spmv{l}"] = ref.spmv(l, x)
                D, Di = ref.diagonal(l)
                out[f"diag{l}"] = D; out[f"dinv{l}"] = Di
                out[f"color{l}"] = ref.color_order(l)
                for kind, iters in SMOOTHERS.items():
                    ug, rg = ref.smooth(l, kind, u0, x, iters, initial_residual=4.0 * x if kind == 2 else None)
                    out[f"smooth{kind}_l{l}_u"] = ug; out[f"smooth{kind}_l{l}_r"] = rg
            for l in range(LEVELS - 1):
                pc, pv = ref.transfer(l, 1)
                out[f"pcol{l}"] = pc; out[f"pw{l}"] = pv[:, :, 0]      # the prolongation blocks are weight * I
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mg_ref.npz"), **out)


CHEB_OUT = os.path.join(ROOT, "tests", "golden", "cheb_ref.npz")
CHEB_ITERS = 4


def main_chebyshev():
    """tests/golden/cheb_ref.npz (f4: -smoother 6): the reference's estimate2norm with the start vector it drew (its seed is the clock: the vector is
    recovered after the call, oracle/mg_ref_shim.cpp), chebyshev_smooth on every level and the V-cycle of the Chebyshev hierarchy"""
    import oracle_binding as orc
    o = scene(orc.OracleSim)
    ref = Reference()
    ref.build(o, 6, 2, 1)              # (the build estimates the norms once itself; the calls below redo it and keep the start vectors)
    out = {"dofs": np.asarray(ref.dofs, dtype=np.int64)}
    for l in range(LEVELS):
        start, lmax, lmin = ref.estimate2norm(l)
        out[f"start{l}"] = start.astype(np.int8); out[f"lmax{l}"] = np.float64(lmax); out[f"lmin{l}"] = np.float64(lmin)
        x, u0 = vectors(ref.dofs[l], 10 + l)
        out[f"cheb_l{l}_u"], out[f"cheb_l{l}_r"] = ref.smooth(l, 6, u0, x, CHEB_ITERS)
        print("level", l, "dofs", ref.dofs[l], "lMax", lmax, "lMin", lmin)
    b, _ = vectors(ref.dofs[0], 1)
    out["vcycle"] = ref.vcycle(b)
    np.savez_compressed(CHEB_OUT, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "chebyshev":
        main_chebyshev()
    else:
        main()
