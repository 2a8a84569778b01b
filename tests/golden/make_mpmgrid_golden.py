"""Generates tests/golden/mpmgrid_ref.npz: the particle <-> grid transfers of rows a3 / a5 / a6 / a7 / a23 run on the REFERENCE'S OWN grid
code - GridState, BSplineWeights, MpmGrid::{iterateKernel, getNumNodes, iterateGrid} (Lib/MPM/MpmGrid.h) over its SPGrid allocator and page
map (Lib/SPGrid/Core) and baseNode / the quadratic weights (Lib/Ziran/Math/Splines/BSplines.h), compiled where they lie into
oracle/_ref/libmpmgrid_ref.so (oracle/mpmgrid_ref_shim.cpp: the particle loops of MpmSimulationBase.cpp:1066-1137, 611-656, 521-532, 891-901,
930-1006 around those calls are written out there, MpmSimulationBase.cpp itself cannot be compiled in this image).
tests/test_oracle_mpmgrid_ref.py compares the oracle's restatement (oracle/hot_oracle.cpp) and the CUDA path with these results.
The inputs are stored with the outputs, so nothing depends on a random-number stream.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_mpmgrid_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmpmgrid_ref.so")
OUT = os.path.join(ROOT, "tests", "golden", "mpmgrid_ref.npz")
DT = 2e-3
# case -> (scene arguments, apic_rpic_ratio, velocity scale of the Newton increment)
CASES = {
    "tiny": (dict(cells=(3, 2, 2), dx=0.05, ppc=3, seed=5), 1.0, 0.1),
    "ragged": (dict(cells=(7, 5, 9), dx=0.02, ppc=5, seed=6), 1.0, 0.1),
    "ragged_rpic": (dict(cells=(7, 5, 9), dx=0.02, ppc=5, seed=6), 0.0, 0.1),
    "dense_cells": (dict(cells=(4, 4, 4), dx=0.04, ppc=40, seed=8), 0.5, 0.1),       # more particles per page than one staging chunk
    "page_corner": (dict(cells=(5, 9, 6), dx=1.0 / 64, ppc=4, seed=9, origin_cells=(30, 13, 61)), 1.0, 30.0),  # straddles page borders; fast: CFL flags
}


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Reference:
    """the reference's grid code on arrays (oracle/mpmgrid_ref_shim.cpp); method names of OracleSim / MpmSimulationB200"""
    ELEMENTS_PER_BLOCK = 32

    def __init__(self, dx, apic_rpic_ratio=1.0, cfl=0.6):
        self.lib = C.CDLL(REF_LIB)
        self.lib.mpmgrid_ref_create.restype = C.c_void_p
        self.lib.mpmgrid_ref_create.argtypes = [C.c_double, C.c_double, C.c_double]
        for n in ("mpmgrid_ref_sort", "mpmgrid_ref_num_pages"):
            getattr(self.lib, n).restype = C.c_long
        self.h = C.c_void_p(self.lib.mpmgrid_ref_create(float(dx), float(apic_rpic_ratio), float(cfl)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mpmgrid_ref_destroy(self.h)
            self.h = None

    def layout(self):
        out = np.zeros(6, dtype=np.int64)
        self.lib.mpmgrid_ref_layout(_p(out))
        return out

    def set_particles(self, X, V, mass, C_, F=None, vol=None, mu=None, lam=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.N = len(mass)
        self._keep = [f(X), f(V), f(mass), f(C_)]
        self.lib.mpmgrid_ref_set_particles(self.h, C.c_long(self.N), *[_p(a) for a in self._keep])

    def sortParticlesAndPolluteGrid(self):
        self.num_groups = int(self.lib.mpmgrid_ref_sort(self.h))
        self.num_pages = int(self.lib.mpmgrid_ref_num_pages(self.h))
        n, g = self.N, self.num_groups
        self._sort = (np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.uint64))
        self._groups = (np.empty(g, dtype=np.int32), np.empty(g, dtype=np.int32), np.empty(g, dtype=np.uint64))
        self._pages = np.empty(self.num_pages, dtype=np.uint64)
        self.lib.mpmgrid_ref_get_sort(self.h, *[_p(a) for a in self._sort + self._groups], _p(self._pages))

    def get_sort(self):
        return self._sort

    def get_groups(self):
        return self._groups

    def get_pages(self):
        return self._pages

    def particlesToGrid(self):
        self.num_nodes = int(self.lib.mpmgrid_ref_p2g(self.h))
        gn = self.num_pages * self.ELEMENTS_PER_BLOCK
        self._grid = (np.empty(gn, dtype=np.int64), np.empty(gn), np.empty((gn, 3)))
        self._id2coord = np.empty((self.num_nodes, 3), dtype=np.int32)
        self.lib.mpmgrid_ref_get_grid(self.h, *[_p(a) for a in self._grid], _p(self._id2coord))
        return self.num_nodes

    def get_grid(self):
        return self._grid

    def get_id2coord(self):
        return self._id2coord

    def set_dv(self, dv):
        self._dv = np.ascontiguousarray(dv, dtype=np.float64)

    def gridToParticles(self, dt):
        flags = (C.c_int * 2)(0, 0)
        self.lib.mpmgrid_ref_g2p(self.h, _p(self._dv), C.c_double(dt), flags)
        return (flags[0], flags[1])

    def get_particles(self):
        n = self.N
        X = np.empty((n, 3)); V = np.empty((n, 3)); Cm = np.empty((n, 9)); G = np.empty((n, 9))
        self.lib.mpmgrid_ref_get_particles(self.h, _p(X), _p(V), _p(Cm), _p(G))
        return dict(X=X, V=V, C=Cm, gradV=G)


def make_scene(name):
    from hot_b200 import scenes
    kw, ratio, vscale = CASES[name]
    sc = scenes.block(**kw)
    return sc, ratio, vscale


def run(sim, inp, dv=None):
    """one sort -> P2G -> G2P pass; returns every product the rows define (dv: Newton increment per DOF, default = the stored one)"""
    sim.set_particles(inp["X"], inp["V"], inp["mass"], inp["C"], inp["F"], inp["vol"], inp["mu"], inp["lam"])
    sim.sortParticlesAndPolluteGrid()
    out = {}
    out["sorter"], out["order"], out["base"] = sim.get_sort()
    out["first"], out["last"], out["blk"] = sim.get_groups()
    out["pages"] = sim.get_pages()
    out["num_nodes"] = np.int64(sim.particlesToGrid())
    out["idx"], out["m"], out["v"] = sim.get_grid()
    out["id2coord"] = sim.get_id2coord()
    sim.set_dv(inp["dv"] if dv is None else dv)
    out["flags"] = np.array(sim.gridToParticles(DT), dtype=np.int32)
    p = sim.get_particles()
    for k in ("X", "V", "C", "gradV"):
        out["p" + k] = p[k]
    return out


def time_slab(path):
    """bench.py's cpu_baseline.reference_code leg (run as a child process): mean seconds of one P2G + G2P(dt = 0) pass over the particles of `path`"""
    import time
    d = np.load(path)
    ref = Reference(float(d["dx"]))
    ref.set_particles(d["X"], d["V"], d["mass"], d["C"])
    ref.sortParticlesAndPolluteGrid()
    times = []
    for it in range(1 + int(d["reps"])):
        t0 = time.perf_counter()
        nn = ref.particlesToGrid()
        ref.set_dv(np.zeros((nn, 3)))
        ref.gridToParticles(0.0)
        if it >= 1:
            times.append(time.perf_counter() - t0)
    print(repr(float(np.mean(times))))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "time":
    time_slab(sys.argv[2])
    sys.exit(0)

if __name__ == "__main__":
    gold = {}
    r0 = Reference(0.1)
    gold["layout"] = r0.layout()
    for name in CASES:
        sc, ratio, vscale = make_scene(name)
        ref = Reference(sc["dx"], apic_rpic_ratio=ratio)
        inp = {k: np.ascontiguousarray(sc[k]) for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam")}
        if name == "dense_cells":
            inp["V"] = inp["V"] + np.array([10.0, 0.0, 0.0])      # faster than half a cell per step, slower than a cell: flags (0, 1)
        # the Newton increment needs the node count: one P2G first
        ref.set_particles(inp["X"], inp["V"], inp["mass"], inp["C"])
        ref.sortParticlesAndPolluteGrid()
        n = ref.particlesToGrid()
        inp["dv"] = vscale * (np.random.default_rng(11).random((n, 3)) - 0.5)
        ref = Reference(sc["dx"], apic_rpic_ratio=ratio)
        out = run(ref, inp)
        gold[name + "/dx"] = np.float64(sc["dx"]); gold[name + "/ratio"] = np.float64(ratio)
        for k, v in inp.items():
            if k in ("F", "vol", "mu", "lam"):
                continue            # not read by the transfers; the test passes identity / ones
            gold[f"{name}/in_{k}"] = v
        for k, v in out.items():
            gold[f"{name}/{k}"] = v
        print(name, "particles", len(inp["mass"]), "groups", len(out["first"]), "pages", len(out["pages"]), "nodes", int(out["num_nodes"]),
              "flags", out["flags"])
    np.savez_compressed(OUT, **gold)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; GridState layout", gold["layout"])
