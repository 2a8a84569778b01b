"""Deterministic synthetic inputs shaped like the scenes BASELINE.json names.

Two particle generators:
  * `poisson_box`: the reference's own recipe, sampleInAnalyticLevelSet = PoissonDisk::setDistanceByParticlesPerCell +
    sampleFromPeriodicData (Lib/Ziran/Math/Geometry/PoissonDisk.h:152-165,185-222, Lib/MPM/MpmInitializationHelper.h:237-252,367-377)
    on the part of the reference's tile file that ships with this package (hot_b200/data/poisson_brick.npz, cut by
    tests/golden/make_poisson_brick.py: tile points with x, z >= 0).  No RNG: the positions are the reference's positions.
    Usable for level sets whose x and z sides are <= 60 min_distance (C1, C2).
  * `block` / `level_set_jitter`: per-cell jitter with a fixed numpy seed for the scenes that need the whole 13 MB tile
    (C3, C4, C5) and for the small parity-test scenes.
Material constants follow CorotatedIsotropic (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h:69-73);
vol = volume / N, m = rho * vol (MpmInitializationHelper.h:367-377).
"""
import itertools
import os

import numpy as np

_BRICK = None


def _brick():
    global _BRICK
    if _BRICK is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "poisson_brick.npz")
        _BRICK = np.load(path)["points"].astype(np.float64)  # the reference reads float and computes in T = double
    return _BRICK


def poisson_box(min_corner, max_corner, dx, ppc):
    """Positions sampleInAnalyticLevelSet(AxisAlignedAnalyticBox(min_corner, max_corner), ., ppc) produces, in its order."""
    lo, hi = np.asarray(min_corner, dtype=np.float64), np.asarray(max_corner, dtype=np.float64)
    md = (dx ** 3 / ppc * (13.0 / 18.0)) ** (1.0 / 3.0)           # setDistanceByParticlesPerCell
    side = hi - lo
    if side[0] > 60 * md or side[2] > 60 * md:
        raise ValueError("poisson_box: the shipped brick of the tile covers x, z sides up to 60 min_distance only")
    ref = 120.0 * md
    n_off = (np.ceil(side / ref) + 1).astype(int)
    new_point = md * _brick() + lo
    out = []
    for it in itertools.product(range(n_off[0]), range(n_off[1]), range(n_off[2])):
        q = new_point + np.asarray(it, dtype=np.float64) * ref
        out.append(np.where(np.all((q >= lo) & (q <= hi), axis=1), np.arange(len(q)), -1))
    idx = np.stack(out, 1)                                          # [tile point, offset]: the reference's loop nest
    sel = idx >= 0
    offs = np.asarray(list(itertools.product(range(n_off[0]), range(n_off[1]), range(n_off[2]))), dtype=np.float64) * ref
    pts = (new_point[:, None, :] + offs[None, :, :])[sel]
    return pts


def perturb_state(X, rng):
    """fixed smooth velocity / affine field + a small random strain so that kernels see F != I, C != 0"""
    n = len(X)
    c = X.mean(0)
    omega = np.array([0.3, 1.0, -0.2])
    V = np.cross(omega, X - c) + 0.05 * np.sin(7.0 * X[:, [1, 2, 0]])
    W = np.array([[0, -omega[2], omega[1]], [omega[2], 0, -omega[0]], [-omega[1], omega[0], 0]])
    Cm = np.tile(W.T.reshape(1, 9), (n, 1)) + 0.02 * (rng.random((n, 9)) - 0.5)  # column-major W
    F = np.tile(np.eye(3).reshape(1, 9), (n, 1)) + 0.05 * (rng.random((n, 9)) - 0.5)
    return V, Cm, F


def lame(E, nu):
    return E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))


def block(cells, dx, ppc=8, origin_cells=(8, 8, 8), rho=1000.0, E=1e5, nu=0.3, seed=0, perturb=True, shuffle=True):
    """Axis-aligned elastic block of `cells` grid cells, lower corner at origin_cells*dx (+0.25 dx so that
    particles do not sit on the cell faces)."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = cells
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    corner = np.stack([ci, cj, ck], -1).reshape(-1, 1, 3).astype(np.float64)
    jit = rng.random((corner.shape[0], ppc, 3))
    X = ((corner + jit).reshape(-1, 3) + np.asarray(origin_cells, dtype=np.float64) + 0.25) * dx
    n = len(X)
    if shuffle:
        X = X[rng.permutation(n)]
    volume = nx * ny * nz * dx ** 3
    vol = np.full(n, volume / n)
    mass = rho * vol
    mu, lam = lame(E, nu)
    V = np.zeros((n, 3)); Cm = np.zeros((n, 9)); F = np.tile(np.eye(3).reshape(1, 9), (n, 1))
    if perturb:
        c = X.mean(0)
        omega = np.array([0.3, 1.0, -0.2])
        V = np.cross(omega, X - c) + 0.05 * np.sin(7.0 * X[:, [1, 2, 0]])
        W = np.array([[0, -omega[2], omega[1]], [omega[2], 0, -omega[0]], [-omega[1], omega[0], 0]])
        Cm = np.tile(W.T.reshape(1, 9), (n, 1)) + 0.02 * (rng.random((n, 9)) - 0.5)  # column-major W
        F = F + 0.05 * (rng.random((n, 9)) - 0.5)
    return dict(X=X, V=V, mass=mass, C=Cm, F=F, vol=vol, mu=np.full(n, mu), lam=np.full(n, lam), dx=dx)


def assemble(parts, dx, seed=0, perturb=True):
    """parts: list of (positions, volume, rho, E, nu) -> scene dict; particle order = the order of the parts"""
    rng = np.random.default_rng(seed)
    X = np.concatenate([p[0] for p in parts])
    vol = np.concatenate([np.full(len(p[0]), p[1] / len(p[0])) for p in parts])
    mass = np.concatenate([np.full(len(p[0]), p[2] * p[1] / len(p[0])) for p in parts])
    mu = np.concatenate([np.full(len(p[0]), lame(p[3], p[4])[0]) for p in parts])
    lam = np.concatenate([np.full(len(p[0]), lame(p[3], p[4])[1]) for p in parts])
    n = len(X)
    if perturb:
        V, Cm, F = perturb_state(X, rng)
    else:
        V = np.zeros((n, 3)); Cm = np.zeros((n, 9)); F = np.tile(np.eye(3).reshape(1, 9), (n, 1))
    return dict(X=np.ascontiguousarray(X), V=V, mass=mass, C=Cm, F=F, vol=vol, mu=mu, lam=lam, dx=dx)


def level_set_jitter(inside, lo_cell, hi_cell, dx, ppc, rng):
    """per-cell jitter inside an analytic level set: `ppc` candidates per cell of the box [lo_cell, hi_cell), kept where inside(x)"""
    ci, cj, ck = np.meshgrid(*[np.arange(a, b) for a, b in zip(lo_cell, hi_cell)], indexing="ij")
    corner = np.stack([ci, cj, ck], -1).reshape(-1, 1, 3).astype(np.float64)
    X = ((corner + rng.random((corner.shape[0], ppc, 3))).reshape(-1, 3) + 0.25) * dx
    return X[inside(X)]


# BASELINE.json configs restated
def config_c1(seed=0):
    """box drop modelled on the soft cube of test 9211 (MultigridInit3D.h:96-125): 18^3 cells at dx = 1/64, ppc 8, rho 1000, E 2.5e4,
    nu .4, sampled with the reference's Poisson-tile recipe (about 38 k particles)"""
    dx = 1.0 / 64
    lo = np.array([20.25, 8.25, 20.25]) * dx
    X = poisson_box(lo, lo + 18 * dx, dx, 8)
    return assemble([(X, (18 * dx) ** 3, 1000.0, 2.5e4, 0.4)], dx, seed)


def config_c2(seed=0, dx=0.12 / 23, E_mid=1e9, copy=0):
    """twisting bar, test 777001 (MultigridInit3D.h:571-664): three stacked boxes 0.12 x 0.3 x 0.12 at (2 +- .06, 2.8 .. 3.7, 2 +- .06),
    E = 1e5 / -cmd0 (1e9) / 1e5, nu .3, rho 2e3, ppc 12, each sampled separately with the reference's Poisson-tile recipe;
    dx = 0.12 / 23 (reference 0.0075) for the ~1 M particles BASELINE.json names: 951 k particles.  copy = k shifts the bar by k bar lengths."""
    half = 0.06
    parts = []
    for y0, E in ((2.8, 1e5), (3.1, E_mid), (3.4, 1e5)):
        y0 += 0.9 * copy          # copy k: the same bar stacked on top of copy k - 1 (weak scaling: one object of N bars end to end)
        X = poisson_box((2 - half, y0, 2 - half), (2 + half, y0 + 0.3, 2 + half), dx, 12)
        parts.append((X, 0.12 * 0.3 * 0.12, 2e3, E, 0.3))
    return assemble(parts, dx, seed)


def config_c3(seed=0, scale=1.0):
    """faceless, test 777011 (MultigridInit3D.h:2477-2583): lion.vdb is unreadable here (no OpenVDB), so the solid is the
    stand-in of SURVEY 8d: union of an analytic sphere and a box of about the lion's volume, same constants (rho 2000, E 5e4, nu .3,
    dx 0.01, ppc 20) -> about 4 M particles at scale 1 (scale shrinks the solid for parity tests)."""
    dx, ppc = 0.01, 20
    rng = np.random.default_rng(seed)
    c = np.array([2.0, 1.0, 2.0])
    r = 0.30 * scale
    blo, bhi = c + np.array([-0.38, -0.36, -0.22]) * scale, c + np.array([0.38, -0.12, 0.22]) * scale

    def inside(X):
        return (np.sum((X - c) ** 2, 1) <= r * r) | np.all((X >= blo) & (X <= bhi), 1)
    lo = np.floor((np.minimum(c - r, blo)) / dx).astype(int) - 1
    hi = np.ceil((np.maximum(c + r, bhi)) / dx).astype(int) + 1
    X = level_set_jitter(inside, lo, hi, dx, ppc, rng)
    n_cells = len(X) / ppc
    return assemble([(X, n_cells * dx ** 3, 2000.0, 5e4, 0.3)], dx, seed)


def config_c5(seed=0, scale=1.0, ppc=12):
    """stiff wheel, test 777019 (MultigridInit3D.h:3278-3413): wheel.vdb replaced by the analytic Torus (AnalyticLevelSet.h:366-408)
    R = .25, r = .06, axis y; rho 2700, nu .33, E = 200e9 (BASELINE.json); dx chosen for about 16 M particles at scale 1."""
    R, r = 0.25 * scale, 0.06 * scale
    volume = 2 * np.pi ** 2 * R * r * r
    dx = (volume * ppc / (1.6e7 * scale ** 3)) ** (1.0 / 3.0)
    rng = np.random.default_rng(seed)
    c = np.array([2.0, 1.0, 2.0])

    def inside(X):
        d = X - c
        q = np.sqrt(d[:, 0] ** 2 + d[:, 2] ** 2) - R
        return q * q + d[:, 1] ** 2 <= r * r
    lo = np.floor((c - np.array([R + r, r, R + r])) / dx).astype(int) - 1
    hi = np.ceil((c + np.array([R + r, r, R + r])) / dx).astype(int) + 1
    X = level_set_jitter(inside, lo, hi, dx, ppc, rng)
    return assemble([(X, volume, 2700.0, 200e9, 0.33)], dx, seed)


def config_c4(seed=0):
    """column 100 x 400 x 25 cells, dx=1/512, ppc 8 -> 8.0 M particles"""
    return block((100, 400, 25), 1.0 / 512, ppc=8, origin_cells=(16, 8, 16), rho=1600.0, E=1e6, nu=0.3, seed=seed)
