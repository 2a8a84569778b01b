"""Deterministic synthetic inputs shaped like the scenes BASELINE.json names.

The reference samples particles from a Poisson-disk tile file (Lib/Ziran/Math/Geometry/PoissonDisk.h:185-222)
that cannot travel to the GPU box, so the generator here is a stratified jitter with a fixed numpy seed:
every cell of the solid gets exactly `ppc` particles.  Material constants follow
CorotatedIsotropic (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h:69-73) and
MpmInitializationHelper.h:367-377 (vol = volume/N, m = rho*vol).
"""
import numpy as np


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


# BASELINE.json configs restated
def config_c1(seed=0):
    """box drop: 18^3 cells, dx=1/64, ppc 8, rho 1000, E 2.5e4, nu .4 -> 46 656 particles"""
    return block((18, 18, 18), 1.0 / 64, ppc=8, origin_cells=(20, 8, 20), rho=1000.0, E=2.5e4, nu=0.4, seed=seed)


def config_c2(seed=0):
    """twisting bar: 0.12 x 0.9 x 0.12 at dx = 0.12/22, ppc 12 -> 22 x 165 x 22 cells, 958 320 particles"""
    return block((22, 165, 22), 0.12 / 22, ppc=12, origin_cells=(16, 16, 16), rho=2000.0, E=1e5, nu=0.3, seed=seed)


def config_c4(seed=0):
    """column 100 x 400 x 25 cells, dx=1/512, ppc 8 -> 8.0 M particles"""
    return block((100, 400, 25), 1.0 / 512, ppc=8, origin_cells=(16, 8, 16), rho=1600.0, E=1e6, nu=0.3, seed=seed)
