"""Pins the oracle's assembled-matrix / Galerkin-multigrid restatement (a15-a20) with the reference's own checks
(matrixSanityCheck: assembled == matrix-free, 1e-10, ImplicitSolver.h:698-739; symmetricSanityCheck / SPDSanityCheck,
SquareMatrix.h:84-194) and with independent scipy.sparse linear algebra (dense-free R A P, triangular solves for the
coloured block Gauss-Seidel, a textbook PCG)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from hot_b200 import scenes


def ell_to_csr(col, val, ncols):
    """fixed-width rows of 3x3 blocks (column-major) -> scalar CSR"""
    n, cs = col.shape
    rows = np.repeat(np.arange(n), cs)
    blocks = val.reshape(n * cs, 3, 3).transpose(0, 2, 1)            # -> [r][c]
    A = sp.bsr_matrix((blocks, col.reshape(-1), np.arange(0, n * cs + 1, cs)), shape=(3 * n, 3 * ncols))
    A = A.tocsr(); A.sum_duplicates()
    return A


def _setup(oracle, cells=(5, 6, 5), project=True, bc=True, seed=4, dt=2e-3):
    sc = scenes.block(cells, 0.04, ppc=6, seed=seed, E=1e4)
    o = oracle.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.set_dt_gravity(dt, (0, -9.8, 0))
    o.set_project(project)
    o.sortParticlesAndPolluteGrid()
    o.particlesToGrid()
    o.backupStrain()
    coord = o.get_id2coord()
    bcn = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32) if bc else np.zeros(0, dtype=np.int32)
    o.set_bc(bcn, P=np.zeros((len(bcn), 9)), dv_bc=np.zeros((len(bcn), 3)))
    rng = np.random.default_rng(seed)
    o.updateState(o.get_dv() + 0.2 * (rng.random((o.num_nodes, 3)) - 0.5))
    return sc, o, bcn


def test_assembled_matrix_equals_matrix_free(oracle):
    sc, o, _ = _setup(oracle, bc=False)
    o.buildMatrix(bcproject=False)
    n = o.num_nodes
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.random((n, 3)) - 0.5
        a, b = o.spmv(0, x), o.multiply(x)
        assert np.abs(a - b).max() < 1e-10 * np.abs(b).max()        # matrixSanityCheck tolerance
    col, val = o.get_matrix()
    A = ell_to_csr(col, val, n)
    assert abs(A - A.T).max() < 1e-10 * abs(A).max()                # symmetricSanityCheck
    w = np.linalg.eigvalsh(A.toarray())
    assert w[0] > 0                                                  # SPDSanityCheck (projected dPdF + mass)
    # slot addressing: col of slot (dx+2)*25+(dy+2)*5+(dz+2) is the node at coord_i - d (ImplicitSolver.h:465-468,538)
    coord = o.get_id2coord()
    nz = np.abs(val).max(axis=2) > 0
    i, s = np.nonzero(nz)
    d = np.stack([s // 25 - 2, (s // 5) % 5 - 2, s % 5 - 2], 1)
    assert (coord[i] - coord[col[i, s]] == d).all()


def test_bc_projected_system(oracle):
    sc, o, bc = _setup(oracle)
    o.buildMatrix(bcproject=True)
    n = o.num_nodes
    col, val = o.get_matrix()
    A = ell_to_csr(col, val, n).toarray()
    dofs = (3 * bc[:, None] + np.arange(3)).reshape(-1)
    assert np.allclose(A[dofs][:, dofs], np.eye(len(dofs)))          # sticky: identity rows / columns
    mask = np.ones(3 * n, bool); mask[dofs] = False
    assert np.abs(A[dofs][:, mask]).max() == 0 and np.abs(A[mask][:, dofs]).max() == 0
    assert np.allclose(A, A.T, atol=1e-10 * np.abs(A).max())
    assert (col >= 0).all() and (col < n).all()                      # padding aliases node 0 / 1 with zero blocks
    # free-free block equals the unprojected operator
    x = np.zeros((n, 3)); x[np.setdiff1d(np.arange(n), bc)] = np.random.default_rng(1).random((n - len(bc), 3))
    y = o.multiply(x); y[bc] = 0
    assert np.abs(o.spmv(0, x) - y).max() < 1e-10 * np.abs(y).max()


def test_block_jacobi_diagonal(oracle):
    sc, o, _ = _setup(oracle, bc=False)
    o.buildMatrix(bcproject=False)
    col, val = o.get_matrix()
    n = o.num_nodes
    Dinv = o.buildDiagonal(Ainv=1).reshape(n, 3, 3).transpose(0, 2, 1)
    D = val[:, 62].reshape(n, 3, 3).transpose(0, 2, 1)
    assert np.allclose(Dinv @ D, np.eye(3), atol=1e-9)


@pytest.fixture(scope="module")
def mg(oracle):
    sc, o, bc = _setup(oracle, cells=(9, 10, 9))
    o.buildMatrix(bcproject=True)
    o.buildMultigrid(levels=3, smoother=5, coarseSolver=2, Ainv=1)
    return o, bc


def test_hierarchy_transfer_operators(mg):
    o, _ = mg
    dofs = o.level_dofs()
    assert len(dofs) == 3 and dofs[0] == o.num_nodes and dofs[0] > dofs[1] > dofs[2] > 0
    for l in range(2):
        cf, cc = o.level_coords(l), o.level_coords(l + 1)
        pc, pv = o.level_matrix(l, 1)
        assert pc.shape[1] == 8
        w = pv[:, :, 0]
        assert np.allclose(w.sum(1), 1.0)                            # trilinear weights: partition of unity
        assert set(np.unique(w)) <= {0.0, 0.125, 0.25, 0.5, 1.0}
        nzr, nzs = np.nonzero(w)
        # slot (a,b,c) of fine node x addresses coarse node x//2 + (a,b,c)
        off = np.stack([nzs // 4, (nzs // 2) % 2, nzs % 2], 1)
        assert (cc[pc[nzr, nzs]] == cf[nzr] // 2 + off).all()
        # coarse ids are assigned in first-touch order (MultigridPreconditioner.h:630-668)
        first = {}
        for r_, s_ in zip(nzr, nzs):
            first.setdefault(int(pc[r_, s_]), len(first))
        assert all(k == v for k, v in first.items())
        P = ell_to_csr(pc, pv, dofs[l + 1])
        rc, rv = o.level_matrix(l, 2)
        assert rc.shape[1] <= 27
        R = ell_to_csr(rc, rv, dofs[l])
        assert abs(R - P.T).max() == 0
        x = np.random.default_rng(l).random((dofs[l], 3))
        assert np.allclose(o.restrict(l, x).reshape(-1), P.T @ x.reshape(-1), rtol=1e-13)
        y = np.random.default_rng(l).random((dofs[l + 1], 3))
        assert np.allclose(o.prolong(l, y).reshape(-1), P @ y.reshape(-1), rtol=1e-13)


def test_galerkin_coarse_matrices(mg):
    o, _ = mg
    dofs = o.level_dofs()
    for l in range(2):
        A = ell_to_csr(*o.level_matrix(l, 0), dofs[l])
        P = ell_to_csr(*o.level_matrix(l, 1), dofs[l + 1])
        col, val = o.level_matrix(l + 1, 0)
        # the structural width can exceed 125 (zero-valued aliases of the padding column, SquareMatrix.h:565-569), the
        # non-zero coarse stencil stays within 5^3
        cc = o.level_coords(l + 1)
        i_, s_ = np.nonzero(np.abs(val).max(axis=2) > 0)
        assert np.abs(cc[i_] - cc[col[i_, s_]]).max() <= 2
        Ac = ell_to_csr(col, val, dofs[l + 1])
        ref = (P.T @ A @ P).tocsr()
        assert abs(Ac - ref).max() < 1e-12 * abs(ref).max()
        D, Di = o.level_diagonal(l + 1)
        n = dofs[l + 1]
        Dm = D.reshape(n, 3, 3).transpose(0, 2, 1)
        blocks = np.stack([ref[3 * i:3 * i + 3, 3 * i:3 * i + 3].toarray() for i in range(n)])
        assert np.allclose(Dm, blocks, rtol=1e-12, atol=1e-12 * np.abs(blocks).max())
        assert np.allclose(Di.reshape(n, 3, 3).transpose(0, 2, 1) @ Dm, np.eye(3), atol=1e-9)


def _gs_reference(A, order, r):
    """one symmetric block-GS iteration with the node sequence `order` (colour, block, position) via triangular solves"""
    n = len(order)
    perm = np.lexsort((order[:, 2], order[:, 1], order[:, 0]))
    dperm = (3 * perm[:, None] + np.arange(3)).reshape(-1)
    Ap = A[dperm][:, dperm].tocsr()
    Dblk = sp.block_diag([Ap[3 * i:3 * i + 3, 3 * i:3 * i + 3].toarray() for i in range(n)], format="csr")
    strictL = sp.tril(Ap - Dblk, k=-1, format="csr")
    strictU = sp.triu(Ap - Dblk, k=1, format="csr")
    rp = r.reshape(-1)[dperm]
    hdu = spla.spsolve((Dblk + strictL).tocsc(), rp)
    hdu = Dblk @ hdu
    du = spla.spsolve((Dblk + strictU).tocsc(), hdu)
    out = np.empty_like(du); out[dperm] = du
    return out.reshape(-1, 3)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_gs_smoother_vs_triangular_solves(mg, level):
    o, _ = mg
    dofs = o.level_dofs()
    n = dofs[level]
    A = ell_to_csr(*o.level_matrix(level, 0), n)
    order = o.color_order(level)
    coord = o.level_coords(level)
    # colouring rule: colour from the parity of the 4^3 block coordinate, position = 1-based first-seen rank
    b = coord >> 2
    assert (order[:, 0] == ((b[:, 0] & 1) << 2 | (b[:, 1] & 1) << 1 | (b[:, 2] & 1))).all()
    rng = np.random.default_rng(level)
    r0 = rng.random((n, 3)) - 0.5
    u, r = o.smooth(level, 5, np.zeros((n, 3)), r0, iterations=1)
    du = _gs_reference(A, order, r0)
    assert np.abs(u - du).max() < 1e-11 * np.abs(du).max()
    assert np.abs(r - (r0.reshape(-1) - A @ du.reshape(-1)).reshape(n, 3)).max() < 1e-11 * np.abs(r0).max()
    # iterations = (times+1)>>1: 2 -> one sweep, 3 -> two sweeps
    u2, _ = o.smooth(level, 5, np.zeros((n, 3)), r0, iterations=2)
    assert np.array_equal(u2, u)
    # a symmetric GS sweep reduces the energy norm of the error
    x = spla.spsolve(A.tocsc(), r0.reshape(-1))
    e0, e1 = x, x - u.reshape(-1)
    assert e1 @ (A @ e1) < e0 @ (A @ e0)


def test_pcg_coarse_solver_vs_textbook(mg):
    o, _ = mg
    dofs = o.level_dofs()
    level = 2
    n = dofs[level]
    A = ell_to_csr(*o.level_matrix(level, 0), n)
    _, Di = o.level_diagonal(level)
    Minv = sp.block_diag(list(Di.reshape(n, 3, 3).transpose(0, 2, 1)), format="csr")
    rng = np.random.default_rng(7)
    r0 = rng.random((n, 3)) - 0.5
    u, r = o.smooth(level, 2, np.zeros((n, 3)), r0, iterations=10000, initial_residual=r0)
    # textbook Jacobi-PCG stopped at z^T r < 0.25 z0^T r0 (MultigridPreconditioner.h:190-226)
    x = np.zeros(3 * n); rr = r0.reshape(-1).copy(); z = Minv @ rr; p = z.copy(); zr = z @ rr; tol = 0.25 * zr; it = 0
    while zr >= tol:
        q = A @ p; a = zr / (q @ p); x += a * p; rr -= a * q; z = Minv @ rr; zr_new = z @ rr; p = z + (zr_new / zr) * p; zr = zr_new; it += 1
    assert it >= 1
    assert np.abs(u.reshape(-1) - x).max() < 1e-11 * np.abs(x).max()
    assert np.abs(r.reshape(-1) - rr).max() < 1e-11 * np.abs(r0).max()


def test_jacobi_smoothers(mg):
    o, _ = mg
    n = o.level_dofs()[1]
    A = ell_to_csr(*o.level_matrix(1, 0), n)
    _, Di = o.level_diagonal(1)
    Minv = sp.block_diag(list(Di.reshape(n, 3, 3).transpose(0, 2, 1)), format="csr")
    r0 = np.random.default_rng(9).random((n, 3)) - 0.5
    u, r = o.smooth(1, 0, np.zeros((n, 3)), r0, iterations=2)
    x = np.zeros(3 * n); rr = r0.reshape(-1).copy()
    for _ in range(2):
        du = 0.1 * (Minv @ rr); x += du; rr -= A @ du
    assert np.allclose(u.reshape(-1), x, rtol=1e-12, atol=1e-14 * np.abs(x).max())
    u, r = o.smooth(1, 1, np.zeros((n, 3)), r0, iterations=2)
    x = np.zeros(3 * n); rr = r0.reshape(-1).copy()
    for _ in range(2):
        du = Minv @ rr; q = A @ du; w = (du @ rr) / (du @ q); x += w * du; rr -= w * q
    assert np.allclose(u.reshape(-1), x, rtol=1e-11, atol=1e-13 * np.abs(x).max())


def test_chebyshev_smoother_and_2norm_estimate(mg):
    """estimate2norm (SquareMatrix.h:375-475) against scipy's largest eigenvalue; chebyshev_smooth (MultigridPreconditioner.h:227-264)
    against the recurrence written out in numpy"""
    import scipy.sparse.linalg as spla
    o, _ = mg
    n = o.level_dofs()[1]
    A = ell_to_csr(*o.level_matrix(1, 0), n)
    lmax, lmin = o.estimate2norm(1)
    exact = spla.eigsh(A, k=1, which="LA", return_eigenvectors=False)[0]
    assert abs(lmax - exact) <= 2e-3 * exact and lmin == lmax / 30           # power iteration stopped at |de| <= 1e-6 e
    _, Di = o.level_diagonal(1)
    Minv = sp.block_diag(list(Di.reshape(n, 3, 3).transpose(0, 2, 1)), format="csr")
    r0 = np.random.default_rng(19).random((n, 3)) - 0.5
    u, r = o.smooth(1, 6, np.zeros((n, 3)), r0, iterations=4)
    d, c = (lmax + lmin) / 2, (lmax - lmin) / 2
    x = np.zeros(3 * n); rr = r0.reshape(-1).copy()
    p = Minv @ rr; alpha = 1 / d; du = p.copy(); q = A @ du; x += alpha * du; rr -= alpha * q
    for cnt in range(1, 4):
        p = Minv @ rr
        beta = 0.5 * c * c * alpha * alpha * (0.5 if cnt > 1 else 1.0)
        alpha = 1 / (d - beta / alpha)
        du = p + beta * du; q = A @ du; x += alpha * du; rr -= alpha * q
    assert np.allclose(u.reshape(-1), x, rtol=1e-11, atol=1e-13 * np.abs(x).max())
    assert np.allclose(r.reshape(-1), rr, rtol=1e-11, atol=1e-13 * np.abs(r0).max())


def test_vcycle_is_spd_preconditioner(oracle):
    sc, o, bc = _setup(oracle, cells=(9, 10, 9))
    o.buildMatrix(bcproject=True)
    o.buildMultigrid(levels=3, smoother=5, coarseSolver=5)            # all-GS V-cycle is a fixed linear operator
    n = o.num_nodes
    A = ell_to_csr(*o.level_matrix(0, 0), n)
    rng = np.random.default_rng(3)
    x, y = rng.random((n, 3)) - 0.5, rng.random((n, 3)) - 0.5
    Mx, My = o.vcycle(x), o.vcycle(y)
    assert abs((y * Mx).sum() - (x * My).sum()) < 1e-9 * abs((y * Mx).sum())   # checkPreconditioningMatrix (LBFGS.h:95-175)
    assert (x * Mx).sum() > 0
    assert np.allclose(o.vcycle(2 * x), 2 * Mx, rtol=1e-10, atol=1e-12 * np.abs(Mx).max())
    # HOT configuration (GS + coarse PCG): one V-cycle contracts the error in the energy norm
    o.buildMultigrid(levels=3, smoother=5, coarseSolver=2)
    b = rng.random((n, 3)) - 0.5
    sol = spla.spsolve(A.tocsc(), b.reshape(-1))
    e = sol - o.vcycle(b).reshape(-1)
    assert e @ (A @ e) < 0.2 * (sol @ (A @ sol))
    t, cg_it = o.vcycle_timing()
    assert t[:3, 0].min() > 0 and cg_it >= 0   # the coarse PCG may exit at once: its target is 0.25 z0.r0 of the RESTRICTED INITIAL residual
