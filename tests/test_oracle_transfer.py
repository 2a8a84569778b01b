"""CPU checks of the oracle's sort / P2G / G2P restatement against independent numpy evaluation and the
conservation properties the transfers must satisfy (the reference ships no golden vectors for them)."""
import numpy as np

from hot_b200 import scenes


def _numpy_p2g(sc):
    """Dense, order-independent evaluation of A.4 with numpy (independent of the oracle's page machinery)."""
    dx = sc["dx"]; X = sc["X"]; inv = 1.0 / dx
    xi = X * inv
    base = np.floor(xi - 0.5).astype(np.int64)
    d0 = xi - base
    w = np.stack([0.5 * (1.5 - d0) ** 2, 0.75 - (d0 - 1) ** 2, 0.5 * (d0 - 0.5) ** 2], -1)  # n,3(axis),3
    lo = base.min(0); span = base.max(0) - lo + 3
    m = np.zeros(span); mv = np.zeros(tuple(span) + (3,))
    Cm = sc["C"].reshape(-1, 3, 3).transpose(0, 2, 1) * sc["mass"][:, None, None]  # column-major -> [r][c]
    for i in range(3):
        for j in range(3):
            for k in range(3):
                ww = w[:, 0, i] * w[:, 1, j] * w[:, 2, k]
                node = base + np.array([i, j, k])
                d = node * dx - X
                contrib = (np.einsum("nrc,nc->nr", Cm, d) + sc["mass"][:, None] * sc["V"]) * ww[:, None]
                ix = tuple((node - lo).T)
                np.add.at(m, ix, ww * sc["mass"])
                np.add.at(mv, ix, contrib)
    return lo, m, mv


def test_sort_keys_groups_pages(oracle):
    sc = scenes.block((6, 5, 7), 0.02, ppc=4, seed=1)
    o = oracle.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.sortParticlesAndPolluteGrid()
    sorter, order, base_off = o.get_sort()
    assert (np.diff(sorter.astype(np.int64)) > 0).all()
    assert sorted(order.tolist()) == list(range(o.N))
    base = np.floor(sc["X"] / sc["dx"] - 0.5).astype(np.int32)
    assert (oracle.linear_offset(base) == base_off).all()
    first, last, blk = o.get_groups()
    assert first[0] == 0 and last[-1] == o.N - 1 and (first[1:] == last[:-1] + 1).all()
    assert (np.diff(blk.astype(np.int64)) > 0).all()
    pages = o.get_pages()
    assert len(set(pages.tolist())) == len(pages)
    assert (pages == oracle.activate(blk << np.uint64(12))).all()


def test_p2g_matches_dense_numpy(oracle):
    sc = scenes.block((5, 6, 4), 0.03, ppc=6, seed=2)
    o = oracle.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.sortParticlesAndPolluteGrid()
    n = o.particlesToGrid()
    idx, m, v = o.get_grid()
    coord = o.get_id2coord()
    lo, dm, dmv = _numpy_p2g(sc)
    assert n == int((dm != 0).sum())
    act = idx >= 0
    c = coord[idx[act]] - lo
    np.testing.assert_allclose(m[act], dm[tuple(c.T)], rtol=1e-13)
    np.testing.assert_allclose(v[act], dmv[tuple(c.T)] / dm[tuple(c.T)][:, None], rtol=1e-11, atol=1e-13)
    # conservation: mass and momentum
    np.testing.assert_allclose(m.sum(), sc["mass"].sum(), rtol=1e-13)
    np.testing.assert_allclose((m[:, None] * v).sum(0), (sc["mass"][:, None] * sc["V"]).sum(0) + 0 * v.sum(0), rtol=1e-9, atol=1e-12)
    # DOF ids are a dense numbering in page-list x element order
    assert (idx[act] == np.arange(n)).all()
    assert (o.buildMassMatrix() == m[act]).all()


def test_g2p_reproduces_affine_field(oracle):
    """APIC round trip: a grid velocity field that is affine in x is interpolated exactly (v_p, C_p = grad v)."""
    sc = scenes.block((4, 4, 4), 0.05, ppc=5, seed=3, perturb=False)
    o = oracle.OracleSim(sc["dx"], apic_rpic_ratio=1.0)
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.sortParticlesAndPolluteGrid()
    n = o.particlesToGrid()
    coord = o.get_id2coord()
    A = np.array([[0.1, -0.3, 0.2], [0.05, 0.2, -0.1], [0.4, 0.0, -0.25]]); b = np.array([0.3, -0.2, 0.1])
    dv = (coord * sc["dx"]) @ A.T + b  # v = 0 after P2G, so new_v = dv
    o.set_dv(dv)
    dt = 1e-3
    flags = o.gridToParticles(dt)
    out = o.get_particles()
    np.testing.assert_allclose(out["V"], sc["X"] @ A.T + b, rtol=1e-11, atol=1e-12)
    G = out["gradV"].reshape(-1, 3, 3).transpose(0, 2, 1)
    np.testing.assert_allclose(G, np.broadcast_to(A, G.shape), rtol=1e-9, atol=1e-10)
    Cp = out["C"].reshape(-1, 3, 3).transpose(0, 2, 1)
    np.testing.assert_allclose(Cp, np.broadcast_to(A, Cp.shape), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(out["X"], sc["X"] + dt * out["V"], rtol=1e-14)
    Fn = out["F"].reshape(-1, 3, 3).transpose(0, 2, 1)
    np.testing.assert_allclose(Fn, np.broadcast_to(np.eye(3) + dt * A, Fn.shape), rtol=1e-10, atol=1e-12)
    assert flags == (0, 0)
