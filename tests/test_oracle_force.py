"""Pins the oracle's constitutive model / force / Hessian restatement (a9-a14) with numpy.linalg and with the
reference's own notion of correctness: the finite-difference test of Lib/Ziran/Sim/DiffTest.h:19-138
(energy <-> residual <-> Hessian apply) and the BC-zero check of ImplicitSolver.h:284-296."""
import numpy as np
import pytest

from hot_b200 import scenes

MU, LAM = scenes.lame(1e4, 0.3)


def _psi_P_numpy(F):
    U, s, Vt = np.linalg.svd(F)
    if np.linalg.det(U) < 0:
        U[:, 2] *= -1; s[2] *= -1
    if np.linalg.det(Vt) < 0:
        Vt[2] *= -1; s[2] *= -1
    R = U @ Vt
    J = np.linalg.det(F)
    psi = MU * ((F - R) ** 2).sum() + 0.5 * LAM * (J - 1) ** 2
    P = 2 * MU * (F - R) + LAM * (J - 1) * J * np.linalg.inv(F).T
    return psi, P, s


def _cases():
    rng = np.random.default_rng(0)
    out = [np.eye(3), np.diag([1.2, 1.2, 0.8]), np.diag([2.0, 1.0, 0.5])]
    for _ in range(40):
        out.append(np.eye(3) + 0.4 * (rng.random((3, 3)) - 0.5))
    for _ in range(10):
        out.append(2.0 * (rng.random((3, 3)) - 0.5))          # large deformation, some inverted
    Q, _ = np.linalg.qr(rng.random((3, 3)))
    out.append(Q * np.linalg.det(Q))                            # pure rotation
    out.append(Q @ np.diag([1.3, 1.3, 0.7]) @ Q.T)              # repeated singular values
    return out


def test_svd_convention_and_reconstruction(oracle):
    for F in _cases():
        c = oracle.constitutive(F, MU, LAM)
        U, s, V = c["U"], c["sigma"], c["V"]
        np.testing.assert_allclose(U @ np.diag(s) @ V.T, F, atol=1e-13)
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-13)
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-13)
        assert np.linalg.det(U) > 0 and np.linalg.det(V) > 0
        assert s[0] >= s[1] - 1e-13 and s[1] >= abs(s[2]) - 1e-13   # ImplicitQRSVD.h:256-352 ordering
        np.testing.assert_allclose(np.sort(np.abs(s)), np.sort(np.linalg.svd(F, compute_uv=False)), rtol=1e-12, atol=1e-14)


def test_psi_and_first_piola_vs_numpy(oracle):
    for F in _cases():
        c = oracle.constitutive(F, MU, LAM)
        psi, P, _ = _psi_P_numpy(F)
        np.testing.assert_allclose(c["psi"], psi, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(c["P"], P, rtol=1e-10, atol=1e-8 * MU)


def test_differential_is_derivative_of_P_unprojected(oracle):
    rng = np.random.default_rng(1)
    for F in _cases()[3:43]:
        dF = rng.random((3, 3)) - 0.5
        c = oracle.constitutive(F, MU, LAM, project=False, dF=dF)
        h = 1e-6
        fd = (_psi_P_numpy(F + h * dF)[1] - _psi_P_numpy(F - h * dF)[1]) / (2 * h)
        np.testing.assert_allclose(c["dP"], fd, rtol=2e-6, atol=2e-6 * MU)
        vec = lambda M: M.T.reshape(9)                            # index i + 3 j
        np.testing.assert_allclose(c["dPdF"] @ vec(dF), vec(c["dP"]), rtol=1e-11, atol=1e-9 * MU)
        np.testing.assert_allclose(c["dPdF"], c["dPdF"].T, atol=1e-9 * MU)


def test_projected_hessian_is_psd_and_consistent(oracle):
    rng = np.random.default_rng(2)
    for F in _cases():
        dF = rng.random((3, 3)) - 0.5
        cp = oracle.constitutive(F, MU, LAM, project=True, dF=dF)
        cu = oracle.constitutive(F, MU, LAM, project=False, dF=dF)
        wp = np.linalg.eigvalsh(cp["dPdF"]); wu = np.linalg.eigvalsh(cu["dPdF"])
        assert wp.min() > -1e-8 * abs(wp).max()
        vec = lambda M: M.T.reshape(9)
        np.testing.assert_allclose(cp["dPdF"] @ vec(dF), vec(cp["dP"]), rtol=1e-11, atol=1e-9 * MU)
        if wu.min() > 1e-6 * abs(wu).max():                      # already PD: projection is the identity map
            np.testing.assert_allclose(cp["dPdF"], cu["dPdF"], atol=1e-9 * abs(wu).max())
        else:                                                    # projection = clamp of the spectrum
            np.testing.assert_allclose(np.sort(wp), np.sort(np.maximum(wu, 0)), atol=1e-8 * abs(wu).max())


def _setup(oracle, project=True, seed=4, cells=(4, 5, 4), gravity=(0, -9.8, 0), dt=2e-3):
    sc = scenes.block(cells, 0.04, ppc=6, seed=seed, E=1e4)
    o = oracle.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    o.set_dt_gravity(dt, gravity)
    o.set_project(project)
    o.sortParticlesAndPolluteGrid()
    o.particlesToGrid()
    o.backupStrain()
    return sc, o


def test_diff_test_energy_residual_hessian(oracle):
    """DiffTest.h: residual = -dE/d(dv), multiply = -d(residual)/d(dv) (unprojected Hessian, no BC)."""
    sc, o = _setup(oracle, project=False)
    n = o.num_nodes
    rng = np.random.default_rng(123)
    o.set_bc(np.zeros(0, dtype=np.int32))
    dv0 = o.get_dv() + 0.3 * (rng.random((n, 3)) - 0.5)
    d = rng.random((n, 3)) - 0.5
    errs_e, errs_h = [], []
    for h in (1e-3, 5e-4):
        ep = o.updateState(dv0 + h * d); rp = o.computeResidual()
        em = o.updateState(dv0 - h * d); rm = o.computeResidual()
        o.updateState(dv0); r0 = o.computeResidual(); Ad = o.multiply(d)
        errs_e.append(abs((ep - em) / (2 * h) + (r0 * d).sum()) / abs((r0 * d).sum()))
        errs_h.append(np.abs((rp - rm) / (2 * h) + Ad).max() / np.abs(Ad).max())
    assert errs_e[0] < 1e-5 and errs_h[0] < 1e-4
    assert errs_e[1] < errs_e[0] * 0.5 or errs_e[1] < 1e-8       # second-order convergence
    assert errs_h[1] < errs_h[0] * 0.5 or errs_h[1] < 1e-8


def test_matrix_free_apply_is_symmetric_and_spd_when_projected(oracle):
    sc, o = _setup(oracle, project=True)
    n = o.num_nodes
    rng = np.random.default_rng(5)
    o.set_bc(np.zeros(0, dtype=np.int32))
    o.updateState(o.get_dv() + 0.5 * (rng.random((n, 3)) - 0.5))
    x, y = rng.random((n, 3)) - 0.5, rng.random((n, 3)) - 0.5
    Ax, Ay = o.multiply(x), o.multiply(y)
    assert abs((y * Ax).sum() - (x * Ay).sum()) < 1e-10 * abs((y * Ax).sum())   # SquareMatrix.h:84-109 symmetry
    assert (x * Ax).sum() > 0                                                    # PD sanity, SquareMatrix.h:129-194


def test_bc_projection_zeroes_residual(oracle):
    sc, o = _setup(oracle)
    coord = o.get_id2coord()
    bc = np.nonzero(coord[:, 1] <= coord[:, 1].min() + 1)[0].astype(np.int32)    # sticky floor nodes
    o.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
    dv = o.get_dv()
    assert (dv[bc] == 0).all() and np.allclose(np.delete(dv, bc, 0), np.array([0, -9.8, 0]) * 2e-3)
    o.updateState()
    r = o.computeResidual()
    assert np.abs(r[bc]).max() < 1e-10                                           # ImplicitSolver.h:284-296
    assert np.abs(np.delete(r, bc, 0)).max() > 0


def test_cn_tolerance_formula(oracle):
    sc, o = _setup(oracle)
    eps, dt = 1e-7, 2e-3
    tol = o.evaluatePerNodeCNTolerance(eps, dt)
    # uniform material: sum_p w m_p = m_i, so tol_i = ||dPdF(I)||_F * eps * 24 dx^2 dt (ImplicitSolver.h:667-696)
    H = oracle.constitutive(np.eye(3), sc["mu"][0], sc["lam"][0])["dPdF"]
    np.testing.assert_allclose(tol, np.linalg.norm(H) * eps * 24 * sc["dx"] ** 2 * dt, rtol=1e-10)


# ---- neo-Hookean extension (hot_set_constitutive_model 1): closed form + finite differences ---------------------------------
def _nh_numpy(F):
    J = np.linalg.det(F)
    FinvT = np.linalg.inv(F).T
    psi = 0.5 * MU * ((F * F).sum() - 3) - MU * np.log(J) + 0.5 * LAM * np.log(J) ** 2
    P = MU * (F - FinvT) + LAM * np.log(J) * FinvT
    return psi, P


def test_neo_hookean_extension_vs_numpy_and_finite_differences(oracle):
    rng = np.random.default_rng(5)
    oracle.set_constitutive_model_global(1)
    try:
        for F in [F for F in _cases()[:43] if np.linalg.det(F) > 0.2]:
            dF = rng.random((3, 3)) - 0.5
            c = oracle.constitutive(F, MU, LAM, project=False, dF=dF)
            psi, P = _nh_numpy(F)
            np.testing.assert_allclose(c["psi"], psi, rtol=1e-11, atol=1e-9)
            np.testing.assert_allclose(c["P"], P, rtol=1e-10, atol=1e-8 * MU)
            h = 1e-6
            fd = (_nh_numpy(F + h * dF)[1] - _nh_numpy(F - h * dF)[1]) / (2 * h)
            np.testing.assert_allclose(c["dP"], fd, rtol=5e-6, atol=5e-6 * MU)
            vec = lambda M: M.T.reshape(9)
            np.testing.assert_allclose(c["dPdF"] @ vec(dF), vec(c["dP"]), rtol=1e-11, atol=1e-9 * MU)
            np.testing.assert_allclose(c["dPdF"], c["dPdF"].T, atol=1e-9 * MU)
            cp = oracle.constitutive(F, MU, LAM, project=True, dF=dF)
            w = np.linalg.eigvalsh(cp["dPdF"])
            assert w.min() > -1e-8 * abs(w).max()
    finally:
        oracle.set_constitutive_model_global(0)
