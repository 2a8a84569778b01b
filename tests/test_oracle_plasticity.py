"""Pins the oracle's return mappings with numpy: VonMisesFixedCorotated::projectStrain
(Lib/Ziran/Physics/PlasticityApplier.cpp:94-131) and SnowPlasticity::projectStrain (:16-50), applied by
gridToParticles right after evolveStrain (Lib/MPM/MpmSimulationBase.cpp:1039-1064)."""
import numpy as np

from hot_b200 import scenes


def _sim(oracle, F, E=1e4, nu=0.3):
    n = len(F)
    rng = np.random.default_rng(3)
    X = 0.5 + 0.2 * rng.random((n, 3))
    mu, lam = scenes.lame(E, nu)
    o = oracle.OracleSim(0.05)
    o.set_particles(X, np.zeros((n, 3)), np.ones(n), np.zeros((n, 9)), _cm(F), np.ones(n), np.full(n, mu), np.full(n, lam))
    return o, mu, lam


def _cm(F):      # (n,3,3) -> (n,9) column-major (Eigen default, Forward.h:10-13)
    return np.ascontiguousarray(np.transpose(F, (0, 2, 1)).reshape(len(F), 9))


def _mat(F9):    # (n,9) column-major -> (n,3,3)
    return np.transpose(F9.reshape(-1, 3, 3), (0, 2, 1))


def _rand_F(n, amp, seed):
    rng = np.random.default_rng(seed)
    return np.eye(3)[None] + amp * (rng.random((n, 3, 3)) - 0.5)


def _kirchhoff(s, mu, lam):
    J = s.prod()
    return 2 * mu * (s - 1) * s + lam * (J - 1) * J


def test_von_mises_return_mapping(oracle):
    F = _rand_F(200, 0.6, 1)
    F[0] = np.eye(3)                                     # inside the yield surface: untouched
    F[1] = np.eye(3) * 1.0001
    o, mu, lam = _sim(oracle, F)
    ys = 300.0
    o.set_plasticity("von_mises", [ys])
    o.applyPlasticity()
    Fn = _mat(o.get_particles()["F"])
    moved = 0
    for a, b in zip(F, Fn):
        U, s, Vt = np.linalg.svd(a)
        s = np.maximum(s, 1e-4)
        tau = _kirchhoff(s, mu, lam)
        dev = tau - tau.mean()
        if np.linalg.norm(dev) <= np.sqrt(2.0 / 3.0) * ys:
            np.testing.assert_array_equal(a, b)
            continue
        moved += 1
        sb = np.linalg.svd(b, compute_uv=False)
        # same rotations (polar factors agree), and the projected stress sits ON the yield surface along the old deviator
        np.testing.assert_allclose(b, U @ np.diag(sb) @ Vt, atol=1e-10)
        # the reference solves each principal stretch with the OLD J (PlasticityApplier.cpp:121-126): check that equation
        J = s.prod()
        tau_new = np.sqrt(2.0 / 3.0) * ys / np.linalg.norm(dev) * dev + tau.mean()
        np.testing.assert_allclose(2 * mu * (sb - 1) * sb + lam * (J - 1) * J, tau_new, rtol=1e-9, atol=1e-9 * mu)
    assert moved > 100
    # idempotent on the untouched ones, and a second application keeps moving only by the J lag
    assert np.isfinite(Fn).all()


def test_snow_clamp_and_hardening(oracle):
    F = _rand_F(200, 0.1, 2)
    o, mu, lam = _sim(oracle, F)
    psi, tc, ts, jmin, jmax = 10.0, 2e-2, 7.5e-3, 0.6, 20.0
    o.set_plasticity("snow", [psi, tc, ts, jmin, jmax])
    o.applyPlasticity()
    Fn = _mat(o.get_particles()["F"])
    Jp, mu_n, lam_n = o.get_plastic_state()
    for i, (a, b) in enumerate(zip(F, Fn)):
        U, s, Vt = np.linalg.svd(a)
        sc = np.clip(s, 1 - tc, 1 + ts)
        np.testing.assert_allclose(b, U @ np.diag(sc) @ Vt, atol=1e-12)
        j = np.clip(1.0 * np.linalg.det(a) / sc.prod(), jmin, jmax)
        np.testing.assert_allclose(Jp[i], j, rtol=1e-12)
        h = np.exp(psi * (1.0 - j))
        np.testing.assert_allclose([mu_n[i], lam_n[i]], [mu * h, lam * h], rtol=1e-12)
    # second application: F already inside the box, only Jp bookkeeping (det F / det Fe = 1) -> nothing changes
    o.applyPlasticity()
    np.testing.assert_allclose(_mat(o.get_particles()["F"]), Fn, atol=1e-13)
    np.testing.assert_allclose(o.get_plastic_state()[0], Jp, rtol=1e-12)


def test_g2p_applies_plasticity_after_evolve_strain(oracle):
    sc = scenes.block((4, 3, 3), 0.05, ppc=4, seed=4)
    outs = []
    for model in ("none", "von_mises"):
        o = oracle.OracleSim(sc["dx"])
        o.set_particles(sc["X"], sc["V"] * 40, sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
        o.set_plasticity(model, [1.0])
        o.sortParticlesAndPolluteGrid(); n = o.particlesToGrid()
        o.set_dv(np.zeros((n, 3)))
        o.gridToParticles(2e-3)
        outs.append(o)
    F0, F1 = outs[0].get_particles()["F"], outs[1].get_particles()["F"]
    assert np.abs(F0 - F1).max() > 1e-6                       # yield stress 1 Pa: every particle is projected
    o = outs[0]
    o.set_plasticity("von_mises", [1.0]); o.applyPlasticity()  # same thing by hand
    np.testing.assert_allclose(o.get_particles()["F"], F1, atol=1e-14)


def test_drucker_prager_return_mapping(oracle):
    """Drucker-Prager extension (Klar et al. 2016): Hencky strain eps = log sigma, yield ||dev eps|| + (3 lam + 2 mu)/(2 mu) tr(eps) alpha <= 0.
    Elastic states untouched, expansion goes to the tip (sigma = 1), the rest lands ON the cone along the old deviator."""
    F = _rand_F(300, 0.5, 7)
    F[0] = np.eye(3) * 0.9                                # hydrostatic compression: inside the cone for any friction angle ...
    F[1] = np.diag([0.95, 0.9, 0.93])                     # ... mild shear under pressure: inside the cone
    F[2] = np.eye(3) * 1.2                                # expansion: tip
    o, mu, lam = _sim(oracle, F)
    phi = 30.0
    o.set_plasticity("drucker_prager", [phi, 0.0])
    o.applyPlasticity()
    Fn = _mat(o.get_particles()["F"])
    alpha = np.sqrt(2.0 / 3.0) * 2 * np.sin(np.radians(phi)) / (3 - np.sin(np.radians(phi)))
    k = (3 * lam + 2 * mu) / (2 * mu)
    kinds = {"elastic": 0, "tip": 0, "cone": 0}
    for a, b in zip(F, Fn):
        U, s, Vt = np.linalg.svd(a)
        eps = np.log(s); tr = eps.sum(); dev = eps - tr / 3; nrm = np.linalg.norm(dev)
        sb = np.linalg.svd(b, compute_uv=False)
        if tr >= 0:
            kinds["tip"] += 1
            np.testing.assert_allclose(sb, 1.0, atol=1e-12)
            continue
        y = nrm + k * tr * alpha
        if y <= 0 or nrm < 1e-14:
            kinds["elastic"] += 1
            np.testing.assert_array_equal(a, b)
            continue
        kinds["cone"] += 1
        np.testing.assert_allclose(b, U @ np.diag(np.exp(eps - y * dev / nrm)) @ Vt, atol=1e-11)
        eb = np.log(np.sort(sb)[::-1]); db = eb - eb.sum() / 3
        assert abs(eb.sum() - tr) < 1e-11                                     # volume-preserving plastic flow
        assert abs(np.linalg.norm(db) + k * eb.sum() * alpha) < 1e-10          # on the yield surface
    assert kinds["elastic"] >= 1 and kinds["tip"] >= 1 and kinds["cone"] >= 50
