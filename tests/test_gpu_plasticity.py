"""GPU parity of the return mappings run by hot_g2p after evolveStrain (Lib/MPM/MpmSimulationBase.cpp:1039-1064;
Lib/Ziran/Physics/PlasticityApplier.cpp:16-50 snow, :94-131 von Mises) against the oracle, through the C ABI.
Tolerance: the projected F is rebuilt from an SVD whose Jacobi sweeps run with fused multiply-adds on the device, so F,
Jp and the hardened Lame parameters compare at 1e-11 relative (fp64)."""
import numpy as np
import pytest

from hot_b200 import scenes

pytestmark = pytest.mark.gpu

MODELS = {"von_mises": [50.0], "snow": [10.0, 2e-2, 7.5e-3, 0.6, 20.0], "drucker_prager": [30.0, 0.0]}   # DP: extension (hot_b200.h)


def _pair(hot, oracle, sc, vscale):
    g = hot.MpmSimulationB200(sc["dx"]); o = oracle.OracleSim(sc["dx"])
    for s in (g, o):
        s.set_particles(sc["X"], sc["V"] * vscale, sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    return g, o


@pytest.mark.parametrize("model", list(MODELS))
def test_plasticity_in_time_steps(hot, oracle, model):
    sc = scenes.block((6, 5, 4), 0.04, ppc=6, seed=12)
    g, o = _pair(hot, oracle, sc, 25.0)
    for s in (g, o):
        s.set_plasticity(model, MODELS[model])
    for step in range(3):                      # state (F, Jp, mu, lambda) is carried through re-sorts
        for s in (g, o):
            s.sortParticlesAndPolluteGrid(); n = s.particlesToGrid()
            s.set_dv(np.zeros((n, 3)))
            s.gridToParticles(2e-3)
        pg, po = g.get_particles(), o.get_particles()
        np.testing.assert_allclose(pg["F"], po["F"], rtol=0, atol=1e-11, err_msg=f"F step {step}")
        for a, b, name in zip(g.get_plastic_state(), o.get_plastic_state(), ("Jp", "mu", "lambda")):
            np.testing.assert_allclose(a, b, rtol=1e-11, err_msg=f"{name} step {step}")
    if model == "snow":
        assert np.abs(g.get_plastic_state()[0] - 1.0).max() > 1e-6      # the test exercised hardening
    else:
        g0, _ = _pair(hot, oracle, sc, 25.0)
        g0.sortParticlesAndPolluteGrid(); n = g0.particlesToGrid(); g0.set_dv(np.zeros((n, 3))); g0.gridToParticles(2e-3)
        g1, _ = _pair(hot, oracle, sc, 25.0)
        g1.set_plasticity(model, MODELS[model])
        g1.sortParticlesAndPolluteGrid(); g1.particlesToGrid(); g1.set_dv(np.zeros((n, 3))); g1.gridToParticles(2e-3)
        assert np.abs(g0.get_particles()["F"] - g1.get_particles()["F"]).max() > 1e-8   # projection happened


def test_apply_plasticity_standalone_and_errors(hot, oracle):
    sc = scenes.block((3, 3, 3), 0.05, ppc=4, seed=2)
    rng = np.random.default_rng(0)
    sc["F"] = np.eye(3).reshape(1, 9) + 0.5 * (rng.random(sc["F"].shape) - 0.5)
    g, o = _pair(hot, oracle, sc, 1.0)
    with pytest.raises(hot.HotError):
        g.set_plasticity(7, [1.0])
    with pytest.raises(hot.HotError):
        g.set_plasticity("von_mises", [-1.0])
    for s in (g, o):
        s.set_plasticity("von_mises", [100.0]); s.applyPlasticity()   # before any sort: original order
    np.testing.assert_allclose(g.get_particles(gradV=False)["F"], o.get_particles()["F"], rtol=0, atol=1e-11)   # no G2P yet: gradV undefined
