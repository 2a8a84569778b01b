"""Pins row f2 (the return mappings at the applyPlasticity hook, Lib/MPM/MpmSimulationBase.cpp:1044-1064) to the REFERENCE'S OWN code:
tests/golden/plasticity_ref.npz was produced by SnowPlasticity<double>::projectStrain and VonMisesFixedCorotated<double,3>::projectStrain of
Lib/Ziran/Physics/PlasticityApplier.cpp, compiled where it lies (oracle/plasticity_ref_shim.cpp -> oracle/_ref/libplasticity_ref.so;
tests/golden/make_plasticity_golden.py).  The oracle's restatement and the CUDA kernel (through the C ABI) must return the same projected F
(1e-11 absolute on entries of order 1, as in tests/test_gpu_plasticity.py: the SVDs differ in their rounding), plastic volume ratio Jp and
hardened Lame parameters (1e-11 relative).  The Drucker-Prager mapping is an extension without reference code and is not part of this file."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_plasticity_golden", os.path.join(ROOT, "tests", "golden", "make_plasticity_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "plasticity_ref.npz"))


def _check(make_sim, name, atol_F):
    model, params, _ = gen.CASES[name]
    F = G[name + "/in_F"]
    n = len(F)
    mu, lam = gen.lame()
    rng = np.random.default_rng(1)
    s = make_sim(0.05)
    X = (8.0 + 4.0 * rng.random((n, 3))) * 0.05               # positions play no part in the mapping
    s.set_particles(X, np.zeros((n, 3)), np.ones(n), np.zeros((n, 9)), F, np.ones(n), np.full(n, mu), np.full(n, lam))
    s.set_plasticity(model, params)
    s.applyPlasticity()                                        # before any sort: original particle order
    Fo = s.get_particles(gradV=False)["F"]
    scale = max(1.0, np.abs(G[name + "/F"]).max())
    assert np.abs(Fo - G[name + "/F"]).max() <= atol_F * scale, (name, np.abs(Fo - G[name + "/F"]).max())
    Jp, mu_o, lam_o = s.get_plastic_state()
    np.testing.assert_allclose(Jp, G[name + "/Jp"], rtol=1e-11, err_msg=name + " Jp")
    np.testing.assert_allclose(mu_o, G[name + "/mu"], rtol=1e-11, err_msg=name + " mu")
    np.testing.assert_allclose(lam_o, G[name + "/lam"], rtol=1e-11, err_msg=name + " lambda")


@pytest.mark.parametrize("name", list(gen.CASES))
def test_oracle_return_mappings_against_reference_code(oracle, name):
    _check(oracle.OracleSim, name, 1e-11)


def test_golden_cases_exercise_the_branches():
    assert 0 < int(G["von_mises_partial/projected"].sum()) < len(G["von_mises_partial/projected"])      # elastic and yielding particles
    assert int(G["von_mises_100/projected"].sum()) > 300
    Jp = G["snow_jp_clamped/Jp"]
    assert (Jp == 0.97).any() and (Jp == 1.02).any() and ((Jp > 0.97) & (Jp < 1.02)).any()            # both clamps and the interior
    assert np.abs(G["snow_default/mu"] / gen.lame()[0] - 1.0).max() > 1e-3                            # hardening happened


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/libplasticity_ref.so not built (needs /root/reference)")
def test_reference_return_mappings_reproduce_the_golden_vectors():
    for name in ("snow_default", "von_mises_large"):
        model, params, _ = gen.CASES[name]
        out = gen.reference(model, params, G[name + "/in_F"])
        for k in ("F", "mu", "lam", "Jp"):
            assert np.array_equal(out[k], G[f"{name}/{k}"]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(gen.MODERATE))
def test_cuda_return_mappings_against_reference_code(hot, name):
    _check(hot.MpmSimulationB200, name, 1e-11)
