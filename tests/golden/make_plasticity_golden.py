"""Generates tests/golden/plasticity_ref.npz: the REFERENCE'S OWN return mappings (row f2) - SnowPlasticity<double>::projectStrain and
VonMisesFixedCorotated<double,3>::projectStrain of Lib/Ziran/Physics/PlasticityApplier.cpp, compiled where it lies (oracle/plasticity_ref_shim.cpp
-> oracle/_ref/libplasticity_ref.so) with the reference's CorotatedIsotropic as the constitutive model - applied to seeded deformation gradients.
tests/test_oracle_plasticity_ref.py compares the oracle's restatement (oracle_force.inl: orc_apply_plasticity) and the CUDA kernel with them.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_plasticity_golden.py"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libplasticity_ref.so")
OUT = os.path.join(ROOT, "tests", "golden", "plasticity_ref.npz")
E, NU = 1e5, 0.3
# case -> (model, parameters, deformation-gradient set)
CASES = {
    "snow_default": ("snow", [10.0, 2e-2, 7.5e-3, 0.6, 20.0], "moderate"),
    "snow_jp_clamped": ("snow", [10.0, 2e-2, 7.5e-3, 0.97, 1.02], "moderate"),          # both Jp clamps taken
    "snow_large": ("snow", [10.0, 2.5e-2, 4.5e-3, 0.6, 20.0], "large"),
    "von_mises_100": ("von_mises", [100.0], "moderate"),
    "von_mises_partial": ("von_mises", [8e3], "moderate"),                              # some particles stay elastic
    "von_mises_large": ("von_mises", [500.0], "large"),                                 # includes singular values clamped at 1e-4 and inverted F
}
MODERATE = ("snow_default", "snow_jp_clamped", "von_mises_100", "von_mises_partial")     # the sets the CUDA test uses (range of tests/test_gpu_plasticity.py)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gradients(kind, n=400, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "moderate":
        return np.eye(3).reshape(1, 9) + 0.5 * (rng.random((n, 9)) - 0.5)
    F = np.eye(3).reshape(1, 9) + 1.6 * (rng.random((n, 9)) - 0.5)
    F[:5] = [np.diag(d).reshape(9) for d in ([1.0, 1.0, 1.0], [1.3, 1.3, 0.6], [2.0, 1e-6, 0.5], [1.0, 1.0, -0.7], [1.01, 0.99, 1.0])]
    return F


def lame():
    return E / (2 * (1 + NU)), E * NU / ((1 + NU) * (1 - 2 * NU))


def reference(model, params, F):
    lib = C.CDLL(REF_LIB)
    n = len(F)
    F = np.ascontiguousarray(F, dtype=np.float64).copy()
    mu0, lam0 = lame()
    mu = np.full(n, mu0); lam = np.full(n, lam0); Jp = np.ones(n)
    proj = np.zeros(n, dtype=np.int32)
    if model == "snow":
        q = np.ascontiguousarray(params, dtype=np.float64)
        lib.zr_plasticity_snow(C.c_long(n), _p(F), _p(mu), _p(lam), _p(Jp), _p(q))
    else:
        lib.zr_plasticity_von_mises(C.c_long(n), _p(F), _p(mu), _p(lam), C.c_double(params[0]), _p(proj))
    return dict(F=F, mu=mu, lam=lam, Jp=Jp, projected=proj)


if __name__ == "__main__":
    gold = {}
    for k, (name, (model, params, kind)) in enumerate(CASES.items()):
        F = gradients(kind, seed=k)
        out = reference(model, params, F)
        gold[name + "/in_F"] = F
        for key, v in out.items():
            gold[f"{name}/{key}"] = v
        print(name, "changed F:", int((np.abs(out["F"] - F).max(1) > 1e-12).sum()), "of", len(F), "| projected flags", int(out["projected"].sum()),
              "| Jp range", out["Jp"].min(), out["Jp"].max())
    np.savez_compressed(OUT, **gold)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
