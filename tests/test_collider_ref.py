"""Pins row a8 (+ f3: collision objects, buildInitialDvAndVnForNewton) to the REFERENCE'S OWN code: tests/golden/collider_ref.npz was produced by
Lib/Ziran/Math/Geometry/{AnalyticLevelSet.cpp, CollisionObject.cpp} compiled where they lie (oracle/collider_ref_shim.cpp ->
oracle/_ref/libcollider_ref.so; tests/golden/make_collider_golden.py).  The host mirror of include/hot_b200_host.hpp (level sets,
AnalyticCollisionObject::detectAndResolveCollision, multiObjectCollision, collisionNodeAt; driver tests/cpp/colliders_ref.cpp, no device call) must
take the same collision decisions at every point and return the same CollisionNode {P, R, R^-1, shouldRotate} and Newton initial guess
(1e-12: another quaternion -> matrix route and R^-1 = R^T instead of the 3 x 3 inverse).  The device kernel (colliders.cu) is compared with this
mirror on a grid in tests/test_gpu_host_cpp.py::test_device_colliders_match_host_evaluation."""
import importlib.util
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_collider_golden", os.path.join(ROOT, "tests", "golden", "make_collider_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "collider_ref.npz"))
SCENES = ["mixed", "slip_corner", "rotated"]


@pytest.fixture(scope="module")
def mirror(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("colliders_ref") / "colliders_ref")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "colliders_ref.cpp"),
                           "-o", exe, "-L", lib, "-lhot_b200", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def _mirror(exe, tmp_path, objs, xi, v):
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<qqd3d", len(objs), len(xi), gen.DT, *gen.GRAVITY))
        for a in (objs, xi, v):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    subprocess.check_call([exe, inp, out])
    rec = np.fromfile(out, dtype=np.float64).reshape(len(xi), 32)
    return dict(collide=rec[:, 0].astype(np.int32), dv=rec[:, 1:4], P=rec[:, 4:13], R=rec[:, 13:22], Rinv=rec[:, 22:31], slip=rec[:, 31].astype(np.int32))


@pytest.mark.parametrize("name", SCENES)
def test_host_collision_objects_against_reference_code(mirror, tmp_path, name):
    g = lambda k: G[f"{name}/{k}"]
    out = _mirror(mirror, tmp_path, g("objects"), g("xi"), g("v"))
    assert np.array_equal(out["collide"], g("collide"))              # the same inside / outside decision at every point
    assert np.array_equal(out["slip"], g("slip"))
    assert g("collide").sum() > 500
    # a slip normal exactly antiparallel to e_x: Eigen's setFromTwoVectors takes the rotation axis from a JacobiSVD null vector there (any axis orthogonal
    # to the normal is valid); not reference code, so R / R^-1 of those points are left out
    keep = ~((g("slip") == 1) & (g("R")[:, 0] < -1.0 + 1e-9))
    assert (~keep).sum() <= 5
    for k in ("dv", "P", "R", "Rinv"):
        a, b = (out[k][keep], g(k)[keep]) if k in ("R", "Rinv") else (out[k], g(k))
        err = np.abs(a - b).max()
        assert err <= 1e-12 * max(1.0, np.abs(b).max()), (name, k, err)


@pytest.mark.parametrize("name", SCENES)
def test_host_max_speed_and_calculate_dt_against_reference_code(mirror, tmp_path, name):
    """AnalyticCollisionObject::evalMaxSpeed (CollisionObject.cpp:201-238) of every object of the scene for four particle boxes - what calculateDt
    (MpmSimulationBase.cpp:789-814) takes its object speed from - and the dt rule itself on a two-particle set"""
    objs = G[name + "/objects"]
    for k, (lo, hi) in enumerate(gen.BOXES):
        inp, out = str(tmp_path / f"in{k}.bin"), str(tmp_path / f"out{k}.bin")
        xi = np.array([lo, hi], dtype=np.float64); v = np.array([[0.3, -0.2, 0.1], [0.0, 0.5, 0.0]])
        with open(inp, "wb") as f:
            f.write(struct.pack("<qqd3d", len(objs), 2, gen.DT, *gen.GRAVITY))
            for a in (objs, xi, v):
                f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        subprocess.check_call([mirror, inp, out, "speed"])
        rec = np.fromfile(out, dtype=np.float64)
        got, dt = rec[:len(objs)], rec[len(objs)]
        ref = G[name + "/max_speed"][k]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (name, k)              # where the reference throws "No bounds available." the mirror throws
        ok = ~np.isnan(ref)
        assert np.abs(got[ok] - ref[ok]).max() <= 1e-12 * max(1.0, np.abs(ref[ok]).max()), (name, k, got, ref)
        grown = G[name + "/max_speed_grown"][k]                      # calculateDt: cfl dx / max(particle speed, object speed over the box grown by 4 dx)
        top = max(np.linalg.norm(v, axis=1).max(), np.nanmax(grown) if ok.any() else 0.0)
        assert abs(dt - 0.6 * (1.0 / 32) / top) <= 1e-12 * dt


def test_golden_scenes_exercise_the_branches():
    m, c = "mixed", "slip_corner"
    assert 0 < G[m + "/slip"].sum() < G[m + "/collide"].sum()                             # sticky and slip nodes
    rank = np.array([np.linalg.matrix_rank(P.reshape(3, 3), tol=1e-9) for P in G[c + "/P"][G[c + "/collide"] == 1]])
    assert (rank == 2).any() and (rank == 1).any() and (rank == 0).any()                     # one, two and three slip normals (Gram-Schmidt)
    free = G[m + "/collide"] == 0
    assert np.allclose(G[m + "/dv"][free], gen.DT * gen.GRAVITY)                           # Newton initial guess of the free nodes


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/libcollider_ref.so not built (needs /root/reference)")
def test_reference_collision_objects_reproduce_the_golden_vectors():
    name = "rotated"
    out = gen.reference(list(G[name + "/objects"]), G[name + "/xi"], G[name + "/v"])
    for k in ("collide", "slip", "dv", "P", "R", "Rinv"):
        assert np.array_equal(out[k], G[f"{name}/{k}"]), k
