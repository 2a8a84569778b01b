"""Generates tests/golden/collider_ref.npz: row a8 (+ f3) on the REFERENCE'S OWN collision-object code - Lib/Ziran/Math/Geometry/AnalyticLevelSet.cpp
(HalfSpace, Sphere, AxisAlignedAnalyticBox, AnalyticBox, CappedCylinder) and CollisionObject.cpp (detectAndResolveCollision with the object transform
and its rates, STICKY / SLIP / SEPARATE / GHOST, friction; multiObjectCollision), compiled where they lie (oracle/collider_ref_shim.cpp ->
oracle/_ref/libcollider_ref.so), with the per-node body of buildInitialDvAndVnForNewton (MpmSimulationBase.cpp:1139-1184) around them - evaluated at
seeded points around seeded scenes of objects.  tests/test_collider_ref.py compares the host mirror of include/hot_b200_host.hpp with these results
(the device kernel is compared with that mirror in tests/test_gpu_host_cpp.py).
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_collider_golden.py"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libcollider_ref.so")
OUT = os.path.join(ROOT, "tests", "golden", "collider_ref.npz")
STICKY, SLIP, SEPARATE, GHOST = 1, 2, 3, 4
HALFSPACE, SPHERE, BOX, CYLINDER, AABOX = 0, 1, 2, 3, 4
GRAVITY = np.array([0.0, -9.8, 0.0])
DT = 2e-3


def obj(type_, shape, p, friction=0.0, shape_q=(1, 0, 0, 0), shape_b=(0, 0, 0), q=(1, 0, 0, 0), s=1.0, b=(0, 0, 0), omega=(0, 0, 0), dsdt=0.0, dbdt=(0, 0, 0)):
    """33 doubles: the layout of oracle/collider_ref_shim.cpp"""
    pp = list(p) + [0.0] * (8 - len(p))
    return np.array([type_, shape, friction] + pp + list(shape_q) + list(shape_b) + list(q) + [s] + list(b) + list(omega) + [dsdt] + list(dbdt), dtype=np.float64)


def scenes():
    L = 1.0
    out = {}
    # the objects of tests/cpp/colliders.cpp: tilted SLIP ground with friction, moving STICKY sphere, rotating STICKY capped cylinder (the twisting bar's
    # clamp, MultigridInit3D.h:628-661), SEPARATE rotated box with friction, SLIP axis-aligned box (two slip normals at its edge with the ground), a GHOST
    out["mixed"] = [
        obj(SLIP, HALFSPACE, [0, 0.12 * L, 0, 0.15, 1.0, -0.1], friction=0.3),
        obj(STICKY, SPHERE, [0, 0, 0, 0.3 * L], b=(0.0, 0.5, 0.5), dbdt=(0.2, -0.1, 0.05)),
        obj(STICKY, CYLINDER, [0.6 * L, 0.25 * L], q=(np.cos(0.35), 0, np.sin(0.35), 0), omega=(0, 2 * np.pi, 0), b=(0.5, 1.0, 0.5), dbdt=(0, 0.3, 0)),
        obj(SEPARATE, BOX, [0.2 * L, 0.15 * L, 0.5 * L], friction=0.5, shape_q=(0.9, 0.1, 0.3, -0.2), shape_b=(1.0, 0.5, 0.5)),
        obj(SLIP, AABOX, [0.0, 0.0, 0.9, 1.0, 0.3, 2.0]),
        obj(GHOST, SPHERE, [0.5, 0.5, 0.5, 10 * L]),
    ]
    # three SLIP planes meeting in a corner (Gram-Schmidt up to dim normals), one scaled and growing SLIP sphere, SEPARATE ground
    out["slip_corner"] = [
        obj(SLIP, HALFSPACE, [0.2, 0, 0, 1.0, 0.1, 0.0]),
        obj(SLIP, HALFSPACE, [0, 0.2, 0, 0.05, 1.0, 0.1], friction=0.2),
        obj(SLIP, HALFSPACE, [0, 0, 0.2, 0.0, -0.1, 1.0]),
        obj(SLIP, SPHERE, [0, 0, 0, 0.2], s=1.5, dsdt=0.4, b=(0.8, 0.8, 0.8), q=(0.8, 0.2, -0.4, 0.3), omega=(1.0, -2.0, 0.5)),
        obj(SEPARATE, HALFSPACE, [0, 0.9, 0, 0.0, -1.0, 0.0], friction=0.7),
    ]
    # rotated capped cylinder (own rotation) under a moving rotating object transform, SLIP; rotated SLIP box
    out["rotated"] = [
        obj(SLIP, CYLINDER, [0.25, 0.5], shape_q=(0.7, 0.3, -0.2, 0.6), shape_b=(0.1, -0.05, 0.0), q=(0.9, -0.3, 0.2, 0.1), b=(0.5, 0.5, 0.5), omega=(0.3, 0.2, -1.0),
            dbdt=(0.1, 0.0, -0.2), friction=0.4),
        obj(SLIP, BOX, [0.3, 0.1, 0.2], shape_q=(0.6, -0.5, 0.4, 0.2), shape_b=(0.2, 0.2, 0.8), s=0.8, b=(0.1, 0.0, 0.0)),
        obj(STICKY, AABOX, [0.7, 0.7, 0.0, 1.2, 1.2, 0.3], dbdt=(0.0, 0.0, 0.5)),
    ]
    return out


def points(seed, n=6000):
    rng = np.random.default_rng(seed)
    xi = -0.1 + 1.3 * rng.random((n, 3))
    xi[: n // 3] = np.round(xi[: n // 3] * 32) / 32            # grid nodes of dx = 1/32 as in the solver
    v = 2.0 * (rng.random((n, 3)) - 0.5)
    v[::7] = 0.0
    return xi, v


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def reference(objs, xi, v):
    lib = C.CDLL(REF_LIB)
    o = np.ascontiguousarray(np.concatenate(objs))
    n = len(xi)
    xi = np.ascontiguousarray(xi); v = np.ascontiguousarray(v)
    collide = np.zeros(n, dtype=np.int32); slip = np.zeros(n, dtype=np.int32)
    dv = np.empty((n, 3)); P = np.empty((n, 9)); R = np.empty((n, 9)); Rinv = np.empty((n, 9))
    g = np.ascontiguousarray(GRAVITY)
    lib.zr_colliders_eval(len(objs), _p(o), C.c_long(n), _p(xi), _p(v), _p(g), C.c_double(DT), _p(collide), _p(dv), _p(P), _p(R), _p(Rinv), _p(slip))
    return dict(collide=collide, dv=dv, P=P, R=R, Rinv=Rinv, slip=slip)


def reference_max_speed(objs, p_min, p_max):
    lib = C.CDLL(REF_LIB)
    o = np.ascontiguousarray(np.concatenate(objs))
    lo = np.ascontiguousarray(p_min, dtype=np.float64); hi = np.ascontiguousarray(p_max, dtype=np.float64)
    out = np.empty(len(objs))
    lib.zr_colliders_max_speed(len(objs), _p(o), _p(lo), _p(hi), _p(out))
    return out


# particle boxes for evalMaxSpeed (calculateDt): around the objects, far from them, inside them
BOXES = [((0.2, 0.2, 0.2), (0.9, 0.9, 0.9)), ((-0.3, 0.0, 0.1), (0.4, 1.4, 1.2)), ((3.0, 3.0, 3.0), (3.5, 3.2, 3.1)), ((0.45, 0.45, 0.45), (0.55, 0.55, 0.55))]


if __name__ == "__main__":
    gold = {}
    for k, (name, objs) in enumerate(scenes().items()):
        xi, v = points(100 + k)
        out = reference(objs, xi, v)
        gold[name + "/objects"] = np.stack(objs); gold[name + "/xi"] = xi; gold[name + "/v"] = v
        for key, val in out.items():
            gold[f"{name}/{key}"] = val
        two = int(((np.abs(out["P"]).sum(1) > 0) & (out["slip"] == 1) & (np.abs(np.linalg.det(out["P"].reshape(-1, 3, 3))) < 1e-12)).sum())
        print(name, "points", len(xi), "colliding", int(out["collide"].sum()), "slip", int(out["slip"].sum()), "sticky", int((out["collide"] - out["slip"]).sum()))
    for name, objs in scenes().items():
        gold[name + "/max_speed"] = np.stack([reference_max_speed(objs, lo, hi) for lo, hi in BOXES])
        g = 4.0 / 32          # calculateDt grows the particles' box by (interpolation_degree + 2) dx, dx = 1/32 in the test
        gold[name + "/max_speed_grown"] = np.stack([reference_max_speed(objs, np.asarray(lo) - g, np.asarray(hi) + g) for lo, hi in BOXES])
        print(name, "max speeds", np.round(gold[name + "/max_speed"], 4).tolist())
    np.savez_compressed(OUT, **gold)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
