"""The C++ host mirror (include/hot_b200_host.hpp: MpmSimulationB200::advanceOneTimeStep, host-evaluated collision
objects = a8, HOTSettings / flag parsing, -smoother function pointers) against the CPU oracle stepping the same scene
with the same reference flags.  The oracle side evaluates a8 in numpy (HalfSpace STICKY / SLIP,
CollisionObject.cpp:108-149,384-452; MpmSimulationBase.cpp:1139-1184)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from hot_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_step(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "host_step")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_step.cpp"), "-o", exe, "-L", lib, "-lhot_b200",
                           f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def _rotate_to_x(a):
    a = a / np.linalg.norm(a)
    v = np.cross(a, [1.0, 0, 0]); c = a[0]
    vx = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + vx + vx @ vx / (1 + c)


def _oracle_step(o, dt, ground, gtype, opts):
    """MultigridSimulation::advanceOneTimeStep with the oracle: sort, P2G, a8 in numpy, BE solve, G2P"""
    o.set_dt_gravity(dt, (0, -9.8, 0))
    o.sortParticlesAndPolluteGrid(); o.particlesToGrid()
    coord = o.get_id2coord(); idx, _, v = o.get_grid()
    n = o.num_nodes
    vn = np.zeros((n, 3)); vn[idx[idx >= 0]] = v[idx >= 0]
    inside = coord[:, 1] * o.dx - ground <= 0
    bc = np.nonzero(inside)[0].astype(np.int32)
    nrm = np.array([0.0, 1.0, 0.0])
    if gtype == 1:       # STICKY: v = 0, P = 0
        vi = np.zeros((len(bc), 3)); P = np.zeros((len(bc), 9)); R = np.tile(np.eye(3).reshape(1, 9), (len(bc), 1)); slip = np.zeros(len(bc), np.int32)
    else:                # SLIP: remove the normal component, P = I - n n^T, R rotates n onto x
        vi = vn[bc] - np.outer(vn[bc] @ nrm, nrm)
        P = np.tile((np.eye(3) - np.outer(nrm, nrm)).T.reshape(1, 9), (len(bc), 1))
        R = np.tile(_rotate_to_x(nrm).T.reshape(1, 9), (len(bc), 1)); slip = np.ones(len(bc), np.int32)
    Rinv = R.reshape(-1, 3, 3).transpose(0, 2, 1).reshape(-1, 9)
    mode = 1 if (opts.get("bcproject") and opts.get("bc_type") == 1) else 0
    o.set_bc(bc, P=P, R=R, Rinv=Rinv, slip=slip, dv_bc=vi - vn[bc], mode=mode)
    kw = {k: v for k, v in opts.items() if k != "bc_type"}
    log = o.backwardEulerStep(**kw)
    o.gridToParticles(dt)
    return log, n, len(bc)


CASES = {
    "hot_sticky": (1, ["-lsolver", "3", "-Ainv", "1", "--project", "--linesearch", "--bcproject", "-mg_level", "3", "-mg_times", "1",
                       "-coarseSolver", "2", "-smoother", "5", "--usecn", "-cneps", "1e-7"],
                   dict(lsolver=3, Ainv=1, project=1, linesearch=1, bcproject=1, mg_level=3, mg_times=1, coarse_solver=2, smoother=5, usecn=1, cneps=1e-7)),
    "pnmf_sticky": (1, ["-lsolver", "2", "-Ainv", "1", "--project", "--linesearch", "--matfree", "--usecn", "-cneps", "1e-7"],
                    dict(lsolver=2, Ainv=1, project=1, linesearch=1, matfree=1, bcproject=0, mg_level=1, mg_times=1, smoother=0, coarse_solver=0,
                         usecn=1, cneps=1e-7)),
    "hot_slip": (2, ["-lsolver", "3", "-Ainv", "1", "--project", "--linesearch", "--bcproject", "-bc", "1", "-mg_level", "3", "-mg_times", "1",
                     "-coarseSolver", "2", "-smoother", "5", "--usecn", "-cneps", "1e-7"],
                 dict(lsolver=3, Ainv=1, project=1, linesearch=1, bcproject=1, bc_type=1, mg_level=3, mg_times=1, coarse_solver=2, smoother=5,
                      usecn=1, cneps=1e-7)),
}


@pytest.mark.parametrize("case", list(CASES))
def test_cpp_host_time_steps_match_oracle(host_step, oracle, tmp_path, case):
    gtype, flags, opts = CASES[case]
    sc = scenes.block((6, 6, 6), 1.0 / 32, ppc=6, origin_cells=(8, 8, 8), rho=1000.0, E=2.5e4, nu=0.4, seed=3)
    sc["V"][:, 1] -= 0.5                                             # falling onto the ground
    ground = 9.0 / 32                                                # one cell into the box: the lowest node layers collide
    n = len(sc["mass"]); steps, dt = 2, 2e-3
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<qd", n, sc["dx"]))
        for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam"):
            f.write(np.ascontiguousarray(sc[k], dtype=np.float64).tobytes())
    out = subprocess.run([host_step, fin, fout, str(steps), repr(dt), repr(ground), str(gtype)] + flags, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    raw = np.fromfile(fout, dtype=np.float64, offset=8)
    X = raw[:3 * n].reshape(n, 3); V = raw[3 * n:6 * n].reshape(n, 3); F = raw[15 * n:24 * n].reshape(n, 9)
    rec = raw[24 * n:].reshape(steps, 5)

    o = oracle.OracleSim(sc["dx"])
    o.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    for s in range(steps):
        log, nn, nbc = _oracle_step(o, dt, ground, gtype, opts)
        assert bool(rec[s, 1]) == log["converged"]      # Newton may stop at its 3-iteration cap like the reference
        assert int(rec[s, 0]) == log["iterations"] and int(rec[s, 2]) == nn and int(rec[s, 3]) == nbc and nbc > 0
        assert abs(rec[s, 4] - log["residual_norm"][-1]) <= 1e-5 * log["residual_norm"][0]
    po = o.get_particles()
    for a, b in ((X, po["X"]), (V, po["V"]), (F, po["F"])):
        assert np.abs(a - b).max() < 1e-6 * np.abs(b).max()


def test_restart_state_layout_and_round_trip(host_step, tmp_path):
    """MpmSimulationB200::writeState / readState: the reference's restart layout (Scene.h:189-206, DataManager.h:263-273,
    DataArray.h:100-105, BinaryIO.h:82-88,167-172), parsed here byte by byte, and a write -> read round trip in the C++ driver"""
    gtype, flags, _ = CASES[next(iter(CASES))]
    sc = scenes.block((4, 4, 4), 1.0 / 32, ppc=4, origin_cells=(8, 8, 8), rho=1000.0, E=2.5e4, nu=0.4, seed=5)
    n = len(sc["mass"])
    fin, fout, frs = str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(tmp_path / "restart_1.dat")
    with open(fin, "wb") as f:
        f.write(struct.pack("<qd", n, sc["dx"]))
        for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam"):
            f.write(np.ascontiguousarray(sc[k], dtype=np.float64).tobytes())
    out = subprocess.run([host_step, fin, fout, "1", "2e-3", repr(9.0 / 32), str(gtype)] + flags, capture_output=True, text=True,
                         env=dict(os.environ, HOT_RESTART_FILE=frs))
    assert out.returncode == 0 and "restart ok" in out.stdout, out.stderr
    raw = np.fromfile(fout, dtype=np.float64, offset=8)
    X = raw[:3 * n].reshape(n, 3); F = raw[15 * n:24 * n].reshape(n, 9)
    b = open(frs, "rb").read()
    pos = 0
    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, b, pos); pos += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v
    assert take("i") == n and take("Q") == 6
    arrays = {}
    for _ in range(6):
        ln = take("Q"); name = b[pos:pos + ln].decode(); pos += ln
        assert take("i") == 7                                               # DisjointRanges lg2_grain_size
        assert take("Q") == 1 and take("Q") == 8 and take("ii") == (0, n)   # one Range [0, n)
        size, eb = take("Q"), take("Q")
        assert size == n
        arrays[name] = (eb, np.frombuffer(b, dtype=np.float64, count=n * eb // 8, offset=pos))
        pos += n * eb
    assert {k: v[0] for k, v in arrays.items()} == {"X": 24, "V": 24, "m": 8, "element measure": 8, "F": 72, "CorotatedIsotropic": 24}
    assert np.array_equal(arrays["X"][1].reshape(n, 3), X) and np.array_equal(arrays["F"][1].reshape(n, 9), F)
    assert np.array_equal(arrays["m"][1], sc["mass"]) and np.array_equal(arrays["element measure"][1], sc["vol"])
    # (24 raw bytes per entry: project flag + padding, mu, lambda - tests/test_restart_ref.py compares with the reference's own writer)
    assert np.array_equal(arrays["CorotatedIsotropic"][1].reshape(n, 3)[:, 1:], np.stack([sc["mu"], sc["lam"]], 1))
    assert take("QQ") == (0, 12) and take("QQ") == (0, 8)                    # empty trimesh / segmesh index vectors
    assert take("QQ") == (n, 72) and len(b) - pos == 72 * n                  # trailing APIC matrices


def test_cpp_host_rejects_unknown_flag(host_step, tmp_path):
    fin = str(tmp_path / "in.bin")
    sc = scenes.block((3, 3, 3), 1.0 / 32, ppc=2, seed=1)
    with open(fin, "wb") as f:
        f.write(struct.pack("<qd", len(sc["mass"]), sc["dx"]))
        for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam"):
            f.write(np.ascontiguousarray(sc[k], dtype=np.float64).tobytes())
    out = subprocess.run([host_step, fin, str(tmp_path / "o.bin"), "1", "1e-3", "0.2", "1", "--l2norm"], capture_output=True, text=True)
    assert out.returncode == 1 and "Unknown flag" in out.stderr      # like the reference


def test_force_helper_mirror(hot, tmp_path):
    """FBasedMpmForceHelperB200 (MpmForceHelperBase.h:18-46 surface) against the ctypes path: strain energy, the per-particle dPdF walk
    in the reference's colour-pass order, Fn, stored / reused Hessians, dPdF(F = I)."""
    exe = str(tmp_path / "force_helper")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "force_helper.cpp"), "-o", exe, "-L", lib, "-lhot_b200",
                           f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    sc = scenes.block((5, 6, 4), 1.0 / 32, ppc=6, origin_cells=(8, 8, 8), rho=1000.0, E=2.5e4, nu=0.4, seed=7)
    sc["mu"][::3] *= 2.0                                             # two materials: the helper evaluates per distinct (mu, lambda)
    n = len(sc["mass"]); dt = 2e-3
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<qd", n, sc["dx"]))
        for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam"):
            f.write(np.ascontiguousarray(sc[k], dtype=np.float64).tobytes())
    out = dict(line.split(" ", 1) for line in subprocess.check_output([exe, str(inp), str(dt)], text=True).strip().splitlines())
    g = hot.MpmSimulationB200(sc["dx"])
    g.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    g.set_dt_gravity(dt, (0, -9.8, 0))
    g.sortParticlesAndPolluteGrid(); g.particlesToGrid()
    g.set_bc(np.zeros(0, dtype=np.int32))
    g.backupStrain(); g.updateState()
    _, F = g.get_stress()
    e = 0.0; sh = 0.0; si = 0.0
    for mu in np.unique(sc["mu"]):
        sel = sc["mu"] == mu
        c = g.corotated_eval(F[sel].reshape(-1, 3, 3).transpose(0, 2, 1), mu, sc["lam"][0], project=True)
        e += (sc["vol"][sel] * c["psi"]).sum(); sh += c["dPdF"][:, 0, 0].sum()
        ci = g.corotated_eval(np.tile(np.eye(3), (int(sel.sum()), 1, 1)), mu, sc["lam"][0], project=True)
        si += ci["dPdF"][:, 0, 0].sum()
    assert int(out["visited"]) == n
    assert abs(float(out["energy"]) - e) <= 1e-11 * abs(e)
    assert abs(float(out["sum_dPdF00"]) - sh) <= 1e-11 * abs(sh)
    assert float(out["sum_dPdF00_reused"]) == float(out["sum_dPdF00"])
    assert abs(float(out["sum_dPdF00_identity"]) - si) <= 1e-11 * abs(si)
    assert abs(float(out["sum_Fn00"]) - sc["F"][:, 0].sum()) <= 1e-12 * n
    sorter, order, _ = g.get_sort()
    first, last, block = g.get_groups()
    colour0 = [gi for gi in range(len(block)) if (int(block[gi]) & 7) == min(int(b) & 7 for b in block)]
    assert int(out["first_visited"]) == order[first[colour0[0]]]


def test_device_colliders_match_host_evaluation(tmp_path):
    """hot_set_colliders + hot_build_bc (a8 on the device: half-space, sphere, rotated box, rotating capped cylinder; STICKY / SLIP /
    SEPARATE / GHOST, friction, moving and rotating objects) against the host evaluation of the C++ mirror on the same grid."""
    exe = str(tmp_path / "colliders")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "colliders.cpp"), "-o", exe, "-L", lib, "-lhot_b200",
                           f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    sc = scenes.block((9, 12, 8), 1.0 / 32, ppc=6, origin_cells=(10, 10, 10), rho=1000.0, E=2.5e4, nu=0.4, seed=11)
    n = len(sc["mass"])
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<qd", n, sc["dx"]))
        for k in ("X", "V", "mass", "C", "F", "vol", "mu", "lam"):
            f.write(np.ascontiguousarray(sc[k], dtype=np.float64).tobytes())
    out = dict(line.split(" ", 1) for line in subprocess.check_output([exe, str(inp), "2e-3"], text=True).strip().splitlines())
    assert out["match"] == "1" and int(out["host_bc"]) == int(out["device_bc"]) > 100
    assert int(out["bad_id"]) == 0 and int(out["bad_slip"]) == 0 and int(out["slip_nodes"]) > 10
    for k in ("err_P", "err_R", "err_Rinv"):
        assert float(out[k]) <= 1e-13, (k, out[k])
    assert float(out["err_dv"]) <= 1e-13 * max(float(out["scale_dv"]), 1.0)
