"""Golden vectors from the REFERENCE'S OWN code for the floating-point rows a4 / a11: quadratic B-spline weights (BSplines.h:55-81),
int_floor (MathTools.h:15-25), the 3x3 implicit-QR SVD (ImplicitQRSVD.h) and the fixed-corotated model (CorotatedIsotropic.h:64-230,
SvdBasedIsotropicHelper.h, EigenDecomposition.h:126-135, DenseExt.h:240-252).

oracle/_ref/libziran_ref.so is built by oracle/Makefile from those headers where they lie under /root/reference/Lib, against the
Eigen / Tick stand-in of oracle/ref_shim/ (neither library exists in this image).  Run here (the reference is not on the GPU box):
    make -C oracle ref && python tests/golden/make_ziran_golden.py
Writes tests/golden/ziran_ref.npz."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libziran_ref.so"))
ref.ziran_ref_int_floor.argtypes = [C.c_double]
vp = C.c_void_p


def p(a):
    return a.ctypes.data_as(vp)


def cases(rng):
    out = [np.eye(3), np.diag([1.2, 1.2, 0.8]), np.diag([2.0, 1.0, 0.5]), np.diag([1.0, 1.0, 1.0 + 1e-9])]
    for _ in range(60):
        out.append(np.eye(3) + 0.4 * (rng.random((3, 3)) - 0.5))
    for _ in range(30):
        out.append(2.0 * (rng.random((3, 3)) - 0.5))              # large deformation, some inverted
    for _ in range(10):
        out.append(np.diag([1.0, 0.3, 0.05]) @ (np.eye(3) + 0.2 * (rng.random((3, 3)) - 0.5)))   # strongly compressed: makePD clamps
    Q, _ = np.linalg.qr(rng.random((3, 3)))
    Q = Q * np.linalg.det(Q)
    out += [Q, Q @ np.diag([1.3, 1.3, 0.7]) @ Q.T, Q @ np.diag([1.1, 0.9, -0.4]), np.zeros((3, 3)), np.diag([1.0, 1.0, 0.0])]
    return np.array(out)


rng = np.random.default_rng(20261017)
F = cases(rng)
n = len(F)
dF = rng.random((n, 3, 3)) - 0.5
E, nu = 1e4, 0.3
mu, lam = C.c_double(), C.c_double()
ref.ziran_ref_lame(C.c_double(E), C.c_double(nu), C.byref(mu), C.byref(lam))
Fc = np.ascontiguousarray(F.transpose(0, 2, 1)).reshape(n, 9)      # column-major per item
dFc = np.ascontiguousarray(dF.transpose(0, 2, 1)).reshape(n, 9)
U = np.empty((n, 9)); V = np.empty((n, 9)); sig = np.empty((n, 3))
ref.ziran_ref_svd3(C.c_long(n), p(Fc), p(U), p(sig), p(V))
out = dict(F=Fc, dF=dFc, mu=mu.value, lam=lam.value, E=E, nu=nu, U=U, sigma=sig, V=V)
for project in (0, 1):
    psi = np.empty(n); P = np.empty((n, 9)); dP = np.empty((n, 9)); H = np.empty((n, 81))
    ref.ziran_ref_corotated(C.c_long(n), mu, lam, project, p(Fc), p(dFc), p(psi), p(P), p(dP), p(H))
    out.update({f"psi_{project}": psi, f"P_{project}": P, f"dP_{project}": dP, f"dPdF_{project}": H})

# B-spline weights: index-space positions incl. cell-boundary neighbourhoods
x = np.concatenate([rng.random(200) * 60 + 2, np.arange(3, 13) + 0.5, np.nextafter(np.arange(3, 13) + 0.5, 0), np.nextafter(np.arange(3, 13) + 0.5, 100),
                    np.arange(3, 8).astype(float), [2.0000000001, 77.49999999999, 4095.25 - 3]])
base = np.empty(len(x), dtype=np.int32); w = np.empty((len(x), 3)); dw = np.empty((len(x), 3))
ref.ziran_ref_bspline2(C.c_long(len(x)), p(x), p(base), p(w), p(dw))
out.update(bs_x=x, bs_base=base, bs_w=w, bs_dw=dw)
fl = np.concatenate([rng.random(50) * 200 - 100, [0.0, -0.0, 1.0, -1.0, 2.5, -2.5, 1e-300, -1e-300, 4094.999999]])
out.update(floor_x=fl, floor_i=np.array([ref.ziran_ref_int_floor(float(v)) for v in fl], dtype=np.int32))
path = os.path.join(ROOT, "tests", "golden", "ziran_ref.npz")
np.savez_compressed(path, **out)
print(path, n, "deformation gradients,", len(x), "spline positions")
