// Small dense kernels of the constitutive model, device side: 3x3 products, cofactor, 3x3 SVD, symmetric
// eigen-projection.  Column-major like Eigen: M(r,c) = a[r + 3c] (Lib/Ziran/CS/Util/Forward.h:10-13).
//
// The reference computes the SVD with an implicit-shift QR iteration (Lib/Ziran/Math/Linear/ImplicitQRSVD.h:354-533)
// and the PSD projection with Eigen::SelfAdjointEigenSolver (Lib/Ziran/Math/Linear/EigenDecomposition.h:126-135).
// Both are data-dependent iterations on <= 9 numbers; on the GPU they are replaced by Jacobi iterations (one-sided
// Hestenes for the SVD - no squaring of the condition number -, two-sided for the symmetric blocks): branch-light,
// register-resident, quadratically convergent, and accurate to working precision for small singular values.
// Sign/sort convention of ImplicitQRSVD.h:256-352 is kept: U, V rotations, sigma0 >= sigma1 >= |sigma2|.
// Everything downstream (psi, P, dP, dPdF) is invariant to the remaining freedom.
#pragma once
#include <cuda_runtime.h>

namespace hot {

__device__ __forceinline__ void mm(const double* A, const double* B, double* C) // C = A B (C must not alias)
{
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) C[r + 3 * c] = A[r] * B[3 * c] + A[r + 3] * B[3 * c + 1] + A[r + 6] * B[3 * c + 2];
}
__device__ __forceinline__ void mm_bt(const double* A, const double* B, double* C) // C = A B^T
{
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) C[r + 3 * c] = A[r] * B[c] + A[r + 3] * B[c + 3] + A[r + 6] * B[c + 6];
}
__device__ __forceinline__ void mm_at(const double* A, const double* B, double* C) // C = A^T B
{
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) C[r + 3 * c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
}
__device__ __forceinline__ double det3(const double* A)
{
    return A[0] * (A[4] * A[8] - A[7] * A[5]) - A[3] * (A[1] * A[8] - A[7] * A[2]) + A[6] * (A[1] * A[5] - A[4] * A[2]);
}
// Lib/Ziran/Math/Linear/DenseExt.h:240-252 cofactorMatrix (= J F^-T)
__device__ __forceinline__ void cofactor3(const double* F, double* A)
{
    A[0] = F[4] * F[8] - F[7] * F[5]; A[3] = F[7] * F[2] - F[1] * F[8]; A[6] = F[1] * F[5] - F[4] * F[2];
    A[1] = F[6] * F[5] - F[3] * F[8]; A[4] = F[0] * F[8] - F[6] * F[2]; A[7] = F[3] * F[2] - F[0] * F[5];
    A[2] = F[3] * F[7] - F[6] * F[4]; A[5] = F[6] * F[1] - F[0] * F[7]; A[8] = F[0] * F[4] - F[3] * F[1];
}

// one Hestenes rotation between columns p and q of A (and of the accumulated right factor W); returns |cos angle|
template <int p, int q>
__device__ __forceinline__ double hestenes(double* A, double* W)
{
    double* ap = A + 3 * p;
    double* aq = A + 3 * q;
    const double alpha = ap[0] * ap[0] + ap[1] * ap[1] + ap[2] * ap[2];
    const double beta = aq[0] * aq[0] + aq[1] * aq[1] + aq[2] * aq[2];
    const double gamma = ap[0] * aq[0] + ap[1] * aq[1] + ap[2] * aq[2];
    const double lim = sqrt(alpha * beta);
    if (gamma == 0.0 || fabs(gamma) <= 1e-17 * lim) return 0.0;
    const double zeta = (beta - alpha) / (2.0 * gamma);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double x = ap[r], y = aq[r];
        ap[r] = c * x - s * y; aq[r] = s * x + c * y;
        const double wx = W[3 * p + r], wy = W[3 * q + r];
        W[3 * p + r] = c * wx - s * wy; W[3 * q + r] = s * wx + c * wy;
    }
    return fabs(gamma) / lim;
}

// F = U diag(sig) V^T, det U = det V = +1, sig0 >= sig1 >= |sig2|
__device__ inline void svd3(const double* F, double* U, double* sig, double* V)
{
    double A[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int q = 0; q < 9; ++q) A[q] = F[q];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = hestenes<0, 1>(A, W);
        off = fmax(off, hestenes<0, 2>(A, W));
        off = fmax(off, hestenes<1, 2>(A, W));
        if (off < 1e-16) break;
    }
    double n0 = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
    double n1 = A[3] * A[3] + A[4] * A[4] + A[5] * A[5];
    double n2 = A[6] * A[6] + A[7] * A[7] + A[8] * A[8];
    // descending order of the column norms (stable, like the CPU restatement's std::sort on 3 items)
    int o0 = 0, o1 = 1, o2 = 2;
    if (n1 > n0) { int t = o0; o0 = o1; o1 = t; double tn = n0; n0 = n1; n1 = tn; }
    if (n2 > n1) { int t = o1; o1 = o2; o2 = t; double tn = n1; n1 = n2; n2 = tn; }
    if (n1 > n0) { int t = o0; o0 = o1; o1 = t; double tn = n0; n0 = n1; n1 = tn; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V[r] = o0 == 0 ? W[r] : (o0 == 1 ? W[3 + r] : W[6 + r]);
        V[3 + r] = o1 == 0 ? W[r] : (o1 == 1 ? W[3 + r] : W[6 + r]);
        V[6 + r] = o2 == 0 ? W[r] : (o2 == 1 ? W[3 + r] : W[6 + r]);
    }
    if (det3(V) < 0.0) {
        V[6] = -V[6]; V[7] = -V[7]; V[8] = -V[8];
    }
    double FV[9];
    mm(F, V, FV);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double l = sqrt(FV[3 * c] * FV[3 * c] + FV[3 * c + 1] * FV[3 * c + 1] + FV[3 * c + 2] * FV[3 * c + 2]);
        sig[c] = l;
#pragma unroll
        for (int r = 0; r < 3; ++r) U[3 * c + r] = l > 0.0 ? FV[3 * c + r] / l : (r == c ? 1.0 : 0.0);
    }
    const double d01 = U[0] * U[3] + U[1] * U[4] + U[2] * U[5];
#pragma unroll
    for (int r = 0; r < 3; ++r) U[3 + r] -= d01 * U[r];
    const double l1 = sqrt(U[3] * U[3] + U[4] * U[4] + U[5] * U[5]);
#pragma unroll
    for (int r = 0; r < 3; ++r) U[3 + r] /= l1;
    U[6] = U[1] * U[5] - U[2] * U[4];
    U[7] = U[2] * U[3] - U[0] * U[5];
    U[8] = U[0] * U[4] - U[1] * U[3];
    sig[2] = U[6] * FV[6] + U[7] * FV[7] + U[8] * FV[8];
}

// makePD (EigenDecomposition.h:126-135): S <- Q max(L,0) Q^T for a symmetric n x n block (n = 2, 3), column-major.
template <int n>
__device__ inline void make_pd(double* S)
{
    double A[n * n], Q[n * n];
#pragma unroll
    for (int i = 0; i < n * n; ++i) {
        A[i] = S[i];
        Q[i] = (i % (n + 1) == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 50; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int p = 0; p < n; ++p)
#pragma unroll
            for (int q = 0; q < n; ++q) {
                const double v = A[p + n * q] * A[p + n * q];
                if (p == q) diag += v;
                else off += v;
            }
        if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
        for (int p = 0; p < n - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p + n * q];
                if (apq == 0.0) continue;
                const double theta = (A[q + n * q] - A[p + n * p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    const double x = A[k + n * p], y = A[k + n * q];
                    A[k + n * p] = c * x - s * y; A[k + n * q] = s * x + c * y;
                }
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    const double x = A[p + n * k], y = A[q + n * k];
                    A[p + n * k] = c * x - s * y; A[q + n * k] = s * x + c * y;
                }
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    const double x = Q[k + n * p], y = Q[k + n * q];
                    Q[k + n * p] = c * x - s * y; Q[k + n * q] = s * x + c * y;
                }
            }
    }
    double l[n];
    bool neg = false;
#pragma unroll
    for (int i = 0; i < n; ++i) {
        l[i] = A[i + n * i];
        if (l[i] < 0.0) { l[i] = 0.0; neg = true; }
    }
    if (!neg) return; // already PSD: the recomposition would only add rounding
#pragma unroll
    for (int c = 0; c < n; ++c)
#pragma unroll
        for (int r = 0; r < n; ++r) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < n; ++k) v += Q[r + n * k] * l[k] * Q[c + n * k];
            S[r + n * c] = v;
        }
}

// MathTools.h:163-175
__device__ __forceinline__ double clamp_small_magnitude(double x, double eps)
{
    if (x < -eps) return x;
    if (x < 0.0) return -eps;
    if (x < eps) return eps;
    return x;
}

// PSD projection of [[a, b], [b, a]]
__device__ __forceinline__ void pd2_equal_diagonal(double* B)
{
    const double l1 = fmax(B[0] + B[1], 0.0), l2 = fmax(B[0] - B[1], 0.0);
    const double d = 0.5 * (l1 + l2), o = 0.5 * (l1 - l2);
    B[0] = B[3] = d;
    B[1] = B[2] = o;
}

// SVD-space Hessian blocks of the fixed-corotated model, CorotatedIsotropic.h:116-144 +
// SvdBasedIsotropicHelper.h:223-247.  A: 3x3 symmetric (column-major 9), B01/B12/B20: 2x2 symmetric (4 each).
struct HessBlocks {
    double A[9], B01[4], B12[4], B20[4];
};
// projectABBlock (SvdBasedIsotropicHelper.h:239-247) = makePD (EigenDecomposition.h:126-135) of the four blocks.  The 2x2 blocks have
// equal diagonal entries [[a, b], [b, a]]: eigenvalues a + b, a - b on the fixed eigenvectors (1, +-1)/sqrt(2), so the clamp is closed-form
// (no iteration, no sqrt / divide).  The 3x3 block is positive definite for all but strongly compressed / inverted states: a Sylvester
// test skips the Jacobi iteration then (fp64 divide / sqrt sequences cost ~1k dependent cycles each on this part).
__device__ inline void project_blocks(HessBlocks& h)
{
    pd2_equal_diagonal(h.B01);
    pd2_equal_diagonal(h.B12);
    pd2_equal_diagonal(h.B20);
    const double m2 = h.A[0] * h.A[4] - h.A[1] * h.A[1];
    const double m3 = h.A[0] * (h.A[4] * h.A[8] - h.A[5] * h.A[5]) - h.A[1] * (h.A[1] * h.A[8] - h.A[5] * h.A[2])
        + h.A[2] * (h.A[1] * h.A[5] - h.A[4] * h.A[2]);
    const double scale = fabs(h.A[0]) + fabs(h.A[4]) + fabs(h.A[8]);
    const bool pd = h.A[0] > 1e-12 * scale && m2 > 1e-12 * scale * scale && m3 > 1e-12 * scale * scale * scale;
    if (!pd) make_pd<3>(h.A);
}
__device__ inline void corotated_blocks(const double* sig, double mu, double lambda, bool project, HessBlocks& h)
{
    const double J = sig[0] * sig[1] * sig[2];
    const double _2mu = mu * 2.0, _lambda = lambda * (J - 1.0), eps = 1e-6;
    const double S0 = sig[1] * sig[2], S1 = sig[0] * sig[2], S2 = sig[0] * sig[1];
    const double psi0 = _2mu * (sig[0] - 1.0) + _lambda * S0;
    const double psi1 = _2mu * (sig[1] - 1.0) + _lambda * S1;
    const double psi2 = _2mu * (sig[2] - 1.0) + _lambda * S2;
    h.A[0] = _2mu + lambda * S0 * S0;
    h.A[4] = _2mu + lambda * S1 * S1;
    h.A[8] = _2mu + lambda * S2 * S2;
    h.A[3] = h.A[1] = _lambda * sig[2] + lambda * S0 * S1;
    h.A[6] = h.A[2] = _lambda * sig[1] + lambda * S0 * S2;
    h.A[7] = h.A[5] = _lambda * sig[0] + lambda * S1 * S2;
    const double m01 = _2mu - _lambda * sig[2], m02 = _2mu - _lambda * sig[1], m12 = _2mu - _lambda * sig[0];
    const double p01 = (psi0 + psi1) / clamp_small_magnitude(sig[0] + sig[1], eps);
    const double p02 = (psi0 + psi2) / clamp_small_magnitude(sig[0] + sig[2], eps);
    const double p12 = (psi1 + psi2) / clamp_small_magnitude(sig[1] + sig[2], eps);
    h.B01[0] = h.B01[3] = (m01 + p01) * 0.5; h.B01[1] = h.B01[2] = (m01 - p01) * 0.5;
    h.B12[0] = h.B12[3] = (m12 + p12) * 0.5; h.B12[1] = h.B12[2] = (m12 - p12) * 0.5;
    h.B20[0] = h.B20[3] = (m02 + p02) * 0.5; h.B20[1] = h.B20[2] = (m02 - p02) * 0.5;
    if (project) project_blocks(h);
}
// Neo-Hookean (EXTENSION: the reference ships CorotatedIsotropic and LinearCorotated only; BASELINE's box-drop configuration names it):
// psi = mu/2 (|F|^2 - 3) - mu log J + lambda/2 log^2 J in the same SVD-based isotropic framework (SvdBasedIsotropicHelper.h):
// psi_i = mu s_i + c / s_i, c = lambda log J - mu; psi_ii = mu + (lambda - c) / s_i^2; psi_ij = lambda / (s_i s_j);
// (psi_i - psi_j) / (s_i - s_j) = mu - c / (s_i s_j)
__device__ inline void neohookean_blocks(const double* sig, double mu, double lambda, bool project, HessBlocks& h)
{
    const double J = sig[0] * sig[1] * sig[2], c = lambda * log(J) - mu, eps = 1e-6;
    const double i0 = 1.0 / sig[0], i1 = 1.0 / sig[1], i2 = 1.0 / sig[2];
    const double psi0 = mu * sig[0] + c * i0, psi1 = mu * sig[1] + c * i1, psi2 = mu * sig[2] + c * i2;
    h.A[0] = mu + (lambda - c) * i0 * i0;
    h.A[4] = mu + (lambda - c) * i1 * i1;
    h.A[8] = mu + (lambda - c) * i2 * i2;
    h.A[3] = h.A[1] = lambda * i0 * i1;
    h.A[6] = h.A[2] = lambda * i0 * i2;
    h.A[7] = h.A[5] = lambda * i1 * i2;
    const double m01 = mu - c * i0 * i1, m02 = mu - c * i0 * i2, m12 = mu - c * i1 * i2;
    const double p01 = (psi0 + psi1) / clamp_small_magnitude(sig[0] + sig[1], eps);
    const double p02 = (psi0 + psi2) / clamp_small_magnitude(sig[0] + sig[2], eps);
    const double p12 = (psi1 + psi2) / clamp_small_magnitude(sig[1] + sig[2], eps);
    h.B01[0] = h.B01[3] = (m01 + p01) * 0.5; h.B01[1] = h.B01[2] = (m01 - p01) * 0.5;
    h.B12[0] = h.B12[3] = (m12 + p12) * 0.5; h.B12[1] = h.B12[2] = (m12 - p12) * 0.5;
    h.B20[0] = h.B20[3] = (m02 + p02) * 0.5; h.B20[1] = h.B20[2] = (m02 - p02) * 0.5;
    if (project) project_blocks(h);
}
// flags: bit 0 = --project (PSD blocks), bits 1.. = constitutive model (0 fixed corotated, 1 neo-Hookean)
__device__ __forceinline__ void model_blocks(const double* sig, double mu, double lambda, int flags, HessBlocks& h)
{
    if ((flags >> 1) == 1) neohookean_blocks(sig, mu, lambda, (flags & 1) != 0, h);
    else corotated_blocks(sig, mu, lambda, (flags & 1) != 0, h);
}
// energy density and first Piola stress of the selected model from F and its SVD (R = U V^T, cof = J F^-T)
__device__ __forceinline__ double model_stress(int model, const double* F, const double* U, const double* sig, const double* V, double mu, double lambda,
    double* P)
{
    const double J = sig[0] * sig[1] * sig[2];
    if (model == 1) {
        const double lj = log(J), c = lambda * lj - mu;
        const double p0 = mu * sig[0] + c / sig[0], p1 = mu * sig[1] + c / sig[1], p2 = mu * sig[2] + c / sig[2];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int r = 0; r < 3; ++r) P[r + 3 * cc] = U[r] * p0 * V[cc] + U[r + 3] * p1 * V[cc + 3] + U[r + 6] * p2 * V[cc + 6];
        return 0.5 * mu * (sig[0] * sig[0] + sig[1] * sig[1] + sig[2] * sig[2] - 3.0) - mu * lj + 0.5 * lambda * lj * lj;
    }
    double R[9], cof[9];
    mm_bt(U, V, R);
    cofactor3(F, cof);
    double n2 = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const double d = F[q] - R[q];
        n2 += d * d;
        P[q] = 2.0 * mu * d + lambda * (J - 1.0) * cof[q]; // firstPiola, CorotatedIsotropic.h:157-160
    }
    return mu * n2 + 0.5 * lambda * (J - 1.0) * (J - 1.0); // psi, :151-155
}
// dPdFOfSigmaContract(Projected), SvdBasedIsotropicHelper.h:256-282: K = dPdF_Sigma : D (both in SVD space)
__device__ __forceinline__ void blocks_contract(const HessBlocks& h, const double* D, double* K)
{
    K[0] = h.A[0] * D[0] + h.A[3] * D[4] + h.A[6] * D[8];
    K[4] = h.A[1] * D[0] + h.A[4] * D[4] + h.A[7] * D[8];
    K[8] = h.A[2] * D[0] + h.A[5] * D[4] + h.A[8] * D[8];
    K[3] = h.B01[0] * D[3] + h.B01[2] * D[1]; // K01
    K[1] = h.B01[1] * D[3] + h.B01[3] * D[1]; // K10
    K[6] = h.B20[0] * D[6] + h.B20[2] * D[2]; // K02
    K[2] = h.B20[1] * D[6] + h.B20[3] * D[2]; // K20
    K[7] = h.B12[0] * D[7] + h.B12[2] * D[5]; // K12
    K[5] = h.B12[1] * D[7] + h.B12[3] * D[5]; // K21
}

} // namespace hot
