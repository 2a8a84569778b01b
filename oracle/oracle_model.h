// TEST INFRASTRUCTURE ONLY — CPU oracle for hot_b200.  Not linked into, imported by, or called from the
// product path (hot_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker / reported baseline.
//
// Parity status: PINNED against the reference's own code, compiled where it lies under /root/reference into oracle/_ref/ (oracle/Makefile,
// golden vectors + generating scripts under tests/golden/, one tests/test_oracle_*_ref.py per library):
//   SPGrid addressing (this header, section 1)                      SPGrid core                                 libspgrid_ref.so
//   B-spline weights, QR-SVD, makePD, fixed-corotated psi / P / dP  BSplines.h, ImplicitQRSVD.h, CorotatedIsotropic.h ...   libziran_ref.so
//   inexact PCG, MINRES                                             InexactConjugateGradient.h, Minres.h        libziran_ref.so
//   Galerkin hierarchy, colouring, smoothers (incl. Chebyshev + estimate2norm), V-cycle   MultigridPreconditioner.h, SquareMatrix.h   libziran_ref.so
//   L-BFGS loop                                                     LBFGS.h                                     libhot_oracle_lbfgsref.so
//   node record, sort + page activation, P2G, DOF numbering, G2P    MpmGrid.h over SPGrid_Page_Map              libmpmgrid_ref.so
//   objective: state update, energy, residual, Hessian apply, CN tolerance, buildMatrix (+ BC projection), buildDiagonal, and whole
//   implicit solves (Newton + PCG / MGPCG, HOT)                     ImplicitSolver.h, ExtendedNewtonsMethod.h, LBFGS.h, ...   libimplicit_ref.so
//   snow / von Mises return mappings                                PlasticityApplier.cpp                       libplasticity_ref.so
//   collision objects, buildInitialDvAndVnForNewton (host mirror)   AnalyticLevelSet.cpp, CollisionObject.cpp   libcollider_ref.so
//   restart files (host mirror)                                     DataManager.h, DataArray.h, BinaryIO.h      librestart_ref.so
//   command-line flags (host mirror)                                CommandLineFlags.h, Configurations.h        libflags_ref.so
// The member functions of MpmSimulationBase / MpmForceBase / FBasedMpmForceHelper themselves cannot be compiled here (Scene / DataManager /
// Particles / TBB containers / Partio absent): their particle loops are written out in the shims around the reference's grid, model and
// objective code (each shim's header lists exactly which lines).  Not in the reference at all, hence without such a pin:
// the extensions (Drucker-Prager, neo-Hookean) - those rest on closed forms, finite differences and numpy / scipy checks.
//
// Section 1: SPGrid addressing restated from Lib/SPGrid/Core/SPGrid_Mask.h:22-52,59-128,150-189,237-245.
// Section 2: quadratic B-spline weights from Lib/Ziran/Math/Splines/BSplines.h:10-29,55-81 and
//            Lib/Ziran/Math/MathTools.h:15-25.
// Section 3: small dense 3x3 helpers (column-major like Eigen, Lib/Ziran/CS/Util/Forward.h:10-13).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

namespace orc {

// ------------------------------------------------------------------------------------------------
// 1. SPGrid addressing
// ------------------------------------------------------------------------------------------------
// A node's byte offset inside the 4096^3 virtual box: the low 12 bits index a byte inside a 4 KB
// page (record bits, then z, y, x element bits, lexicographic), the bits from 12 up are the page
// index with the three axes' remaining coordinate bits interleaved (SPGrid_Mask.h:31-52).
struct Mask {
    int data_bits, block_bits, xb, yb, zb;
    uint64_t xmask, ymask, zmask;
    int elements_per_block;

    explicit Mask(int log2_struct)
    {
        data_bits = log2_struct;
        block_bits = 12 - data_bits;
        zb = block_bits / 3 + (block_bits % 3 > 0 ? 1 : 0);
        yb = block_bits / 3 + (block_bits % 3 > 1 ? 1 : 0);
        xb = block_bits / 3;
        elements_per_block = 1 << block_bits;
        // in-page element bits: z lowest, then y, then x (SPGrid_Mask.h:76-80)
        uint64_t ez = ((1ull << zb) - 1) << data_bits;
        uint64_t ey = ((1ull << yb) - 1) << (data_bits + zb);
        uint64_t ex = ((1ull << xb) - 1) << (data_bits + zb + yb);
        // page bits: z sits on bit positions == s (mod 3), y on s+1, x on s+2, with
        // s = 3 - block_bits%3, restricted to bits >= 12 (SPGrid_Mask.h:39-49).
        int s = 3 - block_bits % 3;
        uint64_t pz = 0, py = 0, px = 0;
        for (int b = 12; b < 64; ++b) {
            int r = ((b - s) % 3 + 3) % 3;
            if (r == 0) pz |= 1ull << b;
            else if (r == 1) py |= 1ull << b;
            else px |= 1ull << b;
        }
        xmask = px | ex;
        ymask = py | ey;
        zmask = pz | ez;
    }

    // deposit the low bits of v into the set bits of mask, lowest first (software pdep;
    // the reference's non-HASWELL Bit_Spread, SPGrid_Utilities.h)
    static uint64_t spread(uint64_t v, uint64_t mask)
    {
        uint64_t r = 0;
        for (int b = 0; b < 64 && mask; ++b)
            if (mask >> b & 1) {
                r |= (v & 1ull) << b;
                v >>= 1;
                mask &= ~(1ull << b);
            }
        return r;
    }
    static uint64_t pack(uint64_t v, uint64_t mask)
    {
        uint64_t r = 0;
        int o = 0;
        for (int b = 0; b < 64; ++b)
            if (mask >> b & 1) r |= (v >> b & 1ull) << o++;
        return r;
    }
    // SPGrid_Mask.h:150-166
    uint64_t linear_offset(int i, int j, int k) const
    {
        return spread((uint64_t)(int64_t)i, xmask) | spread((uint64_t)(int64_t)j, ymask) | spread((uint64_t)(int64_t)k, zmask);
    }
    // SPGrid_Mask.h:176-189
    void linear_to_coord(uint64_t off, int* ijk) const
    {
        ijk[0] = (int)pack(off, xmask);
        ijk[1] = (int)pack(off, ymask);
        ijk[2] = (int)pack(off, zmask);
    }
    // SPGrid_Mask.h:237-245: per-axis carry-isolated addition (plus the "w" lane of unowned bits)
    uint64_t packed_add(uint64_t a, uint64_t b) const
    {
        uint64_t w = ~(xmask | ymask | zmask);
        uint64_t rx = ((a | ~xmask) + (b & xmask)) & xmask;
        uint64_t ry = ((a | ~ymask) + (b & ymask)) & ymask;
        uint64_t rz = ((a | ~zmask) + (b & zmask)) & zmask;
        uint64_t rw = ((a | ~w) + (b & w)) & w;
        return rx | ry | rz | rw;
    }
};

// ------------------------------------------------------------------------------------------------
// 2. B-spline weights
// ------------------------------------------------------------------------------------------------
// MathTools.h:15-25
inline int int_floor(double x)
{
    int i = (int)x;
    return i - (i > x);
}
// MpmGrid.h:67,74 / MpmSimulationBase.cpp:1080-1083: X_index_space = one_over_dx * X, one IEEE rounding.
// volatile pins the unfused evaluation (gcc -ffp-contract=fast would otherwise be free to fuse the
// product into the following "- 0.5"), which is what bit-exact particle->cell indices are defined on.
inline double index_space(double X, double one_over_dx)
{
    volatile double p = one_over_dx * X;
    return p;
}
// BSplines.h:16-20 (degree 2): base node of index-space coordinate x
inline int base_node(double x)
{
    volatile double y = x - 0.5;
    return int_floor(y);
}

// BSplines.h:55-81, one axis.
inline void bspline_axis(double x, int& base, double w[3], double dw[3])
{
    base = base_node(x);
    double d0 = x - base;
    double z = 1.5 - d0;
    w[0] = 0.5 * (z * z);
    double d1 = d0 - 1;
    w[1] = 0.75 - d1 * d1;
    double d2 = 1 - d1;
    double zz = 1.5 - d2;
    w[2] = 0.5 * (zz * zz);
    dw[0] = -z;
    dw[1] = -2.0 * d1;
    dw[2] = zz;
}

// MpmGrid.h:55-78: X_index_space = one_over_dx * X, one_over_dx = 1/dx
struct Spline {
    int base[3];
    double w[3][3], dw[3][3];
    double one_over_dx;
    Spline(const double* X, double dx)
    {
        one_over_dx = 1.0 / dx;
        for (int d = 0; d < 3; ++d) bspline_axis(index_space(X[d], one_over_dx), base[d], w[d], dw[d]);
    }
};

// ------------------------------------------------------------------------------------------------
// 3. 3x3 helpers, column-major: M(r,c) = a[r + 3c]
// ------------------------------------------------------------------------------------------------
inline void mat_mul(const double* A, const double* B, double* C)
{ // C = A B
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) t[r + 3 * c] = A[r] * B[3 * c] + A[r + 3] * B[3 * c + 1] + A[r + 6] * B[3 * c + 2];
    std::memcpy(C, t, sizeof t);
}
inline void mat_mul_bt(const double* A, const double* B, double* C)
{ // C = A B^T
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) t[r + 3 * c] = A[r] * B[c] + A[r + 3] * B[c + 3] + A[r + 6] * B[c + 6];
    std::memcpy(C, t, sizeof t);
}
inline void mat_mul_at(const double* A, const double* B, double* C)
{ // C = A^T B
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) t[r + 3 * c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
    std::memcpy(C, t, sizeof t);
}
inline double det3(const double* A)
{
    return A[0] * (A[4] * A[8] - A[7] * A[5]) - A[3] * (A[1] * A[8] - A[7] * A[2]) + A[6] * (A[1] * A[5] - A[4] * A[2]);
}

} // namespace orc
