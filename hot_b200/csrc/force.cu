// Elastic force model of the hot path on the device: a9 (grid -> particle gradient gather), a10 (F-based force helper),
// a11 (fixed-corotated model), a12 (force rasterisation), a13 (matrix-free Hessian apply), a14 (residual / energy /
// BC projection), a18-prep (per-node CN tolerance).
//
// Reference: MpmForceBase::{evalInterpolantAndGradient, rasterizeForceToTVStack, addScaledForceDifferential,
// updatePositionBasedState} (Lib/MPM/Force/MpmForceBase.cpp:100-153,213-328), FBasedMpmForceHelper
// (Lib/MPM/Force/FBasedMpmForceHelper.cpp:25-160, .h:123-157), CorotatedIsotropic
// (Lib/Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h:110-230), SvdBasedIsotropicHelper.h:223-282,
// ImplicitSolverObjective::{computeResidual, updateState, totalEnergy, multiply, evaluatePerNodeCNTolerance}
// (Projects/multigrid/ImplicitSolver.h:128-155,237-275,667-696,741-763), MassLumpedInertia (Inertia.cpp:16-53).
//
// Re-design for the GPU
//  * updateState = ONE kernel: CTA per page group stages the (vn+dv) node tile in shared memory, each thread gathers its
//    particle's grad v, evolves F, runs the SVD and writes F, vol*P*Fn^T, U, sigma, V and the energy density (the
//    reference runs three colour-serialised / particle-parallel passes with 72-byte round trips between them).
//  * The matrix-free Hessian apply does not redo U^T(.)V / U(.)V^T per particle and per Krylov iteration like
//    firstPiolaDifferential: once per linearisation each particle's Hessian is contracted with Fn on both sides into the
//    symmetric 9x9 map  H~ : grad x_p -> vol * dP(grad x_p Fn) Fn^T  (45 doubles), so an apply is
//    gather(27) -> 81 FMA -> scatter(27) and the same H~ feeds the matrix assembly (a15).
//  * scatters use the shared prep / accumulate / gather-combine skeletons of scatter.cuh: no colour passes, no atomics in
//    the particle loop.
#include "scatter_ws.cuh"
#include "dense3.cuh"
#include "reduce.cuh"
#include <cstdlib>

namespace hot {
namespace {

constexpr int TILE = Geo::TILE;
constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

// grad f_p = sum_i f_i grad w_ip^T over the 27 stencil nodes, field staged as tile[3][TILE]; weight gradients in the
// reference's association (MpmGrid.h:272-291)
__device__ __forceinline__ void gather_gradient(const double* __restrict__ tile, const SplineEval& sp, double one_over_dx, double (&G)[9])
{
    const int tb = tile_base(sp.base[0], sp.base[1], sp.base[2]);
#pragma unroll
    for (int q = 0; q < 9; ++q) G[q] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double wi = sp.w[0][i], dwidxi = one_over_dx * sp.dw[0][i];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double wij = wi * sp.w[1][j], dwijdxi = dwidxi * sp.w[1][j], dwijdxj = wi * one_over_dx * sp.dw[1][j];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int n = tb + (i * Geo::TY + j) * Geo::TZ + k;
                const double wk = sp.w[2][k];
                const double gw[3] = {dwijdxi * wk, dwijdxj * wk, wij * one_over_dx * sp.dw[2][k]};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double nv = tile[r * TILE + n];
#pragma unroll
                    for (int c = 0; c < 3; ++c) G[r + 3 * c] += nv * gw[c];
                }
            }
        }
    }
}

// packed upper triangle of a symmetric 9x9: entry (i <= j) at j(j+1)/2 + i
__device__ __forceinline__ constexpr int tri(int i, int j) { return i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j; }

// ---- a9 + a10 + a11: ImplicitSolverObjective::updateState -------------------------------------------------------
constexpr int US_THREADS = 128;
static_assert(TILE <= 2 * US_THREADS && TILE >= US_THREADS, "tile staged in two rounds");

__global__ void __launch_bounds__(US_THREADS) k_update_state(const int* __restrict__ group_first, const int* __restrict__ tile_dof, size_t ps,
    const double* __restrict__ X, const double* __restrict__ Fn, double* __restrict__ F, const double* __restrict__ vol,
    const double* __restrict__ mu, const double* __restrict__ lam, double* __restrict__ stress, double* __restrict__ Uo, double* __restrict__ Vo,
    double* __restrict__ sigo, double* __restrict__ gradV, double dx, double one_over_dx, double dt, const double* __restrict__ vn,
    const double* __restrict__ dv, double* __restrict__ group_psi, int pf_dist, int model)
{
    __shared__ double tile[3 * TILE];
    __shared__ double s_res[1];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int id0 = tile_dof[(size_t)g * TILE + tid];
    const int id1 = tid + US_THREADS < TILE ? tile_dof[(size_t)g * TILE + tid + US_THREADS] : -1;
    const int first = group_first[g], end = group_first[g + 1];
    int pf_first = 0, pf_end = 0;
    if (pf_dist > 0 && g + pf_dist < (int)gridDim.x) {
        pf_first = group_first[g + pf_dist];
        pf_end = group_first[g + pf_dist + 1];
    }
    // the first particle's rows are requested before the tile gather
    if (first + tid < end) {
#pragma unroll
        for (int d = 0; d < 3; ++d) prefetch_l2(X + d * ps + first + tid);
#pragma unroll
        for (int q = 0; q < 9; ++q) prefetch_l2(Fn + q * ps + first + tid);
    }
    // moveNodes + the field of computeVAndGradV: v_i = vn_i + dv_i  (MpmSimulationBase.cpp:735-747)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int n = tid + u * US_THREADS, id = u ? id1 : id0;
        if (n < TILE) {
            double nv[3] = {0, 0, 0};
            if (id >= 0)
#pragma unroll
                for (int d = 0; d < 3; ++d) nv[d] = vn[3 * (size_t)id + d] + dv[3 * (size_t)id + d];
            tile[n] = nv[0]; tile[TILE + n] = nv[1]; tile[2 * TILE + n] = nv[2];
        }
    }
    __syncthreads();
    double e[1] = {0.0};
    for (int s = first + tid; s < end; s += US_THREADS) {
        if (s + US_THREADS < end) {
#pragma unroll
            for (int d = 0; d < 3; ++d) prefetch_l2(X + d * ps + s + US_THREADS);
#pragma unroll
            for (int q = 0; q < 9; ++q) prefetch_l2(Fn + q * ps + s + US_THREADS);
        }
        SplineEval sp;
        sp.eval(X, ps, s, dx, one_over_dx, true);
        double G[9], A[9], Fo[9], Fnew[9];
        gather_gradient(tile, sp, one_over_dx, G);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            gradV[q * ps + s] = G[q];
            A[q] = dt * G[q];
            Fo[q] = Fn[q * ps + s];
        }
        A[0] += 1.0; A[4] += 1.0; A[8] += 1.0;
        mm(A, Fo, Fnew); // evolveStrain: F = (I + dt gradV) Fn, FBasedMpmForceHelper.cpp:100-114
        double U[9], V[9], sig[3], P[9], T[9];
        svd3(Fnew, U, sig, V);
        const double m_ = mu[s], l_ = lam[s], vo = vol[s];
        e[0] += vo * model_stress(model, Fnew, U, sig, V, m_, l_, P);
        mm_bt(P, Fo, T); // updateImplicitState: vol P Fn^T, FBasedMpmForceHelper.cpp:72-97
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            F[q * ps + s] = Fnew[q];
            stress[q * ps + s] = vo * T[q];
            Uo[q * ps + s] = U[q];
            Vo[q * ps + s] = V[q];
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) sigo[d * ps + s] = sig[d];
    }
    prefetch_rows_l2(X, ps, 3, pf_first, pf_end, tid, US_THREADS);
    prefetch_rows_l2(Fn, ps, 9, pf_first, pf_end, tid, US_THREADS);
    block_sum<1>(e, s_res);
    if (tid == 0) group_psi[g] = s_res[0];
}

// inertia + gravity terms of totalEnergy (ImplicitSolver.h:254-275, Inertia.cpp:16-30): [sum m |dv|^2, sum m g.dv]
struct EnergyNodesF {
    const double *dv, *mass;
    const unsigned char* own; // partitioned object: shared nodes count on one rank
    double g0, g1, g2;
    __device__ void operator()(long i, double (&acc)[2]) const
    {
        if (own && !own[i]) return;
        const double a = dv[3 * i], b = dv[3 * i + 1], c = dv[3 * i + 2], m = mass[i];
        acc[0] += (a * a + b * b + c * c) * m;
        acc[1] += (g0 * a + g1 * b + g2 * c) * m;
    }
};
struct SumF {
    const double* a;
    __device__ void operator()(long i, double (&acc)[1]) const { acc[0] += a[i]; }
};

// ---- contracted particle Hessian H~ (see file header) ---------------------------------------------------------------
__global__ void __launch_bounds__(128) k_build_hessian(long n, size_t ps, const double* __restrict__ Fn, const double* __restrict__ vol,
    const double* __restrict__ mu, const double* __restrict__ lam, const double* __restrict__ Ui, const double* __restrict__ Vi,
    const double* __restrict__ sigi, int project, double* __restrict__ H)
{
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double U[9], V[9], F0[9], sig[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        U[q] = Ui[q * ps + s];
        V[q] = Vi[q * ps + s];
        F0[q] = Fn[q * ps + s];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) sig[d] = sigi[d * ps + s];
    HessBlocks hb;
    model_blocks(sig, mu[s], lam[s], project, hb);
    const double vo = vol[s];
    // FV = Fn V (3x3): D for the basis gradient e_a e_d^T is  (U^T e_a) (row d of Fn V)
    double FV[9];
    mm(F0, V, FV);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int col = a + 3 * d; // column-major index of grad x entry (a, d)
            double D[9], K[9], t[9], dP[9], T[9];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) D[r + 3 * c] = U[a + 3 * r] * FV[d + 3 * c];
            blocks_contract(hb, D, K);
            mm(U, K, t);
            mm_bt(t, V, dP);
            mm_bt(dP, F0, T);
#pragma unroll
            for (int row = 0; row < 9; ++row)
                if (row <= col) H[(size_t)tri(row, col) * ps + s] = vo * T[row];
        }
}

// a13, first half: T_p = dt^2 * H~_p : grad x_p for every particle of a page group (CTA per group like k_update_state: the
// group's x node tile is staged once, one thread per particle).  The scatter half reuses the force rasterisation kernel.
// TMA = true: the X and H~ rows of the group (48 runs, 360 + 24 bytes per particle: the HBM stream of the whole matrix-free
// apply) are staged in shared memory by bulk copies issued by one thread, HG_CAP particles per chunk.
constexpr int HG_CAP = US_THREADS;
constexpr size_t HG_SMEM = (size_t)48 * (HG_CAP + 2) * sizeof(double);
template <bool TMA>
__global__ void __launch_bounds__(US_THREADS, TMA ? 4 : 6) k_hessian_gather(const int* __restrict__ group_first, const int* __restrict__ tile_dof,
    size_t ps, const double* __restrict__ X, const double* __restrict__ H, double dx, double one_over_dx, double scale,
    const double* __restrict__ x, double* __restrict__ Tout, int pf_dist)
{
    __shared__ double tile[3 * TILE];
    __shared__ __align__(8) unsigned long long s_bar;
    extern __shared__ __align__(16) double hg_rows[]; // [48][HG_CAP + 2]
    const int g = blockIdx.x, tid = threadIdx.x;
    if (TMA && tid == 0) mbar_init(&s_bar, 1);
    const int id0 = tile_dof[(size_t)g * TILE + tid];
    const int id1 = tid + US_THREADS < TILE ? tile_dof[(size_t)g * TILE + tid + US_THREADS] : -1;
    const int first = group_first[g], end = group_first[g + 1];
    int pf_first = 0, pf_end = 0;
    if (!TMA && pf_dist > 0 && g + pf_dist < (int)gridDim.x) {
        pf_first = group_first[g + pf_dist];
        pf_end = group_first[g + pf_dist + 1];
    }
    auto issue_chunk = [&](int c0) {
        const int cn = min(HG_CAP, end - c0), start = c0 & ~1, cnt = (c0 + cn - start + 1) & ~1;
        fence_proxy_async();
        mbar_expect_tx(&s_bar, 48u * cnt * 8u);
#pragma unroll
        for (int d = 0; d < 3; ++d) bulk_load(hg_rows + d * (HG_CAP + 2), X + d * ps + start, cnt * 8u, &s_bar);
        for (int q = 0; q < 45; ++q) bulk_load(hg_rows + (3 + q) * (HG_CAP + 2), H + q * ps + start, cnt * 8u, &s_bar);
    };
    if (TMA) {
        __syncthreads(); // barrier initialised
        if (tid == 0 && first < end) issue_chunk(first);
    }
    else if (first + tid < end) {
        // (an L2 prefetch of the particle's 45 H~ rows ahead of the staging was measured: it RAISED the DRAM reads from 381 MB
        //  to 539 MB per launch - prefetched lines were evicted before use - so only the position rows are requested early)
#pragma unroll
        for (int d = 0; d < 3; ++d) prefetch_l2(X + d * ps + first + tid);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int n = tid + u * US_THREADS, id = u ? id1 : id0;
        if (n < TILE) {
            double nv[3] = {0, 0, 0};
            if (id >= 0)
#pragma unroll
                for (int d = 0; d < 3; ++d) nv[d] = x[3 * (size_t)id + d];
            tile[n] = nv[0]; tile[TILE + n] = nv[1]; tile[2 * TILE + n] = nv[2];
        }
    }
    __syncthreads();
    unsigned phase = 0;
    for (int c0 = first; c0 < end; c0 += TMA ? HG_CAP : end - first) {
        const int c1 = TMA ? min(c0 + HG_CAP, end) : end, start = c0 & ~1;
        if (TMA) {
            if (c0 != first) {
                __syncthreads(); // every thread is done with the previous chunk's rows
                if (tid == 0) issue_chunk(c0);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1;
        }
        for (int s = c0 + tid; s < c1; s += US_THREADS) {
            SplineEval sp;
            if (TMA) {
                const double Xp[3] = {hg_rows[s - start], hg_rows[(HG_CAP + 2) + s - start], hg_rows[2 * (HG_CAP + 2) + s - start]};
                sp.eval3(Xp, dx, one_over_dx, true);
            }
            else sp.eval(X, ps, s, dx, one_over_dx, true);
            double G[9], T[9];
            gather_gradient(tile, sp, one_over_dx, G);
#pragma unroll
            for (int q = 0; q < 9; ++q) T[q] = 0.0;
#pragma unroll
            for (int j = 0; j < 9; ++j)
#pragma unroll
                for (int i = 0; i <= j; ++i) {
                    const double h = TMA ? hg_rows[(3 + tri(i, j)) * (HG_CAP + 2) + s - start] : H[(size_t)tri(i, j) * ps + s];
                    T[i] += h * G[j];
                    if (i != j) T[j] += h * G[i];
                }
#pragma unroll
            for (int q = 0; q < 9; ++q) Tout[q * ps + s] = scale * T[q];
        }
    }
    // long distance (the group that takes over this CTA's slot): 48 rows = ~100 KB per group, 6 CTAs per SM x distance in L2
    prefetch_rows_l2(X, ps, 3, pf_first, pf_end, tid, US_THREADS);
    prefetch_rows_l2(H, ps, 45, pf_first, pf_end, tid, US_THREADS);
}

// ---- scatters -------------------------------------------------------------------------------------------------------
// Common record of the two vector scatters (a12 force, a13 Hessian apply): node value = T grad w with a per-particle 3x3 T.
//   record: X(3)  T(9, column-major)
struct TGradScatter {
    static constexpr int NCH = 3;
    // column form (k_column_scatter).  Prepared record (30 doubles, 16-byte words):
    //   [0..5] (wx_t, gx_t) t = 0..2 | [6..11] (wy_t, gy_t) | [12..14] wz [15..17] gz | [18..26] T (column-major) | pad;  g = dw / dx
    static constexpr int REC = 30;
    __device__ __forceinline__ static void prep(const double* X, size_t ps, size_t s, double dx, double one_over_dx, const double (&T)[9], double* __restrict__ r)
    {
        double Xp[3], d0n[3], w[3][3], g[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) Xp[d] = X[d * ps + s];
        prep_weights<true>(Xp, dx, one_over_dx, w, g, d0n);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            sts2(r + 2 * t, w[0][t], g[0][t]);
            sts2(r + 6 + 2 * t, w[1][t], g[1][t]);
        }
        sts2(r + 12, w[2][0], w[2][1]); sts2(r + 14, w[2][2], g[2][0]); sts2(r + 16, g[2][1], g[2][2]);
        sts2(r + 18, T[0], T[1]); sts2(r + 20, T[2], T[3]); sts2(r + 22, T[4], T[5]); sts2(r + 24, T[6], T[7]); sts2(r + 26, T[8], 0.0);
    }
    // node (i, j, k): T grad w,  grad w = (gx_i wy_j wz_k, wx_i gy_j wz_k, wx_i wy_j gz_k)
    __device__ __forceinline__ static void accumulate_col(const double* __restrict__ rec, int i, int j, double (&acc)[3][3])
    {
        double wx, gx, wy, gy, wz[3], gz[3], T[10];
        lds2(rec + 2 * i, wx, gx); lds2(rec + 6 + 2 * j, wy, gy);
        lds2(rec + 12, wz[0], wz[1]); lds2(rec + 14, wz[2], gz[0]); lds2(rec + 16, gz[1], gz[2]);
        lds2(rec + 18, T[0], T[1]); lds2(rec + 20, T[2], T[3]); lds2(rec + 22, T[4], T[5]); lds2(rec + 24, T[6], T[7]); lds2(rec + 26, T[8], T[9]);
        const double c0 = gx * wy, c1 = wx * gy, c2 = wx * wy;
        double ab[3], cc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            ab[r] = fma(T[3 + r], c1, T[r] * c0);
            cc[r] = T[6 + r] * c2;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) acc[k][r] = fma(cc[r], gz[k], fma(ab[r], wz[k], acc[k][r]));
        }
    }
    // plane form: COMPACT record (14 doubles: 42 KB for a full page, 5 CTAs per SM): [0..2] xi - base per axis, [3..11] T, pad;
    // the (cell, plane) thread re-derives weights and weight derivatives in the reference's operation order
    static constexpr int RECP = 14;
    __device__ __forceinline__ static void prep_plane(const double* X, size_t ps, size_t s, double one_over_dx, const double (&T)[9], double* __restrict__ r)
    {
        double d0[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double xi;
            const int b = base_node_of(X[d * ps + s], one_over_dx, &xi);
            d0[d] = xi - (double)b;
        }
        sts2(r + 0, d0[0], d0[1]); sts2(r + 2, d0[2], T[0]);
        sts2(r + 4, T[1], T[2]); sts2(r + 6, T[3], T[4]); sts2(r + 8, T[5], T[6]); sts2(r + 10, T[7], T[8]);
    }
    // the 9 nodes (j, k) of x-plane i
    __device__ __forceinline__ static void accumulate_plane(const double* __restrict__ rec, double one_over_dx, int i, double (&acc)[9][3])
    {
        double v[12];
#pragma unroll
        for (int u = 0; u < 6; ++u) lds2(rec + 2 * u, v[2 * u], v[2 * u + 1]);
        accumulate_rec(v, one_over_dx, i, acc);
    }
    __device__ __forceinline__ static void accumulate_rec(const double (&v)[12], double one_over_dx, int i, double (&acc)[9][3])
    {
        const double d0[3] = {v[0], v[1], v[2]}, T[9] = {v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]};
        double w[3][3], dw[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) bspline_axis(d0[d], w[d], dw[d]);
        const double wx = i == 0 ? w[0][0] : (i == 1 ? w[0][1] : w[0][2]);
        const double gx = one_over_dx * (i == 0 ? dw[0][0] : (i == 1 ? dw[0][1] : dw[0][2]));
        double wy[3], gy[3], wz[3], gz[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            wy[t] = w[1][t]; gy[t] = one_over_dx * dw[1][t];
            wz[t] = w[2][t]; gz[t] = one_over_dx * dw[2][t];
        }
        double tx[3], ty[3], tz[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            tx[r] = T[r] * gx; // column 0 x d/dx part
            ty[r] = T[3 + r] * wx;
            tz[r] = T[6 + r] * wx;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double ab[3], cc[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                ab[r] = fma(ty[r], gy[j], tx[r] * wy[j]);
                cc[r] = tz[r] * wy[j];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int r = 0; r < 3; ++r) acc[j * 3 + k][r] = fma(cc[r], gz[k], fma(ab[r], wz[k], acc[j * 3 + k][r]));
        }
    }
};

// a12: rasterizeForceToTVStack: f_i -= scale * (vol P Fn^T) grad w   (MpmForceBase.cpp:100-153)
struct ForcePolicy {
    static constexpr int NCH = 3;
    struct Args {
        size_t ps;
        const double *X, *stress;
        double dx, one_over_dx, scale;
        const int* g_idx;
        double* out; // DOF vector
    };
    static constexpr int REC = TGradScatter::REC, MINB = 4; // 45 KB of records per CTA
    static constexpr bool PLANE = true; // measured (C2): plane with compact records 0.089 ms (5 CTAs per SM), column 0.103 ms
    __device__ __forceinline__ static void prep(const Args& a, size_t s, double* __restrict__ r)
    {
        double T[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) T[q] = -a.scale * a.stress[q * a.ps + s];
        TGradScatter::prep(a.X, a.ps, s, a.dx, a.one_over_dx, T, r);
    }
    __device__ __forceinline__ static void accumulate_col(const double* __restrict__ rec, int i, int j, double, double, double (&acc)[3][3])
    {
        TGradScatter::accumulate_col(rec, i, j, acc);
    }
    static constexpr int RECP = TGradScatter::RECP, PMINB = 5;
    __device__ __forceinline__ static void prep_plane(const Args& a, size_t s, double* __restrict__ r)
    {
        double T[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) T[q] = -a.scale * a.stress[q * a.ps + s];
        TGradScatter::prep_plane(a.X, a.ps, s, a.one_over_dx, T, r);
    }
    __device__ __forceinline__ static void accumulate_plane(const Args& a, const double* __restrict__ rec, int i, double, double (&acc)[9][3])
    {
        TGradScatter::accumulate_plane(rec, a.one_over_dx, i, acc);
    }
    // ---- ws form (scatter_ws.cuh): 12 raw rows X, T staged by TMA; record = the compact one above (12 doubles = 6 units) in place
    static constexpr int ROWS = 12, UNITS = 6, WS_ID = 1;
    __device__ __forceinline__ static const double* row(const Args& a, int r)
    {
        return r < 3 ? a.X + (size_t)r * a.ps : a.stress + (size_t)(r - 3) * a.ps;
    }
    __device__ __forceinline__ static int prep_ws(const Args& a, const double (&raw)[12], double (&r)[12])
    {
        int cb[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double xi;
            cb[d] = base_node_of(raw[d], a.one_over_dx, &xi);
            r[d] = xi - (double)cb[d];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) r[3 + q] = -a.scale * raw[3 + q];
        return (((cb[0] & (Geo::BX - 1)) << Geo::yb | (cb[1] & (Geo::BY - 1))) << Geo::zb) | (cb[2] & (Geo::BZ - 1));
    }
    __device__ __forceinline__ static void accumulate_rec(const Args& a, const double (&v)[12], int i, double, double (&acc)[9][3])
    {
        TGradScatter::accumulate_rec(v, a.one_over_dx, i, acc);
    }
    static constexpr bool DOF = true;
    __device__ __forceinline__ static void prefetch(const Args& a, int first, int end, int tid, int nt)
    {
        prefetch_rows_l2(a.X, a.ps, 3, first, end, tid, nt);
        prefetch_rows_l2(a.stress, a.ps, 9, first, end, tid, nt);
    }
    __device__ __forceinline__ static void flush1(const Args& a, long id, int ch, double v) { atomicAdd(a.out + 3 * (size_t)id + ch, v); }
};

// a18: nodeCNTol_i += w_ip m_p ||dPdF(F = I)||_F   (ImplicitSolver.h:667-696, FBasedMpmForceHelper.h:123-157)
struct CNTolPolicy {
    static constexpr int NCH = 1;
    struct Args {
        size_t ps;
        const double *X, *M, *mu, *lam;
        double dx, one_over_dx;
        int project;
        const int* g_idx;
        double* out; // per node
    };
    // dPdF = Q blockdiag(A, B01, B12, B20) Q^T with Q orthogonal, so ||dPdF||_F^2 = sum of the block norms^2
    // column form: [0..8] w[3][3], [9] m_p ||dPdF(I)||_F
    static constexpr int REC = 10, MINB = 4;
    static constexpr bool PLANE = false;
    __device__ __forceinline__ static void prep(const Args& a, size_t s, double* __restrict__ r)
    {
        const double one[3] = {1.0, 1.0, 1.0};
        HessBlocks hb;
        model_blocks(one, a.mu[s], a.lam[s], a.project, hb);
        double n2 = 0.0;
#pragma unroll
        for (int q = 0; q < 9; ++q) n2 += hb.A[q] * hb.A[q];
#pragma unroll
        for (int q = 0; q < 4; ++q) n2 += hb.B01[q] * hb.B01[q] + hb.B12[q] * hb.B12[q] + hb.B20[q] * hb.B20[q];
        double Xp[3], d0n[3], w[3][3], g[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) Xp[d] = a.X[d * a.ps + s];
        prep_weights<false>(Xp, a.dx, a.one_over_dx, w, g, d0n);
#pragma unroll
        for (int q = 0; q < 9; ++q) r[q] = w[q / 3][q % 3];
        r[9] = a.M[s] * sqrt(n2);
    }
    __device__ __forceinline__ static void accumulate_col(const double* __restrict__ rec, int i, int j, double, double, double (&acc)[3][1])
    {
        const double v = rec[9] * (rec[i] * rec[3 + j]);
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k][0] = fma(v, rec[6 + k], acc[k][0]);
    }
    static constexpr int RECP = REC, PMINB = 3;
    __device__ __forceinline__ static void prep_plane(const Args& a, size_t s, double* __restrict__ r) { prep(a, s, r); }
    __device__ __forceinline__ static void accumulate_plane(const Args&, const double* __restrict__ rec, int i, double, double (&acc)[9][1])
    {
        const double v = rec[9] * rec[i];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[j * 3 + k][0] = fma(v, rec[3 + j] * rec[6 + k], acc[j * 3 + k][0]);
    }
    static constexpr bool DOF = true;
    __device__ __forceinline__ static void prefetch(const Args&, int, int, int, int) {} // once per step: not worth it
    __device__ __forceinline__ static void flush1(const Args& a, long id, int, double v) { atomicAdd(a.out + id, v); }
};

// r = dt m g - m dv   (the gravity and inertia terms of computeResidual, ImplicitSolver.h:133-145, Inertia.cpp:33-41)
__global__ void k_residual_init(int n, const double* __restrict__ mass, const double* __restrict__ dv, double dtg0, double dtg1, double dtg2,
    double* __restrict__ r)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int i = t / 3, d = t - 3 * i;
    const double m = mass[i];
    r[t] = (d == 0 ? dtg0 : (d == 1 ? dtg1 : dtg2)) * m - m * dv[t];
}
// b = M x   (MassLumpedInertia::addScaledForceDifferential with scale -dt^2, Inertia.cpp:45-53)
__global__ void k_mass_mul(int n, const double* __restrict__ mass, const double* __restrict__ x, double* __restrict__ b)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3 * n) b[t] = mass[t / 3] * x[t];
}
// objective.project (MultigridSimulation.h:104-125)
__global__ void k_bc_project(int n_bc, int mode, const int* __restrict__ node, const int* __restrict__ slip, const double* __restrict__ P,
    double* __restrict__ v)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bc) return;
    double* x = v + 3 * (size_t)node[b];
    if (mode == 1) {
        x[0] = 0.0;
        if (!slip[b]) x[1] = x[2] = 0.0;
    }
    else {
        const double* Pm = P + 9 * (size_t)b;
        const double x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) x[r] = Pm[r] * x0 + Pm[r + 3] * x1 + Pm[r + 6] * x2;
    }
}
// transformResidual / recoverSolution (ImplicitSolver.h:106-125)
__global__ void k_bc_rotate(int n_bc, const int* __restrict__ node, const int* __restrict__ slip, const double* __restrict__ R,
    double* __restrict__ v)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bc || !slip[b]) return;
    double* x = v + 3 * (size_t)node[b];
    const double* Rm = R + 9 * (size_t)b;
    const double x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) x[r] = Rm[r] * x0 + Rm[r + 3] * x1 + Rm[r + 6] * x2;
}
// Newton initial guess (MpmSimulationBase.cpp:1177-1180): dv = g dt on free nodes ...
__global__ void k_fill3(int n, double a, double b, double c, double* __restrict__ v)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int d = t % 3;
    v[t] = d == 0 ? a : (d == 1 ? b : c);
}
// ... and the collider velocity difference on BC nodes
__global__ void k_set_bc_dv(int n_bc, const int* __restrict__ node, const double* __restrict__ dv_bc, double* __restrict__ dv)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bc) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) dv[3 * (size_t)node[b] + d] = dv_bc ? dv_bc[3 * b + d] : 0.0;
}
__global__ void k_add(long n, const double* __restrict__ x, double* __restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
__global__ void k_cn_finish(int n, const double* __restrict__ mass, double factor, double* __restrict__ tol)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tol[i] = mass[i] != 0.0 ? tol[i] * (factor / mass[i]) : 1.0; // (mass 0: a node only other ranks touch)
}

} // namespace

// CorotatedIsotropic<T,3> evaluated per deformation gradient with the device routines of the force kernels (svd3, cofactor3,
// corotated_blocks, blocks_contract): updateScratch + psi + firstPiola + firstPiolaDifferential + firstPiolaDerivative
// (CorotatedIsotropic.h:78-230).  Arrays are column-major per item; any output may be null.
__global__ void k_corotated_eval(long n, const double* __restrict__ F, double mu, double lambda, int project, const double* __restrict__ dF,
    double* __restrict__ psi, double* __restrict__ P, double* __restrict__ dP, double* __restrict__ dPdF, double* __restrict__ Uo,
    double* __restrict__ sigo, double* __restrict__ Vo)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double Fm[9], U[9], V[9], sig[3], Pm[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) Fm[q] = F[9 * t + q];
    svd3(Fm, U, sig, V);
    const double e = model_stress(project >> 1, Fm, U, sig, V, mu, lambda, Pm);
    if (P)
#pragma unroll
        for (int q = 0; q < 9; ++q) P[9 * t + q] = Pm[q];
    if (psi) psi[t] = e;
    if (Uo)
#pragma unroll
        for (int q = 0; q < 9; ++q) { Uo[9 * t + q] = U[q]; Vo[9 * t + q] = V[q]; }
    if (sigo)
#pragma unroll
        for (int d = 0; d < 3; ++d) sigo[3 * t + d] = sig[d];
    if (!dP && !dPdF) return;
    HessBlocks hb;
    model_blocks(sig, mu, lambda, project, hb);
    auto differential = [&](const double* dFm, double* out) { // dP = U (dPdF_Sigma : (U^T dF V)) V^T
        double T1[9], D[9], K[9], T2[9];
        mm_at(U, dFm, T1);
        mm(T1, V, D);
        blocks_contract(hb, D, K);
        mm(U, K, T2);
        mm_bt(T2, V, out);
    };
    if (dP) {
        double d[9], o[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) d[q] = dF[9 * t + q];
        differential(d, o);
#pragma unroll
        for (int q = 0; q < 9; ++q) dP[9 * t + q] = o[q];
    }
    if (dPdF)
        for (int rs = 0; rs < 9; ++rs) {
            double d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, o[9];
            d[rs] = 1.0;
            differential(d, o);
            for (int ij = 0; ij < 9; ++ij) dPdF[81 * t + ij + 9 * rs] = o[ij];
        }
}

// FBasedMpmForceHelper::backupStrain / restoreStrain, FBasedMpmForceHelper.cpp:25-44
int corotated_eval(Sim* s, long n, const double* F, double mu, double lambda, int project, const double* dF, double* psi, double* P,
    double* dP, double* dPdF, double* U, double* sigma, double* V)
{
    if (n <= 0) return 0;
    // (flags: bit 0 project, bits 1.. the handle's constitutive model)
    k_corotated_eval<<<nblk(n), TPB, 0, s->stream>>>(n, F, mu, lambda, (project ? 1 : 0) | (s->constitutive_model << 1), dF, psi, P, dP, dPdF, U, sigma, V);
    HOT_LAUNCHED(s);
    return 0;
}

int backup_strain(Sim* s)
{
    if (!s->sorted) return fail(s, "backupStrain: call hot_sort_and_activate first");
    HOT_CUDA(s->P.Fn.reserve(9 * s->P.stride));
    HOT_CUDA(cudaMemcpyAsync(s->P.Fn.p, s->P.F.p, 9 * s->P.stride * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    s->strain_backed_up = true;
    s->state_valid = s->hessian_valid = false;
    return 0;
}
int restore_strain(Sim* s)
{
    if (!s->strain_backed_up) return fail(s, "restoreStrain: no backup (call hot_backup_strain first)");
    HOT_CUDA(cudaMemcpyAsync(s->P.F.p, s->P.Fn.p, 9 * s->P.stride * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

// host arrays in; also sets the Newton initial guess (buildInitialDvAndVnForNewton, MpmSimulationBase.cpp:1139-1184)
int set_bc(Sim* s, int mode, int n_bc, const int* node_id, const double* P, const double* R, const double* Rinv, const int* slip,
    const double* dv_bc)
{
    if (!s->p2g_done) return fail(s, "hot_set_bc: call hot_p2g first");
    if (n_bc < 0 || (n_bc > 0 && !node_id)) return fail(s, "hot_set_bc: bad BC table");
    for (int b = 0; b < n_bc; ++b)
        if (node_id[b] < 0 || node_id[b] >= s->num_nodes) return fail(s, "hot_set_bc: node id out of range");
    cudaStream_t st = s->stream;
    const size_t nb = n_bc > 0 ? n_bc : 1;
    HOT_CUDA(s->bc_node.reserve(nb));
    HOT_CUDA(s->bc_slip.reserve(nb));
    HOT_CUDA(s->bc_P.reserve(9 * nb));
    HOT_CUDA(s->bc_R.reserve(9 * nb));
    HOT_CUDA(s->bc_Rinv.reserve(9 * nb));
    HOT_CUDA(s->work[0].reserve(3 * nb));
    s->bc_mode = mode;
    s->n_bc = n_bc;
    HOT_CUDA(cudaMemsetAsync(s->bc_slip.p, 0, nb * sizeof(int), st));
    HOT_CUDA(cudaMemsetAsync(s->bc_P.p, 0, 9 * nb * sizeof(double), st));
    HOT_CUDA(cudaMemsetAsync(s->bc_R.p, 0, 9 * nb * sizeof(double), st));
    HOT_CUDA(cudaMemsetAsync(s->bc_Rinv.p, 0, 9 * nb * sizeof(double), st));
    if (n_bc > 0) {
        HOT_CUDA(cudaMemcpyAsync(s->bc_node.p, node_id, n_bc * sizeof(int), cudaMemcpyHostToDevice, st));
        if (slip) HOT_CUDA(cudaMemcpyAsync(s->bc_slip.p, slip, n_bc * sizeof(int), cudaMemcpyHostToDevice, st));
        if (P) HOT_CUDA(cudaMemcpyAsync(s->bc_P.p, P, 9 * (size_t)n_bc * sizeof(double), cudaMemcpyHostToDevice, st));
        if (R) HOT_CUDA(cudaMemcpyAsync(s->bc_R.p, R, 9 * (size_t)n_bc * sizeof(double), cudaMemcpyHostToDevice, st));
        if (Rinv) HOT_CUDA(cudaMemcpyAsync(s->bc_Rinv.p, Rinv, 9 * (size_t)n_bc * sizeof(double), cudaMemcpyHostToDevice, st));
        if (dv_bc) HOT_CUDA(cudaMemcpyAsync(s->work[0].p, dv_bc, 3 * (size_t)n_bc * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    const int nn = s->num_nodes;
    k_fill3<<<nblk(3 * (long)nn), TPB, 0, st>>>(nn, s->gravity[0] * s->dt, s->gravity[1] * s->dt, s->gravity[2] * s->dt, s->dv.p);
    HOT_LAUNCHED(s);
    if (n_bc > 0) {
        k_set_bc_dv<<<nblk(n_bc), TPB, 0, st>>>(n_bc, s->bc_node.p, dv_bc ? s->work[0].p : nullptr, s->dv.p);
        HOT_LAUNCHED(s);
    }
    HOT_CUDA(cudaStreamSynchronize(st)); // the host arrays may be released by the caller
    s->state_valid = s->hessian_valid = false;
    return 0;
}

int bc_project(Sim* s, double* v)
{
    if (s->n_bc <= 0) return 0;
    k_bc_project<<<nblk(s->n_bc), TPB, 0, s->stream>>>(s->n_bc, s->bc_mode, s->bc_node.p, s->bc_slip.p, s->bc_P.p, v);
    HOT_LAUNCHED(s);
    return 0;
}
int bc_rotate(Sim* s, double* v, bool inverse)
{
    if (s->n_bc <= 0 || s->bc_mode != 1) return 0;
    k_bc_rotate<<<nblk(s->n_bc), TPB, 0, s->stream>>>(s->n_bc, s->bc_node.p, s->bc_slip.p, inverse ? s->bc_Rinv.p : s->bc_R.p, v);
    HOT_LAUNCHED(s);
    return 0;
}

// ImplicitSolverObjective::updateState (ImplicitSolver.h:237-252) on the device-resident dv
int update_state(Sim* s, bool want_energy, double* energy)
{
    if (!s->p2g_done) return fail(s, "updateState: call hot_p2g first");
    if (!s->strain_backed_up) return fail(s, "updateState: call hot_backup_strain first (startBackwardEuler, MultigridSimulation.h:167-186)");
    cudaStream_t st = s->stream;
    const size_t ps = s->P.stride;
    HOT_CUDA(s->f_stress.reserve(9 * ps));
    HOT_CUDA(s->f_U.reserve(9 * ps));
    HOT_CUDA(s->f_V.reserve(9 * ps));
    HOT_CUDA(s->f_sig.reserve(3 * ps));
    HOT_CUDA(s->P.gradV.reserve(9 * ps));
    HOT_CUDA(s->group_psi.reserve(s->n_groups));
    {
        KTime t(s, KC_STRESS);
        if (s->g1 > s->g0)
        k_update_state<<<(unsigned)(s->g1 - s->g0), US_THREADS, 0, st>>>(s->group_first.p + s->g0, s->tile_dof.p + (size_t)s->g0 * TILE, ps, s->P.X.p,
            s->P.Fn.p, s->P.F.p, s->P.vol.p, s->P.mu.p, s->P.lam.p, s->f_stress.p, s->f_U.p, s->f_V.p, s->f_sig.p, s->P.gradV.p, s->dx, 1.0 / s->dx,
            s->dt, s->vn.p, s->dv.p, s->group_psi.p, pf_distance(s, 4), s->constitutive_model);
        HOT_LAUNCHED(s);
    }
    s->state_valid = true;
    s->hessian_valid = false;
    if (want_energy) {
        KTime t(s, KC_BLAS1);
        HOT_CUDA(s->red_out.reserve(64));
        // own particles' strain energy + own nodes' inertia / gravity terms, summed over the ranks
        int rc = reduce_to<1>(s, s->g1 - s->g0, SumF{s->group_psi.p}, s->red_out.p + 8, nullptr);
        if (rc) return rc;
        rc = reduce_to<2>(s, s->num_nodes, EnergyNodesF{s->dv.p, s->mass_matrix.p, s->world > 1 ? s->own_node.p : nullptr, s->gravity[0], s->gravity[1], s->gravity[2]},
            s->red_out.p + 9, nullptr);
        if (rc) return rc;
        HOT_CUDA(cudaMemcpyAsync(s->h_red, s->red_out.p + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        HOT_CUDA(cudaStreamSynchronize(st));
        double h[3] = {s->h_red[0], s->h_red[1], s->h_red[2]};
        rc = dist_allreduce_host(s, h, 3, 0);
        if (rc) return rc;
        if (energy) *energy = h[0] + h[1] / 2 - s->dt * h[2];
    }
    return 0;
}

// FBasedMpmForceHelper::totalEnergy (FBasedMpmForceHelper.cpp:116-136): sum_p vol psi(F_p) of the last updateState (this rank's particles)
int strain_energy(Sim* s, double* e)
{
    if (!s->state_valid) return fail(s, "totalEnergy: call hot_update_state first");
    HOT_CUDA(s->red_out.reserve(64));
    if (!s->h_red) HOT_CUDA(cudaMallocHost((void**)&s->h_red, 64 * sizeof(double)));
    int rc = reduce_to<1>(s, s->g1 - s->g0, SumF{s->group_psi.p}, s->red_out.p + 8, nullptr);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->h_red, s->red_out.p + 8, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    *e = s->h_red[0];
    return 0;
}

// One particle->grid scatter of this rank's page groups into a DOF array.  Single GPU: straight into `out` (which already
// holds the node-local terms).  Partitioned: into a zeroed scratch array, summed over the ranks on the interface nodes, then
// added to `out` - so node-local terms are counted once.
template <class Policy>
int scatter_dispatch(Sim* s, const typename Policy::Args& a) { return launch_scatter_best<Policy>(s, a); }
template <>
int scatter_dispatch<CNTolPolicy>(Sim* s, const CNTolPolicy::Args& a) { return launch_scatter<CNTolPolicy>(s, a); } // once per step: round-1 column form
template <class Policy>
int scatter_to_dofs(Sim* s, typename Policy::Args a, double* Policy::Args::*target, double* out, int comps)
{
    cudaStream_t st = s->stream;
    const size_t m = (size_t)comps * s->num_nodes;
    double* dst = out;
    if (s->world > 1) {
        HOT_CUDA(s->scat_tmp.reserve(m > 0 ? m : 1));
        HOT_CUDA(cudaMemsetAsync(s->scat_tmp.p, 0, m * sizeof(double), st));
        dst = s->scat_tmp.p;
    }
    a.*target = dst;
    if (s->g1 > s->g0) {
        int rc = scatter_dispatch<Policy>(s, a);
        if (rc) return rc;
    }
    if (s->world > 1) {
        int rc = dist_exchange_shared(s, dst, comps);
        if (rc) return rc;
        k_add<<<nblk((long)m), TPB, 0, st>>>((long)m, dst, out);
        HOT_LAUNCHED(s);
    }
    return 0;
}

// ImplicitSolverObjective::computeResidual, ImplicitSolver.h:128-155
int compute_residual(Sim* s, double* r)
{
    if (!s->state_valid) return fail(s, "computeResidual: call hot_update_state first");
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    KTime t(s, KC_FORCE);
    k_residual_init<<<nblk(3 * (long)nn), TPB, 0, st>>>(nn, s->mass_matrix.p, s->dv.p, s->dt * s->gravity[0], s->dt * s->gravity[1],
        s->dt * s->gravity[2], r);
    HOT_LAUNCHED(s);
    ForcePolicy::Args a{s->P.stride, s->P.X.p, s->f_stress.p, s->dx, 1.0 / s->dx, s->dt, s->g_idx.p, r};
    int rc = scatter_to_dofs<ForcePolicy>(s, a, &ForcePolicy::Args::out, r, 3);
    if (rc) return rc;
    rc = bc_rotate(s, r, false);
    if (rc) return rc;
    return bc_project(s, r);
}

// MpmSimulationBase::addScaledForces (MpmSimulationBase.cpp:829-833 -> MpmForceBase::rasterizeForceToTVStack): f -= scale (vol P Fn^T) grad w
int add_scaled_forces(Sim* s, double scale, double* f)
{
    if (!s->state_valid) return fail(s, "addScaledForces: call hot_update_state first");
    KTime t(s, KC_FORCE);
    ForcePolicy::Args a{s->P.stride, s->P.X.p, s->f_stress.p, s->dx, 1.0 / s->dx, scale, s->g_idx.p, f};
    return scatter_to_dofs<ForcePolicy>(s, a, &ForcePolicy::Args::out, f, 3);
}

int ensure_hessian(Sim* s)
{
    if (!s->state_valid) return fail(s, "Hessian: call hot_update_state first");
    if (s->hessian_valid) return 0;
    const size_t ps = s->P.stride;
    HOT_CUDA(s->f_H.reserve(45 * ps));
    KTime t(s, KC_STRESS);
    const long np = s->p1 - s->p0, o = s->p0; // own particles: every row pointer shifted by p0
    if (np > 0) {
        k_build_hessian<<<(unsigned)((np + 127) / 128), 128, 0, s->stream>>>(np, ps, s->P.Fn.p + o, s->P.vol.p + o, s->P.mu.p + o, s->P.lam.p + o,
            s->f_U.p + o, s->f_V.p + o, s->f_sig.p + o, model_flags(s), s->f_H.p + o);
        HOT_LAUNCHED(s);
    }
    s->hessian_valid = true;
    return 0;
}

// out += scale * sum_p [H~_p : grad x_p] grad w: per-particle gather kernel, then the force scatter on the 9 doubles it leaves
// behind (two lean kernels beat one fused gather->scatter CTA: the fused form serialises a 45-load stage in front of the
// register-heavy accumulate loop and halves the occupancy of both)
static int hessian_scatter(Sim* s, double scale, const double* x, double* out)
{
    const size_t ps = s->P.stride;
    HOT_CUDA(s->f_T.reserve(9 * ps));
    if (s->g1 > s->g0) {
        static const bool tma = !(getenv("HOT_HG_TMA") && atoi(getenv("HOT_HG_TMA")) == 0); // A/B switch, default on
        if (tma) {
            HOT_FUNC_ATTR_ONCE(s, k_hessian_gather<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HG_SMEM);
            k_hessian_gather<true><<<(unsigned)(s->g1 - s->g0), US_THREADS, HG_SMEM, s->stream>>>(s->group_first.p + s->g0,
                s->tile_dof.p + (size_t)s->g0 * TILE, ps, s->P.X.p, s->f_H.p, s->dx, 1.0 / s->dx, scale, x, s->f_T.p, 0);
        }
        else
            k_hessian_gather<false><<<(unsigned)(s->g1 - s->g0), US_THREADS, 0, s->stream>>>(s->group_first.p + s->g0,
                s->tile_dof.p + (size_t)s->g0 * TILE, ps, s->P.X.p, s->f_H.p, s->dx, 1.0 / s->dx, scale, x, s->f_T.p, pf_distance(s, 3));
        HOT_LAUNCHED(s);
    }
    ForcePolicy::Args a{ps, s->P.X.p, s->f_T.p, s->dx, 1.0 / s->dx, -1.0, s->g_idx.p, out};
    return scatter_to_dofs<ForcePolicy>(s, a, &ForcePolicy::Args::out, out, 3);
}

// ImplicitSolverObjective::multiply with matrix_free, ImplicitSolver.h:741-763
int hessian_apply_mf(Sim* s, const double* x, double* b)
{
    int rc = ensure_hessian(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    KTime t(s, KC_HESSIAN);
    k_mass_mul<<<nblk(3 * (long)nn), TPB, 0, st>>>(nn, s->mass_matrix.p, x, b);
    HOT_LAUNCHED(s);
    return hessian_scatter(s, s->dt * s->dt, x, b);
}

// MpmSimulationBase::addScaledForceDifferentials (-> MpmForceBase::addScaledForceDifferential, MpmForceBase.cpp:261-306):
// f += scale * df(x), df = -K x  (the objective calls it with scale = -dt^2)
int add_scaled_force_differentials(Sim* s, double scale, const double* x, double* f)
{
    int rc = ensure_hessian(s);
    if (rc) return rc;
    KTime t(s, KC_HESSIAN);
    return hessian_scatter(s, -scale, x, f);
}

// ImplicitSolverObjective::evaluatePerNodeCNTolerance, ImplicitSolver.h:667-696
int eval_cn_tolerance(Sim* s, double eps, double dt, double* tol)
{
    if (!s->p2g_done) return fail(s, "evaluatePerNodeCNTolerance: call hot_p2g first");
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    KTime t(s, KC_FORCE);
    HOT_CUDA(cudaMemsetAsync(tol, 0, (size_t)nn * sizeof(double), st));
    CNTolPolicy::Args a{s->P.stride, s->P.X.p, s->P.M.p, s->P.mu.p, s->P.lam.p, s->dx, 1.0 / s->dx, model_flags(s), s->g_idx.p, tol};
    int rc = scatter_to_dofs<CNTolPolicy>(s, a, &CNTolPolicy::Args::out, tol, 1);
    if (rc) return rc;
    k_cn_finish<<<nblk(nn), TPB, 0, st>>>(nn, s->mass_matrix.p, eps * 24 * s->dx * s->dx * dt, tol);
    HOT_LAUNCHED(s);
    return 0;
}

} // namespace hot
