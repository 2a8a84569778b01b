// a6 (APIC P2G) and a23 (G2P + evolveStrain) as one-CTA-per-page-group kernels.
//
// Reference: MpmSimulationBase::particlesToGridHelper<true,false> (Lib/MPM/MpmSimulationBase.cpp:611-656),
// MpmGrid::iterateKernel (Lib/MPM/MpmGrid.h:245-296), computeBSplineWeights degree 2
// (Lib/Ziran/Math/Splines/BSplines.h:55-81), gridToParticlesHelper<true,false,false> (:930-1006),
// constructNewVelocityFromNewtonResult (:891-901), FBasedMpmForceHelper::evolveStrain
// (Lib/MPM/Force/FBasedMpmForceHelper.cpp:100-114).
//
// P2G design (why it is not the reference's 8-colour scatter): fp64 P2G sits near the B200 ridge (128 B and ~300 DFMA-class
// ops per particle; ~17 T DFMA/s vs 6.5 TB/s), so the kernel minimises fp64 issue slots and shared-memory wavefronts and has
// no atomics in the particle loop - see scatter.cuh for the prep / accumulate / gather-combine structure it shares with the
// force and Hessian scatters.  G2P is the matching gather: CTA per page group, the group's (v + dv) node tile staged in
// shared memory through the per-step tile -> DOF table, one thread per particle, tensor-product contraction, F update fused.
#include "scatter_ws.cuh"
#include "dense3.cuh"
#include <cstdlib>

namespace hot {
namespace {

constexpr int E = Geo::E;
constexpr int TILE = Geo::TILE;

// a6: g.m += w m_p ; g.v += w (m_p C_p (x_i - x_p) + m_p v_p)   (MpmSimulationBase.cpp:636-652)
struct P2GPolicy {
    static constexpr int NCH = 4;
    struct Args {
        size_t ps;
        const double *X, *V, *M, *C;
        double dx, one_over_dx;
        size_t gs;
        double *g_m, *g_v;
    };
    // Prepared record (22 doubles, 16-byte words):
    //   [0..2] wx  [3..5] wy | [6..8] wz [9] m | [10..12] A = m v + m C (x_node0 - x_p) | [13..15] Gx [16..18] Gy [19..21] Gz,
    //   G* = dx * m C(:, d): the node value of stencil node (i, j, k) is A + i Gx + j Gy + k Gz.
    static constexpr int REC = 22, MINB = 5; // column form: 33 KB of records per CTA, plane form 66 KB
    static constexpr bool PLANE = true; // measured (C2): plane with compact records 0.096 ms (4 CTAs per SM), column 0.109 ms
    static constexpr bool DOF = false;
    __device__ __forceinline__ static void prefetch(const Args& a, int first, int end, int tid, int nt)
    {
        prefetch_rows_l2(a.X, a.ps, 3, first, end, tid, nt);
        prefetch_rows_l2(a.V, a.ps, 3, first, end, tid, nt);
        prefetch_rows_l2(a.M, a.ps, 1, first, end, tid, nt);
        prefetch_rows_l2(a.C, a.ps, 9, first, end, tid, nt);
    }
    __device__ __forceinline__ static void prep(const Args& a, size_t s, double* __restrict__ r)
    {
        const double m = a.M[s];
        double Xp[3], v[3], Cm[9], d0n[3], w[3][3], g[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            Xp[d] = a.X[d * a.ps + s];
            v[d] = a.V[d * a.ps + s];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) Cm[q] = m * a.C[q * a.ps + s];
        prep_weights<false>(Xp, a.dx, a.one_over_dx, w, g, d0n);
        double A[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) A[c] = m * v[c] + (Cm[c] * d0n[0] + Cm[c + 3] * d0n[1] + Cm[c + 6] * d0n[2]);
        sts2(r + 0, w[0][0], w[0][1]); sts2(r + 2, w[0][2], w[1][0]); sts2(r + 4, w[1][1], w[1][2]);
        sts2(r + 6, w[2][0], w[2][1]); sts2(r + 8, w[2][2], m);
        sts2(r + 10, A[0], A[1]); sts2(r + 12, A[2], a.dx * Cm[0]); sts2(r + 14, a.dx * Cm[1], a.dx * Cm[2]);
        sts2(r + 16, a.dx * Cm[3], a.dx * Cm[4]); sts2(r + 18, a.dx * Cm[5], a.dx * Cm[6]); sts2(r + 20, a.dx * Cm[7], a.dx * Cm[8]);
    }
    __device__ __forceinline__ static void accumulate_col(const double* __restrict__ rec, int i, int j, double di, double dj, double (&acc)[3][4])
    {
        const double wij = rec[i] * rec[3 + j];
        double wz[3], m, A[3], gx[3], gy[3], gz[3];
        lds2(rec + 6, wz[0], wz[1]); lds2(rec + 8, wz[2], m);
        lds2(rec + 10, A[0], A[1]); lds2(rec + 12, A[2], gx[0]); lds2(rec + 14, gx[1], gx[2]);
        lds2(rec + 16, gy[0], gy[1]); lds2(rec + 18, gy[2], gz[0]); lds2(rec + 20, gz[1], gz[2]);
        double b[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) b[c] = fma(dj, gy[c], fma(di, gx[c], A[c]));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double w = wij * wz[k];
            acc[k][0] = fma(w, m, acc[k][0]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double val = k == 0 ? b[c] : (k == 1 ? b[c] + gz[c] : fma(2.0, gz[c], b[c]));
                acc[k][1 + c] = fma(w, val, acc[k][1 + c]);
            }
        }
    }
    // ---- plane form.  COMPACT record (18 doubles: 55 KB for a full page of 384 particles, 4 CTAs per SM instead of 3):
    //   [0..2] xi - base per axis (the argument of the B-spline weights) [3] m | [4..6] A | [7..15] Gx Gy Gz | pad
    // the (cell, plane) thread re-derives the 9 weights in the reference's operation order (bspline_axis).
    static constexpr int RECP = 18, PMINB = 4;
    __device__ __forceinline__ static void prep_plane(const Args& a, size_t s, double* __restrict__ r)
    {
        const double m = a.M[s];
        double d0[3], d0n[3], v[3], Cm[9];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double Xd = a.X[d * a.ps + s];
            v[d] = a.V[d * a.ps + s];
            double xi;
            const int b = base_node_of(Xd, a.one_over_dx, &xi);
            d0[d] = xi - (double)b;
            d0n[d] = (double)b * a.dx - Xd;
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) Cm[q] = m * a.C[q * a.ps + s];
        double A[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) A[c] = m * v[c] + (Cm[c] * d0n[0] + Cm[c + 3] * d0n[1] + Cm[c + 6] * d0n[2]);
        sts2(r + 0, d0[0], d0[1]); sts2(r + 2, d0[2], m);
        sts2(r + 4, A[0], A[1]); sts2(r + 6, A[2], a.dx * Cm[0]); sts2(r + 8, a.dx * Cm[1], a.dx * Cm[2]);
        sts2(r + 10, a.dx * Cm[3], a.dx * Cm[4]); sts2(r + 12, a.dx * Cm[5], a.dx * Cm[6]); sts2(r + 14, a.dx * Cm[7], a.dx * Cm[8]);
    }
    // the 9 nodes (j, k) of x-plane i
    __device__ __forceinline__ static void accumulate_plane(const Args& a, const double* __restrict__ rec, int i, double di, double (&acc)[9][4])
    {
        double v[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) lds2(rec + 2 * u, v[2 * u], v[2 * u + 1]);
        accumulate_rec(a, v, i, di, acc);
    }
    // ---- ws form (scatter_ws.cuh): 16 raw rows X V M C staged by TMA, the compact record above (16 doubles = 8 units) in place
    static constexpr int ROWS = 16, UNITS = 8, WS_ID = 0;
    __device__ __forceinline__ static const double* row(const Args& a, int r)
    {
        return r < 3 ? a.X + (size_t)r * a.ps : (r < 6 ? a.V + (size_t)(r - 3) * a.ps : (r == 6 ? a.M : a.C + (size_t)(r - 7) * a.ps));
    }
    __device__ __forceinline__ static int prep_ws(const Args& a, const double (&raw)[16], double (&r)[16])
    {
        const double m = raw[6];
        double d0n[3], Cm[9];
        int cb[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double xi;
            cb[d] = base_node_of(raw[d], a.one_over_dx, &xi);
            r[d] = xi - (double)cb[d];
            d0n[d] = (double)cb[d] * a.dx - raw[d];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) Cm[q] = m * raw[7 + q];
        r[3] = m;
#pragma unroll
        for (int c = 0; c < 3; ++c) r[4 + c] = m * raw[3 + c] + (Cm[c] * d0n[0] + Cm[c + 3] * d0n[1] + Cm[c + 6] * d0n[2]);
#pragma unroll
        for (int q = 0; q < 9; ++q) r[7 + q] = a.dx * Cm[q];
        return (((cb[0] & (Geo::BX - 1)) << Geo::yb | (cb[1] & (Geo::BY - 1))) << Geo::zb) | (cb[2] & (Geo::BZ - 1));
    }
    __device__ __forceinline__ static void accumulate_rec(const Args&, const double (&v)[16], int i, double di, double (&acc)[9][4])
    {
        const double d0[3] = {v[0], v[1], v[2]}, m = v[3], A[3] = {v[4], v[5], v[6]}, gx[3] = {v[7], v[8], v[9]}, gy[3] = {v[10], v[11], v[12]},
                     gz[3] = {v[13], v[14], v[15]};
        double w[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) bspline_axis(d0[d], w[d], nullptr);
        const double wx = i == 0 ? w[0][0] : (i == 1 ? w[0][1] : w[0][2]);
        double b[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) b[c] = fma(di, gx[c], A[c]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double wij = wx * w[1][j];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double ww = wij * w[2][k];
                acc[j * 3 + k][0] = fma(ww, m, acc[j * 3 + k][0]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double val = k == 0 ? b[c] : (k == 1 ? b[c] + gz[c] : fma(2.0, gz[c], b[c]));
                    acc[j * 3 + k][1 + c] = fma(ww, val, acc[j * 3 + k][1 + c]);
                }
            }
            if (j < 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) b[c] += gy[c];
            }
        }
    }
    __device__ __forceinline__ static void flush1(const Args& a, long n, int ch, double v)
    {
        atomicAdd(ch == 0 ? a.g_m + n : a.g_v + (size_t)(ch - 1) * a.gs + n, v);
    }
};

__global__ void k_step_zero(size_t gn, double* __restrict__ g_m, double* __restrict__ g_v, int* __restrict__ done, int* __restrict__ flags)
{
    const size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < gn) {
        g_m[a] = 0.0;
        g_v[a] = 0.0; g_v[gn + a] = 0.0; g_v[2 * gn + a] = 0.0;
    }
    if (a == 0) {
        *done = 0;
        flags[0] = 0; flags[1] = 0;
    }
}

constexpr int G2P_THREADS = 128;
static_assert(TILE <= 2 * G2P_THREADS && TILE >= G2P_THREADS, "tile staged in two rounds");

// TMA = true: the group's X and F rows (12 runs of <= G2P_CAP particles) are staged in shared memory by bulk copies that one
// thread issues before the tile gather, so the particle loop has no global loads and no scoreboard waits of its own.
constexpr int G2P_CAP = 256;
template <bool TMA>
__global__ void __launch_bounds__(G2P_THREADS, 4) k_g2p(const int* __restrict__ group_first, const int* __restrict__ tile_dof, size_t ps,
    double* __restrict__ X, double* __restrict__ V, double* __restrict__ C, double* __restrict__ F, double* __restrict__ gradV, double dx,
    double one_over_dx, double dt, double apic_rpic_ratio, double cfl, const double* __restrict__ vn, const double* __restrict__ dv,
    int* __restrict__ flags, int pf_dist)
{
    __shared__ double tile[3][TILE];
    __shared__ __align__(16) double s_rows[TMA ? 12 : 1][TMA ? G2P_CAP + 2 : 2];
    __shared__ __align__(8) unsigned long long s_bar;
    const int g = blockIdx.x, tid = threadIdx.x;
    if (TMA && tid == 0) mbar_init(&s_bar, 1);
    // dependent DRAM round trips of a CTA: {group_first, tile_dof} -> {vn + dv gather, X} -> F (L2-prefetched); the tile's DOF ids
    // come from the per-step table instead of the chain group_slot -> nbr8 -> g_idx
    const int id0 = tile_dof[(size_t)g * TILE + tid];
    const int id1 = tid + G2P_THREADS < TILE ? tile_dof[(size_t)g * TILE + tid + G2P_THREADS] : -1;
    const int first = group_first[g], end = group_first[g + 1];
    // the group that takes over this CTA's slot: its rows are requested from L2 at the end (bounds requested now)
    int pf_first = 0, pf_end = 0;
    if (pf_dist > 0 && g + pf_dist < (int)gridDim.x) {
        pf_first = group_first[g + pf_dist];
        pf_end = group_first[g + pf_dist + 1];
    }
    double X0[3] = {0.0, 0.0, 0.0};
    if (!TMA && first + tid < end) {
#pragma unroll
        for (int d = 0; d < 3; ++d) X0[d] = X[d * ps + first + tid];
#pragma unroll
        for (int q = 0; q < 9; ++q) prefetch_l2(F + q * ps + first + tid);
    }
    // chunk [c0, c0 + cn) of the group: the bulk copies start at the even index below c0 and move an even count (16-byte rule)
    auto issue_chunk = [&](int c0) {
        const int cn = min(G2P_CAP, end - c0), start = c0 & ~1, cnt = (c0 + cn - start + 1) & ~1;
        fence_proxy_async();
        mbar_expect_tx(&s_bar, 12u * cnt * 8u);
#pragma unroll
        for (int d = 0; d < 3; ++d) bulk_load(s_rows[d], X + d * ps + start, cnt * 8u, &s_bar);
#pragma unroll
        for (int q = 0; q < 9; ++q) bulk_load(s_rows[3 + q], F + q * ps + start, cnt * 8u, &s_bar);
    };
    if (TMA) {
        __syncthreads(); // barrier initialised
        if (tid == 0 && first < end) issue_chunk(first);
    }
    // new_v = v + dv on the touched tile (constructNewVelocityFromNewtonResult fused into the staging; vn = the normalised v)
    {
        double nv[3] = {0, 0, 0};
        if (id0 >= 0)
#pragma unroll
            for (int d = 0; d < 3; ++d) nv[d] = vn[3 * (size_t)id0 + d] + dv[3 * (size_t)id0 + d];
        tile[0][tid] = nv[0]; tile[1][tid] = nv[1]; tile[2][tid] = nv[2];
        if (tid + G2P_THREADS < TILE) {
            double nw[3] = {0, 0, 0};
            if (id1 >= 0)
#pragma unroll
                for (int d = 0; d < 3; ++d) nw[d] = vn[3 * (size_t)id1 + d] + dv[3 * (size_t)id1 + d];
            tile[0][tid + G2P_THREADS] = nw[0]; tile[1][tid + G2P_THREADS] = nw[1]; tile[2][tid + G2P_THREADS] = nw[2];
        }
    }
    __syncthreads();
    const double D_inverse = 4.0 / (dx * dx); // MpmSimulationBase.cpp:114-118
    const double ca = (apic_rpic_ratio + 1.0) * 0.5, cb = (apic_rpic_ratio - 1.0) * 0.5;
    int fast = 0, half_fast = 0;
    unsigned phase = 0;
    for (int c0 = first; c0 < end; c0 += TMA ? G2P_CAP : end - first) {
    const int c1 = TMA ? min(c0 + G2P_CAP, end) : end, start = c0 & ~1;
    if (TMA) {
        if (c0 != first) {
            __syncthreads(); // every thread is done with the previous chunk's rows
            if (tid == 0) issue_chunk(c0);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
    }
    for (int s = c0 + tid; s < c1; s += G2P_THREADS) {
        double Xp[3], w[3][3], gw[3][3], xm[3][3]; // weights, weight derivatives / dx, x_node - x_p per axis and stencil index
        int tb[3];
        if (TMA) {
#pragma unroll
            for (int d = 0; d < 3; ++d) X0[d] = s_rows[d][s - start];
        }
        // software pipeline: the next particle's position load and F prefetch are in flight during this particle's contraction
        double Xn[3] = {0.0, 0.0, 0.0};
        if (!TMA && s + G2P_THREADS < end) {
#pragma unroll
            for (int d = 0; d < 3; ++d) Xn[d] = X[d * ps + s + G2P_THREADS];
#pragma unroll
            for (int q = 0; q < 9; ++q) prefetch_l2(F + q * ps + s + G2P_THREADS);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            Xp[d] = X0[d];
            double xi, dw[3];
            int b = base_node_of(Xp[d], one_over_dx, &xi);
            bspline_axis(xi - (double)b, w[d], dw);
            const double d0n = (double)b * dx - Xp[d];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                gw[d][i] = one_over_dx * dw[i];
                xm[d][i] = (double)i * dx + d0n;
            }
            int bits = d == 0 ? Geo::xb : (d == 1 ? Geo::yb : Geo::zb);
            tb[d] = b & ((1 << bits) - 1);
        }
        // The 27-node sums  v_p = sum w v,  B = sum w v (x_i - x_p)^T,  grad v = sum v grad w^T  (MpmSimulationBase.cpp:942-1006)
        // are evaluated as a tensor-product contraction z -> y -> x: 441 FMA instead of 27 x 32 (fp64 issue is the bound of this
        // kernel), same terms in a different association.
        double vp[3] = {0, 0, 0}, B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const double wz0 = w[2][0] * xm[2][0], wz1 = w[2][1] * xm[2][1], wz2 = w[2][2] * xm[2][2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double Av[3] = {0, 0, 0}, Agy[3] = {0, 0, 0}, Agz[3] = {0, 0, 0}, Aby[3] = {0, 0, 0}, Abz[3] = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n0 = ((tb[0] + i) * Geo::TY + (tb[1] + j)) * Geo::TZ + tb[2];
                const double wj = w[1][j], gj = gw[1][j], wyj = wj * xm[1][j];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double v0 = tile[r][n0], v1 = tile[r][n0 + 1], v2 = tile[r][n0 + 2];
                    const double a = w[2][0] * v0 + w[2][1] * v1 + w[2][2] * v2;                       // sum_k w_k v
                    const double g = gw[2][0] * v0 + gw[2][1] * v1 + gw[2][2] * v2;                    // sum_k dw_k/dx v
                    const double c = wz0 * v0 + wz1 * v1 + wz2 * v2;                                   // sum_k w_k z_k v
                    Av[r] += wj * a;
                    Agy[r] += gj * a;
                    Agz[r] += wj * g;
                    Aby[r] += wyj * a;
                    Abz[r] += wj * c;
                }
            }
            const double wi = w[0][i], gi = gw[0][i], wxi = wi * xm[0][i];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                vp[r] += wi * Av[r];
                G[r] += gi * Av[r];
                G[r + 3] += wi * Agy[r];
                G[r + 6] += wi * Agz[r];
                B[r] += wxi * Av[r];
                B[r + 3] += wi * Aby[r];
                B[r + 6] += wi * Abz[r];
            }
        }
        double inc = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            V[d * ps + s] = vp[d];
            double step = dt * vp[d];
            X[d * ps + s] = Xp[d] + step;
            inc += step * step;
        }
        if (inc > dx * dx) fast = 1;
        if (inc > dx * dx * 0.25 * (cfl * cfl)) half_fast = 1;
        double Fo[9], A[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            B[q] *= D_inverse;
            Fo[q] = TMA ? s_rows[3 + q][s - start] : F[q * ps + s];
            A[q] = dt * G[q];
            gradV[q * ps + s] = G[q];
        }
        A[0] += 1.0; A[4] += 1.0; A[8] += 1.0;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                C[(r + 3 * cc) * ps + s] = ca * B[r + 3 * cc] + cb * B[cc + 3 * r];
                F[(r + 3 * cc) * ps + s] = A[r] * Fo[3 * cc] + A[r + 3] * Fo[3 * cc + 1] + A[r + 6] * Fo[3 * cc + 2];
            }
#pragma unroll
        for (int d = 0; d < 3; ++d) X0[d] = Xn[d];
    }
    }
    prefetch_rows_l2(X, ps, 3, pf_first, pf_end, tid, G2P_THREADS);
    prefetch_rows_l2(F, ps, 9, pf_first, pf_end, tid, G2P_THREADS);
    if (pf_end > pf_first && tid < 5) prefetch_l2(tile_dof + (size_t)(g + pf_dist) * TILE + tid * 32);
    if (fast) atomicOr(flags, 1);
    if (half_fast) atomicOr(flags + 1, 1);
}

// PlasticityApplier<...>::applyPlasticity -> projectStrain on every particle's F (Lib/Ziran/Physics/PlasticityApplier.h:40-55):
// model 1 VonMisesFixedCorotated::projectStrain (PlasticityApplier.cpp:94-131), model 2 SnowPlasticity::projectStrain (:16-50)
__global__ void k_plasticity(long n, size_t ps, int model, double p0, double p1, double p2, double p3, double p4, double* __restrict__ F,
    double* __restrict__ mu, double* __restrict__ lam, double* __restrict__ Jp)
{
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double Fm[9], U[9], V[9], sig[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) Fm[q] = F[q * ps + s];
    svd3(Fm, U, sig, V);
    double sn[3];
    if (model == 1) {
        const double m_ = mu[s], l_ = lam[s], yield_stress = p0;
#pragma unroll
        for (int d = 0; d < 3; ++d) sig[d] = fmax(1e-4, sig[d]);
        const double J = sig[0] * sig[1] * sig[2];
        double tau[3], tr = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            tau[d] = 2.0 * m_ * (sig[d] - 1.0) * sig[d] + l_ * (J - 1.0) * J;
            tr += tau[d];
        }
        double sdev[3], n2 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            sdev[d] = tau[d] - tr / 3.0;
            n2 += sdev[d] * sdev[d];
        }
        const double s_norm = sqrt(n2), scaled_tauy = sqrt(2.0 / 3.0) * yield_stress; // sqrt(2 / (6 - dim))
        if (s_norm - scaled_tauy <= 0.0) return;
        const double alpha = scaled_tauy / s_norm;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double tau_new = alpha * sdev[d] + tr / 3.0;
            const double b2m4ac = m_ * m_ - 2.0 * m_ * (l_ * (J - 1.0) * J - tau_new);
            sn[d] = (m_ + sqrt(b2m4ac)) / (2.0 * m_);
        }
    }
    else if (model == 3) {
        // Drucker-Prager (extension, not in the reference; BASELINE C4): return mapping of Klar et al. 2016 in Hencky strain,
        // eps = log sigma, yield ||dev eps|| + (3 lambda + 2 mu) / (2 mu) tr(eps) alpha <= 0 with alpha = sqrt(2/3) 2 sin(phi) / (3 - sin(phi));
        // expansion (tr eps > 0) projects to the tip (sigma = 1, optionally shifted by the cohesion p1)
        const double m_ = mu[s], l_ = lam[s], sphi = sin(p0 * 0.017453292519943295), alpha = sqrt(2.0 / 3.0) * 2.0 * sphi / (3.0 - sphi), coh = p1;
        double eps[3], tr = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            eps[d] = log(fmax(sig[d], 1e-6)) - coh;
            tr += eps[d];
        }
        double dev[3], n2 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            dev[d] = eps[d] - tr / 3.0;
            n2 += dev[d] * dev[d];
        }
        const double nrm = sqrt(n2);
        if (tr >= 0.0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) sn[d] = exp(coh);
        }
        else if (nrm == 0.0) return; // hydrostatic compression: inside the cone
        else {
            const double dgamma = nrm + (3.0 * l_ + 2.0 * m_) / (2.0 * m_) * tr * alpha;
            if (dgamma <= 0.0) return; // elastic
#pragma unroll
            for (int d = 0; d < 3; ++d) sn[d] = exp(eps[d] - dgamma * dev[d] / nrm + coh);
        }
    }
    else {
        const double psi = p0, theta_c = p1, theta_s = p2, min_Jp = p3, max_Jp = p4;
        double Fe_det = 1.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            sn[d] = fmax(fmin(sig[d], 1.0 + theta_s), 1.0 - theta_c);
            Fe_det *= sn[d];
        }
        const double Jold = Jp[s];
        double Jnew = Jold * det3(Fm) / Fe_det;
        if (!(Jnew <= max_Jp)) Jnew = max_Jp;
        if (!(Jnew >= min_Jp)) Jnew = min_Jp;
        const double h = exp(psi * (Jold - Jnew));
        mu[s] *= h;
        lam[s] *= h;
        Jp[s] = Jnew;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) F[(r + 3 * c) * ps + s] = U[r] * sn[0] * V[c] + U[r + 3] * sn[1] * V[c + 3] + U[r + 6] * sn[2] * V[c + 6];
}

} // namespace

int apply_plasticity(Sim* s)
{
    if (s->plastic_model == 0) return 0;
    const long np = s->p1 > s->p0 ? s->p1 - s->p0 : 0, o = s->p0;
    if (np <= 0) return 0;
    KTime t(s, KC_STRESS);
    const double* q = s->plastic_param;
    k_plasticity<<<(unsigned)((np + 127) / 128), 128, 0, s->stream>>>(np, s->P.stride, s->plastic_model, q[0], q[1], q[2], q[3], q[4], s->P.F.p + o,
        s->P.mu.p + o, s->P.lam.p + o, s->P.Jp.p + o);
    HOT_LAUNCHED(s);
    return 0;
}

int p2g(Sim* s)
{
    if (!s->sorted) return fail(s, "hot_p2g: call hot_sort_and_activate first (MultigridSimulation.h:235-297 step order)");
    cudaStream_t st = s->stream;
    const size_t gn = s->g_stride;
    // the pages are re-zeroed so the call is repeatable (the reference zeroes them in the sort, :1128-1136)
    // one launch zeroes the four grid channels, the numbering's done-counter and the G2P flags (was four memsets)
    HOT_CUDA(s->flags.reserve(2));
    HOT_CUDA(s->dcount.reserve(16));
    k_step_zero<<<(unsigned)((gn + 255) / 256), 256, 0, st>>>(gn, s->g_m.p, s->g_v.p, s->dcount.p + 12, s->flags.p);
    HOT_LAUNCHED(s);
    s->flags_zeroed = true;
    {
        KTime t(s, KC_P2G);
        P2GPolicy::Args a{s->P.stride, s->P.X.p, s->P.V.p, s->P.M.p, s->P.C.p, s->dx, 1.0 / s->dx, gn, s->g_m.p, s->g_v.p};
        if (s->g1 > s->g0) {
            int rc = launch_scatter_best<P2GPolicy>(s, a);
            if (rc) return rc;
        }
    }
    int rc = dist_p2g_exchange(s); // shared pages: partial sums exchanged between the sharing ranks, added in rank order
    if (rc) return rc;
    {
        KTime t(s, KC_NUMBER);
        rc = number_nodes(s, false);
    }
    if (rc) return rc;
    rc = dist_after_numbering(s);
    if (rc) return rc;
    s->p2g_done = true;
    return 0;
}

int g2p(Sim* s, double dt, int* flags)
{
    if (!s->p2g_done) return fail(s, "hot_g2p: call hot_p2g first");
    cudaStream_t st = s->stream;
    HOT_CUDA(s->flags.reserve(2));
    HOT_CUDA(s->P.gradV.reserve(9 * s->P.stride));
    if (!s->flags_zeroed) HOT_CUDA(cudaMemsetAsync(s->flags.p, 0, 2 * sizeof(int), st)); // (hot_p2g's zero pass covers the first G2P after it)
    s->flags_zeroed = false;
    {
        KTime t(s, KC_G2P);
        static const bool tma = !(getenv("HOT_G2P_TMA") && atoi(getenv("HOT_G2P_TMA")) == 0); // A/B switch, default on
        if (s->g1 > s->g0)
        (tma ? k_g2p<true> : k_g2p<false>)<<<(unsigned)(s->g1 - s->g0), G2P_THREADS, 0, st>>>(s->group_first.p + s->g0, s->tile_dof.p + (size_t)s->g0 * TILE, s->P.stride, s->P.X.p,
            s->P.V.p, s->P.C.p, s->P.F.p, s->P.gradV.p, s->dx, 1.0 / s->dx, dt, s->apic_rpic_ratio, s->cfl, s->vn.p, s->dv.p, s->flags.p,
            pf_distance(s, 4));
        HOT_LAUNCHED(s);
    }
    if (dt != 0.0) { // evolveStrain is followed by applyPlasticity (MpmSimulationBase.cpp:1039-1041)
        int rc = apply_plasticity(s);
        if (rc) return rc;
    }
    if (flags) {
        HOT_CUDA(cudaMemcpyAsync(s->hcount + 6, s->flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        HOT_CUDA(cudaStreamSynchronize(st));
        flags[0] = s->hcount[6];
        flags[1] = s->hcount[7];
    }
    if (dt != 0.0) { // positions moved: the reference re-sorts at the start of the next step
        s->sorted = false;
        s->p2g_done = false;
    }
    return 0;
}

} // namespace hot
