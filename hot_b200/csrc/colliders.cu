// a8 on the device: analytic collision objects evaluated per grid node, the CollisionNode table and the Newton
// initial guess built without a host round trip.
//
// Reference: MpmSimulationBase::buildInitialDvAndVnForNewton (Lib/MPM/MpmSimulationBase.cpp:1139-1184),
// AnalyticCollisionObject::{detectAndResolveCollision, multiObjectCollision} (Lib/Ziran/Math/Geometry/CollisionObject.cpp:108-149,
// 384-452), the analytic level sets HalfSpace / Sphere / AnalyticBox (+ AxisAlignedAnalyticBox) / CappedCylinder
// (Lib/Ziran/Math/Geometry/AnalyticLevelSet.{h:122-310, cpp:259-304,353-368,435-452,504-539}), RotationExtractor<T,3>::rotate
// (Lib/MPM/MpmSimulationBase.h:270-281).
//
// One thread per DOF node evaluates ALL objects in order (the reference's loop: first STICKY hit wins, SLIP / SEPARATE normals are
// Gram-Schmidt'ed); colliding nodes are flagged and compacted in node order (the reference's concurrent_vector order is
// nondeterministic; nothing depends on it).  What crosses the boundary per step: the object table (a few hundred bytes).
#include "../../include/hot_b200.h"
#include "sim.h"
#include <cub/cub.cuh>

namespace hot {
namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

__device__ __forceinline__ void mat_vec(const double* M, const double* x, double* y) // column-major 3x3
{
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = M[r] * x[0] + M[r + 3] * x[1] + M[r + 6] * x[2];
}
__device__ __forceinline__ void mat_t_vec(const double* M, const double* x, double* y)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = M[3 * r] * x[0] + M[3 * r + 1] * x[1] + M[3 * r + 2] * x[2];
}

// queryInside(X, phi, N) of the level set in ITS material space; N only where the node is inside (phi <= 0)
__device__ bool level_set_inside(const hot_collider& o, const double* X, double* N)
{
    if (o.shape == HOT_SHAPE_HALFSPACE) { // AnalyticLevelSet.cpp:272-287
        const double phi = o.p[3] * (X[0] - o.p[0]) + o.p[4] * (X[1] - o.p[1]) + o.p[5] * (X[2] - o.p[2]);
        N[0] = o.p[3]; N[1] = o.p[4]; N[2] = o.p[5];
        return phi <= 0.0;
    }
    if (o.shape == HOT_SHAPE_SPHERE) { // Sphere::queryInside, :435-452 (strict <, normal e_x at the centre)
        const double d[3] = {X[0] - o.p[0], X[1] - o.p[1], X[2] - o.p[2]};
        const double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (!(d2 < o.p[3] * o.p[3])) return false;
        const double l = sqrt(d2);
        if (l < 1e-7) { N[0] = 1.0; N[1] = 0.0; N[2] = 0.0; }
        else { N[0] = d[0] / l; N[1] = d[1] / l; N[2] = d[2] / l; }
        return true;
    }
    // AnalyticBox / CappedCylinder: own rigid transform, X_primitive = R^-1 (X - b); normal = R * d(phi)/dX_primitive
    const double xm[3] = {X[0] - o.shape_b[0], X[1] - o.shape_b[1], X[2] - o.shape_b[2]};
    double Xp[3], Np[3] = {0.0, 0.0, 0.0};
    mat_t_vec(o.shape_R, xm, Xp);
    double phi;
    if (o.shape == HOT_SHAPE_BOX) { // :504-539; inside: phi = max_i (|X_i| - h_i), gradient on the arg-max axis
        const double d[3] = {fabs(Xp[0]) - o.p[0], fabs(Xp[1]) - o.p[1], fabs(Xp[2]) - o.p[2]};
        int a = 0;
        if (d[1] > d[a]) a = 1;
        if (d[2] > d[a]) a = 2;
        const double q[3] = {fmax(d[0], 0.0), fmax(d[1], 0.0), fmax(d[2], 0.0)};
        phi = fmin(d[a], 0.0) + sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        if (!(phi <= 0.0)) return false;
        Np[a] = Xp[a] < 0.0 ? -1.0 : 1.0;
    }
    else { // CappedCylinder along y, AnalyticLevelSet.h:258-274
        const double rxz = sqrt(Xp[0] * Xp[0] + Xp[2] * Xp[2]);
        const double d0 = rxz - o.p[0], d1 = fabs(Xp[1]) - 0.5 * o.p[1];
        const double q0 = fmax(d0, 0.0), q1 = fmax(d1, 0.0);
        phi = fmin(fmax(d0, d1), 0.0) + sqrt(q0 * q0 + q1 * q1);
        if (!(phi <= 0.0)) return false;
        if (d0 >= d1) {
            if (rxz > 0.0) { Np[0] = Xp[0] / rxz; Np[2] = Xp[2] / rxz; }
            else Np[0] = 1.0;
        }
        else Np[1] = Xp[1] < 0.0 ? -1.0 : 1.0;
    }
    mat_vec(o.shape_R, Np, N);
    return true;
}

// AnalyticCollisionObject::detectAndResolveCollision, CollisionObject.cpp:384-452 (material velocity 0)
__device__ bool detect_and_resolve(const hot_collider& o, const double* x, double* v, double* n)
{
    const double xb[3] = {x[0] - o.b[0], x[1] - o.b[1], x[2] - o.b[2]};
    const double one_over_s = 1.0 / o.s;
    double Xr[3], X[3], N[3];
    mat_t_vec(o.R, xb, Xr);
#pragma unroll
    for (int d = 0; d < 3; ++d) X[d] = Xr[d] * one_over_s;
    if (!level_set_inside(o, X, N)) return false;
    const double k = o.dsdt * one_over_s;
    const double vo[3] = {o.omega[1] * xb[2] - o.omega[2] * xb[1] + k * xb[0] + o.dbdt[0], o.omega[2] * xb[0] - o.omega[0] * xb[2] + k * xb[1] + o.dbdt[1],
        o.omega[0] * xb[1] - o.omega[1] * xb[0] + k * xb[2] + o.dbdt[2]};
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] -= vo[d];
    if (o.type == HOT_COLLIDER_STICKY) v[0] = v[1] = v[2] = 0.0;
    else {
        mat_vec(o.R, N, n);
        const double dn = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
        if (o.type == HOT_COLLIDER_SLIP || dn < 0.0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) v[d] -= n[d] * dn;
            if (o.friction != 0.0 && dn < 0.0) { // kinematic friction
                const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                if (-dn * o.friction < l) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) v[d] += v[d] / l * dn * o.friction;
                }
                else v[0] = v[1] = v[2] = 0.0;
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] += vo[d];
    return true;
}

// RotationExtractor<T,3>::rotate = Quaternion::setFromTwoVectors(a, e_x) as a matrix (column-major)
__device__ void rotate_to_x(const double* a_in, double* R)
{
    const double l = sqrt(a_in[0] * a_in[0] + a_in[1] * a_in[1] + a_in[2] * a_in[2]);
    const double a[3] = {a_in[0] / l, a_in[1] / l, a_in[2] / l};
    const double c = a[0];
    if (c < -1 + 1e-12) {
        const double Rp[9] = {-1, 0, 0, 0, -1, 0, 0, 0, 1};
#pragma unroll
        for (int q = 0; q < 9; ++q) R[q] = Rp[q];
        return;
    }
    const double v[3] = {0.0, a[2], -a[1]}, k = 1.0 / (1.0 + c);
    R[0] = 1 + k * (-v[1] * v[1] - v[2] * v[2]); R[1] = v[2] + k * v[0] * v[1]; R[2] = -v[1] + k * v[0] * v[2];
    R[3] = -v[2] + k * v[0] * v[1]; R[4] = 1 + k * (-v[0] * v[0] - v[2] * v[2]); R[5] = v[0] + k * v[1] * v[2];
    R[6] = v[1] + k * v[0] * v[2]; R[7] = -v[0] + k * v[1] * v[2]; R[8] = 1 + k * (-v[0] * v[0] - v[1] * v[1]);
}

// per node: multiObjectCollision (CollisionObject.cpp:108-149) + the CollisionNode of MpmSimulationBase.cpp:1158-1172, dense by node id
__global__ void k_collide_nodes(int nn, const int* __restrict__ coord, const double* __restrict__ vn, double dx, int n_obj,
    const hot_collider* __restrict__ objs, int* __restrict__ flag, double* __restrict__ P, double* __restrict__ R, int* __restrict__ slip,
    double* __restrict__ dv_bc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const double xi[3] = {coord[3 * i] * dx, coord[3 * i + 1] * dx, coord[3 * i + 2] * dx};
    const double old_v[3] = {vn[3 * i], vn[3 * i + 1], vn[3 * i + 2]};
    double vi[3] = {old_v[0], old_v[1], old_v[2]}, wn[3] = {0.0, 0.0, 0.0}, nb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool any = false;
    int slip_count = 0;
    for (int k = 0; k < n_obj; ++k) {
        const hot_collider& o = objs[k];
        if (o.type == HOT_COLLIDER_GHOST) continue;
        double n[3] = {0.0, 0.0, 0.0};
        if (!detect_and_resolve(o, xi, vi, n)) continue;
        any = true;
        if (o.type == HOT_COLLIDER_STICKY) {
            wn[0] = wn[1] = wn[2] = 0.0;
#pragma unroll
            for (int q = 0; q < 9; ++q) nb[q] = (q % 4 == 0) ? 1.0 : 0.0;
            break;
        }
        for (int c = 0; c < slip_count; ++c) {
            const double d = nb[3 * c] * n[0] + nb[3 * c + 1] * n[1] + nb[3 * c + 2] * n[2];
#pragma unroll
            for (int q = 0; q < 3; ++q) n[q] -= d * nb[3 * c + q];
        }
        wn[0] = n[0]; wn[1] = n[1]; wn[2] = n[2];
        const double l = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (l != 0.0) {
#pragma unroll
            for (int q = 0; q < 3; ++q) nb[3 * slip_count + q] = n[q] / l;
            if (++slip_count == 3) break;
        }
    }
    flag[i] = any;
    if (!any) return;
    const bool is_slip = wn[0] != 0.0 || wn[1] != 0.0 || wn[2] != 0.0;
    slip[i] = is_slip;
    double Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (is_slip) rotate_to_x(wn, Rm);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double kk = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) kk += nb[r + 3 * q] * nb[c + 3 * q];
            P[9 * (size_t)i + r + 3 * c] = (r == c ? 1.0 : 0.0) - kk; // I - K K^T
            R[9 * (size_t)i + r + 3 * c] = Rm[r + 3 * c];
        }
#pragma unroll
    for (int d = 0; d < 3; ++d) dv_bc[3 * (size_t)i + d] = vi[d] - old_v[d];
}

// dense-by-node -> compact BC table (node order) + Newton initial guess on the BC nodes
__global__ void k_compact_bc(int nn, const int* __restrict__ flag, const int* __restrict__ pos, const double* __restrict__ P, const double* __restrict__ R,
    const int* __restrict__ slip, const double* __restrict__ dv_bc, int* __restrict__ bc_node, int* __restrict__ bc_slip, double* __restrict__ bc_P,
    double* __restrict__ bc_R, double* __restrict__ bc_Rinv, double* __restrict__ dv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn || !flag[i]) return;
    const int b = pos[i];
    bc_node[b] = i;
    bc_slip[b] = slip[i];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            bc_P[9 * (size_t)b + r + 3 * c] = P[9 * (size_t)i + r + 3 * c];
            bc_R[9 * (size_t)b + r + 3 * c] = R[9 * (size_t)i + r + 3 * c];
            bc_Rinv[9 * (size_t)b + r + 3 * c] = R[9 * (size_t)i + c + 3 * r]; // rotation: inverse = transpose
        }
#pragma unroll
    for (int d = 0; d < 3; ++d) dv[3 * (size_t)i + d] = dv_bc[3 * (size_t)i + d];
}
__global__ void k_fill_dv(int n, double a, double b, double c, double* __restrict__ v)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int d = t % 3;
    v[t] = d == 0 ? a : (d == 1 ? b : c);
}

} // namespace

int set_colliders(Sim* s, int n, const hot_collider* objs)
{
    if (n < 0 || (n > 0 && !objs)) return fail(s, "hot_set_colliders: bad object table");
    std::vector<hot_collider> h(objs, objs + n);
    for (auto& o : h) {
        if (o.type < HOT_COLLIDER_STICKY || o.type > HOT_COLLIDER_GHOST) return fail(s, "hot_set_colliders: type must be STICKY 1, SLIP 2, SEPARATE 3 or GHOST 4");
        if (o.shape < HOT_SHAPE_HALFSPACE || o.shape > HOT_SHAPE_CAPPED_CYLINDER) return fail(s, "hot_set_colliders: unknown shape");
        if (!(o.s != 0.0)) return fail(s, "hot_set_colliders: scale s must be non-zero (CollisionObject.cpp:405)");
        if (o.shape == HOT_SHAPE_HALFSPACE) { // HalfSpace normalises its outward normal (AnalyticLevelSet.cpp:259-263)
            const double l = std::sqrt(o.p[3] * o.p[3] + o.p[4] * o.p[4] + o.p[5] * o.p[5]);
            if (!(l > 0.0)) return fail(s, "hot_set_colliders: zero half-space normal");
            for (int d = 3; d < 6; ++d) o.p[d] /= l;
        }
    }
    HOT_CUDA(s->colliders.reserve((size_t)(n > 0 ? n : 1) * sizeof(hot_collider)));
    if (n > 0) HOT_CUDA(cudaMemcpyAsync(s->colliders.p, h.data(), (size_t)n * sizeof(hot_collider), cudaMemcpyHostToDevice, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream)); // `h` goes out of scope
    s->n_colliders = n;
    return 0;
}

// buildInitialDvAndVnForNewton with the device-resident objects: BC table + initial guess, returns the number of collision nodes
int build_bc_from_colliders(Sim* s, int mode, int* n_bc_out)
{
    if (!s->p2g_done) return fail(s, "hot_build_bc: call hot_p2g first");
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    const size_t m = nn > 0 ? nn : 1;
    HOT_CUDA(s->col_coord.reserve(3 * m));
    HOT_CUDA(s->col_flag.reserve(m));
    HOT_CUDA(s->col_pos.reserve(m));
    HOT_CUDA(s->col_slip.reserve(m));
    HOT_CUDA(s->col_P.reserve(9 * m));
    HOT_CUDA(s->col_R.reserve(9 * m));
    HOT_CUDA(s->col_dv.reserve(3 * m));
    HOT_CUDA(s->dcount.reserve(16));
    int rc = fill_id2coord(s, s->col_coord.p);
    if (rc) return rc;
    k_fill_dv<<<nblk(3 * (long)nn), TPB, 0, st>>>(nn, s->gravity[0] * s->dt, s->gravity[1] * s->dt, s->gravity[2] * s->dt, s->dv.p);
    HOT_LAUNCHED(s);
    k_collide_nodes<<<nblk(nn), TPB, 0, st>>>(nn, s->col_coord.p, s->vn.p, s->dx, s->n_colliders, reinterpret_cast<const hot_collider*>(s->colliders.p),
        s->col_flag.p, s->col_P.p, s->col_R.p, s->col_slip.p, s->col_dv.p);
    HOT_LAUNCHED(s);
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, s->col_flag.p, s->col_pos.p, nn, st);
    HOT_CUDA(s->cub_tmp.reserve(bytes + 16));
    HOT_CUDA(cub::DeviceScan::ExclusiveSum(s->cub_tmp.p, bytes, s->col_flag.p, s->col_pos.p, nn, st));
    s->launches++;
    int last[2] = {0, 0};
    if (nn > 0) {
        HOT_CUDA(cudaMemcpyAsync(s->hcount + 8, s->col_pos.p + nn - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        HOT_CUDA(cudaMemcpyAsync(s->hcount + 9, s->col_flag.p + nn - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        HOT_CUDA(cudaStreamSynchronize(st));
        last[0] = s->hcount[8]; last[1] = s->hcount[9];
    }
    const int n_bc = last[0] + last[1];
    const size_t nb = n_bc > 0 ? n_bc : 1;
    HOT_CUDA(s->bc_node.reserve(nb));
    HOT_CUDA(s->bc_slip.reserve(nb));
    HOT_CUDA(s->bc_P.reserve(9 * nb));
    HOT_CUDA(s->bc_R.reserve(9 * nb));
    HOT_CUDA(s->bc_Rinv.reserve(9 * nb));
    s->bc_mode = mode;
    s->n_bc = n_bc;
    if (n_bc > 0) {
        k_compact_bc<<<nblk(nn), TPB, 0, st>>>(nn, s->col_flag.p, s->col_pos.p, s->col_P.p, s->col_R.p, s->col_slip.p, s->col_dv.p, s->bc_node.p, s->bc_slip.p,
            s->bc_P.p, s->bc_R.p, s->bc_Rinv.p, s->dv.p);
        HOT_LAUNCHED(s);
    }
    s->state_valid = s->hessian_valid = false;
    if (n_bc_out) *n_bc_out = n_bc;
    return 0;
}

} // namespace hot
