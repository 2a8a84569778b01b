// TEST INFRASTRUCTURE ONLY — CPU oracle for hot_b200 (see oracle_model.h for the parity status).
//
// A dependency-free C++17 + OpenMP restatement of the reference's implicit-MPM hot path that keeps the
// reference's data layout where it affects bandwidth (128-byte AoS GridState in 4 KB pages, particle
// attribute arrays in original order accessed through particle_order) and its parallel schedule
// (u64 key sort -> page groups -> 8 sequential colour passes with a parallel loop over the groups of
// one colour), so that it doubles as the "reference CPU path" baseline of bench.py.
// Every function cites the reference file:line it follows.  The exported orc_* entry points have the same
// argument meaning as the hot_* C-ABI in include/hot_b200.h so tests drive both with one harness.
#include "oracle_model.h"
#include <algorithm>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>
#include <omp.h>

using namespace orc;

namespace {

// Lib/MPM/MpmGrid.h:14-34 (T=double): 128-byte record; v doubles as the momentum accumulator in P2G.
struct GridState {
    double v[3];
    double m;
    double new_v[3];
    int64_t idx;
    double padding[6];
    double phase_field;
    double phase_field_multiplier;
};
static_assert(sizeof(GridState) == 128, "GridState<double,3> is 128 bytes");

struct Sim {
    Mask mask{7};
    double dx = 1, apic_rpic_ratio = 1, cfl = 0.6;
    double dt = 0, gravity[3] = {0, 0, 0};
    std::string err;

    // particles, original order (AoS like Eigen StdVector<TV>/<TM>)
    int64_t N = 0;
    std::vector<double> X, V, mass, C, F, vol, mu, lambda, gradV, Jp;
    int plastic_model = 0;
    double plastic_param[5] = {0, 0, 0, 0, 0};

    // a5 outputs (MpmSimulationBase.h:98-112)
    std::vector<uint64_t> sorter;
    std::vector<int> order;
    std::vector<uint64_t> base_offset;
    std::vector<std::pair<int, int>> groups;
    std::vector<uint64_t> block_offset; // key>>32 per group

    // page map: list in first-Set order + flat lookup page id -> slot
    std::vector<uint64_t> pages;
    std::vector<int32_t> slot_of_page;
    std::vector<GridState> grid; // pages.size() * elements_per_block
    uint64_t lin27[27];

    int num_nodes = 0;
    double dpdf_norm_max = -1; // computeCharacteristicNorm's function-static cache (MultigridSimulation.h:131-133), per particle set here
    std::vector<double> dv, vn, mass_matrix;
    void* force_state = nullptr; // oracle_force.inl
    void* matrix_state = nullptr; // oracle_matrix.inl

    GridState* node_at(uint64_t offset)
    {
        int32_t s = slot_of_page[offset >> 12];
        return &grid[(size_t)s * mask.elements_per_block + ((offset & 0xfff) >> mask.data_bits)];
    }

    // MpmGrid.h:245-296 (dim==3): i outer, k inner; weights multiplied in exactly this association.
    template <class OP>
    void iterate_kernel(const Spline& sp, uint64_t base_off, const OP& op)
    {
        const double one_over_dx = sp.one_over_dx;
        int coord[3];
        for (int i = 0; i < 3; ++i) {
            double wi = sp.w[0][i];
            double dwidxi = one_over_dx * sp.dw[0][i];
            coord[0] = sp.base[0] + i;
            for (int j = 0; j < 3; ++j) {
                double wj = sp.w[1][j];
                double wij = wi * wj;
                double dwijdxi = dwidxi * wj;
                double dwijdxj = wi * one_over_dx * sp.dw[1][j];
                coord[1] = sp.base[1] + j;
                for (int k = 0; k < 3; ++k) {
                    coord[2] = sp.base[2] + k;
                    double wk = sp.w[2][k];
                    double wijk = wij * wk;
                    double dw[3] = {dwijdxi * wk, dwijdxj * wk, wij * one_over_dx * sp.dw[2][k]};
                    uint64_t off = mask.packed_add(base_off, lin27[i * 9 + j * 3 + k]);
                    op(coord, wijk, dw, *node_at(off));
                }
            }
        }
    }

    // MpmSimulationBase.h:251-264: 8 sequential colours, parallel over the groups of one colour.
    template <class OP>
    void for_colored_groups(const OP& op)
    {
        for (uint64_t color = 0; color < 8; ++color) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int g = 0; g < (int)groups.size(); ++g) {
                if ((block_offset[g] & 7) != color) continue;
                op(g);
            }
        }
    }
};

int fail(Sim* s, const char* msg)
{
    s->err = msg;
    return -1;
}

} // namespace

extern "C" {

// ---- SPGrid addressing entry points (validated against oracle/_ref and tests/golden) -----------------
void orc_mask_info(int fp32, int* out6, unsigned long long* masks3)
{
    Mask m(fp32 ? 6 : 7);
    out6[0] = m.data_bits; out6[1] = m.block_bits; out6[2] = m.xb; out6[3] = m.yb; out6[4] = m.zb; out6[5] = m.elements_per_block;
    masks3[0] = m.xmask; masks3[1] = m.ymask; masks3[2] = m.zmask;
}
void orc_linear_offset(int fp32, long n, const int* ijk, unsigned long long* out)
{
    Mask m(fp32 ? 6 : 7);
    for (long a = 0; a < n; ++a) out[a] = m.linear_offset(ijk[3 * a], ijk[3 * a + 1], ijk[3 * a + 2]);
}
void orc_linear_to_coord(int fp32, long n, const unsigned long long* off, int* ijk)
{
    Mask m(fp32 ? 6 : 7);
    for (long a = 0; a < n; ++a) m.linear_to_coord(off[a], ijk + 3 * a);
}
void orc_packed_add(int fp32, long n, const unsigned long long* a, const unsigned long long* b, unsigned long long* out)
{
    Mask m(fp32 ? 6 : 7);
    for (long q = 0; q < n; ++q) out[q] = m.packed_add(a[q], b[q]);
}
// page activation of MpmSimulationBase.cpp:1100-1125 on a bare list of group offsets (fp32 or fp64 mask)
long orc_activate(int fp32, long n_groups, const unsigned long long* group_offsets, unsigned long long* out_blocks, long cap)
{
    Mask m(fp32 ? 6 : 7);
    std::vector<uint64_t> list;
    std::vector<uint64_t> seen; // sorted set of page ids is enough for a test helper
    auto set_page = [&](uint64_t off) {
        uint64_t p = off >> 12;
        auto it = std::lower_bound(seen.begin(), seen.end(), p);
        if (it == seen.end() || *it != p) {
            seen.insert(it, p);
            list.push_back(p << 12);
        }
    };
    int x = 1 << m.xb, y = 1 << m.yb, z = 1 << m.zb;
    for (long g = 0; g < n_groups; ++g) {
        set_page(group_offsets[g]);
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) set_page(m.packed_add(group_offsets[g], m.linear_offset(x * i, y * j, z * k)));
    }
    for (long b = 0; b < (long)list.size() && b < cap; ++b) out_blocks[b] = list[b];
    return (long)list.size();
}

// ---- simulation object ---------------------------------------------------------------------------------
void* orc_create(double dx, double apic_rpic_ratio, double cfl)
{
    Sim* s = new Sim;
    s->dx = dx;
    s->apic_rpic_ratio = apic_rpic_ratio;
    s->cfl = cfl;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) s->lin27[i * 9 + j * 3 + k] = s->mask.linear_offset(i, j, k);
    return s;
}
void orc_destroy(void* h) { delete (Sim*)h; }
const char* orc_last_error(void* h) { return ((Sim*)h)->err.c_str(); }
int orc_num_threads() { return omp_get_max_threads(); }

int orc_set_dt_gravity(void* h, double dt, const double* g)
{
    Sim* s = (Sim*)h;
    s->dt = dt;
    for (int d = 0; d < 3; ++d) s->gravity[d] = g[d];
    return 0;
}

// Particles (Lib/Ziran/Math/Geometry/Particles.h:8-45 + the F/vol/model attributes the force helper reads,
// Lib/MPM/MpmParticleHandleBase.cpp:280-289).  All arrays AoS in original particle order; 3x3 column-major.
int orc_set_particles(void* h, long n, const double* X, const double* V, const double* mass, const double* C, const double* F,
    const double* vol, const double* mu, const double* lambda)
{
    Sim* s = (Sim*)h;
    s->N = n;
    s->X.assign(X, X + 3 * n);
    s->V.assign(V, V + 3 * n);
    s->mass.assign(mass, mass + n);
    s->C.assign(C, C + 9 * n);
    s->F.assign(F, F + 9 * n);
    s->vol.assign(vol, vol + n);
    s->mu.assign(mu, mu + n);
    s->lambda.assign(lambda, lambda + n);
    s->gradV.assign(9 * n, 0.0);
    s->Jp.assign(n, 1.0);
    s->dpdf_norm_max = -1;
    return 0;
}
int orc_get_particles(void* h, double* X, double* V, double* C, double* F, double* gradV)
{
    Sim* s = (Sim*)h;
    if (X) std::copy(s->X.begin(), s->X.end(), X);
    if (V) std::copy(s->V.begin(), s->V.end(), V);
    if (C) std::copy(s->C.begin(), s->C.end(), C);
    if (F) std::copy(s->F.begin(), s->F.end(), F);
    if (gradV) std::copy(s->gradV.begin(), s->gradV.end(), gradV);
    return 0;
}

// a5: MpmSimulationBase::sortParticlesAndPolluteGrid, Lib/MPM/MpmSimulationBase.cpp:1066-1137
int orc_sort_and_activate(void* h)
{
    Sim* s = (Sim*)h;
    const Mask& mk = s->mask;
    const int index_bits = 32 - mk.block_bits;
    const long n = s->N;
    if (n >= (1l << index_bits)) return fail(s, "particle count must be < 2^index_bits (MpmSimulationBase.cpp:1072)");
    s->sorter.resize(n);
    s->order.resize(n);
    s->base_offset.resize(n);
    const double one_over_dx = 1.0 / s->dx;
    int bad = 0;
#pragma omp parallel for reduction(| : bad)
    for (long i = 0; i < n; ++i) {
        int b[3];
        for (int d = 0; d < 3; ++d) {
            b[d] = base_node(index_space(s->X[3 * i + d], one_over_dx));
            if (b[d] < 0 || b[d] + 2 >= 4096) bad = 1;
        }
        uint64_t off = mk.linear_offset(b[0], b[1], b[2]);
        s->sorter[i] = ((off >> mk.data_bits) << index_bits) + (uint64_t)i;
    }
    if (bad) return fail(s, "particle outside the 4096^3 SPGrid box (MpmGrid.h:109,127)");
    std::sort(s->sorter.begin(), s->sorter.end());

    s->groups.clear();
    s->block_offset.clear();
    int last = 0;
    for (long i = 0; i < n; ++i)
        if (i == n - 1 || (s->sorter[i] >> 32) != (s->sorter[i + 1] >> 32)) {
            s->groups.emplace_back(last, (int)i);
            s->block_offset.push_back(s->sorter[i] >> 32);
            last = (int)i + 1;
        }

    // page map Clear(): forget the previous list and reset its lookup entries
    for (uint64_t p : s->pages) s->slot_of_page[p >> 12] = -1;
    s->pages.clear();
    auto set_page = [&](uint64_t off) {
        uint64_t p = off >> 12;
        if (p >= s->slot_of_page.size()) {
            size_t cap = s->slot_of_page.size() ? s->slot_of_page.size() : 1024;
            while (cap <= p) cap *= 2;
            s->slot_of_page.resize(cap, -1);
        }
        if (s->slot_of_page[p] < 0) {
            s->slot_of_page[p] = (int32_t)s->pages.size();
            s->pages.push_back(p << 12);
        }
    };
    const int bx = 1 << mk.xb, by = 1 << mk.yb, bz = 1 << mk.zb;
    uint64_t nb[8];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) nb[i * 4 + j * 2 + k] = mk.linear_offset(bx * i, by * j, bz * k);
    for (long i = 0; i < n; ++i) {
        s->order[i] = (int)(s->sorter[i] & ((1ull << index_bits) - 1));
        uint64_t off = (s->sorter[i] >> index_bits) << mk.data_bits;
        s->base_offset[s->order[i]] = off;
        if (i == n - 1 || (s->sorter[i] >> 32) != (s->sorter[i + 1] >> 32)) {
            set_page(off);
            for (int q = 0; q < 8; ++q) set_page(mk.packed_add(off, nb[q]));
        }
    }
    // zero the pages, idx = -1 (MpmSimulationBase.cpp:1128-1136)
    const int E = mk.elements_per_block;
    s->grid.resize(s->pages.size() * (size_t)E);
    std::memset(s->grid.data(), 0, s->grid.size() * sizeof(GridState));
    for (auto& g : s->grid) g.idx = -1;
    s->num_nodes = 0;
    return 0;
}

long orc_num_particles(void* h) { return ((Sim*)h)->N; }
long orc_num_groups(void* h) { return (long)((Sim*)h)->groups.size(); }
long orc_num_pages(void* h) { return (long)((Sim*)h)->pages.size(); }
int orc_num_nodes(void* h) { return ((Sim*)h)->num_nodes; }

int orc_get_sort(void* h, unsigned long long* sorter, int* order, unsigned long long* base_offset)
{
    Sim* s = (Sim*)h;
    if (sorter) std::copy(s->sorter.begin(), s->sorter.end(), sorter);
    if (order) std::copy(s->order.begin(), s->order.end(), order);
    if (base_offset) std::copy(s->base_offset.begin(), s->base_offset.end(), base_offset);
    return 0;
}
int orc_get_groups(void* h, int* first, int* last, unsigned long long* block_offset)
{
    Sim* s = (Sim*)h;
    for (size_t g = 0; g < s->groups.size(); ++g) {
        if (first) first[g] = s->groups[g].first;
        if (last) last[g] = s->groups[g].second;
        if (block_offset) block_offset[g] = s->block_offset[g];
    }
    return 0;
}
int orc_get_pages(void* h, unsigned long long* offsets)
{
    Sim* s = (Sim*)h;
    std::copy(s->pages.begin(), s->pages.end(), offsets);
    return 0;
}

// a6 + a7: particlesToGridHelper<true,false> (MpmSimulationBase.cpp:611-656), then getNumNodes
// (MpmGrid.h:148-161) and the mass normalisation of particlesToGrid (MpmSimulationBase.cpp:521-532).
int orc_p2g(void* h, int* n_nodes)
{
    Sim* s = (Sim*)h;
    const double dx = s->dx;
    // pages are re-zeroed so the call is repeatable (the reference zeroes them in the sort, :1128-1136)
    std::memset(s->grid.data(), 0, s->grid.size() * sizeof(GridState));
    for (auto& g : s->grid) g.idx = -1;

    s->for_colored_groups([&](int grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            const double* Xp = &s->X[3 * i];
            double mass = s->mass[i];
            double momentum[3] = {mass * s->V[3 * i], mass * s->V[3 * i + 1], mass * s->V[3 * i + 2]};
            double Cm[9];
            for (int q = 0; q < 9; ++q) Cm[q] = mass * s->C[9 * i + q];
            Spline sp(Xp, dx);
            s->iterate_kernel(sp, s->base_offset[i], [&](const int* node, double w, const double*, GridState& g) {
                // velocity_density * [xi - xp; 1] * w, 4x4 times 4-vector (:640-652)
                double d[3] = {node[0] * dx - Xp[0], node[1] * dx - Xp[1], node[2] * dx - Xp[2]};
                double delta[4];
                for (int r = 0; r < 3; ++r) delta[r] = (Cm[r] * d[0] + Cm[r + 3] * d[1] + Cm[r + 6] * d[2] + momentum[r] * 1.0) * w;
                delta[3] = (mass * 1.0) * w;
                g.m += delta[3];
                g.v[0] += delta[0];
                g.v[1] += delta[1];
                g.v[2] += delta[2];
            });
        }
    });

    // getNumNodes: serial scan, page list order x element order
    int total = 0;
    for (auto& g : s->grid)
        if (g.m != 0) g.idx = total++;
    s->num_nodes = total;
    // normalise (iterateGrid visits idx>=0 nodes only; the others stay zero)
#pragma omp parallel for
    for (long a = 0; a < (long)s->grid.size(); ++a) {
        GridState& g = s->grid[a];
        if (g.idx >= 0) {
            if (g.m != 0) {
                g.v[0] /= g.m; g.v[1] /= g.m; g.v[2] /= g.m;
            }
            else
                g.v[0] = g.v[1] = g.v[2] = 0;
        }
    }
    s->dv.assign(3 * (size_t)total, 0.0);
    s->vn.assign(3 * (size_t)total, 0.0);
    s->mass_matrix.assign(total, 0.0); // buildMassMatrix, MpmSimulationBase.cpp:817-826
    for (auto& g : s->grid)
        if (g.idx >= 0) {
            s->mass_matrix[g.idx] = g.m;
            for (int d = 0; d < 3; ++d) s->vn[3 * g.idx + d] = g.v[d];
        }
    if (n_nodes) *n_nodes = total;
    return 0;
}

// grid read-back, page-list order x element order
int orc_get_grid(void* h, long long* idx, double* m, double* v)
{
    Sim* s = (Sim*)h;
    for (size_t a = 0; a < s->grid.size(); ++a) {
        if (idx) idx[a] = s->grid[a].idx;
        if (m) m[a] = s->grid[a].m;
        if (v) {
            v[3 * a] = s->grid[a].v[0]; v[3 * a + 1] = s->grid[a].v[1]; v[3 * a + 2] = s->grid[a].v[2];
        }
    }
    return 0;
}

// node coordinates per DOF id (ImplicitSolver.h id2coord, filled through grid.iterateGrid)
int orc_get_id2coord(void* h, int* coord)
{
    Sim* s = (Sim*)h;
    const int E = s->mask.elements_per_block;
    for (size_t p = 0; p < s->pages.size(); ++p) {
        int base[3];
        s->mask.linear_to_coord(s->pages[p], base);
        for (int e = 0; e < E; ++e) {
            const GridState& g = s->grid[p * E + e];
            if (g.idx < 0) continue;
            int c[3];
            s->mask.linear_to_coord((uint64_t)e << s->mask.data_bits, c);
            for (int d = 0; d < 3; ++d) coord[3 * g.idx + d] = base[d] + c[d];
        }
    }
    return 0;
}

// buildMassMatrix, MpmSimulationBase.cpp:817-826
int orc_get_mass_matrix(void* h, double* mass)
{
    Sim* s = (Sim*)h;
    for (auto& g : s->grid)
        if (g.idx >= 0) mass[g.idx] = g.m;
    return 0;
}

int orc_set_dv(void* h, const double* dv)
{
    Sim* s = (Sim*)h;
    s->dv.assign(dv, dv + 3 * (size_t)s->num_nodes);
    return 0;
}

// constructNewVelocityFromNewtonResult, MpmSimulationBase.cpp:891-901
static void construct_new_velocity(Sim* s)
{
#pragma omp parallel for
    for (long a = 0; a < (long)s->grid.size(); ++a) {
        GridState& g = s->grid[a];
        g.new_v[0] = g.new_v[1] = g.new_v[2] = 0;
        if (g.idx >= 0)
            for (int d = 0; d < 3; ++d) g.new_v[d] = g.v[d] + s->dv[3 * g.idx + d];
    }
}

int orc_apply_plasticity(void* h);
// a23: gridToParticlesHelper<true,false,false> (MpmSimulationBase.cpp:930-1006) preceded by
// constructNewVelocityFromNewtonResult (:891-901) and followed by evolveStrain
// (FBasedMpmForceHelper.cpp:100-114).  flags[0] = faster than a grid cell, flags[1] = faster than half.
int orc_g2p(void* h, double dt, int* flags)
{
    Sim* s = (Sim*)h;
    construct_new_velocity(s);
    const double dx = s->dx;
    const double D_inverse = 4.0 / (dx * dx); // MpmSimulationBase.cpp:114-118 (quadratic)
    const double r = s->apic_rpic_ratio;
    int fast = 0, half_fast = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : fast, half_fast)
    for (int grp = 0; grp < (int)s->groups.size(); ++grp) {
        for (int sidx = s->groups[grp].first; sidx <= s->groups[grp].second; ++sidx) {
            int i = s->order[sidx];
            double* Xp = &s->X[3 * i];
            double picV[3] = {0, 0, 0};
            double Bp[9] = {0}, gradVp[9] = {0};
            Spline sp(Xp, dx);
            s->iterate_kernel(sp, s->base_offset[i], [&](const int* node, double w, const double* dw, GridState& g) {
                for (int d = 0; d < 3; ++d) picV[d] += w * g.new_v[d];
                double xm[3] = {node[0] * dx - Xp[0], node[1] * dx - Xp[1], node[2] * dx - Xp[2]};
                for (int c = 0; c < 3; ++c)
                    for (int rr = 0; rr < 3; ++rr) {
                        Bp[rr + 3 * c] += w * g.new_v[rr] * xm[c];
                        gradVp[rr + 3 * c] += g.new_v[rr] * dw[c];
                    }
            });
            for (int d = 0; d < 3; ++d) s->V[3 * i + d] = picV[d];
            double CC[9];
            for (int q = 0; q < 9; ++q) CC[q] = Bp[q] * D_inverse;
            for (int c = 0; c < 3; ++c)
                for (int rr = 0; rr < 3; ++rr)
                    s->C[9 * i + rr + 3 * c] = ((r + 1) * 0.5) * CC[rr + 3 * c] + ((r - 1) * 0.5) * CC[c + 3 * rr];
            std::copy(gradVp, gradVp + 9, &s->gradV[9 * i]);
            double inc = 0;
            for (int d = 0; d < 3; ++d) {
                double incr = dt * picV[d];
                Xp[d] += incr;
                inc += incr * incr;
            }
            double dx2 = dx * dx;
            if (inc > dx2) fast = 1;
            if (inc > dx2 * 0.25 * (s->cfl * s->cfl)) half_fast = 1;
            // evolveStrain: F = (I + dt gradV) F
            double A[9];
            for (int q = 0; q < 9; ++q) A[q] = dt * gradVp[q];
            A[0] += 1; A[4] += 1; A[8] += 1;
            mat_mul(A, &s->F[9 * i], &s->F[9 * i]);
        }
    }
    if (flags) {
        flags[0] = fast;
        flags[1] = half_fast;
    }
    if (dt != 0.0) orc_apply_plasticity(h); // MpmSimulationBase.cpp:1039-1041
    return 0;
}

} // extern "C"

#include "oracle_force.inl"
#include "oracle_matrix.inl"
#include "oracle_solver.inl"
