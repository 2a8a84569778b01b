// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// The particle <-> grid transfers of one MPM substep run on the REFERENCE'S OWN grid code: Lib/MPM/MpmGrid.h is compiled where it
// lies (GridState<T,dim> :14-34, BSplineWeights :56-79, MpmGrid::{getNumNodes :148-161, iterateGrid :204-245, iterateKernel :247-300})
// on top of the reference's SPGrid allocator / page map (Lib/SPGrid/Core) and its B-spline header (Lib/Ziran/Math/Splines/BSplines.h),
// against the Eigen / TBB stand-ins of ref_shim/ (parallel_for runs serially; the 8-colour schedule makes the result independent of that).
// Built by oracle/Makefile into oracle/_ref/libmpmgrid_ref.so.
//
// What IS the reference's code here: the 128-byte node record and the mmap-ed 4096^3 SPGrid array it lives in, the page map
// (Set_Page / Update_Block_Offsets / Get_Blocks), baseNode + the quadratic weights, the 27-node kernel walk with its weight / weight-gradient
// products and Packed_Add offsets, the DOF numbering scan and the "valid node" iteration.
// What is NOT: the member functions of MpmSimulationBase (Lib/MPM/MpmSimulationBase.cpp needs Scene / DataManager / Particles / TBB
// containers / Partio and cannot be compiled here).  Their particle loops — a few lines each around the grid calls above — are written
// out below with the reference's own types, statement for statement, each citing the lines it follows:
//   sortParticlesAndPolluteGrid          MpmSimulationBase.cpp:1066-1137   (rows a5, a2)
//   particlesToGridHelper<true,false>    MpmSimulationBase.cpp:611-656     (row a6)
//   particlesToGrid tail                 MpmSimulationBase.cpp:521-532     (row a7: getNumNodes + normalisation)
//   constructNewVelocityFromNewtonResult MpmSimulationBase.cpp:891-901
//   gridToParticlesHelper<true,false,false> MpmSimulationBase.cpp:930-1006 (row a23, without evolveStrain / plasticity)
// tests/golden/make_mpmgrid_golden.py writes tests/golden/mpmgrid_ref.npz from this library; tests/test_oracle_mpmgrid_ref.py compares the
// oracle's restatement (hot_oracle.cpp) and the CUDA path with it.
#include <immintrin.h>
#include <algorithm>
#include <array>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>
#include <tbb/tbb.h>
#include <Ziran/CS/Util/Debug.h>
#include <MPM/MpmGrid.h>

using namespace ZIRAN;

namespace {
typedef double T;
constexpr int dim = 3;
typedef Vector<T, dim> TV;
typedef Vector<int, dim> IV;
typedef Matrix<T, dim, dim> TM;
typedef Vector<T, 4> TV4;
typedef Matrix<T, 4, 4> TM4;
typedef MpmGrid<T, dim>::SparseMask SparseMask;
constexpr int interpolation_degree = MpmGrid<T, dim>::interpolation_degree;

struct RefSim {
    MpmGrid<T, dim> grid;
    T dx = 0, apic_rpic_ratio = 1, cfl = 0.6, D_inverse = 0;
    int count = 0, num_nodes = 0;
    std::vector<TV> X, V;
    std::vector<T> mass;
    std::vector<TM> C, scratch_gradV;
    // MpmSimulationBase.h: the sort products
    std::vector<uint64_t> particle_base_offset, particle_sorter;
    std::vector<int> particle_order;
    std::vector<std::pair<int, int>> particle_group;
    std::vector<uint64_t> block_offset;
};
} // namespace

extern "C" {

// row a3: layout of the node record as the reference's compiler lays it out {sizeof, offsetof v, m, new_v, idx, elements_per_block}
void mpmgrid_ref_layout(long long* out)
{
    typedef GridState<T, dim> G;
    out[0] = (long long)sizeof(G);
    out[1] = (long long)offsetof(G, v);
    out[2] = (long long)offsetof(G, m);
    out[3] = (long long)offsetof(G, new_v);
    out[4] = (long long)offsetof(G, idx);
    out[5] = (long long)SparseMask::elements_per_block;
}

void* mpmgrid_ref_create(double dx, double apic_rpic_ratio, double cfl)
{
    RefSim* s = new RefSim();
    s->dx = dx;
    s->apic_rpic_ratio = apic_rpic_ratio;
    s->cfl = cfl;
    s->D_inverse = 4 / (dx * dx); // MpmSimulationBase.cpp:114-118, quadratic
    return s;
}

void mpmgrid_ref_destroy(void* h) { delete (RefSim*)h; }

// X, V: n x 3; C: n x 9 column-major (the oracle's buffer layout)
void mpmgrid_ref_set_particles(void* h, long n, const double* X, const double* V, const double* mass, const double* C)
{
    RefSim* s = (RefSim*)h;
    s->count = (int)n;
    s->X.resize(n); s->V.resize(n); s->mass.assign(mass, mass + n); s->C.resize(n); s->scratch_gradV.resize(n);
    for (long i = 0; i < n; ++i) {
        for (int d = 0; d < dim; ++d) { s->X[i](d) = X[3 * i + d]; s->V[i](d) = V[3 * i + d]; }
        for (int c = 0; c < dim; ++c)
            for (int r = 0; r < dim; ++r) s->C[i](r, c) = C[9 * i + r + 3 * c];
        s->scratch_gradV[i] = TM::Zero();
    }
}

void mpmgrid_ref_get_particles(void* h, double* X, double* V, double* C, double* gradV)
{
    RefSim* s = (RefSim*)h;
    for (int i = 0; i < s->count; ++i) {
        for (int d = 0; d < dim; ++d) { X[3 * i + d] = s->X[i](d); V[3 * i + d] = s->V[i](d); }
        for (int c = 0; c < dim; ++c)
            for (int r = 0; r < dim; ++r) { C[9 * i + r + 3 * c] = s->C[i](r, c); gradV[9 * i + r + 3 * c] = s->scratch_gradV[i](r, c); }
    }
}

// MpmSimulationBase.cpp:1066-1137 (tbb::parallel_sort -> std::sort: the keys are unique, so the order is the same)
long mpmgrid_ref_sort(void* h)
{
    RefSim* s = (RefSim*)h;
    MpmGrid<T, dim>& grid = s->grid;
    const int count = s->count;
    auto& particle_sorter = s->particle_sorter;
    auto& particle_order = s->particle_order;
    auto& particle_base_offset = s->particle_base_offset;
    auto& particle_group = s->particle_group;
    auto& block_offset = s->block_offset;

    constexpr int index_bits = (32 - SparseMask::block_bits);
    ZIRAN_ASSERT(count < (1 << index_bits));
    particle_base_offset.resize(count);
    particle_sorter.resize(count);
    particle_order.resize(count);

    T one_over_dx = (T)1 / s->dx;
    tbb::parallel_for(0, count, [&](int i) {
        uint64_t offset = SparseMask::Linear_Offset(to_std_array(
            baseNode<interpolation_degree, T, dim>(s->X[i] * one_over_dx)));
        particle_sorter[i] = ((offset >> SparseMask::data_bits) << index_bits) + i;
    });

    std::sort(particle_sorter.begin(), particle_sorter.end());

    particle_group.clear();
    block_offset.clear();
    int last_index = 0;
    for (int i = 0; i < count; ++i)
        if (i == count - 1 || (particle_sorter[i] >> 32) != (particle_sorter[i + 1] >> 32)) {
            particle_group.push_back(std::make_pair(last_index, i));
            block_offset.push_back(particle_sorter[i] >> 32);
            last_index = i + 1;
        }

    grid.page_map->Clear();
    for (int i = 0; i < count; ++i) {
        particle_order[i] = (int)(particle_sorter[i] & ((1ll << index_bits) - 1));
        uint64_t offset = (particle_sorter[i] >> index_bits) << SparseMask::data_bits;
        particle_base_offset[particle_order[i]] = offset;
        if (i == count - 1 || (particle_sorter[i] >> 32) != (particle_sorter[i + 1] >> 32)) {
            grid.page_map->Set_Page(offset);
            auto x = 1 << SparseMask::block_xbits;
            auto y = 1 << SparseMask::block_ybits;
            auto z = 1 << SparseMask::block_zbits;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b)
                    for (int c = 0; c < 2; ++c)
                        grid.page_map->Set_Page(SparseMask::Packed_Add(
                            offset, SparseMask::Linear_Offset(x * a, y * b, z * c)));
        }
    }
    grid.page_map->Update_Block_Offsets();

    auto grid_array = grid.grid->Get_Array();
    auto blocks = grid.page_map->Get_Blocks();
    for (int b = 0; b < (int)blocks.second; ++b) {
        auto base_offset = blocks.first[b];
        std::memset(&grid_array(base_offset), 0, (size_t)(1 << MpmGrid<T, dim>::log2_page));
        GridState<T, dim>* g = reinterpret_cast<GridState<T, dim>*>(&grid_array(base_offset));
        for (int i = 0; i < (int)SparseMask::elements_per_block; ++i)
            g[i].idx = -1;
    }
    return (long)particle_group.size();
}

long mpmgrid_ref_num_pages(void* h) { return (long)((RefSim*)h)->grid.page_map->Get_Blocks().second; }

void mpmgrid_ref_get_sort(void* h, unsigned long long* sorter, int* order, unsigned long long* base, int* first, int* last,
                          unsigned long long* blk, unsigned long long* pages)
{
    RefSim* s = (RefSim*)h;
    for (int i = 0; i < s->count; ++i) { sorter[i] = s->particle_sorter[i]; order[i] = s->particle_order[i]; base[i] = s->particle_base_offset[i]; }
    for (size_t g = 0; g < s->particle_group.size(); ++g) { first[g] = s->particle_group[g].first; last[g] = s->particle_group[g].second; blk[g] = s->block_offset[g]; }
    auto blocks = s->grid.page_map->Get_Blocks();
    for (int b = 0; b < (int)blocks.second; ++b) pages[b] = blocks.first[b];
}

// particlesToGridHelper<USE_APIC_BLEND_RPIC = true, USE_MPM_DEGREE_ONE = false> (MpmSimulationBase.cpp:611-656) + :521-532
int mpmgrid_ref_p2g(void* h)
{
    RefSim* s = (RefSim*)h;
    MpmGrid<T, dim>& grid = s->grid;
    const T dx = s->dx;
    for (uint64_t color = 0; color < (1 << dim); ++color) {
        tbb::parallel_for(0, (int)s->particle_group.size(), [&](int group_idx) {
            if ((s->block_offset[group_idx] & ((1 << dim) - 1)) != color)
                return;
            for (int idx = s->particle_group[group_idx].first; idx <= s->particle_group[group_idx].second; ++idx) {
                int i = s->particle_order[idx];
                TV& Xp = s->X[i];
                T mass = s->mass[i];
                TV momentum = s->mass[i] * s->V[i];
                TM C = TM::Zero();
                C = mass * s->C[i];
                TM4 velocity_density = TM4::Zero();
                velocity_density.template block<dim, dim>(0, 0) = C;       // topLeftCorner<dim, dim>()
                velocity_density.template block<dim, 1>(0, dim) = momentum; // topRightCorner<dim, 1>()
                velocity_density(3, 3) = mass;
                BSplineWeights<T, dim> spline(Xp, dx);
                grid.iterateKernel(spline, s->particle_base_offset[i], [&](const IV& node, T w, const TV& dw, GridState<T, dim>& g) {
                    TV4 xi_minus_xp = TV4::Zero();
                    xi_minus_xp.template block<dim, 1>(0, 0) = node.template cast<T>() * dx - Xp;
                    xi_minus_xp(3) = 1;
                    TV4 velocity_delta = velocity_density * xi_minus_xp * w;
                    g.m += velocity_delta(3);
                    g.v += velocity_delta.template block<dim, 1>(0, 0);
                });
            }
        });
    }
    s->num_nodes = grid.getNumNodes();
    grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        if (g.m != 0) {
            g.v /= g.m;
        }
        else {
            g.v = TV::Zero();
        }
    });
    return s->num_nodes;
}

// read-back in page-list x element order (all nodes of the activated pages) and, through MpmGrid::iterateGrid, the coordinate of every DOF
void mpmgrid_ref_get_grid(void* h, long long* idx, double* m, double* v, int* id2coord)
{
    RefSim* s = (RefSim*)h;
    auto grid_array = s->grid.grid->Get_Array();
    auto blocks = s->grid.page_map->Get_Blocks();
    const int E = (int)SparseMask::elements_per_block;
    for (int b = 0; b < (int)blocks.second; ++b) {
        GridState<T, dim>* g = reinterpret_cast<GridState<T, dim>*>(&grid_array(blocks.first[b]));
        for (int e = 0; e < E; ++e) {
            size_t a = (size_t)b * E + e;
            idx[a] = g[e].idx;
            m[a] = g[e].m;
            for (int d = 0; d < dim; ++d) v[3 * a + d] = g[e].v(d);
        }
    }
    s->grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        for (int d = 0; d < dim; ++d) id2coord[3 * g.idx + d] = node(d);
    });
}

// constructNewVelocityFromNewtonResult (MpmSimulationBase.cpp:891-901) + gridToParticlesHelper<true, false, false> (:930-1006);
// evolveStrain / applyPlasticity (force helper) are not part of this library.  flags = {faster than a cell, faster than half a cell}
void mpmgrid_ref_g2p(void* h, const double* dv, double dt, int* flags)
{
    RefSim* s = (RefSim*)h;
    MpmGrid<T, dim>& grid = s->grid;
    const T dx = s->dx;
    const T apic_rpic_ratio = s->apic_rpic_ratio, cfl = s->cfl, D_inverse = s->D_inverse;
    grid.iterateGrid([&](IV node, GridState<T, dim>& g) {
        TV d;
        for (int q = 0; q < dim; ++q) d(q) = dv[3 * g.idx + q];
        g.new_v = g.v + d;
    });
    bool faster_than_grid_cell = false, faster_than_half_grid_cell = false;
    tbb::parallel_for(0, (int)s->particle_group.size(), [&](int group_idx) {
        for (int idx = s->particle_group[group_idx].first; idx <= s->particle_group[group_idx].second; ++idx) {
            bool local_faster_than_grid_cell = false;
            bool local_faster_than_half_grid_cell = false;
            int i = s->particle_order[idx];
            TV& Xp = s->X[i];
            TV picV = TV::Zero();
            BSplineWeights<T, dim> spline(Xp, dx);
            TM Bp = TM::Zero();
            TM& gradVp = s->scratch_gradV[i];
            gradVp = TM::Zero();
            grid.iterateKernel(spline, s->particle_base_offset[i], [&](IV node, T w, TV dw, GridState<T, dim>& g) {
                picV += w * g.new_v;
                TV xi_minus_xp = node.template cast<T>() * dx - Xp;
                Bp.noalias() += w * g.new_v * xi_minus_xp.transpose();
                gradVp.noalias() += g.new_v * dw.transpose();
            });
            s->V[i] = picV;
            TM CC = Bp * D_inverse;
            s->C[i] = ((apic_rpic_ratio + 1) * (T)0.5) * CC + ((apic_rpic_ratio - 1) * (T)0.5) * CC.transpose();
            TV increment = dt * picV;
            s->X[i] += increment;
            T inc = increment.squaredNorm();
            T dx2 = dx * dx;
            local_faster_than_grid_cell = local_faster_than_grid_cell + (inc > dx2);
            local_faster_than_half_grid_cell = local_faster_than_half_grid_cell + (inc > dx2 * (T)0.25 * (cfl * cfl));
            if (local_faster_than_half_grid_cell) faster_than_half_grid_cell = true;
            if (local_faster_than_grid_cell) faster_than_grid_cell = true;
        }
    });
    flags[0] = faster_than_grid_cell;
    flags[1] = faster_than_half_grid_cell;
}

} // extern "C"
