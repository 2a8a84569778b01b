// hot_b200 internal state shared by the .cu translation units (not part of the C ABI).
//
// HBM layout (all device memory, SURVEY 8a rows a2/a3 re-designed for the GPU):
//  * particles: component-major SoA in SORTED order (key order of a5), so that a page group is one
//    contiguous run in every attribute array and CTA loads are fully coalesced.  `orig_id[s]` maps
//    a sorted slot back to the reference's original particle index (== particle_order).
//  * grid: the SPGrid virtual-memory trick (mmap of 4096^3 x 128 B, SPGrid_Allocator_Base.h:35-41) is
//    replaced by a compact page table: pages are stored in the reference's first-Set order
//    (SPGrid_Page_Map.h:61-70), `slot*E + e` addresses node e (in-page memory order, z fastest) of
//    page `slot`; the 128-byte AoS GridState record (MpmGrid.h:14-34, half of it dead padding) becomes
//    separate channel arrays m / v[3] / idx.
//  * solver vectors (dv, residual, x, b ...) are DOF-indexed "TVStack" arrays n_nodes x 3.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

struct hot_collider; // include/hot_b200.h
// mirror of hot_transport (include/hot_b200.h) so that this header does not need the C ABI header
struct hot_transport_fwd {
    void* user = nullptr;
    int (*all_reduce)(void* user, double* dev, long count, int op) = nullptr;
    int (*all_gather)(void* user, const void* dev_send, void* dev_recv, long bytes_per_rank) = nullptr;
    int (*neighbor_exchange)(void* user, int n_peers, const int* peers, double* const* send, double* const* recv, const long* count) = nullptr;
};

namespace hot {

// ---- SPGrid geometry for GridState<double,3> (128 B record): data_bits 7, block 2x4x4 -------------
// Lib/SPGrid/Core/SPGrid_Mask.h:31-52 evaluated for log2_struct=7, dim=3, log2_page=12.
struct Geo {
    static constexpr int data_bits = 7;
    static constexpr int block_bits = 5;
    static constexpr int xb = 1, yb = 2, zb = 2;
    static constexpr int BX = 1 << xb, BY = 1 << yb, BZ = 1 << zb;
    static constexpr int E = 1 << block_bits; // nodes per page
    static constexpr int index_bits = 32 - block_bits; // MpmSimulationBase.cpp:1071
    static constexpr uint64_t xmask = 0x9249249249249800ull;
    static constexpr uint64_t ymask = 0x4924924924924600ull;
    static constexpr uint64_t zmask = 0x2492492492492180ull;
    // touched node tile of the particles of one page: (B+2) per axis
    static constexpr int TX = BX + 2, TY = BY + 2, TZ = BZ + 2;
    static constexpr int TILE = TX * TY * TZ;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // grow-only; contents are NOT preserved
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        release();
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void swap(DevBuf& o)
    {
        T* tp = p; p = o.p; o.p = tp;
        size_t tc = cap; cap = o.cap; o.cap = tc;
    }
};

// particle attribute set, component-major SoA (attribute a, component c at a[c*stride + s])
struct ParticleSoA {
    DevBuf<double> X, V, M, C, F, Fn, vol, mu, lam, gradV;
    DevBuf<double> Jp; // SnowPlasticity::Jp (PlasticityApplier.h:57-80), 1 per particle
    DevBuf<int> orig_id;
    size_t stride = 0;
    cudaError_t reserve(size_t n);
    void swap(ParticleSoA& o);
};

// kernel classes of hot_get_timings (same order as HOT_K_* in include/hot_b200.h)
enum KernelClass { KC_SORT = 0, KC_P2G, KC_NUMBER, KC_G2P, KC_GATHER, KC_STRESS, KC_FORCE, KC_HESSIAN, KC_ASSEMBLE, KC_SPMV, KC_GS,
    KC_TRANSFER, KC_BLAS1, KC_COUNT };

// One multigrid level (level 0 = the assembled system of a15).  Block rows of the reference's fixed-width
// SquareMatrix (entryCol/entryVal, SquareMatrix.h:27-35) are re-laid for warp-per-row kernels:
//   slot s = (dx+2)*25 + (dy+2)*5 + (dz+2) addresses the neighbour at coord_i - (dx,dy,dz) on EVERY level
//   (ImplicitSolver.h:465-468; the Galerkin product keeps the 5^3 stencil, so coarse rows need no hash maps);
//   col[i*128 + s]            neighbour DOF id (i itself where the neighbour does not exist; 125..127 padding)
//   val[(i*9 + q)*128 + s]    entry q (column-major 3x3) of block s: a lane reads 32 consecutive slots of one q,
//                             so a warp streams its 9 KB row with fully coalesced 256-byte requests.
struct MGLevel {
    static constexpr int W = 128; // padded row width (125 slots)
    int n = 0;
    DevBuf<int> coord; // 3n
    DevBuf<uint64_t> key_sorted; // coordinate keys ascending ...
    DevBuf<int> id_sorted; // ... and the node id of each
    DevBuf<int> col;
    DevBuf<double> val;
    DevBuf<double> diag, dinv; // 9n each: D_i (column-major) and its inverse (block or entry-wise, -Ainv)
    // 8-colour 4^3-block Gauss-Seidel schedule (MultigridPreconditioner.h:582-605)
    DevBuf<int> gs_colrank; // rank of col[i*128+s], same layout as col
    // per-direction row stream of the sweeps (multigrid.cu: k_gs_stream): chunk offsets per sweep position, codes, values
    DevBuf<int> gs_pblock, gs_off[2], gs_code[2];
    DevBuf<double> gs_sval[2];
    int gs_chunks[2] = {0, 0};
    // block-inverse form (multigrid.cu: k_gx_stream / k_gx_inverse): per-direction stream [ext rows | inverse] per half block,
    // offsets by sweep position in direction order (the residual update reads the full forward row stream gs_*[0])
    DevBuf<int> gx_off[2];
    DevBuf<double> gx_data[2]; // chunk records: 9 x 32 values + 32 codes (2 432 bytes)
    int gx_chunks[2] = {0, 0};
    DevBuf<int> gs_seq, gs_rank, gs_block_start; // node ids in sweep order; rank of a node; block b = gs_seq[start[b]..start[b+1])
    double lMax = 1e2, lMin = 1e-8; // SquareMatrix::lMax / lMin (SquareMatrix.h:37), set by estimate2norm for the Chebyshev smoother
    int n_blocks = 0;
    int color_first_block[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    // transfer to the next coarser level: P rows (8 parents per fine node), R = P^T rows (<= 27 children, padded to 32)
    DevBuf<int> pcol, rcol;
    DevBuf<double> pw, rw;
    // V-cycle work vectors (MultigridOperator::residuals/initialResiduals/sols/dus/dAus/tmps)
    DevBuf<double> residual, initial_residual, sol, du, dAu, tmp;
};


struct Sim;
// CUDA-event timing of one kernel class on the handle's stream (enabled by hot_timing)
struct KTimers {
    bool on = false;
    struct Pending { int c; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    double ms[KC_COUNT] = {0};
    long long count[KC_COUNT] = {0};
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void collect()
    {
        for (auto& p : pending) {
            cudaEventSynchronize(p.b);
            float t = 0; cudaEventElapsedTime(&t, p.a, p.b);
            ms[p.c] += t; count[p.c]++;
            pool.push_back(p.a); pool.push_back(p.b);
        }
        pending.clear();
    }
    ~KTimers() { collect(); for (auto e : pool) cudaEventDestroy(e); }
};

struct Sim {
    KTimers timers;
    int device = 0;
    cudaStream_t stream = 0;
    std::string err;
    long long launches = 0;

    double dx = 1, apic_rpic_ratio = 1, cfl = 0.6;
    double dt = 0, gravity[3] = {0, 0, 0};

    // ---- particles
    long N = 0;
    ParticleSoA P, Palt; // current (sorted after a5) and the double buffer used by the reorder
    bool sorted = false;

    // ---- a5 outputs
    DevBuf<uint64_t> keys, keys_alt; // particle_sorter (sorted keys live in `keys` after the sort)
    DevBuf<int> perm, perm_alt;
    DevBuf<int> group_first; // n_groups + 1 entries (last = N)
    DevBuf<uint64_t> group_block; // block_offset (key >> 32)
    DevBuf<int> group_slot; // page slot of each group's page
    DevBuf<int> cell_start; // n_groups x (E+1): first sorted particle of every cell of the group's page
    long n_groups = 0;
    // page table
    long n_pages = 0;
    DevBuf<uint32_t> page_id; // page ids (offset >> 12) in first-Set order
    DevBuf<uint32_t> pid_sorted; // the same ids ascending ...
    DevBuf<int> slot_sorted; // ... and the slot of each
    DevBuf<int> nbr8; // n_pages x 8: slot of page + (i,j,k) block, -1 if absent
    // scratch
    DevBuf<unsigned char> cub_tmp;
    DevBuf<uint32_t> cand_key, cand_key_alt;
    DevBuf<int> cand_val, cand_val_alt, head_flag, scratch_i;
    DevBuf<int> dcount; // small device counters
    int* hcount = nullptr; // pinned mirror

    // ---- grid channels, n_pages*E entries
    DevBuf<double> g_m, g_v; // g_v: 3 channels, channel-major
    DevBuf<int> g_idx; // DOF id or -1
    size_t g_stride = 0;
    int num_nodes = 0;
    DevBuf<int> dof_slot; // inverse of g_idx
    DevBuf<int> tile_dof; // n_groups x Geo::TILE: DOF id of every node of a page group's (B+2)^3 tile (sort.cu: k_tile_dof)
    // work items of the persistent scatter (scatter_ws.cuh), built once per sort (sort.cu: build_scatter_items)
    DevBuf<unsigned char> ws_items; // WsItem records: item g = (first chunk of) page group g, extra chunks of oversized groups behind
    DevBuf<int> ws_order, ws_key, ws_key_alt, ws_idx; // this rank's items by decreasing particle count (+ sort scratch)
    DevBuf<int> ws_count; // [0] own items [1] extra chunks [2] work counter [3] finished teams [4] error flag
    long ws_own_items = 0;
    bool ws_ok = false; // false: a cell exceeds the item format (scatters fall back to the round-1 skeleton)
    bool ws_attr_set[4] = {false, false, false, false};
    int n_sm = 0;
    bool p2g_done = false;

    // ---- DOF vectors
    DevBuf<double> dv, vn, mass_matrix;

    int pf_dist = -1; // L2 prefetch distance of the per-group kernels in CTAs (-1: default = one wave of resident CTAs; 0: off; HOT_PF_DIST)
    DevBuf<int> flags; // g2p CFL flags
    bool flags_zeroed = false; // hot_p2g's zero pass already cleared flags / the numbering's done-counter
    // plasticity applied after G2P + evolveStrain (MpmSimulationBase.cpp:1039-1064): 0 none, 1 VonMisesFixedCorotated, 2 SnowPlasticity
    int plastic_model = 0;
    double plastic_param[5] = {0, 0, 0, 0, 0};

    // ---- force model state (force.cu); particle arrays are in sorted order and live for one time step
    bool project_pd = true; // CorotatedIsotropic::project (CorotatedIsotropic.h:60)
    int constitutive_model = 0; // 0 CorotatedIsotropic (the reference's model), 1 neo-Hookean (extension, hot_set_constitutive_model)
    bool strain_backed_up = false, state_valid = false, hessian_valid = false;
    DevBuf<double> f_stress; // vol P Fn^T, 9 rows
    DevBuf<double> f_U, f_V, f_sig; // SVD of the trial F
    DevBuf<double> f_T; // 9 rows: per-particle result of the Hessian gather (a13)
    DevBuf<double> f_H; // 45 rows: packed upper triangle of the contracted particle Hessian (see force.cu)
    DevBuf<double> group_psi; // per-group sum of vol*psi
    DevBuf<double> red_partial, red_out; // deterministic two-stage reductions
    double* h_red = nullptr; // pinned mirror of red_out
    // BC table (a8 output): CollisionNode{node_id,P,R,Rinv,shouldRotate}, CollisionObject.h:16-45
    int bc_mode = 0, n_bc = 0;
    DevBuf<int> bc_node, bc_slip;
    DevBuf<double> bc_P, bc_R, bc_Rinv;
    // device-resident collision objects (colliders.cu) and the dense-by-node scratch of the BC build
    DevBuf<unsigned char> colliders;
    int n_colliders = 0;
    DevBuf<int> col_coord, col_flag, col_pos, col_slip;
    DevBuf<double> col_P, col_R, col_dv;
    DevBuf<double> cn_tol; // per-node CN tolerance (a18)
    DevBuf<double> work[8]; // DOF-sized scratch vectors of the host-buffer entry points
    DevBuf<unsigned char> l2_flush; // 256 MiB written between the timed repetitions of hot_op_bench / hot_vcycle_bench

    // ---- assembled system + multigrid hierarchy (matrix.cu, multigrid.cu)
    std::vector<MGLevel*> levels; // levels[0] = assembled matrix
    bool matrix_built = false, mg_built = false, matrix_bcproject = true;
    DevBuf<uint64_t> mg_ckey, mg_ckey_sorted, mg_heads_key; // coarsening scratch (multigrid.cu: coarsen)
    DevBuf<int> mg_cpos, mg_cpos_sorted, mg_heads_pos, mg_order;
    DevBuf<int> bc_of; // node -> BC table row or -1
    DevBuf<int> asm_nbr27, asm_group_of_slot; // row-gather assembly (matrix.cu): 27 page neighbours per page, page slot -> page group
    DevBuf<double> asm_wrec; // 64 per particle: H~ (45), w[3][3], dw[3][3] / dx, pad
    DevBuf<double> diag_mf; // 9n: inverse diagonal blocks of the matrix-free operator (buildDiagonal)
    // HOTSettings (Projects/multigrid/Configurations.h:18-42)
    int mg_smoother = 5, mg_coarse = 2, mg_Ainv = 1, mg_levels = 3, mg_times = 1, mg_levelscale = 0;
    double mg_topomega = 0.1;
    double mg_cneps = 0.0; // HOTSettings::cneps of the running solve: tolerance of the coarsest-level smoother
    double dpdf_norm_max = -1.0; // computeCharacteristicNorm's function-static cache (MultigridSimulation.h:131-133): first solve after hot_set_particles
    // ---- one object over several GPUs (dist.cu): this rank's particles only; pages activated by >= 2 ranks are shared, scatter
    // results are summed over the sharers after every scatter, reductions count a shared node on its lowest-ranked sharer
    int rank = 0, world = 1;
    void* nccl_comm = nullptr; // ncclComm_t (hot_comm_init_nccl) ...
    hot_transport_fwd transport; // ... or the caller's callbacks (hot_set_partition)
    bool has_transport = false;
    long g0 = 0, g1 = 0, p0 = 0, p1 = 0; // page groups / sorted particles the particle kernels run on: all of this rank's
    std::vector<int> nbr_rank; // neighbours (ranks sharing pages with this one), ascending
    std::vector<long> nbr_off, nbr_cnt; // their segments of the exchange lists, in pages
    long x_total = 0; // exchange pages over all neighbours
    int n_sh = 0; // shared local pages
    int n_owned_nodes = 0;
    long global_nodes = 0; // nodes of the whole object = sum of the ranks' owned nodes
    bool own_valid = false;
    DevBuf<int> x_counts, x_slot, sh_slot, sh_ptr, sh_entry, sh_owned;
    DevBuf<int> sh_auth; // per shared page: exchange-list entry where the AUTHORITY's values arrive, -1 when this rank is the authority
    bool dot_plain = false; // vec_dot without the ownership mask / all-reduce (replicated coarse levels of a partitioned object)
    bool ghost_ring = false; // hot_set_ghost_ring: hold the 27-neighbourhood of the shared pages too (assembled-matrix / multigrid path)
    long n_base_pages = 0; // pages this rank's own particles activate (the first n_base_pages slots); the rest are ghost pages
    DevBuf<uint32_t> x_pids;
    DevBuf<double> x_send, x_recv, x_scalars;
    DevBuf<unsigned char> own_node; // 1: this rank counts the node in dots / norms
    // peer-memory transport of the shared-page exchange (dist.cu, NVLink P2P through cudaIpc; needs the NCCL communicator only for
    // the set-up): every rank owns a receive arena [flags: one u64 per source rank | parity 0 | parity 1]; the pack kernel of a
    // sharer stores its partial sums straight into the neighbours' arenas and raises its flag there, the unpack kernel waits on
    // the flags of this rank's neighbours - no collective kernel, no host work between the two launches
    int xp_state = 0; // 0 untried, 1 in use, -1 unavailable (no P2P between the ranks' devices, or HOT_XCHG=nccl)
    void* xp_mem = nullptr; // own arena (cudaMalloc, exported with cudaIpcGetMemHandle)
    size_t xp_cap = 0; // doubles per parity
    std::vector<void*> xp_peer; // peers' arenas as mapped here, by rank (nullptr: not a neighbour yet / self)
    std::vector<unsigned char> xp_handles; // world x 64 bytes: the handle every xp_peer entry was opened from
    std::vector<long> xp_peer_cap; // peers' capacities
    std::vector<long> xp_peer_off; // per neighbour j: where this rank's segment starts in that neighbour's arena, in pages
    long xp_gmax_pages = 0; // largest exchange list over the ranks (sizes every arena alike)
    unsigned long long xp_seq = 0; // exchanges done; the same on every rank (exchanges are collective)
    DevBuf<unsigned int> xp_done; // CTA counter of the pack kernel
    DevBuf<double> scat_tmp;
    DevBuf<double> sv[32]; // solver work vectors (solver.cu)
    bool dv0_valid = false;
    DevBuf<double> dv0_keep; // the accepted iterate of the last hot_backward_euler_step (hot_get_dv0)
    double vc_ms[10][4]; // per-level [smooth, restrict, prolongate, merge] of the last timed V-cycle
    int last_cg_iters = 0;
    ~Sim();
};

inline int model_flags(const Sim* s) { return (s->project_pd ? 1 : 0) | (s->constitutive_model << 1); } // kernels' `project` argument
int fail(Sim* s, const std::string& msg);
int cuda_fail(Sim* s, cudaError_t e, const char* what);

#define HOT_CUDA(call)                                                 \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return cuda_fail(s, e__, #call);       \
    } while (0)

// cudaFuncSetAttribute applies to the CURRENT device: remember per device what has been set (a process may open handles on
// several GPUs; the calling thread's current device must be the handle's, hot_create sets it)
#define HOT_FUNC_ATTR_ONCE(s, func, attr, value)                                    \
    do {                                                                           \
        static bool done__[64] = {false};                                          \
        const int d__ = (s)->device & 63;                                          \
        if (!done__[d__]) {                                                        \
            HOT_CUDA(cudaFuncSetAttribute(func, attr, value));                     \
            done__[d__] = true;                                                    \
        }                                                                          \
    } while (0)

#define HOT_LAUNCHED(s)                                                \
    do {                                                               \
        (s)->launches++;                                               \
        cudaError_t e__ = cudaGetLastError();                          \
        if (e__ != cudaSuccess) return cuda_fail(s, e__, "kernel launch"); \
    } while (0)

struct KTime { // RAII: times everything launched on s->stream in its scope under class c
    Sim* s; cudaEvent_t a = nullptr; int c;
    KTime(Sim* s_, int c_) : s(s_), c(c_)
    {
        if (s->timers.on) { a = s->timers.get(); cudaEventRecord(a, s->stream); }
    }
    ~KTime()
    {
        if (a) { cudaEvent_t b = s->timers.get(); cudaEventRecord(b, s->stream); s->timers.pending.push_back({c, a, b}); }
    }
};

// sort.cu
int sort_and_activate(Sim* s);
int number_nodes(Sim* s, bool flags_ready = false); // a7, after the P2G scatter (flags_ready: head_flag already holds the non-zero node flags)
int build_scatter_items(Sim* s); // work items of the persistent scatter, after the sort and the partition (dist_after_sort)
// transfer.cu
int p2g(Sim* s);
int g2p(Sim* s, double dt, int* flags);
int apply_plasticity(Sim* s);
// force.cu -- all pointers are DEVICE pointers to DOF vectors (n_nodes x 3)
int corotated_eval(Sim* s, long n, const double* F, double mu, double lambda, int project, const double* dF, double* psi, double* P,
    double* dP, double* dPdF, double* U, double* sigma, double* V); // device arrays
int strain_energy(Sim* s, double* e);
int backup_strain(Sim* s);
int restore_strain(Sim* s);
int set_bc(Sim* s, int mode, int n_bc, const int* node_id, const double* P, const double* R, const double* Rinv, const int* slip,
    const double* dv_bc);
int update_state(Sim* s, bool want_energy, double* energy);
int compute_residual(Sim* s, double* r);
int bc_project(Sim* s, double* v);
int bc_rotate(Sim* s, double* v, bool inverse);
int ensure_hessian(Sim* s);
int add_scaled_forces(Sim* s, double scale, double* f);
int add_scaled_force_differentials(Sim* s, double scale, const double* x, double* f);
int hessian_apply_mf(Sim* s, const double* x, double* b);
int eval_cn_tolerance(Sim* s, double eps, double dt, double* tol);
// dist.cu
int comm_unique_id(void* out128);
int comm_init_nccl(Sim* s, int rank, int world, const void* id128);
void comm_destroy(Sim* s);
void share_tables(int rank, int world, int max_pages, const int* counts, const uint32_t* all_pids, const int* slot_sorted, std::vector<int>& nbr_rank,
    std::vector<long>& nbr_off, std::vector<long>& nbr_cnt, std::vector<int>& x_slot, std::vector<int>& sh_slot, std::vector<int>& sh_ptr,
    std::vector<int>& sh_entry, std::vector<int>& sh_owned, std::vector<int>* sh_rank = nullptr);
void halo_pages(int rank, int world, int max_pages, const int* counts, const uint32_t* all_pids, std::vector<uint32_t>& ext); // ghost ring of a rank
void page_authority(int world, int max_pages, const int* counts, const uint32_t* all_pids, int n, const uint32_t* pids, int* auth);
int rebuild_neighbours(Sim* s); // sort.cu: nbr8 after the page table grew
int dist_takeover_shared(Sim* s, double* v, int comps); // every holder of a shared page takes the authority's values of a DOF array
int dist_exchange_rows(Sim* s, double* val, int stride, int c0, int nc); // sum of components [c0, c0 + nc) of a DOF array with `stride` per node;
int dist_after_sort(Sim* s); // shared-page tables of the current sort
int dist_p2g_exchange(Sim* s); // complete (m, mv) on the shared pages
int dist_after_numbering(Sim* s); // node ownership for reductions
int dist_allreduce_buffer(Sim* s, double* dev, long count, int op); // a few device scalars, in place
int dist_exchange_shared(Sim* s, double* v, int comps); // sum over the sharers on the shared nodes of a DOF array with `comps` per node
int dist_allreduce_host(Sim* s, double* host, int count, int op); // a few host scalars
int dist_all_gather_host(Sim* s, const void* mine, void* all, long bytes); // `bytes` host bytes per rank
int dist_all_gather_dev(Sim* s, const void* send, void* recv, long bytes); // device buffers
// colliders.cu
int set_colliders(Sim* s, int n, const ::hot_collider* objs);
int build_bc_from_colliders(Sim* s, int mode, int* n_bc);
// matrix.cu
int fill_id2coord(Sim* s, int* coord_dev);
int build_matrix(Sim* s, bool bcproject);
int build_diagonal_mf(Sim* s, int Ainv);
int apply_block_diag(Sim* s, int n, const double* D9, const double* x, double* y);
// multigrid.cu
int build_mg(Sim* s, int levels, int smoother, int coarse_solver, int Ainv, int times, int levelscale, double topomega);
int level_spmv(Sim* s, int level, const double* x, double* b);
int level_restrict(Sim* s, int level, const double* fine, double* coarse);
int level_prolong(Sim* s, int level, const double* coarse, double* fine);
int level_smooth(Sim* s, int level, int kind, double* u, double* r, int iterations, double tolerance);
int vcycle(Sim* s, const double* in, double* out, bool timed);
int build_coord_map(Sim* s, MGLevel& L);
int level_estimate_2norm(Sim* s, int level, double* lmax_lmin); // SquareMatrix::estimate2norm of a level, [lMax, lMin]
int columns_from_coords(Sim* s, MGLevel& L); // partitioned assembly: column ids of the non-zero blocks from the node coordinates
// vector ops (multigrid.cu)
int vec_axpy(Sim* s, long n, double a, const double* x, double* y); // y += a x
int vec_axpy_dev(Sim* s, long n, const double* num, const double* den, double sign, const double* x, double* y); // y += sign*(num/den) x
int vec_xpay_dev(Sim* s, long n, const double* x, const double* num, const double* den, double* y); // y = x + (num/den) y
int vec_copy(Sim* s, long n, const double* x, double* y);
int vec_zero(Sim* s, long n, double* y);
int vec_scale(Sim* s, long n, double a, double* y);
int vec_dot(Sim* s, long n, const double* a, const double* b, double* dev_out, double* host_out);

// ---- device helpers ----------------------------------------------------------------------------------
#ifdef __CUDACC__
// software pdep (the reference's non-HASWELL Bit_Spread, SPGrid_Utilities.h:77-343)
__host__ __device__ inline uint64_t bit_spread(uint32_t v, uint64_t mask)
{
    uint64_t r = 0;
    while (mask) {
        uint64_t low = mask & (~mask + 1);
        if (v & 1u) r |= low;
        v >>= 1;
        mask ^= low;
    }
    return r;
}
__host__ __device__ inline uint32_t bit_pack(uint64_t v, uint64_t mask)
{
    uint32_t r = 0;
    int o = 0;
    while (mask) {
        uint64_t low = mask & (~mask + 1);
        if (v & low) r |= 1u << o;
        ++o;
        mask ^= low;
    }
    return r;
}
// SPGrid_Mask.h:150-166
__host__ __device__ inline uint64_t linear_offset(int i, int j, int k)
{
    return bit_spread((uint32_t)i, Geo::xmask) | bit_spread((uint32_t)j, Geo::ymask) | bit_spread((uint32_t)k, Geo::zmask);
}
// SPGrid_Mask.h:237-245
__host__ __device__ inline uint64_t packed_add(uint64_t a, uint64_t b)
{
    const uint64_t w = ~(Geo::xmask | Geo::ymask | Geo::zmask);
    uint64_t rx = ((a | ~Geo::xmask) + (b & Geo::xmask)) & Geo::xmask;
    uint64_t ry = ((a | ~Geo::ymask) + (b & Geo::ymask)) & Geo::ymask;
    uint64_t rz = ((a | ~Geo::zmask) + (b & Geo::zmask)) & Geo::zmask;
    uint64_t rw = ((a | ~w) + (b & w)) & w;
    return rx | ry | rz | rw;
}
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: contiguous runs of the sorted particle rows are
// staged into shared memory by ONE thread without occupying registers or scoreboard slots; source, destination and size are
// multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_addr(b)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_test(unsigned long long* b, unsigned parity) // non-blocking
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_addr(b)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch of the sorted-particle run [first, end) of `nrows` consecutive SoA rows (row stride ps), one 128-byte line per
// thread and trip.  Used at long distance: a CTA requests the rows of the CTA that will occupy its slot next, so that the
// DRAM round trip of a work item's compulsory loads is paid while an earlier work item computes.
__device__ __forceinline__ void prefetch_rows_l2(const double* base, size_t ps, int nrows, int first, int end, int tid, int nthreads)
{
    if (end <= first) return;
    const int nl = ((end - first) * 8 + 127) / 128 + 1; // + 1: the run is not line-aligned
    for (int t = tid; t < nrows * nl; t += nthreads) {
        const int r = t / nl, l = t - r * nl;
        const char* p = (const char*)(base + (size_t)r * ps + first) + (size_t)l * 128;
        if (p < (const char*)(base + (size_t)r * ps + end)) prefetch_l2(p);
    }
}
// MathTools.h:15-25 + BSplines.h:16-20.  The product X*one_over_dx and the -0.5 are evaluated unfused
// (IEEE round-to-nearest each) so that particle->cell indices are bit-exact against the CPU evaluation.
__device__ inline int base_node_of(double X, double one_over_dx, double* x_index_space)
{
    double x = __dmul_rn(X, one_over_dx);
    double y = __dadd_rn(x, -0.5);
    int i = (int)y;
    *x_index_space = x;
    return i - (i > y);
}
#endif

} // namespace hot
