// a5 + a7 on the device: particle keys, u64 radix sort, page groups, page activation in the reference's
// first-Set order, neighbour-page table, DOF numbering.
//
// Reference: MpmSimulationBase::sortParticlesAndPolluteGrid (Lib/MPM/MpmSimulationBase.cpp:1066-1137),
// SPGrid_Page_Map::Set_Page / Get_Blocks (Lib/SPGrid/Core/SPGrid_Page_Map.h:61-96),
// MpmGrid::getNumNodes (Lib/MPM/MpmGrid.h:148-161).
//
// The reference's serial loops (group detection, Set_Page in particle order, getNumNodes scan) are
// re-expressed as data-parallel primitives that reproduce the serial ORDER exactly:
//   page list  = first occurrence order of the sequence  [P_g, P_g+n(0,0,0) .. P_g+n(1,1,1)]_g
//              = stable sort by page id -> run heads (min position) -> sort heads by position;
//   DOF ids    = exclusive scan of (m != 0) over (page list order x in-page element order).
// Radix sort / scan / select come from CUB (CUDA toolkit headers); everything else is ours.
#include "scatter_ws.cuh"
#include <cub/cub.cuh>

namespace hot {

int fail(Sim* s, const std::string& msg)
{
    s->err = msg;
    return -1;
}
int cuda_fail(Sim* s, cudaError_t e, const char* what)
{
    s->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return -2;
}

cudaError_t ParticleSoA::reserve(size_t n)
{
    size_t st = (n + 31) & ~(size_t)31; // 256-byte aligned component rows
    cudaError_t e;
    if ((e = X.reserve(3 * st)) != cudaSuccess) return e;
    if ((e = V.reserve(3 * st)) != cudaSuccess) return e;
    if ((e = M.reserve(st)) != cudaSuccess) return e;
    if ((e = C.reserve(9 * st)) != cudaSuccess) return e;
    if ((e = F.reserve(9 * st)) != cudaSuccess) return e;
    if ((e = vol.reserve(st)) != cudaSuccess) return e;
    if ((e = mu.reserve(st)) != cudaSuccess) return e;
    if ((e = lam.reserve(st)) != cudaSuccess) return e;
    if ((e = Jp.reserve(st)) != cudaSuccess) return e;
    if ((e = orig_id.reserve(st)) != cudaSuccess) return e;
    stride = st;
    return cudaSuccess;
}
void ParticleSoA::swap(ParticleSoA& o)
{
    X.swap(o.X); V.swap(o.V); M.swap(o.M); C.swap(o.C); F.swap(o.F);
    vol.swap(o.vol); mu.swap(o.mu); lam.swap(o.lam); orig_id.swap(o.orig_id); Jp.swap(o.Jp);
    size_t t = stride; stride = o.stride; o.stride = t;
}

namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

// MpmSimulationBase.cpp:1080-1085
__global__ void k_make_keys(long n, const double* __restrict__ X, size_t stride, const int* __restrict__ orig_id,
    double one_over_dx, uint64_t* __restrict__ keys, int* __restrict__ vals, int* __restrict__ bad)
{
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int b[3];
    bool oob = false;
    for (int d = 0; d < 3; ++d) {
        double X_d = X[d * stride + s], xi;
        b[d] = base_node_of(X_d, one_over_dx, &xi);
        if (!(xi == xi) || !(fabs(xi) < 1e9) || b[d] < 0 || b[d] + 2 >= 4096) oob = true;
    }
    if (oob) {
        atomicOr(bad, 1);
        b[0] = b[1] = b[2] = 0;
    }
    uint64_t off = linear_offset(b[0], b[1], b[2]);
    keys[s] = ((off >> Geo::data_bits) << Geo::index_bits) + (uint64_t)orig_id[s];
    vals[s] = (int)s;
}

// permute the persistent particle attributes into sorted order
__global__ void k_reorder(long n, const int* __restrict__ perm, size_t ss, size_t ds,
    const double* __restrict__ sX, const double* __restrict__ sV, const double* __restrict__ sM, const double* __restrict__ sC,
    const double* __restrict__ sF, const double* __restrict__ svol, const double* __restrict__ smu, const double* __restrict__ slam,
    const int* __restrict__ sid, const double* __restrict__ sJp,
    double* __restrict__ dX, double* __restrict__ dV, double* __restrict__ dM, double* __restrict__ dC, double* __restrict__ dF,
    double* __restrict__ dvol, double* __restrict__ dmu, double* __restrict__ dlam, int* __restrict__ did, double* __restrict__ dJp)
{
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    long p = perm[s];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        dX[c * ds + s] = sX[c * ss + p];
        dV[c * ds + s] = sV[c * ss + p];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        dC[c * ds + s] = sC[c * ss + p];
        dF[c * ds + s] = sF[c * ss + p];
    }
    dM[s] = sM[p];
    dvol[s] = svol[p];
    dmu[s] = smu[p];
    dlam[s] = slam[p];
    did[s] = sid[p];
    dJp[s] = sJp[p];
}

__global__ void k_group_flags(long n, const uint64_t* __restrict__ keys, int* __restrict__ flag)
{
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    flag[s] = (s == 0) || ((keys[s] >> 32) != (keys[s - 1] >> 32));
}

// per group: block_offset and the 8 page candidates of MpmSimulationBase.cpp:1104-1124.
// Position t = 8*g + (i*4 + j*2 + k) reproduces the serial Set_Page order (Set_Page(P) itself == q 0).
__global__ void k_group_pages(long n_groups, long n, const uint64_t* __restrict__ keys, int* __restrict__ group_first,
    uint64_t* __restrict__ group_block, uint32_t* __restrict__ cand_key, int* __restrict__ cand_val)
{
    long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    if (g == n_groups) {
        group_first[g] = (int)n;
        return;
    }
    uint64_t key = keys[group_first[g]];
    group_block[g] = key >> 32;
    uint64_t off = (key >> 32) << 12;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint64_t nb = linear_offset(Geo::BX * (q >> 2), Geo::BY * ((q >> 1) & 1), Geo::BZ * (q & 1));
        cand_key[g * 8 + q] = (uint32_t)(packed_add(off, nb) >> 12);
        cand_val[g * 8 + q] = (int)(g * 8 + q);
    }
}

__global__ void k_head_flags(long n, const uint32_t* __restrict__ k, int* __restrict__ flag)
{
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    flag[s] = (s == 0) || (k[s] != k[s - 1]);
}

__global__ void k_iota(long n, int* __restrict__ v)
{
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) v[s] = (int)s;
}

// order[slot] = index j into the ascending page id list
__global__ void k_finish_pages(long n_pages, const int* __restrict__ order, const uint32_t* __restrict__ pid_sorted,
    uint32_t* __restrict__ page_id, int* __restrict__ slot_sorted)
{
    long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_pages) return;
    int j = order[slot];
    page_id[slot] = pid_sorted[j];
    slot_sorted[j] = (int)slot;
}

__device__ inline int find_slot(uint32_t pid, long n_pages, const uint32_t* __restrict__ pid_sorted, const int* __restrict__ slot_sorted)
{
    long lo = 0, hi = n_pages;
    while (lo < hi) {
        long mid = (lo + hi) >> 1;
        if (pid_sorted[mid] < pid) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n_pages && pid_sorted[lo] == pid) ? slot_sorted[lo] : -1;
}

__global__ void k_neighbours(long n_pages, const uint32_t* __restrict__ page_id, const uint32_t* __restrict__ pid_sorted,
    const int* __restrict__ slot_sorted, int* __restrict__ nbr8)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pages * 8) return;
    long slot = t >> 3;
    int q = (int)(t & 7);
    uint64_t off = (uint64_t)page_id[slot] << 12;
    uint64_t nb = linear_offset(Geo::BX * (q >> 2), Geo::BY * ((q >> 1) & 1), Geo::BZ * (q & 1));
    uint64_t sum = packed_add(off, nb);
    // a neighbour that wrapped around the 4096^3 box is not a neighbour
    bool wrapped = (q & 4 && bit_pack(off, Geo::xmask) + Geo::BX >= 4096u) || (q & 2 && bit_pack(off, Geo::ymask) + Geo::BY >= 4096u)
        || (q & 1 && bit_pack(off, Geo::zmask) + Geo::BZ >= 4096u);
    nbr8[t] = wrapped ? -1 : find_slot((uint32_t)(sum >> 12), n_pages, pid_sorted, slot_sorted);
}

__global__ void k_group_slots(long n_groups, const uint64_t* __restrict__ group_block, long n_pages,
    const uint32_t* __restrict__ pid_sorted, const int* __restrict__ slot_sorted, int* __restrict__ group_slot)
{
    long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    group_slot[g] = find_slot((uint32_t)group_block[g], n_pages, pid_sorted, slot_sorted);
}

// first sorted particle of every in-page cell of a group (lower bound on key >> index_bits)
__global__ void k_cell_start(long n_groups, const int* __restrict__ group_first, const uint64_t* __restrict__ group_block,
    const uint64_t* __restrict__ keys, int* __restrict__ cell_start)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups * (Geo::E + 1)) return;
    long g = t / (Geo::E + 1);
    int c = (int)(t - g * (Geo::E + 1));
    int lo = group_first[g], hi = group_first[g + 1];
    uint64_t want = (group_block[g] << Geo::block_bits) + (uint64_t)c; // page id . cell
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((keys[mid] >> Geo::index_bits) < want) lo = mid + 1;
        else hi = mid;
    }
    cell_start[t] = lo;
}

// node carries mass (flag) -> one 32-bit mask per page (a page is one warp: Geo::E == 32 nodes); the last CTA to finish turns
// the per-page counts into exclusive offsets in page-list order and leaves the node count in *total
constexpr int NM_WARPS = 8;
constexpr int NM_ITEMS = 8;
static_assert(Geo::E == 32, "one warp per page");
__global__ void __launch_bounds__(32 * NM_WARPS) k_page_masks(int n_pages, const double* __restrict__ m, const int* __restrict__ flag,
    unsigned* __restrict__ mask, int* __restrict__ base, int* __restrict__ done, int* __restrict__ total, volatile int* total_host)
{
    const int page = blockIdx.x * NM_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    __shared__ int s_cnt[NM_WARPS];
    if (page < n_pages) {
        const size_t a = (size_t)page * Geo::E + lane;
        const unsigned b = __ballot_sync(0xffffffffu, flag ? flag[a] != 0 : m[a] != 0.0);
        if (lane == 0) {
            mask[page] = b;
            s_cnt[threadIdx.x >> 5] = __popc(b);
        }
    }
    else if (lane == 0) s_cnt[threadIdx.x >> 5] = 0;
    __shared__ int s_last;
    __shared__ int s_part[32 * NM_WARPS];
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < NM_WARPS; ++w) t += s_cnt[w];
        base[blockIdx.x] = t; // node count of this CTA's NM_WARPS pages; turned into an exclusive offset by the last CTA
        __threadfence();
        s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    n_pages = gridDim.x; // from here on: exclusive scan over the per-CTA counts (8x fewer than pages)
    // exclusive scan of base[0..n) by this CTA: super-tiles of T x NM_ITEMS counts, all loads of a super-tile in flight
    // together (coalesced), one warp-shuffle block scan per row of T
    constexpr int T = 32 * NM_WARPS;
    const int tid = threadIdx.x, warp = tid >> 5;
    __shared__ int s_carry;
    if (tid == 0) s_carry = 0;
    for (int t0 = 0; t0 < n_pages; t0 += T * NM_ITEMS) {
        int c[NM_ITEMS];
#pragma unroll
        for (int it = 0; it < NM_ITEMS; ++it) {
            const int p = t0 + it * T + tid;
            c[it] = p < n_pages ? ((volatile int*)base)[p] : 0;
        }
#pragma unroll
        for (int it = 0; it < NM_ITEMS; ++it) {
            int incl = c[it];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            __syncthreads(); // s_part / s_carry of the previous row consumed
            if (lane == 31) s_part[warp] = incl;
            __syncthreads();
            int off = s_carry;
            for (int w = 0; w < warp; ++w) off += s_part[w];
            const int p = t0 + it * T + tid;
            if (p < n_pages) base[p] = off + incl - c[it];
            __syncthreads();
            if (tid == T - 1) s_carry = off + incl;
        }
    }
    __syncthreads();
    if (tid == 0) {
        *total = s_carry;
        *total_host = s_carry; // mapped pinned word the host spins on (no D2H copy, no stream synchronize on the step path)
        __threadfence_system();
    }
}

// MpmGrid.h:148-161 (idx), MpmSimulationBase.cpp:523-531 (v /= m), :817-826 (mass_matrix)
// (mask bit = node carries mass; in a partitioned run a rank numbers nodes whose mass only other ranks hold: m == 0 there)
__global__ void __launch_bounds__(32 * NM_WARPS) k_number_and_normalise(int n_pages, size_t gs, const unsigned* __restrict__ mask,
    const int* __restrict__ base, double* __restrict__ m, double* __restrict__ v, int* __restrict__ idx, int* __restrict__ dof_slot,
    double* __restrict__ mass_matrix, double* __restrict__ vn)
{
    const int warp = threadIdx.x >> 5, page = blockIdx.x * NM_WARPS + warp, lane = threadIdx.x & 31;
    if (page >= n_pages) return;
    const size_t a = (size_t)page * Geo::E + lane;
    // offset of this page = exclusive offset of the CTA (k_page_masks) + the node counts of the CTA's earlier pages
    const int pw = blockIdx.x * NM_WARPS + lane;
    const unsigned mw = (lane < NM_WARPS && pw < n_pages) ? mask[pw] : 0u;
    int before = lane < warp ? __popc(mw) : 0;
#pragma unroll
    for (int o = NM_WARPS / 2; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    before = __shfl_sync(0xffffffffu, before, 0); // lanes 0..NM_WARPS-1 hold the sum
    const unsigned b = __shfl_sync(0xffffffffu, mw, warp);
    if (b >> lane & 1u) {
        const int id = base[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
        const double mm = m[a];
        idx[a] = id;
        dof_slot[id] = (int)a;
        mass_matrix[id] = mm;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double q = mm != 0.0 ? v[d * gs + a] / mm : 0.0;
            v[d * gs + a] = q;
            vn[3 * (size_t)id + d] = q;
        }
    }
    else {
        idx[a] = -1;
    }
}

// tile_dof[g][n]: DOF id of node n of the (B+2)^3 tile of page group g, -1 where the node carries no mass / its page is absent.
// Built once per step; the per-group gather kernels (G2P, updateState, Hessian gather) and the scatter flushes read it with one
// coalesced load instead of the dependent chain group_slot -> nbr8 -> g_idx.
__global__ void __launch_bounds__(160) k_tile_dof(const int* __restrict__ group_slot, const int* __restrict__ nbr8, const int* __restrict__ g_idx,
    int* __restrict__ tile_dof)
{
    __shared__ int s_nbr[8];
    const int g = blockIdx.x, n = threadIdx.x;
    if (n < 8) s_nbr[n] = nbr8[(size_t)group_slot[g] * 8 + n];
    __syncthreads();
    if (n >= Geo::TILE) return;
    const int tz = n % Geo::TZ, ty = (n / Geo::TZ) % Geo::TY, tx = n / (Geo::TZ * Geo::TY);
    const int q = ((tx >= Geo::BX) << 2) | ((ty >= Geo::BY) << 1) | (tz >= Geo::BZ);
    const int e = (((tx & (Geo::BX - 1)) << Geo::yb | (ty & (Geo::BY - 1))) << Geo::zb) | (tz & (Geo::BZ - 1));
    const int slot = s_nbr[q];
    tile_dof[(size_t)g * Geo::TILE + n] = slot < 0 ? -1 : g_idx[(size_t)slot * Geo::E + e];
}

// Work items of the persistent scatter (scatter_ws.cuh), one warp per page group, lane = in-page cell.  A group with more than
// WS_CAP particles is cut greedily into cell ranges (its first chunk keeps item index g, the others are appended behind the
// groups).  Per item: cells ranked by decreasing particle count, jagged-diagonal offsets jd[p] = sum_c min(n_c, p), cell starts
// relative to the run, the 8 neighbour page slots, and a sort key (WS_CAP - count for this rank's groups) for the dispatch order.
__global__ void k_scatter_items(int n_groups, int g0, int g1, int max_items, const int* __restrict__ cell_start, const int* __restrict__ group_slot,
    const int* __restrict__ nbr8, WsItem* __restrict__ items, int* __restrict__ keys, int* __restrict__ count)
{
    const int g = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const unsigned full = 0xffffffffu;
    const int* cs = cell_start + (size_t)g * (Geo::E + 1);
    const int start = cs[lane], n = cs[lane + 1] - start;
    if (__any_sync(full, n > WS_MAXC - 1)) { // a cell beyond the item format: the scatters keep the round-1 skeleton this step
        if (lane == 0) atomicOr(count + 4, 1);
        return;
    }
    int chunk = 0, run = 0, myc = 0;
    for (int c = 0; c < Geo::E; ++c) {
        const int nc = __shfl_sync(full, n, c);
        if (run + nc > WS_CAP && run > 0) { ++chunk; run = 0; }
        run += nc;
        if (c == lane) myc = chunk;
    }
    const int slot = group_slot[g];
    for (int k = 0; k <= chunk; ++k) {
        int idx = g;
        if (k > 0) {
            if (lane == 0) idx = n_groups + atomicAdd(count + 1, 1);
            idx = __shfl_sync(full, idx, 0);
            if (idx >= max_items) {
                if (lane == 0) atomicOr(count + 4, 2);
                return;
            }
        }
        const bool act = myc == k;
        const int na = act ? n : 0;
        int rank = 0, j0 = 0, j1 = 0, cnt = 0;
        for (int c = 0; c < Geo::E; ++c) {
            const int nc = __shfl_sync(full, na, c);
            rank += (nc > na) || (nc == na && c < lane);
            j0 += min(nc, lane);
            j1 += min(nc, lane + 32);
            cnt += nc;
        }
        const unsigned am = __ballot_sync(full, act);
        const int c0 = __ffs(am) - 1, c1 = 32 - __clz(am);
        const int first = __shfl_sync(full, start, c0);
        WsItem* it = items + idx;
        if (lane == 0) {
            it->first = first; it->count = cnt; it->g = g; it->cells = c0 | (c1 << 8);
            const bool own = g >= g0 && g < g1;
            keys[idx] = own ? WS_CAP - cnt : 0x7fff;
            if (own) atomicAdd(count, 1);
        }
        if (lane < 8) it->nbr[lane] = nbr8[(size_t)slot * 8 + lane];
        it->cnt[rank] = (unsigned short)na;
        it->order[rank] = (unsigned char)lane;
        it->rank[lane] = (unsigned char)rank;
        it->cs[lane] = (unsigned short)(act ? start - first : 0);
        it->jd[lane] = (unsigned short)j0;
        it->jd[lane + 32] = (unsigned short)j1;
    }
}

template <class F>
int with_tmp(Sim* s, F f)
{
    size_t bytes = 0;
    cudaError_t e = f((void*)nullptr, bytes);
    if (e != cudaSuccess) return cuda_fail(s, e, "cub size query");
    e = s->cub_tmp.reserve(bytes + 16);
    if (e != cudaSuccess) return cuda_fail(s, e, "cub temp alloc");
    e = f((void*)s->cub_tmp.p, bytes);
    if (e != cudaSuccess) return cuda_fail(s, e, "cub run");
    s->launches += 1; // CUB launches several kernels per call; counted once (conservative)
    return 0;
}

} // namespace

int sort_and_activate(Sim* s)
{
    const long n = s->N;
    cudaStream_t st = s->stream;
    if (n <= 0) return fail(s, "no particles");
    if (n >= (1l << Geo::index_bits)) return fail(s, "particle count must be < 2^index_bits (MpmSimulationBase.cpp:1072)");
    KTime timer(s, KC_SORT);
    HOT_CUDA(s->keys.reserve(n));
    HOT_CUDA(s->keys_alt.reserve(n));
    HOT_CUDA(s->perm.reserve(n));
    HOT_CUDA(s->perm_alt.reserve(n));
    HOT_CUDA(s->head_flag.reserve(n > 64 ? n : 64));
    HOT_CUDA(s->group_first.reserve(n + 1));
    HOT_CUDA(s->dcount.reserve(16));
    HOT_CUDA(s->Palt.reserve(n));
    if (!s->hcount) HOT_CUDA(cudaMallocHost((void**)&s->hcount, 32 * sizeof(int)));

    HOT_CUDA(cudaMemsetAsync(s->dcount.p, 0, 8 * sizeof(int), st));
    const double one_over_dx = 1.0 / s->dx;
    k_make_keys<<<nblk(n), TPB, 0, st>>>(n, s->P.X.p, s->P.stride, s->P.orig_id.p, one_over_dx, s->keys_alt.p, s->perm_alt.p, s->dcount.p);
    HOT_LAUNCHED(s);
    int rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, s->keys_alt.p, s->keys.p, s->perm_alt.p, s->perm.p, (int)n, 0, 64, st);
    });
    if (rc) return rc;
    k_reorder<<<nblk(n), TPB, 0, st>>>(n, s->perm.p, s->P.stride, s->Palt.stride, s->P.X.p, s->P.V.p, s->P.M.p, s->P.C.p, s->P.F.p,
        s->P.vol.p, s->P.mu.p, s->P.lam.p, s->P.orig_id.p, s->P.Jp.p, s->Palt.X.p, s->Palt.V.p, s->Palt.M.p, s->Palt.C.p, s->Palt.F.p,
        s->Palt.vol.p, s->Palt.mu.p, s->Palt.lam.p, s->Palt.orig_id.p, s->Palt.Jp.p);
    HOT_LAUNCHED(s);
    s->P.swap(s->Palt);

    // page groups: maximal runs of equal key>>32 (MpmSimulationBase.cpp:1089-1097)
    k_group_flags<<<nblk(n), TPB, 0, st>>>(n, s->keys.p, s->head_flag.p);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<int>(0), s->head_flag.p, s->group_first.p, s->dcount.p + 1, (int)n, st);
    });
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->hcount, s->dcount.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    if (s->hcount[0]) return fail(s, "particle outside the 4096^3 SPGrid box or non-finite position (MpmGrid.h:109,127)");
    const long G = s->hcount[1];
    s->n_groups = G;

    // page activation
    HOT_CUDA(s->group_block.reserve(G));
    HOT_CUDA(s->group_slot.reserve(G));
    HOT_CUDA(s->cand_key.reserve(8 * G));
    HOT_CUDA(s->cand_key_alt.reserve(8 * G));
    HOT_CUDA(s->cand_val.reserve(8 * G));
    HOT_CUDA(s->cand_val_alt.reserve(8 * G));
    HOT_CUDA(s->scratch_i.reserve(8 * G));
    HOT_CUDA(s->head_flag.reserve(8 * G));
    k_group_pages<<<nblk(G + 1), TPB, 0, st>>>(G, n, s->keys.p, s->group_first.p, s->group_block.p, s->cand_key.p, s->cand_val.p);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, s->cand_key.p, s->cand_key_alt.p, s->cand_val.p, s->cand_val_alt.p, (int)(8 * G), 0, 32, st);
    });
    if (rc) return rc;
    k_head_flags<<<nblk(8 * G), TPB, 0, st>>>(8 * G, s->cand_key_alt.p, s->head_flag.p);
    HOT_LAUNCHED(s);
    // run heads: ascending page ids -> cand_key, their first position -> cand_val
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, s->cand_key_alt.p, s->head_flag.p, s->cand_key.p, s->dcount.p + 2, (int)(8 * G), st);
    });
    if (rc) return rc;
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, s->cand_val_alt.p, s->head_flag.p, s->cand_val.p, s->dcount.p + 2, (int)(8 * G), st);
    });
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->hcount + 2, s->dcount.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    const long NP = s->hcount[2];
    s->n_pages = NP;
    HOT_CUDA(s->page_id.reserve(NP));
    HOT_CUDA(s->pid_sorted.reserve(NP));
    HOT_CUDA(s->slot_sorted.reserve(NP));
    HOT_CUDA(s->nbr8.reserve(8 * NP));
    HOT_CUDA(cudaMemcpyAsync(s->pid_sorted.p, s->cand_key.p, NP * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    // sort the heads by first position -> first-Set order
    k_iota<<<nblk(NP), TPB, 0, st>>>(NP, s->scratch_i.p);
    HOT_LAUNCHED(s);
    rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, (const uint32_t*)s->cand_val.p, s->cand_key_alt.p, s->scratch_i.p, s->cand_val_alt.p, (int)NP, 0, 32, st);
    });
    if (rc) return rc;
    k_finish_pages<<<nblk(NP), TPB, 0, st>>>(NP, s->cand_val_alt.p, s->pid_sorted.p, s->page_id.p, s->slot_sorted.p);
    HOT_LAUNCHED(s);
    k_neighbours<<<nblk(NP * 8), TPB, 0, st>>>(NP, s->page_id.p, s->pid_sorted.p, s->slot_sorted.p, s->nbr8.p);
    HOT_LAUNCHED(s);
    k_group_slots<<<nblk(G), TPB, 0, st>>>(G, s->group_block.p, NP, s->pid_sorted.p, s->slot_sorted.p, s->group_slot.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(s->cell_start.reserve((size_t)G * (Geo::E + 1)));
    k_cell_start<<<nblk(G * (Geo::E + 1)), TPB, 0, st>>>(G, s->group_first.p, s->group_block.p, s->keys.p, s->cell_start.p);
    HOT_LAUNCHED(s);

    // the partition may add ghost pages to the table (dist.cu): the grid arrays are sized after it
    s->num_nodes = 0;
    rc = dist_after_sort(s);
    if (rc) return rc;
    // zero the pages, idx = -1 (MpmSimulationBase.cpp:1128-1136)
    const size_t gn = (size_t)s->n_pages * Geo::E;
    HOT_CUDA(s->g_m.reserve(gn));
    HOT_CUDA(s->g_v.reserve(3 * gn));
    HOT_CUDA(s->g_idx.reserve(gn));
    HOT_CUDA(s->dof_slot.reserve(gn));
    s->g_stride = gn;
    HOT_CUDA(cudaMemsetAsync(s->g_m.p, 0, gn * sizeof(double), st));
    HOT_CUDA(cudaMemsetAsync(s->g_v.p, 0, 3 * gn * sizeof(double), st));
    HOT_CUDA(cudaMemsetAsync(s->g_idx.p, 0xff, gn * sizeof(int), st));
    s->sorted = true;
    s->p2g_done = false;
    return build_scatter_items(s);
}

int rebuild_neighbours(Sim* s)
{
    const long NP = s->n_pages;
    HOT_CUDA(s->nbr8.reserve(8 * (size_t)NP));
    k_neighbours<<<nblk(NP * 8), TPB, 0, s->stream>>>(NP, s->page_id.p, s->pid_sorted.p, s->slot_sorted.p, s->nbr8.p);
    HOT_LAUNCHED(s);
    return 0;
}

int build_scatter_items(Sim* s)
{
    cudaStream_t st = s->stream;
    const long G = s->n_groups, cap = 2 * G + 1024;
    s->ws_ok = false;
    s->ws_own_items = 0;
    if (!s->n_sm) HOT_CUDA(cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, s->device));
    HOT_CUDA(s->ws_items.reserve((size_t)cap * sizeof(WsItem)));
    HOT_CUDA(s->ws_key.reserve(cap));
    HOT_CUDA(s->ws_key_alt.reserve(cap));
    HOT_CUDA(s->ws_idx.reserve(cap));
    HOT_CUDA(s->ws_order.reserve(cap));
    HOT_CUDA(s->ws_count.reserve(8));
    HOT_CUDA(cudaMemsetAsync(s->ws_count.p, 0, 8 * sizeof(int), st));
    k_scatter_items<<<nblk(G * 32), TPB, 0, st>>>((int)G, (int)s->g0, (int)s->g1, (int)cap, s->cell_start.p, s->group_slot.p, s->nbr8.p,
        reinterpret_cast<WsItem*>(s->ws_items.p), s->ws_key.p, s->ws_count.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(cudaMemcpyAsync(s->hcount + 24, s->ws_count.p, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    if (s->hcount[28]) return 0; // keep the round-1 skeleton for this step
    const long total = G + s->hcount[25];
    k_iota<<<nblk(total), TPB, 0, st>>>(total, s->ws_idx.p);
    HOT_LAUNCHED(s);
    int rc = with_tmp(s, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, (const unsigned*)s->ws_key.p, (unsigned*)s->ws_key_alt.p, s->ws_idx.p, s->ws_order.p, (int)total, 0, 16, st);
    });
    if (rc) return rc;
    s->ws_own_items = s->hcount[24];
    s->ws_ok = true;
    return 0;
}

int number_nodes(Sim* s, bool flags_ready)
{
    cudaStream_t st = s->stream;
    const size_t gn = s->g_stride;
    const int NP = (int)s->n_pages;
    HOT_CUDA(s->head_flag.reserve(gn));
    HOT_CUDA(s->scratch_i.reserve(2 * (size_t)NP + 2));
    HOT_CUDA(s->mass_matrix.reserve(gn));
    HOT_CUDA(s->vn.reserve(3 * gn));
    HOT_CUDA(s->dv.reserve(3 * gn));
    // two launches: per-page node masks with the exclusive scan over the pages done by the last CTA to finish, then numbering
    // + normalisation (the serial running count of MpmGrid.h:148-161 over pages in list order x in-page element order)
    unsigned* mask = (unsigned*)s->scratch_i.p;
    int* base = s->scratch_i.p + NP;
    if (!s->flags_zeroed) HOT_CUDA(cudaMemsetAsync(s->dcount.p + 12, 0, sizeof(int), st)); // (hot_p2g's zero pass resets the done-counter too)
    volatile int* hn = s->hcount + 4; // pinned + mapped (UVA): written by the last CTA of k_page_masks
    *hn = -1;
    k_page_masks<<<(NP + NM_WARPS - 1) / NM_WARPS, 32 * NM_WARPS, 0, st>>>(NP, s->g_m.p, flags_ready ? s->head_flag.p : nullptr, mask, base,
        s->dcount.p + 12, s->dcount.p + 13, hn);
    HOT_LAUNCHED(s);
    k_number_and_normalise<<<(NP + NM_WARPS - 1) / NM_WARPS, 32 * NM_WARPS, 0, st>>>(NP, gn, mask, base, s->g_m.p, s->g_v.p, s->g_idx.p, s->dof_slot.p,
        s->mass_matrix.p, s->vn.p);
    HOT_LAUNCHED(s);
    HOT_CUDA(s->tile_dof.reserve((size_t)s->n_groups * Geo::TILE));
    if (s->n_groups > 0) {
        k_tile_dof<<<(unsigned)s->n_groups, 160, 0, st>>>(s->group_slot.p, s->nbr8.p, s->g_idx.p, s->tile_dof.p);
        HOT_LAUNCHED(s);
    }
    // the node count arrives while k_number_and_normalise is still running: the host goes on enqueueing without a stream sync
    for (long spin = 0; *hn < 0; ++spin) {
        if ((spin & 0xfff) == 0xfff) {
            const cudaError_t q = cudaStreamQuery(st);
            if (q != cudaErrorNotReady) {
                if (q != cudaSuccess) return cuda_fail(s, q, "number_nodes");
                if (*hn < 0) return fail(s, "number_nodes: the node count never arrived");
            }
        }
    }
    s->num_nodes = *hn;
    HOT_CUDA(cudaMemsetAsync(s->dv.p, 0, 3 * (size_t)s->num_nodes * sizeof(double), st));
    return 0;
}

} // namespace hot
