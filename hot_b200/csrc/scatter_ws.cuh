// Persistent, TMA-staged particle->grid scatter (a6 P2G, a12 force rasterisation, a13 Hessian-apply scatter).
//
// Round-1's one-CTA-per-page-group scatter (scatter.cuh) sat at 0.20 of the HBM roofline: ncu showed neither DRAM nor the fp64 pipe
// busy - a page group's serial chain  metadata -> particle rows (DRAM) -> prep -> barrier -> accumulate -> park -> combine  had
// nothing to overlap with, 20 % of the instructions were the combine's nested loops and the accumulate loop ran into 2-way
// shared-memory bank conflicts.  This kernel keeps the arithmetic (prep / accumulate / gather-combine, no atomics in the particle
// loop, one RED per node, channel and work item) and rebuilds everything around it:
//   * ONE persistent CTA per SM, four TEAMS of 96 threads (32 cells x 3 x-planes).  A team fetches work items (page groups, or
//     cell ranges of a page group with more than WS_CAP particles) from a global counter in order of decreasing particle count:
//     dynamic scheduling, short tail;
//   * an item's particle rows (16 contiguous runs of the sorted SoA rows for P2G), its 368-byte metadata block and, for DOF
//     targets, its tile -> DOF row are staged by TMA bulk copies (cp.async.bulk, SASS UBLKCP) onto the team's mbarrier.  The
//     next item's index and run bounds are fetched while the current item computes and the copies are issued the moment the
//     team's buffer is free, so a team never waits on a dependent chain of global loads;
//   * prep runs IN PLACE: every thread pulls its <= 4 particles' raw values out of the staged rows into registers, the team
//     syncs, and the 128-byte records go back into the same 49 KB in JAGGED-DIAGONAL order (slot p of every cell side by side,
//     cells ranked by decreasing particle count, 16-byte units XOR-swizzled).  The (cell, plane) threads of the accumulate loop
//     then read CONSECUTIVE records: bank-conflict free, and lanes of a warp run loops of similar length;
//   * team-wide named barriers (bar.sync id, 96) instead of __syncthreads: the four teams of a CTA are at different phases and
//     fill each other's stalls;
//   * the combine is straight-line code: warp w of a team owns 8 (tx, ty) columns of the (B+2)^3 tile, lane = (tz, channel); every
//     shared-memory offset is a compile-time constant plus one per-lane term (no index arithmetic, no loops).
// The sort (sort.cu: build_scatter_items) prepares the per-item metadata once per time step; every scatter of the step uses it.
#pragma once
#include "scatter.cuh"
#include <type_traits>

namespace hot {

constexpr int WS_CAP = 384;            // particles per work item
constexpr int WS_ROWCAP = WS_CAP + 2;  // doubles per staged raw row: bulk copies start at an even index and move an even count
constexpr int WS_MAXC = 64;            // jagged-diagonal offsets held per item: a cell holds <= 63 particles
constexpr int WS_HALF = 3 * Geo::E;    // 96 threads (cell, x-plane); a team has two such halves, one per channel pair
constexpr int WS_TEAMS = 4, WS_TT = 2 * WS_HALF, WS_THREADS = WS_TEAMS * WS_TT;
constexpr int WS_BUF = 16 * WS_ROWCAP * 8;   // 49 408 B: 16 raw rows, then WS_CAP records of 128 B, then the parked sums
constexpr int WS_NCHP = 4;                   // parked channels per (node, thread), padded
struct WsItem {
    int first, count, g, cells;    // sorted-particle run [first, first + count), page group, cell range c0 | c1 << 8
    int nbr[8];                    // page slots of the 8 pages the group's tile touches (-1: absent)
    unsigned short jd[WS_MAXC];    // jd[p] = records in slots < p = sum over cells of min(count, p)
    unsigned short cnt[Geo::E];    // particle count by rank
    unsigned short cs[Geo::E];     // first particle of every cell, relative to `first`
    unsigned char rank[Geo::E];    // cell -> rank (decreasing count)
    unsigned char order[Geo::E];   // rank -> cell
};
static_assert(sizeof(WsItem) == 368 && sizeof(WsItem) % 16 == 0, "bulk-copied as one block");
constexpr int WS_TEAM_BYTES = ((WS_BUF + (int)sizeof(WsItem) + Geo::TILE * 4 + 16 + 127) / 128) * 128;
constexpr int WS_SMEM = WS_TEAMS * WS_TEAM_BYTES;
static_assert(WS_BUF % 128 == 0 && 9 * WS_HALF * WS_NCHP * 8 <= WS_BUF && WS_CAP * 128 <= WS_BUF, "buffer holds rows, records and parked sums");
static_assert(Geo::TILE * 4 % 16 == 0, "tile_dof rows are bulk-copied");
static_assert(WS_CAP <= 2 * WS_TT, "prep: two particles per thread");

__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(WS_TT) : "memory"); }

// sums of the tile columns (tx = 0..3, TY) for lane (tz, ch): park index =
//   ((j*3+k)*96 + ((cx*4+cy)*4 + cz)*3 + i)*4 + ch,  cx = tx - i, cy = TY - j, cz = tz - k
template <int TY>
__device__ __forceinline__ void ws_combine(const double* __restrict__ park, int lane_off, bool pk0, bool pk1, bool pk2, double (&out)[4])
{
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) {
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int cx = tx - i;
            if (cx < 0 || cx >= Geo::BX) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int cy = TY - j;
                if (cy < 0 || cy >= Geo::BY) continue;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int off = ((j * 3 + k) * WS_HALF + (cx * Geo::BY + cy) * Geo::BZ * 3 + i) * WS_NCHP - k * 3 * WS_NCHP;
                    const bool pk = k == 0 ? pk0 : (k == 1 ? pk1 : pk2);
                    if (pk) sum += park[off + lane_off];
                }
            }
        }
        out[tx] = sum;
    }
}

// Policy interface (ws form):
//   NCH, DOF, Args, flush1 as in scatter.cuh and
//   static constexpr int ROWS (raw SoA rows staged per particle), UNITS (16-byte units of a record, <= 8)
//   __device__ static const double* row(const Args&, int r)                       base pointer of raw row r (sorted order)
//   __device__ static int prep_ws(const Args&, const double (&raw)[ROWS], double (&rec)[2 * UNITS])   -> in-page cell of the particle
//   __device__ static void accumulate_rec(const Args&, const double (&rec)[2 * UNITS], int i, double di, double (&acc)[9][NCH])
// A team's two halves run the same accumulate loop; each parks only ITS channel pair, so the compiler drops the other pair's
// arithmetic: 18 accumulators per thread instead of 36 (80 registers, 24 warps per SM), half the serial work per thread.
// DBG: clock64 stamps of every warp's phases summed into dbg[0..8] (HOT_WS_DEBUG; profiling aid, never the bench path)
template <class Policy, bool DBG = false>
__global__ void __launch_bounds__(WS_THREADS, 1) k_scatter_ws(typename Policy::Args args, const WsItem* __restrict__ items,
    const int* __restrict__ order, const int* __restrict__ n_work, int* __restrict__ ctr, const int* __restrict__ tile_dof,
    unsigned long long* dbg = nullptr)
{
    long long t_[9], acc_t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#define WS_STAMP(k) do { if (DBG) t_[k] = clock64(); } while (0)
    constexpr int NCH = Policy::NCH, ROWS = Policy::ROWS, UNITS = Policy::UNITS;
    static_assert(Geo::BX == 2 && Geo::BY == 4 && Geo::BZ == 4, "the combine is written for the 2x4x4 page");
    static_assert(NCH >= 2 && NCH <= 4, "two channel pairs");
    extern __shared__ __align__(128) unsigned char ws_smem[];
    const int tid = threadIdx.x, team = tid / WS_TT, tt = tid - team * WS_TT;
    unsigned char* const base = ws_smem + (size_t)team * WS_TEAM_BYTES;
    double* const rows = reinterpret_cast<double*>(base);
    WsItem* const item = reinterpret_cast<WsItem*>(base + WS_BUF);
    int* const s_dof = reinterpret_cast<int*>(base + WS_BUF + sizeof(WsItem));
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(base + WS_BUF + sizeof(WsItem) + Geo::TILE * 4);
    volatile int* const s_it = reinterpret_cast<volatile int*>(bar + 1);
    const unsigned buf_s = smem_addr(base);
    const bool elected = tt == 0;
    const int nw = *n_work;

    // next work item of this team: index and run bounds (held in the elected thread's registers until the buffer is free)
    int n_it = -1, n_first = 0, n_count = 0, n_g = 0;
    auto fetch = [&]() {
        const int w = atomicAdd(ctr, 1);
        n_it = w < nw ? order[w] : -1;
        if (n_it >= 0) {
            const int4 h = *reinterpret_cast<const int4*>(items + n_it);
            n_first = h.x; n_count = h.y; n_g = h.z;
        }
    };
    auto issue = [&]() {
        const int start = n_first & ~1, cnt = (n_first + n_count - start + 1) & ~1;
        fence_proxy_async();
        mbar_expect_tx(bar, (unsigned)(ROWS * cnt * 8 + (int)sizeof(WsItem) + (Policy::DOF ? Geo::TILE * 4 : 0)));
#pragma unroll
        for (int r = 0; r < ROWS; ++r) bulk_load(rows + r * WS_ROWCAP, Policy::row(args, r) + start, (unsigned)cnt * 8u, bar);
        bulk_load(item, items + n_it, (unsigned)sizeof(WsItem), bar);
        if (Policy::DOF) bulk_load(s_dof, tile_dof + (size_t)n_g * Geo::TILE, Geo::TILE * 4u, bar);
    };
    if (elected) {
        mbar_init(bar, 1);
        fetch();
        if (n_it >= 0) issue();
        *s_it = n_it;
    }
    const int half = tt / WS_HALF, tl = tt - half * WS_HALF; // warp-uniform: a half is three warps
    const int r_ = tl / 3, i_ = tl - 3 * r_;
    const double di = (double)i_;
    const int wteam = tt >> 5, lane = tt & 31;
    // combine role: warp -> tile row ty (2, 3, 1 | 0, 5, 4: the heavy rows first), lane -> (tz, ch)
    const int cty = wteam == 0 ? 2 : (wteam == 1 ? 3 : (wteam == 2 ? 1 : (wteam == 3 ? 0 : (wteam == 4 ? 5 : 4))));
    const int ctz = lane / WS_NCHP, cch = lane - ctz * WS_NCHP;
    const bool c_on = ctz < Geo::TZ && cch < NCH;
    const int lane_off = 3 * WS_NCHP * ctz + cch;
    const bool pk0 = c_on && ctz < Geo::BZ, pk1 = c_on && ctz >= 1 && ctz - 1 < Geo::BZ, pk2 = c_on && ctz >= 2 && ctz - 2 < Geo::BZ;
    unsigned phase = 0;
    for (;;) {
        WS_STAMP(0);
        team_sync(team); // s_it visible; every thread is done with the previous item's metadata
        const int it = *s_it;
        if (it < 0) break;
        if (elected) fetch(); // the item after this one: latency hidden behind this item's work
        mbar_wait(bar, phase);
        phase ^= 1;
        WS_STAMP(1);
        const int count = item->count, off0 = item->first & 1;
        // ---- prep, in place: raw values -> registers | team barrier | records -> jagged-diagonal slots of the same buffer
        double raw[2][ROWS];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = tt + k * WS_TT;
            if (q < count) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) raw[k][r] = rows[r * WS_ROWCAP + off0 + q];
            }
        }
        team_sync(team);
        WS_STAMP(2);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = tt + k * WS_TT;
            if (q < count) {
                double rec[2 * UNITS];
                const int c = Policy::prep_ws(args, raw[k], rec);
                const int p = q - (int)item->cs[c];
                const int ri = (int)item->jd[p] + (int)item->rank[c];
                const unsigned a = buf_s + ((unsigned)ri << 7), x = (unsigned)((ri + p) & 7) << 4;
#pragma unroll
                for (int u = 0; u < UNITS; ++u)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + (((unsigned)u << 4) ^ x)), "d"(rec[2 * u]), "d"(rec[2 * u + 1]) : "memory");
            }
        }
        team_sync(team);
        WS_STAMP(3);
        // ---- accumulate: thread (rank r_, plane i_) walks the slots of its cell; lanes read consecutive records.  Then park its
        // channel pair: [node (j,k)][cell * 3 + plane][channel]
        const int n = item->cnt[r_], cell = item->order[r_];
        double* const park = rows + (size_t)(cell * 3 + i_) * WS_NCHP + 2 * half;
        auto accumulate_and_park = [&](auto H) {
            constexpr int C0 = 2 * decltype(H)::value, C1 = C0 + 1 < NCH ? C0 + 1 : C0;
            double acc[9][NCH];
#pragma unroll
            for (int a = 0; a < 9; ++a)
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) acc[a][ch] = 0.0;
            for (int p = 0; p < n; ++p) {
                const int ri = (int)item->jd[p] + r_;
                const unsigned a = buf_s + ((unsigned)ri << 7), x = (unsigned)((ri + p) & 7) << 4;
                double rec[2 * UNITS];
#pragma unroll
                for (int u = 0; u < UNITS; ++u)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rec[2 * u]), "=d"(rec[2 * u + 1]) : "r"(a + (((unsigned)u << 4) ^ x)));
                Policy::accumulate_rec(args, rec, i_, di, acc);
            }
            WS_STAMP(4);
            team_sync(team); // records dead
            WS_STAMP(5);
#pragma unroll
            for (int a = 0; a < 9; ++a) sts2(park + (size_t)a * WS_HALF * WS_NCHP, acc[a][C0], C1 != C0 ? acc[a][C1] : 0.0);
        };
        if (half == 0) accumulate_and_park(std::integral_constant<int, 0>());
        else accumulate_and_park(std::integral_constant<int, 1>());
        team_sync(team);
        WS_STAMP(6);
        // ---- combine: straight-line code per tile row
        double out[4];
        switch (cty) {
        case 0: ws_combine<0>(rows, lane_off, pk0, pk1, pk2, out); break;
        case 1: ws_combine<1>(rows, lane_off, pk0, pk1, pk2, out); break;
        case 2: ws_combine<2>(rows, lane_off, pk0, pk1, pk2, out); break;
        case 3: ws_combine<3>(rows, lane_off, pk0, pk1, pk2, out); break;
        case 4: ws_combine<4>(rows, lane_off, pk0, pk1, pk2, out); break;
        default: ws_combine<5>(rows, lane_off, pk0, pk1, pk2, out); break;
        }
        // flush targets (read before the next item's copies may overwrite the metadata)
        long tgt[4];
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) {
            long t = -1;
            if (c_on) {
                if (Policy::DOF) t = s_dof[(tx * Geo::TY + cty) * Geo::TZ + ctz];
                else {
                    const int q = ((tx >= Geo::BX) << 2) | ((cty >= Geo::BY) << 1) | (ctz >= Geo::BZ);
                    const int e = (((tx & (Geo::BX - 1)) << Geo::yb | (cty & (Geo::BY - 1))) << Geo::zb) | (ctz & (Geo::BZ - 1));
                    const int slot = item->nbr[q];
                    t = slot < 0 ? -1 : (long)slot * Geo::E + e;
                }
            }
            tgt[tx] = t;
        }
        WS_STAMP(7);
        team_sync(team); // parked sums and metadata consumed: the buffer is free
        if (elected) {
            if (n_it >= 0) issue();
            *s_it = n_it;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (tgt[o] >= 0 && out[o] != 0.0) Policy::flush1(args, tgt[o], cch, out[o]);
        WS_STAMP(8);
        if (DBG) {
#pragma unroll
            for (int k = 1; k < 9; ++k) acc_t[k] += t_[k] - t_[k - 1];
            acc_t[0] += 1;
        }
    }
    if (DBG && lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) atomicAdd(dbg + k, (unsigned long long)acc_t[k]);
    }
#undef WS_STAMP
    // the last team to leave re-arms the work counter for the next launch
    if (elected) {
        __threadfence();
        if (atomicAdd(ctr + 1, 1) == (int)gridDim.x * WS_TEAMS - 1) {
            ctr[0] = 0;
            ctr[1] = 0;
        }
    }
}

template <class Policy>
int launch_scatter_ws(Sim* s, const typename Policy::Args& a)
{
    if (s->ws_own_items <= 0) return 0;
    if (!s->ws_attr_set[Policy::WS_ID]) {
        HOT_CUDA(cudaFuncSetAttribute(k_scatter_ws<Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
        s->ws_attr_set[Policy::WS_ID] = true;
    }
    const long teams = (s->ws_own_items + WS_TEAMS - 1) / WS_TEAMS;
    const unsigned grid = (unsigned)(teams < s->n_sm ? teams : s->n_sm);
    static int dbg_runs = getenv("HOT_WS_DEBUG") ? atoi(getenv("HOT_WS_DEBUG")) : 0;
    if (dbg_runs > 0) {
        --dbg_runs;
        HOT_CUDA(cudaFuncSetAttribute(k_scatter_ws<Policy, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
        unsigned long long* d = nullptr;
        unsigned long long h[9] = {0};
        HOT_CUDA(cudaMalloc((void**)&d, sizeof h));
        HOT_CUDA(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s->stream);
        k_scatter_ws<Policy, true><<<grid, WS_THREADS, WS_SMEM, s->stream>>>(a, reinterpret_cast<const WsItem*>(s->ws_items.p), s->ws_order.p,
            s->ws_count.p, s->ws_count.p + 2, Policy::DOF ? s->tile_dof.p : nullptr, d);
        cudaEventRecord(e1, s->stream);
        HOT_CUDA(cudaStreamSynchronize(s->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        HOT_CUDA(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
        cudaFree(d);
        const double n = h[0] ? (double)h[0] : 1.0; // warp-items
        fprintf(stderr, "[ws dbg] NCH %d: %.1f us, grid %u, %llu warp-items; mean cycles per item: top-sync+tma wait %.0f | raw loads+sync %.0f | records+sync %.0f | "
                        "accumulate %.0f | sync %.0f | park+sync %.0f | combine+targets %.0f | sync+issue+flush %.0f | total %.0f\n",
            Policy::NCH, ms * 1e3, grid, h[0], h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n,
            (h[1] + h[2] + h[3] + h[4] + h[5] + h[6] + h[7] + h[8]) / n);
        s->launches++;
        return 0;
    }
    k_scatter_ws<Policy><<<grid, WS_THREADS, WS_SMEM, s->stream>>>(a, reinterpret_cast<const WsItem*>(s->ws_items.p), s->ws_order.p,
        s->ws_count.p, s->ws_count.p + 2, Policy::DOF ? s->tile_dof.p : nullptr);
    HOT_LAUNCHED(s);
    return 0;
}

// which skeleton a scatter uses: the policy's measured default of scatter.cuh (CTA per page group, register-staged rows),
// HOT_SCATTER = plane | column to force one of its two forms, HOT_SCATTER = ws for the persistent TMA form above.
// Measured on C2 (P2G, B200): plane form 81 us, persistent TMA form 101 us, a CTA-per-item variant of it with register-staged
// rows 85 us (profiles/r2_scatter_experiments.md) - the persistent form stays selectable for A/B runs.
template <class Policy>
int launch_scatter_best(Sim* s, const typename Policy::Args& a)
{
    static const bool ws = getenv("HOT_SCATTER") != nullptr && getenv("HOT_SCATTER")[0] == 'w';
    if (s->ws_ok && ws) return launch_scatter_ws<Policy>(s, a);
    return launch_scatter<Policy>(s, a);
}

} // namespace hot
