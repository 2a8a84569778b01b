// Row (e): one object over the GPUs of a box, one process per GPU, particles partitioned, grid pages shared at the seams.
//
// The reference has no distributed path; this is the design of SURVEY 8e:
//  * every rank holds ITS particles only (whatever the caller gives it: bench.py cuts the object into slabs) and is, locally, a
//    complete single-GPU object: own sort, own page list (its particles' pages and their +1 neighbours,
//    MpmSimulationBase.cpp:1104-1124), own DOF numbering.  Nothing is replicated, nothing scales with the whole object;
//  * a page that two or more ranks activate is SHARED: after every particle->grid scatter (P2G mass / momentum, forces, Hessian
//    products, CN tolerances, block diagonals) the partial sums on shared pages are exchanged between the sharing ranks only
//    (grouped ncclSend / ncclRecv over NVLink, one message per neighbour) and added in ascending rank order on every sharer, so
//    all sharers hold bit-identical totals.  Gathers (G2P, updateState, Hessian gather) then need no communication;
//  * the shared-page tables are built once per sort: all-gather of the ranks' ascending page-id lists, intersections on the host
//    (a few thousand ids), per-neighbour slot lists + a CSR of (sharer, offset) per shared page on the device;
//  * DOF vectors are local (own + shared nodes, shared ones replicated and consistent); dots / norms / energies count a shared
//    node on its lowest-ranked sharer only (own_node mask) and all-reduce 1-3 scalars;
//  * particles stay with their rank while they move (correct for any distribution: sharing is recomputed every sort);
//    re-balancing by migration is a performance matter, not one of correctness.
// Transport: NCCL inside the library (hot_comm_init_nccl; the symbols are resolved with dlopen so the library loads without NCCL),
// or the caller's callbacks (hot_set_partition with a hot_transport: the one-GPU tests run several ranks on one device through gloo).
#include "../../include/hot_b200.h"
#include "sim.h"
#include "reduce.cuh"
#include <algorithm>
#include <cub/cub.cuh>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

namespace hot {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static const char* load_nccl()
{
    if (g_nccl.lib) return nullptr;
    // the process may already carry an NCCL (torch bundles one): reuse it, else load the system library
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 not found (dlopen)";
#define HOT_NCCL_SYM(name)                                                         \
    *reinterpret_cast<void**>(&g_nccl.name) = dlsym(lib, "nccl" #name);            \
    if (!g_nccl.name) return "NCCL symbol nccl" #name " missing";
    HOT_NCCL_SYM(GetUniqueId) HOT_NCCL_SYM(CommInitRank) HOT_NCCL_SYM(CommDestroy) HOT_NCCL_SYM(AllReduce) HOT_NCCL_SYM(AllGather)
    HOT_NCCL_SYM(Send) HOT_NCCL_SYM(Recv) HOT_NCCL_SYM(GroupStart) HOT_NCCL_SYM(GroupEnd) HOT_NCCL_SYM(GetErrorString)
#undef HOT_NCCL_SYM
    g_nccl.lib = lib;
    return nullptr;
}

#define HOT_NCCL(call)                                                                                        \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != ncclSuccess) return fail(s, std::string("NCCL error: ") + g_nccl.GetErrorString(r__) + " in " #call); \
    } while (0)

int comm_unique_id(void* out128)
{
    if (const char* e = load_nccl()) {
        fprintf(stderr, "hot_comm_unique_id: %s\n", e);
        return -1;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "the C ABI passes the id as 128 bytes");
    return g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(out128)) == ncclSuccess ? 0 : -1;
}
int comm_init_nccl(Sim* s, int rank, int world, const void* id128)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(s, "hot_comm_init_nccl: need 0 <= rank < world");
    if (const char* e = load_nccl()) return fail(s, std::string("hot_comm_init_nccl: ") + e);
    if (s->nccl_comm) {
        g_nccl.CommDestroy((ncclComm_t)s->nccl_comm);
        s->nccl_comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c;
    HOT_CUDA(cudaSetDevice(s->device));
    HOT_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
    s->nccl_comm = c;
    s->rank = rank;
    s->world = world;
    s->has_transport = false;
    s->sorted = false;
    s->p2g_done = false;
    return 0;
}
static void xp_release(Sim* s)
{
    for (void*& m : s->xp_peer) {
        if (m) cudaIpcCloseMemHandle(m);
        m = nullptr;
    }
    if (s->xp_mem) cudaFree(s->xp_mem);
    s->xp_mem = nullptr;
    s->xp_cap = 0;
    s->xp_state = 0;
    s->xp_seq = 0;
}
void comm_destroy(Sim* s)
{
    if (s->xp_mem || !s->xp_peer.empty()) {
        cudaStreamSynchronize(s->stream);
        xp_release(s);
        s->xp_peer.clear();
    }
    if (s->nccl_comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)s->nccl_comm);
    s->nccl_comm = nullptr;
}

namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

// ---- transport: NCCL or the caller's callbacks; every operation is enqueued on (or ordered with) the handle's stream ---------
int comm_all_reduce(Sim* s, double* dev, long count, int op)
{
    if (s->nccl_comm) {
        HOT_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, op ? ncclMax : ncclSum, (ncclComm_t)s->nccl_comm, s->stream));
        return 0;
    }
    if (!s->has_transport) return fail(s, "partitioned run without a transport (hot_comm_init_nccl or hot_set_partition)");
    return s->transport.all_reduce(s->transport.user, dev, count, op) ? fail(s, "transport all_reduce failed") : 0;
}
int comm_all_gather(Sim* s, const void* send, void* recv, long bytes_per_rank)
{
    if (s->nccl_comm) {
        HOT_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, ncclChar, (ncclComm_t)s->nccl_comm, s->stream));
        return 0;
    }
    if (!s->has_transport) return fail(s, "partitioned run without a transport (hot_comm_init_nccl or hot_set_partition)");
    return s->transport.all_gather(s->transport.user, send, recv, bytes_per_rank) ? fail(s, "transport all_gather failed") : 0;
}
// send[j] -> peer j, recv[j] <- peer j, `count[j]` doubles each way (the shared-page lists are symmetric)
int comm_exchange(Sim* s, int n_peers, const int* peers, double* const* send, double* const* recv, const long* count)
{
    if (n_peers <= 0) return 0;
    if (s->nccl_comm) {
        HOT_NCCL(g_nccl.GroupStart());
        for (int j = 0; j < n_peers; ++j) {
            HOT_NCCL(g_nccl.Send(send[j], (size_t)count[j], ncclDouble, peers[j], (ncclComm_t)s->nccl_comm, s->stream));
            HOT_NCCL(g_nccl.Recv(recv[j], (size_t)count[j], ncclDouble, peers[j], (ncclComm_t)s->nccl_comm, s->stream));
        }
        HOT_NCCL(g_nccl.GroupEnd());
        return 0;
    }
    if (!s->has_transport) return fail(s, "partitioned run without a transport (hot_comm_init_nccl or hot_set_partition)");
    return s->transport.neighbor_exchange(s->transport.user, n_peers, peers, send, recv, count) ? fail(s, "transport neighbor_exchange failed") : 0;
}

// ---- pack / unpack of shared pages ------------------------------------------------------------------------------------------
// exchange buffer layout: [exchange page k][component c][element e]; a component is a grid channel (m, v0, v1, v2) or a
// component of a DOF vector looked up through g_idx (inactive node: 0)
struct PageSource {
    int comps;
    const double *g_m, *g_v; // grid channels (P2G) ...
    size_t gs;
    const int* g_idx; // ... or a DOF array: components [c0, c0 + comps) of `stride` per node
    const double* v;
    int stride, c0;
    __device__ double load(int slot, int c, int e) const
    {
        const size_t a = (size_t)slot * Geo::E + e;
        if (!g_idx) return c == 0 ? g_m[a] : g_v[(size_t)(c - 1) * gs + a];
        const int id = g_idx[a];
        return id >= 0 ? v[(size_t)id * stride + c0 + c] : 0.0;
    }
};
__global__ void k_pack_shared(long n_x, PageSource src, const int* __restrict__ x_slot, double* __restrict__ buf)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_x * src.comps * Geo::E) return;
    const int e = (int)(t % Geo::E), c = (int)((t / Geo::E) % src.comps);
    const long k = t / ((long)Geo::E * src.comps);
    buf[t] = src.load(x_slot[k], c, e);
}

// ---- peer-memory transport (NVLink P2P) ---------------------------------------------------------------------------------------
// Arena of a rank: XP_FLAGS u64 flags (flag[r] = number of the last exchange whose data from rank r has landed), then two data
// areas of xp_cap doubles (exchange n uses area n & 1: a sender can only be one exchange ahead of a receiver, because its pack of
// exchange n+2 comes after its unpack of n+1, which waited for the receiver's pack of n+1, which came after the receiver's unpack of n).
constexpr int XP_MAX_NBR = 16;
constexpr size_t XP_FLAGS = 64; // u64 words, >= world
struct XpPeers {
    int n;
    int rank[XP_MAX_NBR]; // neighbour ranks
    long off[XP_MAX_NBR], cnt[XP_MAX_NBR]; // their segments of this rank's exchange list, in pages
    long peer_off[XP_MAX_NBR]; // where this rank's segment starts over there, in pages
    double* peer_data[XP_MAX_NBR]; // the neighbour's data area of this exchange
    unsigned long long* peer_flag[XP_MAX_NBR]; // this rank's flag over there
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// pack + send: every value goes straight into the sharer's arena; the last CTA to finish raises this rank's flag at every neighbour
__global__ void k_pack_peer(long n_x, PageSource src, const int* __restrict__ x_slot, XpPeers pr, unsigned long long seq, unsigned int* done)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_x * src.comps * Geo::E) {
        const int e = (int)(t % Geo::E), c = (int)((t / Geo::E) % src.comps);
        const long k = t / ((long)Geo::E * src.comps);
        int j = 0;
        while (j + 1 < pr.n && k >= pr.off[j + 1]) ++j;
        const long at = ((pr.peer_off[j] + (k - pr.off[j])) * src.comps + c) * Geo::E + e;
        pr.peer_data[j][at] = src.load(x_slot[k], c, e);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < pr.n) st_release_sys(pr.peer_flag[threadIdx.x], seq);
        if (threadIdx.x == 0) *done = 0;
    }
}
struct XpWait {
    int n;
    int rank[XP_MAX_NBR];
    const unsigned long long* flags; // this rank's own flags; nullptr: nothing to wait for (NCCL / callback transport)
    unsigned long long seq;
};
// every sharer adds the partial sums in ascending rank order (entry -1 = this rank's own partial): identical totals everywhere
// sh_auth == nullptr: every sharer adds the partial sums in ascending rank order (entry -1 = this rank's own partial): identical
// totals everywhere.  sh_auth != nullptr: take-over - the holders of a page replace their values by the authority's.
__global__ void k_unpack_shared(int n_sh, int comps, const int* __restrict__ sh_slot, const int* __restrict__ sh_ptr, const int* __restrict__ sh_entry,
    const int* __restrict__ sh_auth, const double* recv, double* __restrict__ g_m, double* __restrict__ g_v, size_t gs, const int* __restrict__ g_idx,
    double* __restrict__ v, int stride, int c0, XpWait w)
{
    if (w.flags) { // peer-memory transport: the neighbours' partial sums of this exchange have landed when their flags say so
        if (threadIdx.x < w.n)
            while (ld_acquire_sys(w.flags + w.rank[threadIdx.x]) < w.seq) __nanosleep(64);
        __syncthreads();
    }
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n_sh * comps * Geo::E) return;
    const int e = (int)(t % Geo::E), c = (int)((t / Geo::E) % comps), p = (int)(t / ((long)Geo::E * comps));
    const size_t a = (size_t)sh_slot[p] * Geo::E + e;
    double* dst;
    if (!g_idx) dst = c == 0 ? g_m + a : g_v + (size_t)(c - 1) * gs + a;
    else {
        const int id = g_idx[a];
        if (id < 0) return;
        dst = v + (size_t)id * stride + c0 + c;
    }
    if (sh_auth) {
        const int en = sh_auth[p];
        if (en >= 0) *dst = __ldcg(recv + ((size_t)en * comps + c) * Geo::E + e);
        return;
    }
    const double mine = *dst;
    double sum = 0.0;
    bool first = true;
    for (int q = sh_ptr[p]; q < sh_ptr[p + 1]; ++q) {
        const int en = sh_entry[q];
        const double x = en < 0 ? mine : __ldcg(recv + ((size_t)en * comps + c) * Geo::E + e);
        sum = first ? x : sum + x;
        first = false;
    }
    *dst = sum;
}
// a node counts in dots / norms on the lowest-ranked sharer of its page
__global__ void k_own_nodes(int n_sh, const int* __restrict__ sh_slot, const int* __restrict__ sh_owned, const int* __restrict__ g_idx,
    unsigned char* __restrict__ own)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n_sh * Geo::E) return;
    const int p = (int)(t / Geo::E), e = (int)(t % Geo::E);
    const int id = g_idx[(size_t)sh_slot[p] * Geo::E + e];
    if (id >= 0) own[id] = (unsigned char)sh_owned[p];
}

struct CountOwnF {
    const unsigned char* own;
    __device__ void operator()(long i, double (&acc)[1]) const { acc[0] += own[i] ? 1.0 : 0.0; }
};

// ---- peer transport set-up (host; collective over the ranks, called from dist_after_sort and when an exchange needs more room) ----
// all-gather of `bytes` host bytes per rank through the active transport
int host_all_gather(Sim* s, const void* mine, void* all, long bytes)
{
    const int W = s->world;
    HOT_CUDA(s->x_pids.reserve((size_t)(W + 1) * ((bytes + 3) / 4)));
    unsigned char* d = (unsigned char*)s->x_pids.p;
    HOT_CUDA(cudaMemcpyAsync(d + (size_t)W * bytes, mine, bytes, cudaMemcpyHostToDevice, s->stream));
    int rc = comm_all_gather(s, d + (size_t)W * bytes, d, bytes);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(all, d, (size_t)W * bytes, cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
// (re)allocate this rank's arena for `doubles` per data area and (re)open every rank's arena; all ranks take the same decisions
int xp_open(Sim* s, size_t doubles)
{
    const int W = s->world;
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    for (void*& m : s->xp_peer) { // nobody may keep a mapping of an arena that is about to be freed
        if (m) cudaIpcCloseMemHandle(m);
        m = nullptr;
    }
    s->xp_peer.assign(W, nullptr);
    char dummy = 0, dall[64];
    if (W > 64) return fail(s, "peer transport: at most 64 ranks");
    int rc = host_all_gather(s, &dummy, dall, 1); // barrier: every rank has closed its mappings
    if (rc) return rc;
    if (s->xp_mem) cudaFree(s->xp_mem);
    s->xp_mem = nullptr;
    const size_t want = doubles + doubles / 4 + 4096;
    bool ok = cudaMalloc(&s->xp_mem, XP_FLAGS * sizeof(unsigned long long) + 2 * want * sizeof(double)) == cudaSuccess;
    cudaIpcMemHandle_t hd;
    std::memset(&hd, 0, sizeof hd);
    // (device-wide synchronise: the flags must be zero before any peer can reach them; the handle's stream may be non-blocking)
    if (ok) ok = cudaMemset(s->xp_mem, 0, XP_FLAGS * sizeof(unsigned long long)) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess
            && cudaIpcGetMemHandle(&hd, s->xp_mem) == cudaSuccess;
    if (!ok) cudaGetLastError();
    s->xp_cap = ok ? want : 0;
    s->xp_seq = 0; // fresh flags everywhere
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    struct Rec { unsigned char handle[64]; long cap; long ok; };
    Rec me;
    std::memcpy(me.handle, &hd, 64);
    me.cap = (long)s->xp_cap;
    me.ok = ok ? 1 : 0;
    std::vector<Rec> all(W);
    rc = host_all_gather(s, &me, all.data(), sizeof(Rec));
    if (rc) return rc;
    s->xp_peer_cap.assign(W, 0);
    long good = 1;
    for (int r = 0; r < W; ++r) good &= all[r].ok;
    for (int r = 0; r < W && good; ++r) {
        s->xp_peer_cap[r] = all[r].cap;
        if (r == s->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all[r].handle, 64);
        if (cudaIpcOpenMemHandle(&s->xp_peer[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            s->xp_peer[r] = nullptr;
            good = 0;
        }
    }
    // usable only when it works on every rank
    char mine_ok = good ? 1 : 0, oks[64];
    rc = host_all_gather(s, &mine_ok, oks, 1);
    if (rc) return rc;
    for (int r = 0; r < W; ++r) good &= oks[r];
    if (!good) {
        for (void*& m : s->xp_peer) {
            if (m) cudaIpcCloseMemHandle(m);
            m = nullptr;
        }
        if (s->xp_mem) cudaFree(s->xp_mem);
        s->xp_mem = nullptr;
        s->xp_cap = 0;
        s->xp_state = -1;
        return 0;
    }
    s->xp_state = 1;
    return 0;
}
// after the share tables of a sort: every rank learns where its segments start in its neighbours' arenas
int xp_after_sort(Sim* s)
{
    if (!s->nccl_comm || s->xp_state < 0) return 0;
    if (s->xp_state == 0) {
        const char* e = getenv("HOT_XCHG");
        if (e && !strcmp(e, "nccl")) {
            s->xp_state = -1;
            return 0;
        }
    }
    const int W = s->world;
    // row r of the table: off_for[src] = first page of src's segment in r's exchange list (-1: not a neighbour); last entry: r's total
    std::vector<long> mine(W + 1, -1), all((size_t)W * (W + 1));
    for (size_t j = 0; j < s->nbr_rank.size(); ++j) mine[s->nbr_rank[j]] = s->nbr_off[j];
    mine[W] = s->x_total;
    int rc = host_all_gather(s, mine.data(), all.data(), (long)((W + 1) * sizeof(long)));
    if (rc) return rc;
    s->xp_gmax_pages = 0;
    for (int r = 0; r < W; ++r) s->xp_gmax_pages = std::max(s->xp_gmax_pages, all[(size_t)r * (W + 1) + W]);
    s->xp_peer_off.assign(s->nbr_rank.size(), 0);
    for (size_t j = 0; j < s->nbr_rank.size(); ++j) s->xp_peer_off[j] = all[(size_t)s->nbr_rank[j] * (W + 1) + s->rank];
    if (s->xp_state == 0 || s->xp_gmax_pages * 4 * Geo::E > (long)s->xp_cap) { // room for the 4-channel P2G exchange at least
        rc = xp_open(s, (size_t)std::max(1L, s->xp_gmax_pages) * 4 * Geo::E);
        if (rc) return rc;
    }
    HOT_CUDA(s->xp_done.reserve(4));
    HOT_CUDA(cudaMemsetAsync(s->xp_done.p, 0, 4 * sizeof(unsigned int), s->stream));
    return 0;
}

int exchange_pages(Sim* s, const PageSource& src, double* g_m, double* g_v, double* v, bool takeover = false)
{
    if (s->world <= 1) return 0;
    cudaStream_t st = s->stream;
    const int n_nbr = (int)s->nbr_rank.size();
    bool peer = s->xp_state == 1 && n_nbr <= XP_MAX_NBR;
    if (s->xp_state == 1 && s->xp_gmax_pages * src.comps * Geo::E > (long)s->xp_cap) { // the same on every rank: grow together
        int rc = xp_open(s, (size_t)s->xp_gmax_pages * src.comps * Geo::E);
        if (rc) return rc;
        peer = s->xp_state == 1 && n_nbr <= XP_MAX_NBR;
    }
    if (s->xp_state == 1) ++s->xp_seq; // counted on every rank, whether or not it has neighbours
    if (s->x_total <= 0) return 0;
    const long cnt = s->x_total * src.comps * Geo::E;
    XpWait w;
    w.n = 0;
    w.flags = nullptr;
    w.seq = 0;
    const double* recv = nullptr;
    if (peer) {
        XpPeers pr;
        pr.n = n_nbr;
        w.n = n_nbr;
        w.seq = s->xp_seq;
        w.flags = (const unsigned long long*)s->xp_mem;
        const size_t area = (s->xp_seq & 1) * s->xp_cap;
        for (int j = 0; j < n_nbr; ++j) {
            const int r = s->nbr_rank[j];
            pr.rank[j] = w.rank[j] = r;
            pr.off[j] = s->nbr_off[j];
            pr.cnt[j] = s->nbr_cnt[j];
            pr.peer_off[j] = s->xp_peer_off[j];
            unsigned long long* base = (unsigned long long*)s->xp_peer[r];
            pr.peer_flag[j] = base + s->rank;
            pr.peer_data[j] = (double*)(base + XP_FLAGS) + (s->xp_seq & 1) * (size_t)s->xp_peer_cap[r];
        }
        k_pack_peer<<<nblk(cnt), TPB, 0, st>>>(s->x_total, src, s->x_slot.p, pr, s->xp_seq, s->xp_done.p);
        HOT_LAUNCHED(s);
        recv = (const double*)((unsigned long long*)s->xp_mem + XP_FLAGS) + area;
    }
    else {
        HOT_CUDA(s->x_send.reserve(cnt));
        HOT_CUDA(s->x_recv.reserve(cnt));
        k_pack_shared<<<nblk(cnt), TPB, 0, st>>>(s->x_total, src, s->x_slot.p, s->x_send.p);
        HOT_LAUNCHED(s);
        std::vector<double*> sp(n_nbr), rp(n_nbr);
        std::vector<long> cn(n_nbr);
        for (int j = 0; j < n_nbr; ++j) {
            sp[j] = s->x_send.p + s->nbr_off[j] * src.comps * Geo::E;
            rp[j] = s->x_recv.p + s->nbr_off[j] * src.comps * Geo::E;
            cn[j] = s->nbr_cnt[j] * src.comps * Geo::E;
        }
        int rc = comm_exchange(s, n_nbr, s->nbr_rank.data(), sp.data(), rp.data(), cn.data());
        if (rc) return rc;
        recv = s->x_recv.p;
    }
    const long un = (long)s->n_sh * src.comps * Geo::E;
    k_unpack_shared<<<nblk(un), TPB, 0, st>>>(s->n_sh, src.comps, s->sh_slot.p, s->sh_ptr.p, s->sh_entry.p, takeover ? s->sh_auth.p : nullptr, recv, g_m, g_v,
        src.gs, src.g_idx, v, src.stride, src.c0, w);
    HOT_LAUNCHED(s);
    return 0;
}

} // namespace

// Which of this rank's pages other ranks activate too, from everyone's ascending page-id list (all_pids: world rows of max_pages,
// counts[r] valid entries each; slot_sorted: local slot of this rank's i-th smallest page id).
//   per neighbour r (ascending): the shared pages in ascending id order - the same order on both sides - as local slots, appended
//                                to x_slot (segment nbr_off / nbr_cnt): the exchange list;
//   per shared local page:       the contributions in ascending rank order, sh_entry = index into the concatenated exchange list
//                                (where that neighbour's partial sum arrives) or -1 for this rank's own partial; sh_owned = 1 when
//                                this rank is the lowest sharer (it counts the page's nodes in reductions).
void share_tables(int rank, int world, int max_pages, const int* counts, const uint32_t* all_pids, const int* slot_sorted, std::vector<int>& nbr_rank,
    std::vector<long>& nbr_off, std::vector<long>& nbr_cnt, std::vector<int>& x_slot, std::vector<int>& sh_slot, std::vector<int>& sh_ptr,
    std::vector<int>& sh_entry, std::vector<int>& sh_owned, std::vector<int>* sh_rank)
{
    nbr_rank.clear(); nbr_off.clear(); nbr_cnt.clear(); x_slot.clear(); sh_slot.clear(); sh_entry.clear(); sh_owned.clear();
    if (sh_rank) sh_rank->clear();
    sh_ptr.assign(1, 0);
    const uint32_t* my = all_pids + (size_t)rank * max_pages;
    const int nm = counts[rank];
    std::vector<std::vector<std::pair<int, int>>> sharers(nm); // per local page (ascending-id index): (rank, index in the exchange list)
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        const uint32_t* other = all_pids + (size_t)r * max_pages;
        const int no = counts[r];
        const long off = (long)x_slot.size();
        int i = 0, j = 0;
        while (i < nm && j < no) {
            if (my[i] < other[j]) ++i;
            else if (my[i] > other[j]) ++j;
            else {
                sharers[i].emplace_back(r, (int)x_slot.size());
                x_slot.push_back(slot_sorted[i]);
                ++i; ++j;
            }
        }
        if ((long)x_slot.size() > off) {
            nbr_rank.push_back(r);
            nbr_off.push_back(off);
            nbr_cnt.push_back((long)x_slot.size() - off);
        }
    }
    for (int i = 0; i < nm; ++i) {
        if (sharers[i].empty()) continue;
        sh_slot.push_back(slot_sorted[i]);
        bool self_done = false;
        for (const auto& pr : sharers[i]) { // neighbours were visited in ascending rank order
            if (!self_done && pr.first > rank) {
                sh_entry.push_back(-1);
                if (sh_rank) sh_rank->push_back(rank);
                self_done = true;
            }
            sh_entry.push_back(pr.second);
            if (sh_rank) sh_rank->push_back(pr.first);
        }
        if (!self_done) {
            sh_entry.push_back(-1);
            if (sh_rank) sh_rank->push_back(rank);
        }
        sh_ptr.push_back((int)sh_entry.size());
        sh_owned.push_back(sharers[i].front().first > rank ? 1 : 0);
    }
}


// ---- ghost ring for the assembled-matrix / multigrid path (host logic, exported for the CPU tests) ----------------------------
// Block rows (a15) reach two nodes around a node: a row on a page that another rank's particles also touch has columns on pages
// this rank never activated.  With the ghost ring on, a rank additionally holds every page of the 27-neighbourhood of its SHARED
// pages that some rank activates (no particles of its own there: its partial sums are zero, the page is simply one more shared
// page, summed and numbered like the others).  Rows on the rank's own ("base") pages then have all their columns locally.
// Authority of a page = the rank whose values count (reductions, Galerkin sums, the values every other holder takes over after a
// Gauss-Seidel colour phase or an SpMV): the lowest rank that has the page AND the other page of its 4^3 Gauss-Seidel block (the
// x-neighbour with which it forms the block) among its base pages; if nobody has both, the halves do not couple and the lowest
// base holder of the page takes it.
static inline void page_origin(uint32_t pid, int& x, int& y, int& z)
{
    const uint64_t off = (uint64_t)pid << 12;
    x = (int)bit_pack(off, Geo::xmask); y = (int)bit_pack(off, Geo::ymask); z = (int)bit_pack(off, Geo::zmask);
}
static inline uint32_t page_at(int x, int y, int z) { return (uint32_t)(linear_offset(x, y, z) >> 12); }

struct PageHolders { // page id -> bit mask of the ranks that activate it (world <= 64)
    std::vector<uint32_t> pid; // ascending, unique
    std::vector<uint64_t> mask;
    uint64_t find(uint32_t p) const
    {
        auto it = std::lower_bound(pid.begin(), pid.end(), p);
        return (it != pid.end() && *it == p) ? mask[it - pid.begin()] : 0ull;
    }
};
static PageHolders page_holders(int world, int max_pages, const int* counts, const uint32_t* all_pids)
{
    std::vector<std::pair<uint32_t, int>> pr;
    for (int r = 0; r < world; ++r)
        for (int i = 0; i < counts[r]; ++i) pr.emplace_back(all_pids[(size_t)r * max_pages + i], r);
    std::sort(pr.begin(), pr.end());
    PageHolders h;
    for (const auto& q : pr) {
        if (h.pid.empty() || h.pid.back() != q.first) { h.pid.push_back(q.first); h.mask.push_back(0ull); }
        h.mask.back() |= 1ull << q.second;
    }
    return h;
}
// ghost pages of `rank`, ascending
void halo_pages(int rank, int world, int max_pages, const int* counts, const uint32_t* all_pids, std::vector<uint32_t>& ext)
{
    ext.clear();
    const PageHolders h = page_holders(world, max_pages, counts, all_pids);
    const uint32_t* my = all_pids + (size_t)rank * max_pages;
    const uint64_t me = 1ull << rank;
    for (int i = 0; i < counts[rank]; ++i) {
        if ((h.find(my[i]) & ~me) == 0) continue; // not shared
        int x, y, z;
        page_origin(my[i], x, y, z);
        for (int q = 0; q < 27; ++q) {
            const int nx = x + Geo::BX * (q / 9 - 1), ny = y + Geo::BY * ((q / 3) % 3 - 1), nz = z + Geo::BZ * (q % 3 - 1);
            if (nx < 0 || ny < 0 || nz < 0 || nx >= 4096 || ny >= 4096 || nz >= 4096) continue;
            const uint32_t np = page_at(nx, ny, nz);
            const uint64_t m = h.find(np);
            if (m != 0 && !(m & me)) ext.push_back(np);
        }
    }
    std::sort(ext.begin(), ext.end());
    ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
}
// authority rank of page `pid` from the BASE lists
static int page_authority_of(const PageHolders& h, uint32_t pid)
{
    const uint64_t m = h.find(pid);
    if (!m) return -1;
    int x, y, z;
    page_origin(pid, x, y, z);
    const uint64_t mp = h.find(page_at(x ^ Geo::BX, y, z)); // the other half of the 4^3 block
    const uint64_t both = m & mp;
    const uint64_t pick = both ? both : m;
    return __builtin_ctzll(pick);
}
void page_authority(int world, int max_pages, const int* counts, const uint32_t* all_pids, int n, const uint32_t* pids, int* auth)
{
    const PageHolders h = page_holders(world, max_pages, counts, all_pids);
    for (int i = 0; i < n; ++i) auth[i] = page_authority_of(h, pids[i]);
}

// shared-page tables for the pages of the current sort
int dist_after_sort(Sim* s)
{
    s->g0 = 0; s->g1 = s->n_groups; s->p0 = 0; s->p1 = s->N;
    s->x_total = 0; s->n_sh = 0; s->n_owned_nodes = 0;
    s->nbr_rank.clear(); s->nbr_off.clear(); s->nbr_cnt.clear();
    s->own_valid = false;
    if (s->world <= 1) return 0;
    cudaStream_t st = s->stream;
    const int W = s->world;
    // 1. everyone's page count, then everyone's ascending page-id list (padded to the largest)
    HOT_CUDA(s->x_counts.reserve(2 * (size_t)W));
    int np_local = (int)s->n_pages;
    HOT_CUDA(cudaMemcpyAsync(s->x_counts.p + W, &np_local, sizeof(int), cudaMemcpyHostToDevice, st));
    int rc = comm_all_gather(s, s->x_counts.p + W, s->x_counts.p, sizeof(int));
    if (rc) return rc;
    std::vector<int> counts(W);
    HOT_CUDA(cudaMemcpyAsync(counts.data(), s->x_counts.p, W * sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    const int maxp = *std::max_element(counts.begin(), counts.end());
    HOT_CUDA(s->x_pids.reserve((size_t)(W + 1) * maxp));
    uint32_t* mine = s->x_pids.p + (size_t)W * maxp;
    HOT_CUDA(cudaMemsetAsync(mine, 0xff, (size_t)maxp * sizeof(uint32_t), st));
    HOT_CUDA(cudaMemcpyAsync(mine, s->pid_sorted.p, (size_t)s->n_pages * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    rc = comm_all_gather(s, mine, s->x_pids.p, (long)maxp * sizeof(uint32_t));
    if (rc) return rc;
    std::vector<uint32_t> all((size_t)W * maxp);
    std::vector<int> slot_sorted(s->n_pages);
    HOT_CUDA(cudaMemcpyAsync(all.data(), s->x_pids.p, all.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaMemcpyAsync(slot_sorted.data(), s->slot_sorted.p, slot_sorted.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    std::vector<int> sh_auth;
    std::vector<int> x_slot, sh_slot, sh_ptr, sh_entry, sh_owned, sh_rank;
    s->n_base_pages = s->n_pages;
    if (s->ghost_ring) {
        // 1b. ghost ring: every rank derives everybody's ghost pages from the base lists (same inputs, same result everywhere)
        std::vector<std::vector<uint32_t>> ext(W);
        int maxe = 0;
        for (int r = 0; r < W; ++r) {
            halo_pages(r, W, maxp, counts.data(), all.data(), ext[r]);
            maxe = std::max(maxe, counts[r] + (int)ext[r].size());
        }
        // this rank's page table grows by its ghost pages (slots n_pages ..., no particles, no page groups)
        const std::vector<uint32_t>& mine_ext = ext[s->rank];
        const long NP0 = s->n_pages, NE = (long)mine_ext.size(), NP1 = NP0 + NE;
        if (NE > 0) {
            DevBuf<uint32_t> grown;
            HOT_CUDA(grown.reserve((size_t)NP1));
            HOT_CUDA(cudaMemcpyAsync(grown.p, s->page_id.p, (size_t)NP0 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            HOT_CUDA(cudaMemcpyAsync(grown.p + NP0, mine_ext.data(), (size_t)NE * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            HOT_CUDA(cudaStreamSynchronize(st));
            s->page_id.swap(grown);
        }
        // merged ascending lists of every rank; this rank's with its slots
        std::vector<uint32_t> all2((size_t)W * maxe, 0xffffffffu);
        std::vector<int> counts2(W), slot2(NP1);
        for (int r = 0; r < W; ++r) {
            const uint32_t* base = all.data() + (size_t)r * maxp;
            uint32_t* out = all2.data() + (size_t)r * maxe;
            int i = 0, j = 0, k = 0;
            const int nb = counts[r], ne = (int)ext[r].size();
            while (i < nb || j < ne) {
                const bool take_base = j >= ne || (i < nb && base[i] < ext[r][j]);
                if (r == s->rank) slot2[k] = take_base ? slot_sorted[i] : (int)(NP0 + j);
                out[k++] = take_base ? base[i++] : ext[r][j++];
            }
            counts2[r] = k;
        }
        s->n_pages = NP1;
        HOT_CUDA(s->pid_sorted.reserve((size_t)NP1));
        HOT_CUDA(s->slot_sorted.reserve((size_t)NP1));
        HOT_CUDA(cudaMemcpyAsync(s->pid_sorted.p, all2.data() + (size_t)s->rank * maxe, (size_t)NP1 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        HOT_CUDA(cudaMemcpyAsync(s->slot_sorted.p, slot2.data(), (size_t)NP1 * sizeof(int), cudaMemcpyHostToDevice, st));
        HOT_CUDA(cudaStreamSynchronize(st));
        rc = rebuild_neighbours(s); // nbr8 over the grown table
        if (rc) return rc;
        share_tables(s->rank, W, maxe, counts2.data(), all2.data(), slot2.data(), s->nbr_rank, s->nbr_off, s->nbr_cnt, x_slot, sh_slot, sh_ptr, sh_entry,
            sh_owned, &sh_rank);
        // authority per shared page (from the BASE lists): who counts the page, and where its values arrive in the exchange list
        const PageHolders h = page_holders(W, maxp, counts.data(), all.data());
        std::vector<int> slot_to_sorted(NP1);
        for (long k = 0; k < NP1; ++k) slot_to_sorted[slot2[k]] = (int)k;
        const uint32_t* mine2 = all2.data() + (size_t)s->rank * maxe;
        sh_auth.assign(sh_slot.size(), -1);
        for (size_t p = 0; p < sh_slot.size(); ++p) {
            const int a = page_authority_of(h, mine2[slot_to_sorted[sh_slot[p]]]);
            sh_owned[p] = a == s->rank ? 1 : 0;
            if (a != s->rank)
                for (int q = sh_ptr[p]; q < sh_ptr[p + 1]; ++q)
                    if (sh_rank[q] == a) sh_auth[p] = sh_entry[q];
            if (a != s->rank && sh_auth[p] < 0) return fail(s, "ghost ring: the authority of a shared page is not among its sharers");
        }
    }
    else {
        // 2. + 3. intersections and the CSR of contributions (host logic, also exported for the CPU tests: hot_share_tables)
        share_tables(s->rank, W, maxp, counts.data(), all.data(), slot_sorted.data(), s->nbr_rank, s->nbr_off, s->nbr_cnt, x_slot, sh_slot, sh_ptr, sh_entry,
            sh_owned, nullptr);
        sh_auth.assign(sh_slot.size(), -1); // (lowest sharer counts; no take-over exchange without the ghost ring)
    }
    s->x_total = (long)x_slot.size();
    s->n_sh = (int)sh_slot.size();
    auto up = [&](DevBuf<int>& d, const std::vector<int>& h) -> cudaError_t {
        cudaError_t e = d.reserve(h.size() > 0 ? h.size() : 1);
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    };
    HOT_CUDA(up(s->x_slot, x_slot));
    HOT_CUDA(up(s->sh_slot, sh_slot));
    HOT_CUDA(up(s->sh_ptr, sh_ptr));
    HOT_CUDA(up(s->sh_entry, sh_entry));
    HOT_CUDA(up(s->sh_owned, sh_owned));
    HOT_CUDA(up(s->sh_auth, sh_auth));
    HOT_CUDA(cudaStreamSynchronize(st)); // the host vectors go out of scope
    return xp_after_sort(s);
}

// P2G of a partitioned object: complete mass / momentum on the shared pages (all sharers end with identical values)
int dist_p2g_exchange(Sim* s)
{
    if (s->world <= 1) return 0;
    KTime t(s, KC_TRANSFER);
    PageSource src{4, s->g_m.p, s->g_v.p, s->g_stride, nullptr, nullptr, 0, 0};
    return exchange_pages(s, src, s->g_m.p, s->g_v.p, nullptr);
}

// after the numbering: which nodes this rank counts in reductions
int dist_after_numbering(Sim* s)
{
    if (s->world <= 1) {
        s->n_owned_nodes = s->num_nodes;
        return 0;
    }
    if (s->own_valid) return 0; // same sort -> same pages, same numbering
    cudaStream_t st = s->stream;
    const int nn = s->num_nodes;
    HOT_CUDA(s->own_node.reserve(nn > 0 ? nn : 1));
    HOT_CUDA(cudaMemsetAsync(s->own_node.p, 1, (size_t)nn, st));
    if (s->n_sh > 0) {
        k_own_nodes<<<nblk((long)s->n_sh * Geo::E), TPB, 0, st>>>(s->n_sh, s->sh_slot.p, s->sh_owned.p, s->g_idx.p, s->own_node.p);
        HOT_LAUNCHED(s);
    }
    // owned nodes of this rank and nodes of the whole object (N_n of the convergence tests, ImplicitSolver.h:192-202)
    HOT_CUDA(s->x_scalars.reserve(64));
    int rc = reduce_to<1>(s, nn, CountOwnF{s->own_node.p}, s->x_scalars.p, nullptr);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->x_scalars.p + 1, s->x_scalars.p, sizeof(double), cudaMemcpyDeviceToDevice, st));
    rc = comm_all_reduce(s, s->x_scalars.p + 1, 1, 0);
    if (rc) return rc;
    double h[2];
    HOT_CUDA(cudaMemcpyAsync(h, s->x_scalars.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    s->n_owned_nodes = (int)h[0];
    s->global_nodes = (long)h[1];
    s->own_valid = true;
    return 0;
}

// sum over the sharers on the shared nodes of a DOF array with `comps` per node
int dist_exchange_shared(Sim* s, double* v, int comps)
{
    if (s->world <= 1) return 0;
    KTime t(s, KC_TRANSFER);
    PageSource src{comps, nullptr, nullptr, 0, s->g_idx.p, v, comps, 0};
    return exchange_pages(s, src, nullptr, nullptr, v);
}

// every holder of a shared page replaces its values by those of the page's authority (after a Gauss-Seidel colour phase, an SpMV
// ... on the assembled-matrix path, where only the authority's rows have all their columns)
int dist_takeover_shared(Sim* s, double* v, int comps)
{
    if (s->world <= 1) return 0;
    if (!s->ghost_ring) return fail(s, "take-over exchange without the ghost ring (hot_set_ghost_ring)");
    KTime t(s, KC_TRANSFER);
    PageSource src{comps, nullptr, nullptr, 0, s->g_idx.p, v, comps, 0};
    return exchange_pages(s, src, nullptr, nullptr, v, true);
}

// sum over the sharers of components [c0, c0 + nc) of a DOF array with `stride` doubles per node (block rows in slices)
int dist_exchange_rows(Sim* s, double* val, int stride, int c0, int nc)
{
    if (s->world <= 1) return 0;
    KTime t(s, KC_TRANSFER);
    PageSource src{nc, nullptr, nullptr, 0, s->g_idx.p, val, stride, c0};
    return exchange_pages(s, src, nullptr, nullptr, val);
}

int dist_allreduce_buffer(Sim* s, double* dev, long count, int op)
{
    if (s->world <= 1 || count <= 0) return 0;
    return comm_all_reduce(s, dev, count, op);
}

int dist_all_gather_host(Sim* s, const void* mine, void* all, long bytes)
{
    if (s->world <= 1) {
        std::memcpy(all, mine, (size_t)bytes);
        return 0;
    }
    return host_all_gather(s, mine, all, bytes);
}
int dist_all_gather_dev(Sim* s, const void* send, void* recv, long bytes) { return comm_all_gather(s, send, recv, bytes); }

int dist_allreduce_host(Sim* s, double* host, int count, int op)
{
    if (s->world <= 1) return 0;
    HOT_CUDA(s->x_scalars.reserve(64));
    if (count > 64) return fail(s, "dist_allreduce_host: at most 64 scalars");
    HOT_CUDA(cudaMemcpyAsync(s->x_scalars.p, host, count * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    int rc = comm_all_reduce(s, s->x_scalars.p, count, op);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(host, s->x_scalars.p, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

} // namespace hot
