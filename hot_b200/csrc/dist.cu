// Row (e): one object partitioned over the GPUs of a box.
//
// The reference has no distributed path; this is the B200-side design of SURVEY 8e, first stage:
//  * every rank holds all particle POSITIONS and runs the same sort / page activation / DOF numbering, so keys, page list
//    and DOF ids are bit-identical to the single-GPU (and reference) result on every rank;
//  * the sorted page-group list is cut into `world` contiguous ranges balanced by particle count (the key order is Morton
//    over pages, so ranges are spatially compact); a rank runs the particle kernels (P2G, G2P, updateState, Hessian build,
//    force / Hessian / tolerance scatters) only on its groups;
//  * a node belongs to the rank whose groups touch its page first; because the page list is in first-touch order, owned DOF
//    ids are one contiguous range per rank, so reductions over owned nodes are plain sub-ranges;
//  * scatter results are summed across ranks ONLY on interface nodes (nodes of pages touched by >= 2 ranks): pack ->
//    all-reduce -> unpack.  The all-reduce itself is the caller's (torch.distributed / NCCL over NVLink in bench.py and the
//    tests, ncclAllReduce in a C++ host) through a callback on a caller-owned device buffer: the library stays free of a
//    communicator and of any process-group assumption.
//  * grid-side vector algebra is replicated (N_n << N_p), dots are taken over owned nodes and all-reduced.
#include "sim.h"
#include <cub/cub.cuh>
#include <algorithm>

namespace hot {
namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

// every page a group's particles can touch (its own page and the 7 +1 neighbours) gets the group's rank bit
__global__ void k_page_touch(long n_groups, const int* __restrict__ group_slot, const int* __restrict__ group_rank, const int* __restrict__ nbr8,
    unsigned* __restrict__ mask)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups * 8) return;
    const long g = t >> 3;
    const int slot = nbr8[(size_t)group_slot[g] * 8 + (t & 7)];
    if (slot >= 0) atomicOr(mask + slot, 1u << group_rank[g]);
}
// per node: owner rank = lowest touching rank (first touch), interface = touched by >= 2 ranks
__global__ void k_node_owner(long gn, const int* __restrict__ g_idx, const unsigned* __restrict__ mask, int rank, int* __restrict__ iface_flag,
    int* __restrict__ range /* [0] min own id, [1] max own id */)
{
    const long a = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= gn) return;
    const int id = g_idx[a];
    if (id < 0) return;
    const unsigned m = mask[a / Geo::E];
    iface_flag[id] = __popc(m) >= 2;
    if (m && (__ffs(m) - 1) == rank) {
        atomicMin(range, id);
        atomicMax(range + 1, id);
    }
}
// grid channels of the interface pages <-> exchange buffer: [page k][channel 0..3][element e]
__global__ void k_pack_pages(int n_ip, const int* __restrict__ ipage, size_t gs, const double* __restrict__ g_m, const double* __restrict__ g_v,
    double* __restrict__ buf)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n_ip * 4 * Geo::E) return;
    const int e = (int)(t % Geo::E), ch = (int)((t / Geo::E) % 4), k = (int)(t / (4 * Geo::E));
    const size_t a = (size_t)ipage[k] * Geo::E + e;
    buf[t] = ch == 0 ? g_m[a] : g_v[(size_t)(ch - 1) * gs + a];
}
__global__ void k_unpack_pages(int n_ip, const int* __restrict__ ipage, size_t gs, const double* __restrict__ buf, double* __restrict__ g_m,
    double* __restrict__ g_v)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n_ip * 4 * Geo::E) return;
    const int e = (int)(t % Geo::E), ch = (int)((t / Geo::E) % 4), k = (int)(t / (4 * Geo::E));
    const size_t a = (size_t)ipage[k] * Geo::E + e;
    if (ch == 0) g_m[a] = buf[t];
    else g_v[(size_t)(ch - 1) * gs + a] = buf[t];
}
// per page: bit e set when node e carries mass (as a double: exact for 32 bits; the MAX over ranks is the complete mask
// because every rank that touches a page holds its complete mass after the interface exchange, the others hold 0)
// node masks of the pages that exactly ONE rank touches, contributed by that rank (the others add 0: a SUM all-reduce carries
// them next to the interface pages' mass / momentum); pages touched by >= 2 ranks get their mask from the summed mass instead
__global__ void k_page_nonzero(long n_pages, const double* __restrict__ g_m, const unsigned* __restrict__ mask, int rank, double* __restrict__ out)
{
    const long pg = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pg >= n_pages) return;
    unsigned bits = 0;
    if (mask[pg] == (1u << rank))
        for (int e = 0; e < Geo::E; ++e) bits |= (g_m[(size_t)pg * Geo::E + e] != 0.0 ? 1u : 0u) << e;
    out[pg] = (double)bits;
}
__global__ void k_flags_from_bits(long gn, const double* __restrict__ bits, const unsigned* __restrict__ mask, const double* __restrict__ g_m,
    int* __restrict__ flag)
{
    const long a = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= gn) return;
    const long pg = a / Geo::E;
    flag[a] = __popc(mask[pg]) >= 2 ? (g_m[a] != 0.0) : ((((unsigned)bits[pg]) >> (a % Geo::E)) & 1u);
}
__global__ void k_ipage_flags(long n_pages, const unsigned* __restrict__ mask, int* __restrict__ flag)
{
    const long pg = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pg < n_pages) flag[pg] = __popc(mask[pg]) >= 2;
}
__global__ void k_pack(int n, int comps, const int* __restrict__ dof, const double* __restrict__ v, double* __restrict__ buf)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * comps) return;
    const int k = (int)(t / comps), c = (int)(t - (long)k * comps);
    buf[t] = v[(size_t)dof[k] * comps + c];
}
__global__ void k_unpack(int n, int comps, const int* __restrict__ dof, const double* __restrict__ buf, double* __restrict__ v)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * comps) return;
    const int k = (int)(t / comps), c = (int)(t - (long)k * comps);
    v[(size_t)dof[k] * comps + c] = buf[t];
}

int reserve_xbuf(Sim* s, long count)
{
    if (count <= s->xbuf_cap) return 0;
    if (!s->allreduce) return fail(s, "distributed run without an all-reduce callback (hot_set_partition)");
    if (s->allreduce(s->allreduce_user, 2, count) != 0 || s->xbuf_cap < count || !s->xbuf)
        return fail(s, "the all-reduce callback did not provide an exchange buffer of the requested size (hot_set_exchange_buffer)");
    return 0;
}

} // namespace

int dist_after_sort(Sim* s)
{
    s->g0 = 0; s->g1 = s->n_groups; s->p0 = 0; s->p1 = s->N;
    s->dof0 = 0; s->dof1 = 0; s->n_iface = 0;
    s->iface_valid = false;
    if (s->world <= 1) return 0;
    // balanced contiguous split of the page groups by particle count (host: n_groups + 1 ints)
    std::vector<int> first((size_t)s->n_groups + 1);
    HOT_CUDA(cudaMemcpyAsync(first.data(), s->group_first.p, first.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<long> cut(s->world + 1, 0);
    cut[s->world] = s->n_groups;
    for (int r = 1; r < s->world; ++r) {
        const long target = (long)((double)s->N * r / s->world);
        cut[r] = std::lower_bound(first.begin(), first.end(), (int)target) - first.begin();
        if (cut[r] > s->n_groups) cut[r] = s->n_groups;
        if (cut[r] < cut[r - 1]) cut[r] = cut[r - 1];
    }
    std::vector<int> gr(s->n_groups);
    for (int r = 0; r < s->world; ++r)
        for (long g = cut[r]; g < cut[r + 1]; ++g) gr[g] = r;
    HOT_CUDA(s->group_rank.reserve(s->n_groups));
    HOT_CUDA(cudaMemcpyAsync(s->group_rank.p, gr.data(), gr.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    s->g0 = cut[s->rank]; s->g1 = cut[s->rank + 1];
    s->p0 = first[s->g0]; s->p1 = first[s->g1];
    // which ranks touch which page, and the interface pages (touched by >= 2 ranks)
    cudaStream_t st = s->stream;
    HOT_CUDA(s->page_mask.reserve(s->n_pages));
    HOT_CUDA(s->iface_page.reserve(s->n_pages));
    HOT_CUDA(s->head_flag.reserve(s->n_pages));
    HOT_CUDA(s->dcount.reserve(16));
    HOT_CUDA(cudaMemsetAsync(s->page_mask.p, 0, s->n_pages * sizeof(unsigned), st));
    k_page_touch<<<nblk(s->n_groups * 8), TPB, 0, st>>>(s->n_groups, s->group_slot.p, s->group_rank.p, s->nbr8.p, s->page_mask.p);
    HOT_LAUNCHED(s);
    k_ipage_flags<<<nblk(s->n_pages), TPB, 0, st>>>(s->n_pages, s->page_mask.p, s->head_flag.p);
    HOT_LAUNCHED(s);
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, cub::CountingInputIterator<int>(0), s->head_flag.p, s->iface_page.p, s->dcount.p, (int)s->n_pages, st);
    HOT_CUDA(s->cub_tmp.reserve(bytes + 16));
    HOT_CUDA(cub::DeviceSelect::Flagged(s->cub_tmp.p, bytes, cub::CountingInputIterator<int>(0), s->head_flag.p, s->iface_page.p, s->dcount.p,
        (int)s->n_pages, st));
    s->launches++;
    HOT_CUDA(cudaMemcpyAsync(s->hcount + 20, s->dcount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    s->n_iface_pages = s->hcount[20];
    return 0;
}

// P2G of a partitioned object: sum the partial mass / momentum of the interface pages, then agree on which nodes carry mass
// (one 32-bit mask per page) so that every rank computes the same DOF numbering as a single GPU would
int dist_p2g_exchange(Sim* s, int* node_flags /* n_pages * E, out */)
{
    cudaStream_t st = s->stream;
    const long cnt = (long)s->n_iface_pages * 4 * Geo::E;
    KTime t(s, KC_TRANSFER);
    // ONE sum all-reduce: [interface pages: m, mv] [one node mask per page from the page's only toucher]
    int rc = reserve_xbuf(s, cnt + s->n_pages);
    if (rc) return rc;
    if (cnt > 0) {
        k_pack_pages<<<nblk(cnt), TPB, 0, st>>>(s->n_iface_pages, s->iface_page.p, s->g_stride, s->g_m.p, s->g_v.p, s->xbuf);
        HOT_LAUNCHED(s);
    }
    k_page_nonzero<<<nblk(s->n_pages), TPB, 0, st>>>(s->n_pages, s->g_m.p, s->page_mask.p, s->rank, s->xbuf + cnt);
    HOT_LAUNCHED(s);
    if (s->allreduce(s->allreduce_user, 0, cnt + s->n_pages) != 0) return fail(s, "all-reduce callback failed");
    if (cnt > 0) {
        k_unpack_pages<<<nblk(cnt), TPB, 0, st>>>(s->n_iface_pages, s->iface_page.p, s->g_stride, s->xbuf, s->g_m.p, s->g_v.p);
        HOT_LAUNCHED(s);
    }
    k_flags_from_bits<<<nblk((long)s->g_stride), TPB, 0, st>>>((long)s->g_stride, s->xbuf + cnt, s->page_mask.p, s->g_m.p, node_flags);
    HOT_LAUNCHED(s);
    return 0;
}

int dist_after_numbering(Sim* s)
{
    if (s->world <= 1) {
        s->dof0 = 0; s->dof1 = s->num_nodes;
        return 0;
    }
    if (s->iface_valid) return 0; // same sort -> same pages, same numbering
    cudaStream_t st = s->stream;
    const long gn = (long)s->g_stride;
    const int nn = s->num_nodes;
    HOT_CUDA(s->head_flag.reserve(nn > 0 ? nn : 1));
    HOT_CUDA(s->iface_dof.reserve(nn > 0 ? nn : 1));
    HOT_CUDA(s->dcount.reserve(16));
    HOT_CUDA(cudaMemsetAsync(s->head_flag.p, 0, (size_t)nn * sizeof(int), st));
    const int init[2] = {0x7fffffff, -1};
    HOT_CUDA(cudaMemcpyAsync(s->dcount.p + 8, init, sizeof init, cudaMemcpyHostToDevice, st));
    k_node_owner<<<nblk(gn), TPB, 0, st>>>(gn, s->g_idx.p, s->page_mask.p, s->rank, s->head_flag.p, s->dcount.p + 8);
    HOT_LAUNCHED(s);
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, cub::CountingInputIterator<int>(0), s->head_flag.p, s->iface_dof.p, s->dcount.p, nn, st);
    HOT_CUDA(s->cub_tmp.reserve(bytes + 16));
    HOT_CUDA(cub::DeviceSelect::Flagged(s->cub_tmp.p, bytes, cub::CountingInputIterator<int>(0), s->head_flag.p, s->iface_dof.p, s->dcount.p, nn, st));
    s->launches++;
    HOT_CUDA(cudaMemcpyAsync(s->hcount + 16, s->dcount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaMemcpyAsync(s->hcount + 17, s->dcount.p + 8, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    HOT_CUDA(cudaStreamSynchronize(st));
    s->n_iface = s->hcount[16];
    s->dof0 = s->hcount[18] >= 0 ? s->hcount[17] : 0;
    s->dof1 = s->hcount[18] >= 0 ? s->hcount[18] + 1 : 0;
    s->iface_valid = true;
    return 0;
}

int dist_allreduce_buffer(Sim* s, double* dev, long count, int op)
{
    if (s->world <= 1 || count <= 0) return 0;
    int rc = reserve_xbuf(s, count);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->xbuf, dev, count * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    if (s->allreduce(s->allreduce_user, op, count) != 0) return fail(s, "all-reduce callback failed");
    HOT_CUDA(cudaMemcpyAsync(dev, s->xbuf, count * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

int dist_exchange_iface(Sim* s, double* v, int comps)
{
    if (s->world <= 1 || s->n_iface <= 0) return 0;
    const long count = (long)s->n_iface * comps;
    int rc = reserve_xbuf(s, count);
    if (rc) return rc;
    KTime t(s, KC_TRANSFER);
    k_pack<<<nblk(count), TPB, 0, s->stream>>>(s->n_iface, comps, s->iface_dof.p, v, s->xbuf);
    HOT_LAUNCHED(s);
    if (s->allreduce(s->allreduce_user, 0, count) != 0) return fail(s, "all-reduce callback failed");
    k_unpack<<<nblk(count), TPB, 0, s->stream>>>(s->n_iface, comps, s->iface_dof.p, s->xbuf, v);
    HOT_LAUNCHED(s);
    return 0;
}

int dist_allreduce_host(Sim* s, double* host, int count, int op)
{
    if (s->world <= 1) return 0;
    int rc = reserve_xbuf(s, count);
    if (rc) return rc;
    HOT_CUDA(cudaMemcpyAsync(s->xbuf, host, count * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (s->allreduce(s->allreduce_user, op, count) != 0) return fail(s, "all-reduce callback failed");
    HOT_CUDA(cudaMemcpyAsync(host, s->xbuf, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

} // namespace hot
