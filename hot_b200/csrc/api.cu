// extern "C" entry points of include/hot_b200.h: argument checking, host<->device marshalling, dispatch.
#include "api_internal.h"
#include <algorithm>

using namespace hot;

namespace {

constexpr int TPB = 256;
inline int nblk(long n) { return (int)((n + TPB - 1) / TPB); }

thread_local std::string g_create_error;

// AoS (n x comps, original order) -> component-major SoA rows
__global__ void k_aos_to_soa(long n, int comps, const double* __restrict__ aos, size_t stride, double* __restrict__ soa)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * comps) return;
    long i = t / comps;
    int c = (int)(t - i * comps);
    soa[c * stride + i] = aos[t];
}
// SoA rows in sorted order -> AoS in original order
__global__ void k_soa_to_aos(long n, int comps, const double* __restrict__ soa, size_t stride, const int* __restrict__ orig_id,
    double* __restrict__ aos)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * comps) return;
    long s = t / comps;
    int c = (int)(t - s * comps);
    aos[(size_t)orig_id[s] * comps + c] = soa[c * stride + s];
}
// AoS in original order -> SoA rows in the CURRENT particle order (sorted slot s holds original particle orig_id[s])
__global__ void k_aos_to_soa_perm(long n, int comps, const double* __restrict__ aos, size_t stride, const int* __restrict__ orig_id,
    double* __restrict__ soa)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * comps) return;
    long s = t / comps;
    int c = (int)(t - s * comps);
    soa[c * stride + s] = aos[(size_t)orig_id[s] * comps + c];
}
__global__ void k_fill1(long n, double a, double* v)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) v[t] = a;
}
__global__ void k_iota_i(long n, int* v)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) v[t] = (int)t;
}
// SPGrid_Mask of GridState<float,3> (64-byte record: data_bits 6, block 4x4x4; Lib/MPM/MpmGrid.h:14-34, SPGrid_Mask.h:31-52).  The
// compute kernels are built for the reference binary's T = double (main.cpp:12); the float geometry is served for addressing only.
struct GeoF32 {
    static constexpr uint64_t xmask = 0x4924924924924c00ull, ymask = 0x2492492492492300ull, zmask = 0x92492492492490c0ull;
};
template <class G>
__device__ inline uint64_t packed_add_g(uint64_t a, uint64_t b)
{
    const uint64_t w = ~(G::xmask | G::ymask | G::zmask);
    return (((a | ~G::xmask) + (b & G::xmask)) & G::xmask) | (((a | ~G::ymask) + (b & G::ymask)) & G::ymask)
        | (((a | ~G::zmask) + (b & G::zmask)) & G::zmask) | (((a | ~w) + (b & w)) & w);
}
template <class G>
__global__ void k_mask_ops(int op, long n, const int* __restrict__ ijk_in, const unsigned long long* __restrict__ a,
    const unsigned long long* __restrict__ b, unsigned long long* __restrict__ out, int* __restrict__ ijk_out)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (op == 0)
        out[t] = bit_spread((uint32_t)ijk_in[3 * t], G::xmask) | bit_spread((uint32_t)ijk_in[3 * t + 1], G::ymask)
            | bit_spread((uint32_t)ijk_in[3 * t + 2], G::zmask);
    else if (op == 1) {
        ijk_out[3 * t] = (int)bit_pack(a[t], G::xmask);
        ijk_out[3 * t + 1] = (int)bit_pack(a[t], G::ymask);
        ijk_out[3 * t + 2] = (int)bit_pack(a[t], G::zmask);
    }
    else out[t] = packed_add_g<G>(a[t], b[t]);
}
// particle_sorter / particle_order / particle_base_offset views of the device state
__global__ void k_export_sort(long n, const uint64_t* __restrict__ keys, int* __restrict__ order, unsigned long long* __restrict__ base_offset)
{
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint64_t key = keys[s];
    int o = (int)(key & ((1ull << Geo::index_bits) - 1));
    order[s] = o;
    base_offset[o] = (key >> Geo::index_bits) << Geo::data_bits;
}
__global__ void k_export_groups(long G, const int* __restrict__ group_first, int* __restrict__ first, int* __restrict__ last)
{
    long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    first[g] = group_first[g];
    last[g] = group_first[g + 1] - 1;
}
__global__ void k_export_pages(long NP, const uint32_t* __restrict__ page_id, unsigned long long* __restrict__ out)
{
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < NP) out[p] = (unsigned long long)page_id[p] << 12;
}
__global__ void k_export_grid(long n, size_t gs, const int* __restrict__ idx, const double* __restrict__ v, long long* __restrict__ idx64,
    double* __restrict__ v_aos)
{
    long a = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    idx64[a] = idx[a];
    v_aos[3 * a] = v[a];
    v_aos[3 * a + 1] = v[gs + a];
    v_aos[3 * a + 2] = v[2 * gs + a];
}

template <class T>
int d2h(hot_sim* s, T* host, const T* dev, size_t n)
{
    HOT_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

} // namespace

extern "C" {

hot_sim* hot_create(double dx, double apic_rpic_ratio, double cfl, int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fprintf(stderr, "hot_create: no usable CUDA device (%s); hot_b200 has no CPU fallback\n", cudaGetErrorString(e));
        return nullptr;
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        fprintf(stderr, "hot_create: cudaSetDevice(%d) failed\n", device);
        return nullptr;
    }
    if (!(dx > 0)) {
        fprintf(stderr, "hot_create: dx must be positive\n");
        return nullptr;
    }
    hot_sim* s = new hot_sim;
    cudaGetDevice(&s->device);
    s->dx = dx;
    s->apic_rpic_ratio = apic_rpic_ratio;
    s->cfl = cfl;
    return s;
}

void hot_destroy(hot_sim* s)
{
    if (!s) return;
    comm_destroy(s);
    if (s->copy_in) cudaStreamDestroy(s->copy_in);
    if (s->copy_out) cudaStreamDestroy(s->copy_out);
    for (cudaEvent_t e : {s->ev_in_done, s->ev_in_free, s->ev_out_ready, s->ev_out_done})
        if (e) cudaEventDestroy(e);
    if (s->hcount) cudaFreeHost(s->hcount);
    if (s->h_red) cudaFreeHost(s->h_red);
    delete s;
}

const char* hot_last_error(hot_sim* s) { return s ? s->err.c_str() : "null handle"; }

int hot_set_stream(hot_sim* s, void* cuda_stream)
{
    s->stream = (cudaStream_t)cuda_stream;
    return 0;
}
int hot_synchronize(hot_sim* s)
{
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
long long hot_launch_count(hot_sim* s) { return s->launches; }

int hot_timing(hot_sim* s, int enable)
{
    s->timers.collect();
    s->timers.on = enable != 0;
    if (enable == 2) // reset
        for (int c = 0; c < KC_COUNT; ++c) s->timers.ms[c] = 0, s->timers.count[c] = 0;
    return 0;
}
int hot_get_timings(hot_sim* s, int n, double* ms_total, long long* counts)
{
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    s->timers.collect();
    for (int c = 0; c < n && c < KC_COUNT; ++c) {
        if (ms_total) ms_total[c] = s->timers.ms[c];
        if (counts) counts[c] = s->timers.count[c];
    }
    return KC_COUNT;
}
const char* hot_timing_name(int c)
{
    static const char* names[KC_COUNT] = {"sort", "p2g", "number_nodes", "g2p", "gather", "stress", "force", "hessian_apply", "assemble",
        "spmv", "gs_smooth", "transfer", "blas1"};
    return (c >= 0 && c < KC_COUNT) ? names[c] : "";
}

static int mask_op(hot_sim* s, int op, long n, const int* ijk_in, const unsigned long long* a, const unsigned long long* b,
    unsigned long long* out, int* ijk_out, bool f32 = false)
{
    if (n <= 0) return 0;
    HOT_CUDA(s->stage_u.reserve(3 * n));
    HOT_CUDA(s->stage_i.reserve(3 * n));
    unsigned long long *da = s->stage_u.p, *db = s->stage_u.p + n, *dout = s->stage_u.p + 2 * n;
    if (ijk_in) HOT_CUDA(cudaMemcpyAsync(s->stage_i.p, ijk_in, 3 * n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    if (a) HOT_CUDA(cudaMemcpyAsync(da, a, n * sizeof(*a), cudaMemcpyHostToDevice, s->stream));
    if (b) HOT_CUDA(cudaMemcpyAsync(db, b, n * sizeof(*b), cudaMemcpyHostToDevice, s->stream));
    if (f32) k_mask_ops<GeoF32><<<nblk(n), TPB, 0, s->stream>>>(op, n, s->stage_i.p, da, db, dout, s->stage_i.p);
    else k_mask_ops<Geo><<<nblk(n), TPB, 0, s->stream>>>(op, n, s->stage_i.p, da, db, dout, s->stage_i.p);
    HOT_LAUNCHED(s);
    if (out) return d2h(s, out, dout, n);
    return d2h(s, ijk_out, s->stage_i.p, 3 * n);
}
int hot_linear_offset(hot_sim* s, long n, const int* ijk, unsigned long long* out) { return mask_op(s, 0, n, ijk, nullptr, nullptr, out, nullptr); }
int hot_linear_to_coord(hot_sim* s, long n, const unsigned long long* off, int* ijk) { return mask_op(s, 1, n, nullptr, off, nullptr, nullptr, ijk); }
int hot_packed_add(hot_sim* s, long n, const unsigned long long* a, const unsigned long long* b, unsigned long long* out)
{
    return mask_op(s, 2, n, nullptr, a, b, out, nullptr);
}
int hot_linear_offset_f32(hot_sim* s, long n, const int* ijk, unsigned long long* out) { return mask_op(s, 0, n, ijk, nullptr, nullptr, out, nullptr, true); }
int hot_linear_to_coord_f32(hot_sim* s, long n, const unsigned long long* off, int* ijk) { return mask_op(s, 1, n, nullptr, off, nullptr, nullptr, ijk, true); }
int hot_packed_add_f32(hot_sim* s, long n, const unsigned long long* a, const unsigned long long* b, unsigned long long* out)
{
    return mask_op(s, 2, n, nullptr, a, b, out, nullptr, true);
}

int hot_set_particles(hot_sim* s, long n, const double* X, const double* V, const double* mass, const double* C, const double* F,
    const double* vol, const double* mu, const double* lambda)
{
    if (n <= 0) return fail(s, "hot_set_particles: n must be positive");
    if (!X || !V || !mass || !C || !F || !vol || !mu || !lambda) return fail(s, "hot_set_particles: null attribute array");
    HOT_CUDA(s->P.reserve(n));
    HOT_CUDA(s->stage.reserve(28 * (size_t)n));
    s->N = n;
    cudaStream_t st = s->stream;
    struct Item { const double* h; int comps; double* soa; };
    Item items[] = {{X, 3, s->P.X.p}, {V, 3, s->P.V.p}, {mass, 1, s->P.M.p}, {C, 9, s->P.C.p}, {F, 9, s->P.F.p}, {vol, 1, s->P.vol.p},
        {mu, 1, s->P.mu.p}, {lambda, 1, s->P.lam.p}};
    size_t off = 0;
    for (const Item& it : items) {
        double* d = s->stage.p + off;
        HOT_CUDA(cudaMemcpyAsync(d, it.h, (size_t)n * it.comps * sizeof(double), cudaMemcpyHostToDevice, st));
        if (it.comps == 1)
            HOT_CUDA(cudaMemcpyAsync(it.soa, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        else {
            k_aos_to_soa<<<nblk(n * it.comps), TPB, 0, st>>>(n, it.comps, d, s->P.stride, it.soa);
            HOT_LAUNCHED(s);
        }
        off += (size_t)n * it.comps;
    }
    k_iota_i<<<nblk(n), TPB, 0, st>>>(n, s->P.orig_id.p);
    HOT_LAUNCHED(s);
    k_fill1<<<nblk(n), TPB, 0, st>>>(n, 1.0, s->P.Jp.p); // SnowPlasticity::Jp starts at 1
    HOT_LAUNCHED(s);
    s->sorted = false;
    s->p2g_done = false;
    s->dpdf_norm_max = -1.0; // new material parameters: computeCharacteristicNorm's cache starts over
    s->dv0_valid = false;
    HOT_CUDA(cudaStreamSynchronize(st)); // the caller's arrays may be released / overwritten when the call returns (pinned memory is copied asynchronously)
    return 0;
}

// ---- pipelined particle state: upload of the next step and download of the previous one overlap the current step -------------
static int ensure_copy_streams(hot_sim* s)
{
    if (s->copy_in) return 0;
    HOT_CUDA(cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking));
    HOT_CUDA(cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&s->ev_in_done, &s->ev_in_free, &s->ev_out_ready, &s->ev_out_done}) HOT_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return 0;
}
int hot_upload_state_async(hot_sim* s, const double* X, const double* V, const double* C, const double* F)
{
    const long n = s->N;
    if (n <= 0) return fail(s, "hot_upload_state_async: call hot_set_particles first (masses, volumes and material parameters stay resident)");
    if (!X || !V || !C || !F) return fail(s, "hot_upload_state_async: null array");
    if (s->in_pending) return fail(s, "hot_upload_state_async: the previous upload has not been committed (hot_commit_state)");
    int rc = ensure_copy_streams(s);
    if (rc) return rc;
    HOT_CUDA(s->stage_in.reserve(24 * (size_t)n));
    if (s->in_free_recorded) HOT_CUDA(cudaStreamWaitEvent(s->copy_in, s->ev_in_free, 0)); // the last commit has read the staging area
    const double* h[4] = {X, V, C, F};
    const int comps[4] = {3, 3, 9, 9};
    size_t off = 0;
    for (int k = 0; k < 4; ++k) {
        HOT_CUDA(cudaMemcpyAsync(s->stage_in.p + off, h[k], (size_t)n * comps[k] * sizeof(double), cudaMemcpyHostToDevice, s->copy_in));
        off += (size_t)n * comps[k];
    }
    HOT_CUDA(cudaEventRecord(s->ev_in_done, s->copy_in));
    s->in_pending = true;
    return 0;
}
int hot_commit_state(hot_sim* s)
{
    if (!s->in_pending) return fail(s, "hot_commit_state: no upload in flight (hot_upload_state_async)");
    const long n = s->N;
    cudaStream_t st = s->stream;
    HOT_CUDA(cudaStreamWaitEvent(st, s->ev_in_done, 0));
    double* soa[4] = {s->P.X.p, s->P.V.p, s->P.C.p, s->P.F.p};
    const int comps[4] = {3, 3, 9, 9};
    size_t off = 0;
    for (int k = 0; k < 4; ++k) {
        k_aos_to_soa_perm<<<nblk(n * comps[k]), TPB, 0, st>>>(n, comps[k], s->stage_in.p + off, s->P.stride, s->P.orig_id.p, soa[k]);
        HOT_LAUNCHED(s);
        off += (size_t)n * comps[k];
    }
    HOT_CUDA(cudaEventRecord(s->ev_in_free, st));
    s->in_free_recorded = true;
    s->in_pending = false;
    s->sorted = false;
    s->p2g_done = false;
    s->state_valid = s->hessian_valid = false;
    return 0;
}
int hot_download_state_async(hot_sim* s, double* X, double* V, double* C, double* F)
{
    const long n = s->N;
    if (n <= 0) return fail(s, "hot_download_state_async: no particles");
    if (!X || !V || !C || !F) return fail(s, "hot_download_state_async: null array");
    int rc = ensure_copy_streams(s);
    if (rc) return rc;
    HOT_CUDA(s->stage_out.reserve(24 * (size_t)n));
    cudaStream_t st = s->stream;
    if (s->out_pending) HOT_CUDA(cudaStreamWaitEvent(st, s->ev_out_done, 0)); // the previous download has drained the staging area
    double* h[4] = {X, V, C, F};
    const double* soa[4] = {s->P.X.p, s->P.V.p, s->P.C.p, s->P.F.p};
    const int comps[4] = {3, 3, 9, 9};
    size_t off = 0;
    for (int k = 0; k < 4; ++k) {
        k_soa_to_aos<<<nblk(n * comps[k]), TPB, 0, st>>>(n, comps[k], soa[k], s->P.stride, s->P.orig_id.p, s->stage_out.p + off);
        HOT_LAUNCHED(s);
        off += (size_t)n * comps[k];
    }
    HOT_CUDA(cudaEventRecord(s->ev_out_ready, st));
    HOT_CUDA(cudaStreamWaitEvent(s->copy_out, s->ev_out_ready, 0));
    off = 0;
    for (int k = 0; k < 4; ++k) {
        HOT_CUDA(cudaMemcpyAsync(h[k], s->stage_out.p + off, (size_t)n * comps[k] * sizeof(double), cudaMemcpyDeviceToHost, s->copy_out));
        off += (size_t)n * comps[k];
    }
    HOT_CUDA(cudaEventRecord(s->ev_out_done, s->copy_out));
    s->out_pending = true;
    return 0;
}
int hot_wait_download(hot_sim* s)
{
    if (!s->out_pending) return 0;
    HOT_CUDA(cudaEventSynchronize(s->ev_out_done));
    s->out_pending = false;
    return 0;
}

int hot_get_particles(hot_sim* s, double* X, double* V, double* C, double* F, double* gradV)
{
    const long n = s->N;
    if (n <= 0) return fail(s, "hot_get_particles: no particles");
    if (gradV && s->P.gradV.cap < 9 * s->P.stride) return fail(s, "hot_get_particles: gradV is only defined after hot_g2p");
    HOT_CUDA(s->stage.reserve(28 * (size_t)n));
    cudaStream_t st = s->stream;
    struct Item { double* h; int comps; const double* soa; };
    Item items[] = {{X, 3, s->P.X.p}, {V, 3, s->P.V.p}, {C, 9, s->P.C.p}, {F, 9, s->P.F.p}, {gradV, 9, s->P.gradV.p}};
    size_t off = 0;
    for (const Item& it : items) {
        if (!it.h) continue;
        if (off + (size_t)n * it.comps > 28 * (size_t)n) { // staging holds 28 doubles per particle
            HOT_CUDA(cudaStreamSynchronize(st));
            off = 0;
        }
        double* d = s->stage.p + off;
        k_soa_to_aos<<<nblk(n * it.comps), TPB, 0, st>>>(n, it.comps, it.soa, s->P.stride, s->P.orig_id.p, d);
        HOT_LAUNCHED(s);
        HOT_CUDA(cudaMemcpyAsync(it.h, d, (size_t)n * it.comps * sizeof(double), cudaMemcpyDeviceToHost, st));
        off += (size_t)n * it.comps;
    }
    HOT_CUDA(cudaStreamSynchronize(st));
    return 0;
}
long hot_num_particles(hot_sim* s) { return s->N; }

int hot_sort_and_activate(hot_sim* s) { return sort_and_activate(s); }
long hot_num_groups(hot_sim* s) { return s->n_groups; }
long hot_num_pages(hot_sim* s) { return s->n_pages; }

int hot_get_sort(hot_sim* s, unsigned long long* sorter, int* order, unsigned long long* base_offset)
{
    if (!s->sorted) return fail(s, "hot_get_sort: call hot_sort_and_activate first");
    const long n = s->N;
    if (sorter) {
        int rc = d2h(s, sorter, (const unsigned long long*)s->keys.p, n);
        if (rc) return rc;
    }
    if (order || base_offset) {
        HOT_CUDA(s->stage_i.reserve(n));
        HOT_CUDA(s->stage_u.reserve(n));
        k_export_sort<<<nblk(n), TPB, 0, s->stream>>>(n, s->keys.p, s->stage_i.p, s->stage_u.p);
        HOT_LAUNCHED(s);
        if (order) {
            int rc = d2h(s, order, s->stage_i.p, n);
            if (rc) return rc;
        }
        if (base_offset) return d2h(s, base_offset, s->stage_u.p, n);
    }
    return 0;
}

int hot_get_groups(hot_sim* s, int* first, int* last, unsigned long long* block_offset)
{
    if (!s->sorted) return fail(s, "hot_get_groups: call hot_sort_and_activate first");
    const long G = s->n_groups;
    HOT_CUDA(s->stage_i.reserve(2 * G));
    k_export_groups<<<nblk(G), TPB, 0, s->stream>>>(G, s->group_first.p, s->stage_i.p, s->stage_i.p + G);
    HOT_LAUNCHED(s);
    int rc = 0;
    if (first) rc = d2h(s, first, s->stage_i.p, G);
    if (!rc && last) rc = d2h(s, last, s->stage_i.p + G, G);
    if (!rc && block_offset) rc = d2h(s, block_offset, (const unsigned long long*)s->group_block.p, G);
    return rc;
}

int hot_get_pages(hot_sim* s, unsigned long long* offsets)
{
    if (!s->sorted) return fail(s, "hot_get_pages: call hot_sort_and_activate first");
    HOT_CUDA(s->stage_u.reserve(s->n_pages));
    k_export_pages<<<nblk(s->n_pages), TPB, 0, s->stream>>>(s->n_pages, s->page_id.p, s->stage_u.p);
    HOT_LAUNCHED(s);
    return d2h(s, offsets, s->stage_u.p, s->n_pages);
}

int hot_p2g(hot_sim* s, int* n_nodes)
{
    int rc = p2g(s);
    if (rc) return rc;
    if (n_nodes) *n_nodes = s->num_nodes;
    return 0;
}
int hot_num_nodes(hot_sim* s) { return s->num_nodes; }

int hot_get_grid(hot_sim* s, long long* idx, double* m, double* v)
{
    if (!s->p2g_done) return fail(s, "hot_get_grid: call hot_p2g first");
    const long gn = (long)s->g_stride;
    HOT_CUDA(s->stage.reserve(3 * gn));
    HOT_CUDA(s->stage_u.reserve(gn));
    k_export_grid<<<nblk(gn), TPB, 0, s->stream>>>(gn, s->g_stride, s->g_idx.p, s->g_v.p, (long long*)s->stage_u.p, s->stage.p);
    HOT_LAUNCHED(s);
    int rc = 0;
    if (idx) rc = d2h(s, idx, (const long long*)s->stage_u.p, gn);
    if (!rc && m) rc = d2h(s, m, s->g_m.p, gn);
    if (!rc && v) rc = d2h(s, v, s->stage.p, 3 * gn);
    return rc;
}

int hot_get_id2coord(hot_sim* s, int* coord)
{
    if (!s->p2g_done) return fail(s, "hot_get_id2coord: call hot_p2g first");
    const int n = s->num_nodes;
    if (n == 0) return 0;
    HOT_CUDA(s->stage_i.reserve(3 * (size_t)n));
    int rc = fill_id2coord(s, s->stage_i.p);
    if (rc) return rc;
    return d2h(s, coord, s->stage_i.p, 3 * (size_t)n);
}

int hot_get_mass_matrix(hot_sim* s, double* mass)
{
    if (!s->p2g_done) return fail(s, "hot_get_mass_matrix: call hot_p2g first");
    return d2h(s, mass, s->mass_matrix.p, s->num_nodes);
}

int hot_set_dv(hot_sim* s, const double* dv)
{
    if (!s->p2g_done) return fail(s, "hot_set_dv: call hot_p2g first");
    HOT_CUDA(cudaMemcpyAsync(s->dv.p, dv, 3 * (size_t)s->num_nodes * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream)); // caller's buffer may be released when the call returns
    return 0;
}

int hot_g2p(hot_sim* s, double dt, int* flags) { return g2p(s, dt, flags); }

// ---- plasticity (rank 2 of SURVEY 8f): return mapping applied by hot_g2p after evolveStrain -------------------------------
int hot_set_plasticity(hot_sim* s, int model, const double* params)
{
    if (model < 0 || model > 3) return fail(s, "hot_set_plasticity: model must be 0 (none), 1 (VonMisesFixedCorotated), 2 (SnowPlasticity) or 3 (Drucker-Prager extension)");
    if (model && !params) return fail(s, "hot_set_plasticity: null parameters");
    if (model == 1 && !(params[0] >= 0)) return fail(s, "yield_stress must be non-negative (PlasticityApplier.cpp:99)");
    s->plastic_model = model;
    if (model == 3 && !(params[0] >= 0 && params[0] < 90)) return fail(s, "Drucker-Prager: friction angle in degrees, [0, 90)");
    for (int k = 0; k < 5; ++k) s->plastic_param[k] = (model == 2 || (model == 1 && k == 0) || (model == 3 && k < 2)) ? params[k] : 0.0;
    return 0;
}
int hot_apply_plasticity(hot_sim* s)
{
    if (s->N <= 0) return fail(s, "hot_apply_plasticity: no particles");
    if (!s->sorted) { s->p0 = 0; s->p1 = s->N; }
    return apply_plasticity(s);
}
// per-particle plastic state in original order: SnowPlasticity::Jp and the (hardened) mu / lambda; any pointer may be NULL
int hot_get_plastic_state(hot_sim* s, double* Jp, double* mu, double* lambda)
{
    const long n = s->N;
    if (n <= 0) return fail(s, "hot_get_plastic_state: no particles");
    HOT_CUDA(s->stage.reserve(28 * (size_t)n));
    struct Item { double* h; const double* soa; };
    Item items[] = {{Jp, s->P.Jp.p}, {mu, s->P.mu.p}, {lambda, s->P.lam.p}};
    size_t off = 0;
    for (const Item& it : items) {
        if (!it.h) continue;
        double* d = s->stage.p + off;
        k_soa_to_aos<<<nblk(n), TPB, 0, s->stream>>>(n, 1, it.soa, s->P.stride, s->P.orig_id.p, d);
        HOT_LAUNCHED(s);
        HOT_CUDA(cudaMemcpyAsync(it.h, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        off += (size_t)n;
    }
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
int hot_set_plastic_state(hot_sim* s, const double* Jp)
{
    if (s->N <= 0 || s->sorted) return fail(s, "hot_set_plastic_state: call right after hot_set_particles (original particle order)");
    HOT_CUDA(cudaMemcpyAsync(s->P.Jp.p, Jp, (size_t)s->N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

// ---- row (e): one object over several GPUs (dist.cu) ---------------------------------------------------------------
int hot_comm_unique_id(void* id128) { return comm_unique_id(id128); }
int hot_comm_init_nccl(hot_sim* s, int rank, int world, const void* id128)
{
    if (!id128) return fail(s, "hot_comm_init_nccl: null id");
    return comm_init_nccl(s, rank, world, id128);
}
int hot_set_partition(hot_sim* s, int rank, int world, const hot_transport* t)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(s, "hot_set_partition: need 0 <= rank < world");
    if (world > 1 && (!t || !t->all_reduce || !t->all_gather || !t->neighbor_exchange))
        return fail(s, "hot_set_partition: world > 1 needs the three transport callbacks (or use hot_comm_init_nccl)");
    comm_destroy(s);
    s->rank = rank;
    s->world = world;
    s->has_transport = world > 1;
    if (world > 1) {
        s->transport.user = t->user;
        s->transport.all_reduce = t->all_reduce;
        s->transport.all_gather = t->all_gather;
        s->transport.neighbor_exchange = t->neighbor_exchange;
    }
    s->sorted = false; // the shared-page tables are built by the next hot_sort_and_activate
    s->p2g_done = false;
    return 0;
}
int hot_get_partition(hot_sim* s, long* out8)
{
    if (s->n_pages <= 0) return fail(s, "hot_get_partition: call hot_sort_and_activate first");
    // (the tables describe the last sort, also after hot_g2p moved the particles)
    out8[0] = s->rank; out8[1] = s->world; out8[2] = (long)s->nbr_rank.size(); out8[3] = s->n_sh;
    out8[4] = s->x_total; out8[5] = s->num_nodes > 0 ? s->n_owned_nodes : -1; out8[6] = s->num_nodes > 0 ? (s->world > 1 ? s->global_nodes : s->num_nodes) : -1;
    out8[7] = s->N;
    return 0;
}
int hot_set_constitutive_model(hot_sim* s, int model)
{
    if (model != 0 && model != 1) return fail(s, "hot_set_constitutive_model: 0 (CorotatedIsotropic) or 1 (neo-Hookean extension)");
    s->constitutive_model = model;
    s->state_valid = s->hessian_valid = false;
    s->matrix_built = s->mg_built = false;
    s->dpdf_norm_max = -1.0;
    return 0;
}
int hot_set_ghost_ring(hot_sim* s, int on)
{
    s->ghost_ring = on != 0;
    s->sorted = false; // takes effect with the next hot_sort_and_activate
    s->p2g_done = false;
    return 0;
}
int hot_halo_pages(int rank, int world, int max_pages, const int* counts, const unsigned* all_pids, int* n_out, unsigned* out)
{
    if (world < 1 || world > 64 || rank < 0 || rank >= world || !counts || !all_pids || !n_out) return -1;
    std::vector<uint32_t> ext;
    halo_pages(rank, world, max_pages, counts, all_pids, ext);
    *n_out = (int)ext.size();
    if (out) std::copy(ext.begin(), ext.end(), out);
    return 0;
}
int hot_page_authority(int world, int max_pages, const int* counts, const unsigned* all_pids, int n, const unsigned* pids, int* auth)
{
    if (world < 1 || world > 64 || !counts || !all_pids || (n > 0 && (!pids || !auth))) return -1;
    page_authority(world, max_pages, counts, all_pids, n, pids, auth);
    return 0;
}
int hot_get_transport(hot_sim* s)
{
    if (s->world <= 1) return 0;
    if (s->xp_state == 1) return 3;
    return s->nccl_comm ? 1 : (s->has_transport ? 2 : 0);
}
// the host logic of the shared-page tables on host arrays (no device, no handle): what dist_after_sort runs after its all-gather.
// Output arrays are caller-allocated for the worst case: nbr_* world entries, x_slot (world - 1) * counts[rank], sh_slot / sh_owned
// counts[rank], sh_ptr counts[rank] + 1, sh_entry world * counts[rank].
int hot_share_tables(int rank, int world, int max_pages, const int* counts, const unsigned* all_pids, const int* slot_sorted, int* n_nbr, int* nbr_rank,
    long* nbr_off, long* nbr_cnt, int* n_x, int* x_slot, int* n_sh, int* sh_slot, int* sh_ptr, int* sh_entry, int* sh_owned)
{
    if (world < 1 || rank < 0 || rank >= world || !counts || !all_pids || !slot_sorted) return -1;
    std::vector<int> nr, xs, ss, sp, se, so;
    std::vector<long> no, nc;
    share_tables(rank, world, max_pages, counts, all_pids, slot_sorted, nr, no, nc, xs, ss, sp, se, so, nullptr);
    *n_nbr = (int)nr.size(); *n_x = (int)xs.size(); *n_sh = (int)ss.size();
    std::copy(nr.begin(), nr.end(), nbr_rank); std::copy(no.begin(), no.end(), nbr_off); std::copy(nc.begin(), nc.end(), nbr_cnt);
    std::copy(xs.begin(), xs.end(), x_slot); std::copy(ss.begin(), ss.end(), sh_slot); std::copy(sp.begin(), sp.end(), sh_ptr);
    std::copy(se.begin(), se.end(), sh_entry); std::copy(so.begin(), so.end(), sh_owned);
    return 0;
}
// raw copies between host and device memory on the handle's stream, synchronous (for transport callbacks written in a host
// language that has no CUDA binding of its own: tests/dist_worker.py moves the exchange buffers through gloo with these)
int hot_memcpy_d2h(hot_sim* s, void* host, const void* dev, long bytes)
{
    HOT_CUDA(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
int hot_memcpy_h2d(hot_sim* s, void* dev, const void* host, long bytes)
{
    HOT_CUDA(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

// ---- force model (force.cu): host-buffer wrappers around the device-resident operators ---------------------------
static int upload_dof(hot_sim* s, DevBuf<double>& buf, const double* host, size_t per_node = 3)
{
    const size_t n = per_node * (size_t)s->num_nodes;
    HOT_CUDA(buf.reserve(n > 0 ? n : 1));
    if (host && n) HOT_CUDA(cudaMemcpyAsync(buf.p, host, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return 0;
}

int hot_set_dt_gravity(hot_sim* s, double dt, const double* g)
{
    if (!(dt >= 0)) return fail(s, "hot_set_dt_gravity: dt must be non-negative");
    s->dt = dt;
    for (int d = 0; d < 3; ++d) s->gravity[d] = g ? g[d] : 0.0;
    s->state_valid = s->hessian_valid = false;
    return 0;
}
int hot_set_project(hot_sim* s, int project)
{
    s->project_pd = project != 0;
    s->hessian_valid = false;
    return 0;
}
int hot_set_bc(hot_sim* s, int mode, int n_bc, const int* node_id, const double* P, const double* R, const double* Rinv, const int* slip,
    const double* dv_bc)
{
    return set_bc(s, mode, n_bc, node_id, P, R, Rinv, slip, dv_bc);
}
int hot_set_colliders(hot_sim* s, int n, const hot_collider* objects) { return set_colliders(s, n, objects); }
int hot_build_bc(hot_sim* s, int mode, int* n_bc) { return build_bc_from_colliders(s, mode, n_bc); }
int hot_get_bc(hot_sim* s, int* n_bc, int* node_id, double* P, double* R, double* Rinv, int* slip)
{
    const size_t n = (size_t)s->n_bc;
    if (n_bc) *n_bc = s->n_bc;
    if (n == 0) return 0;
    int rc = 0;
    if (node_id) rc = d2h(s, node_id, s->bc_node.p, n);
    if (!rc && slip) rc = d2h(s, slip, s->bc_slip.p, n);
    if (!rc && P) rc = d2h(s, P, s->bc_P.p, 9 * n);
    if (!rc && R) rc = d2h(s, R, s->bc_R.p, 9 * n);
    if (!rc && Rinv) rc = d2h(s, Rinv, s->bc_Rinv.p, 9 * n);
    return rc;
}
int hot_get_dv(hot_sim* s, double* dv)
{
    if (!s->p2g_done) return fail(s, "hot_get_dv: call hot_p2g first");
    return d2h(s, dv, s->dv.p, 3 * (size_t)s->num_nodes);
}
// CorotatedIsotropic<T,3>::{updateScratch, psi, firstPiola, firstPiolaDifferential, firstPiolaDerivative} on n deformation gradients
int hot_corotated_eval(hot_sim* s, long n, const double* F, double mu, double lambda, int project, const double* dF, double* psi, double* P,
    double* dP, double* dPdF, double* U, double* sigma, double* V)
{
    if (n <= 0 || !F) return fail(s, "hot_corotated_eval: need n > 0 deformation gradients");
    if (dP && !dF) return fail(s, "hot_corotated_eval: dP requested without dF");
    if ((U != nullptr) != (V != nullptr)) return fail(s, "hot_corotated_eval: U and V come together");
    // layout of the staging area: F 9 | dF 9 | psi 1 | P 9 | dP 9 | dPdF 81 | U 9 | sigma 3 | V 9  = 139 doubles per item
    HOT_CUDA(s->stage.reserve(139 * (size_t)n));
    double* d = s->stage.p;
    double *dF_ = d + 9 * n, *psi_ = d + 18 * n, *P_ = d + 19 * n, *dP_ = d + 28 * n, *H_ = d + 37 * n, *U_ = d + 118 * n, *s_ = d + 127 * n, *V_ = d + 130 * n;
    HOT_CUDA(cudaMemcpyAsync(d, F, 9 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (dF) HOT_CUDA(cudaMemcpyAsync(dF_, dF, 9 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    int rc = corotated_eval(s, n, d, mu, lambda, project, dF ? dF_ : nullptr, psi ? psi_ : nullptr, P ? P_ : nullptr, dP ? dP_ : nullptr,
        dPdF ? H_ : nullptr, U ? U_ : nullptr, sigma ? s_ : nullptr, V ? V_ : nullptr);
    if (rc) return rc;
    struct Item { double* h; const double* dev; size_t per; };
    const Item items[] = {{psi, psi_, 1}, {P, P_, 9}, {dP, dP_, 9}, {dPdF, H_, 81}, {U, U_, 9}, {sigma, s_, 3}, {V, V_, 9}};
    for (const Item& it : items)
        if (it.h) HOT_CUDA(cudaMemcpyAsync(it.h, it.dev, it.per * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
int hot_backup_strain(hot_sim* s) { return backup_strain(s); }
int hot_restore_strain(hot_sim* s) { return restore_strain(s); }

int hot_update_state(hot_sim* s, const double* dv, double* energy)
{
    if (!s->p2g_done) return fail(s, "hot_update_state: call hot_p2g first");
    if (dv) {
        HOT_CUDA(cudaMemcpyAsync(s->dv.p, dv, 3 * (size_t)s->num_nodes * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        HOT_CUDA(cudaStreamSynchronize(s->stream)); // caller's buffer may be released when the call returns
    }
    return update_state(s, energy != nullptr, energy);
}

int hot_get_stress(hot_sim* s, double* vPFnT, double* F)
{
    if (!s->state_valid) return fail(s, "hot_get_stress: call hot_update_state first");
    const long n = s->N;
    HOT_CUDA(s->stage.reserve(28 * (size_t)n));
    struct Item { double* h; const double* soa; };
    Item items[] = {{vPFnT, s->f_stress.p}, {F, s->P.F.p}};
    size_t off = 0;
    for (const Item& it : items) {
        if (!it.h) continue;
        double* d = s->stage.p + off;
        k_soa_to_aos<<<nblk(n * 9), TPB, 0, s->stream>>>(n, 9, it.soa, s->P.stride, s->P.orig_id.p, d);
        HOT_LAUNCHED(s);
        HOT_CUDA(cudaMemcpyAsync(it.h, d, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        off += (size_t)n * 9;
    }
    HOT_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int hot_strain_energy(hot_sim* s, double* energy)
{
    if (!energy) return fail(s, "hot_strain_energy: null output");
    return strain_energy(s, energy);
}
// Fn of FBasedMpmForceHelper (the strain saved by backupStrain), original particle order
int hot_get_strain_backup(hot_sim* s, double* Fn)
{
    if (!s->strain_backed_up) return fail(s, "hot_get_strain_backup: call hot_backup_strain first");
    const long n = s->N;
    HOT_CUDA(s->stage.reserve(28 * (size_t)n));
    k_soa_to_aos<<<nblk(n * 9), TPB, 0, s->stream>>>(n, 9, s->P.Fn.p, s->P.stride, s->P.orig_id.p, s->stage.p);
    HOT_LAUNCHED(s);
    return d2h(s, Fn, s->stage.p, 9 * (size_t)n);
}

int hot_compute_residual(hot_sim* s, double* r)
{
    int rc = upload_dof(s, s->work[1], nullptr);
    if (rc) return rc;
    rc = compute_residual(s, s->work[1].p);
    if (rc) return rc;
    return d2h(s, r, s->work[1].p, 3 * (size_t)s->num_nodes);
}

int hot_project(hot_sim* s, double* v)
{
    if (!s->p2g_done) return fail(s, "hot_project: call hot_p2g first");
    int rc = upload_dof(s, s->work[1], v);
    if (rc) return rc;
    rc = bc_project(s, s->work[1].p);
    if (rc) return rc;
    return d2h(s, v, s->work[1].p, 3 * (size_t)s->num_nodes);
}

int hot_hessian_apply_mf(hot_sim* s, const double* x, double* b)
{
    if (!s->state_valid) return fail(s, "hot_hessian_apply_mf: call hot_update_state first");
    int rc = upload_dof(s, s->work[1], x);
    if (rc) return rc;
    rc = upload_dof(s, s->work[2], nullptr);
    if (rc) return rc;
    rc = hessian_apply_mf(s, s->work[1].p, s->work[2].p);
    if (rc) return rc;
    return d2h(s, b, s->work[2].p, 3 * (size_t)s->num_nodes);
}

int hot_add_scaled_forces(hot_sim* s, double scale, double* f)
{
    if (!s->state_valid) return fail(s, "hot_add_scaled_forces: call hot_update_state first");
    int rc = upload_dof(s, s->work[1], f);
    if (rc) return rc;
    rc = add_scaled_forces(s, scale, s->work[1].p);
    if (rc) return rc;
    return d2h(s, f, s->work[1].p, 3 * (size_t)s->num_nodes);
}
int hot_add_scaled_force_differentials(hot_sim* s, double scale, const double* x, double* f)
{
    if (!s->state_valid) return fail(s, "hot_add_scaled_force_differentials: call hot_update_state first");
    int rc = upload_dof(s, s->work[1], x);
    if (!rc) rc = upload_dof(s, s->work[2], f);
    if (rc) return rc;
    rc = add_scaled_force_differentials(s, scale, s->work[1].p, s->work[2].p);
    if (rc) return rc;
    return d2h(s, f, s->work[2].p, 3 * (size_t)s->num_nodes);
}

int hot_eval_cn_tolerance(hot_sim* s, double eps, double dt, double* tol)
{
    if (!s->p2g_done) return fail(s, "hot_eval_cn_tolerance: call hot_p2g first");
    HOT_CUDA(s->cn_tol.reserve(s->num_nodes > 0 ? s->num_nodes : 1));
    int rc = eval_cn_tolerance(s, eps, dt, s->cn_tol.p);
    if (rc) return rc;
    if (tol) return d2h(s, tol, s->cn_tol.p, (size_t)s->num_nodes);
    return 0;
}

} // extern "C"
