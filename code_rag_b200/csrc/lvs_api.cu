// C ABI of liblattice_b200.so (see include/lvs.h for the contract and the reference call each entry replaces).
#include "../../include/lvs.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cudaTypedefs.h>

#include "aux_kernels.cuh"
#include "exchange_kernel.cuh"
#include "finalize_kernel.cuh"
#include "gemm_topk_kernel.cuh"
#include "rank_kernel.cuh"
#include "host_util.h"
#include "scan_launch.cuh"

using namespace lvs;

// ------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(e__ == cudaErrorMemoryAllocation ? LVS_ENOMEM : LVS_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                    \
    } while (0)

// ------------------------------------------------------------------------------------------------------
// library state
// ------------------------------------------------------------------------------------------------------
struct LibState {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
};
static LibState g_lib;
static std::mutex g_lib_mu;

// The CUDA device is a per-thread setting and the adapter calls in from worker threads (asyncio.to_thread): every entry point
// that touches CUDA first makes the calling thread current on the device lvs_init bound this process to.
static thread_local int t_bound_device = -1;
static inline void bind_thread() {
    if (g_lib.ready && t_bound_device != g_lib.device) {
        if (cudaSetDevice(g_lib.device) == cudaSuccess) t_bound_device = g_lib.device;
    }
}

struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
};

static int ensure_dev(Scratch& s, size_t bytes) {
    if (s.bytes >= bytes) return LVS_OK;
    if (s.p) cudaFree(s.p);
    s.p = nullptr; s.bytes = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CU(cudaMalloc(&s.p, want));
    s.bytes = want;
    return LVS_OK;
}
static int ensure_pinned(Scratch& s, size_t bytes) {
    if (s.bytes >= bytes) return LVS_OK;
    if (s.p) cudaFreeHost(s.p);
    s.p = nullptr; s.bytes = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CU(cudaHostAlloc(&s.p, want, cudaHostAllocMapped));   // device-visible (zero-copy) as well as DMA-able
    s.bytes = want;
    return LVS_OK;
}

constexpr int kEventRing = 256;
constexpr int kSubmitSlots = 4;

struct lvs_collection {
    std::string name;
    int dim = 0;
    int storage = 0;
    int metric = 0;
    int n_cols = 0;
    uint32_t row_bytes = 0;
    uint32_t chunks_per_row = 0;
    uint32_t q_stride = 0;       // floats per fp32 query row (padded like a storage row)
    int64_t capacity = 0;
    int64_t n_rows = 0;
    int64_t row_base = 0;
    uint64_t search_counter = 0;   // 64-bit: the replay count is counter - epoch[row] and must never wrap
    float h_norm_stats[2] = {0.f, 0.f};   // host mirror of d_max_norm: max ||row||, max | ||row|| - 1 |

    uint8_t* d_vec = nullptr;
    uint8_t* d_live = nullptr;
    uint64_t* d_epoch = nullptr;
    uint64_t* d_tie = nullptr;
    float* d_inv_norm = nullptr;
    uint32_t* d_codes[kMaxFilterCols] = {nullptr};
    float* d_max_norm = nullptr;
    PwProgram* d_pw = nullptr;
    uint32_t* d_counter = nullptr;   // small scratch counters (16 x u32)

    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr};
    float last_ms[4] = {0, 0, 0, 0};
    int last_launches = 0;
    int last_kind = 0;

    Scratch s_qraw, s_q64, s_q32, s_qnorm, s_keys, s_mins, s_flags, s_res, s_stage_dev, s_misc, s_cand;
    Scratch s_gkeys, s_gtops, s_gdrops, s_qb16, s_tickets, s_dbg, s_geps, s_xlocal;
    cudaStream_t last_stream = nullptr; bool last_stream_valid = false; cudaEvent_t order_ev = nullptr;
    uint32_t launch_seq = 0;          // fused scan launches so far (ScanParams::seq)
    bool last_ready_armed = false;    // the last search_core call armed the host-polled completion word
    uint32_t ready_seq = 0;
    uint32_t tile_base[2] = {0, 0};   // what the two tile counters (launch parity) stand at
    Scratch s_qstage;                 // host queries staged by CTA 0, one slot per launch parity
    int opt_dyn_tiles = 1;
    int opt_dbg_times = 0;            // 1: the scan kernel stamps its phases (lvs_last_kernel_phases)
    int opt_inline_query = 1;         // a host query of up to kInlineQueryBytes rides in the kernel's parameter block ...
    int opt_inline_max_mb = 1024;     // ... when the shard is at most this large
    InlineQueries iq;                 // ... assembled here (under the collection's lock)
    Scratch s_dbg_times;
    Scratch h_pin, h_pin2, h_flags;

    // ring of event pairs around the scan launches (read back by lvs_scan_times after a synchronisation)
    cudaEvent_t ring_ev[2 * kEventRing] = {nullptr};
    double ring_bytes[kEventRing] = {0};
    uint64_t ring_pos = 0;
    cudaEvent_t first_scan_start = nullptr, first_scan_end = nullptr;
    int opt_timing = 0;          // 1 = CUDA events around every scan launch (lvs_scan_times / lvs_last_search_timing); this also
                                 // serialises consecutive searches, i.e. switches the programmatic-launch overlap off
    int opt_pdl = 1;
    int last_kpl = 0;

    // pipelined host API (lvs_search_submit / lvs_search_wait)
    struct Slot {
        bool in_use = false;
        int Q = 0, k = 0, dtype = 0, kpl = 0;
        uint64_t base = 0;
        int kind = 1;          // 1 = K1 scan, 2 = K2 tensor-core path (its flagged queries get the K1 fallback in finish_locked)
        bool has_want = false;
        uint32_t want[kMaxFilterCols];
        Scratch h;
        Scratch d;             // device twin of h for batches (staged == true)
        bool staged = false;
        bool poll = false;     // completion is published by the kernel in the slot's `ready` word (no event)
        uint32_t ready_val = 0;
        bool sharded = false;
        size_t qbytes = 0;
        cudaEvent_t done = nullptr;
    } slots[kSubmitSlots];

    int opt_stage_kb = 72;
    int opt_stages = 0;   // 0 = as many as fit
    int opt_grid = 0;     // 0 = one CTA per SM
    int opt_force_kpl = 0;
    int opt_gemm_min_q = 3;   // batches of at least this many queries take the tensor-core path (K2).  One K2 pass (any Q <= 128) costs
                              // about one single-query K1 pass (2.2 ms on 10M x 768 bf16); K1 with 3-4 bf16 queries per pass is
                              // FMA-issue-bound (3.2-3.4 ms) and needs two passes from 5 queries on.  fp32 shards: from 5 queries.
    int opt_path = 0;         // 0 auto, 1 force K1 scan, 2 force K2 (when eligible)
    int opt_gemm_dbg = 0;
    int opt_gemm_stages = 0;
    int opt_gemm_stages_b = -1;   // pair form: corpus buffers of the split rings (-1 = default, 0 = one combined ring)
    int opt_gemm_no_unit = 0;
    int opt_gemm_keep = 16;
    int opt_gemm_no_pair = 0;
    int opt_gemm_no_tf32 = 0;

    // per-row ranking attributes + lower-cased entity-name pool (fused search -> rank path, lvs_search_rank)
    int64_t rk_cap = 0;              // rows covered by the attribute columns
    uint32_t* d_rk_key = nullptr; uint32_t* d_rk_file = nullptr; uint32_t* d_rk_cent = nullptr; uint32_t* d_rk_name = nullptr;
    int32_t* d_rk_clen = nullptr; uint8_t* d_rk_flags = nullptr;
    uint32_t* d_name_off = nullptr; uint8_t* d_name_bytes = nullptr;
    int64_t n_names = 0, names_cap = 0, name_bytes_used = 0, name_bytes_cap = 0;
    Scratch s_rk_dev, h_rk_pin;
    cudaEvent_t rk_ev[3] = {nullptr, nullptr, nullptr};

    std::mutex mu;
};

// ------------------------------------------------------------------------------------------------------
// numpy pairwise_sum program
// ------------------------------------------------------------------------------------------------------
static int pw_build(PwProgram& pg, int off, int n) {
    if (n <= 128) {
        if (pg.n_leaves >= kPwMaxLeaves) return -1;
        const int id = pg.n_leaves++;
        pg.leaf_off[id] = (uint16_t)off;
        pg.leaf_len[id] = (uint16_t)n;
        return id;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int l = pw_build(pg, off, n2);
    if (l < 0) return -1;
    const int r = pw_build(pg, off + n2, n - n2);
    if (r < 0) return -1;
    if (pg.n_steps >= kPwMaxLeaves) return -1;
    pg.step_dst[pg.n_steps] = (uint8_t)l;
    pg.step_src[pg.n_steps] = (uint8_t)r;
    pg.n_steps++;
    return l;
}

// steps a perfect binary tree over leaves [lo, lo + cnt) would have produced (post-order), compared with the program's
static int pw_expect_perfect(const PwProgram& pg, int lo, int cnt, int* step, bool* same) {
    if (cnt == 1) return lo;
    const int l = pw_expect_perfect(pg, lo, cnt / 2, step, same);
    const int r = pw_expect_perfect(pg, lo + cnt / 2, cnt / 2, step, same);
    if (*step >= pg.n_steps || pg.step_dst[*step] != l || pg.step_src[*step] != r) *same = false;
    ++*step;
    return l;
}
// the shapes np_norm_f32's register form handles (finalize_kernel.cuh): dim 384 / 768 / 1024 and the like
static void pw_classify(PwProgram& pg) {
    pg.regular = pg.n_leaves <= 8 ? 1 : 0;
    for (int i = 0; i < pg.n_leaves; ++i) if (pg.leaf_len[i] < 8 || pg.leaf_len[i] % 8 != 0 || pg.leaf_off[i] % 8 != 0) pg.regular = 0;
    pg.balanced = 0;
    if (pg.regular && (pg.n_leaves & (pg.n_leaves - 1)) == 0) {
        int step = 0; bool same = true;
        pw_expect_perfect(pg, 0, pg.n_leaves, &step, &same);
        pg.balanced = same && step == pg.n_steps ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------------
// library entry points
// ------------------------------------------------------------------------------------------------------
extern "C" int lvs_version(void) { return 100; }

extern "C" const char* lvs_last_error(void) { return g_err.c_str(); }

extern "C" int lvs_init(int device) {
    std::lock_guard<std::mutex> lk(g_lib_mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(LVS_ESTATE, "no CUDA device visible (%s); lattice-b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(LVS_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(LVS_ESTATE, "device %d is sm_%d%d; lattice-b200 kernels are built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    g_lib.device = device;
    g_lib.sm_count = prop.multiProcessorCount;
    g_lib.cc_major = prop.major;
    g_lib.cc_minor = prop.minor;
    g_lib.smem_optin = prop.sharedMemPerBlockOptin;
    g_lib.ready = true;
    t_bound_device = device;
    return LVS_OK;
}

extern "C" int lvs_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_lib_mu);
    g_lib.ready = false;
    return LVS_OK;
}

extern "C" int lvs_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem, int64_t* free_mem) {
    bind_thread();
    if (!g_lib.ready) return fail(LVS_ESTATE, "lvs_init() has not been called");
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = g_lib.sm_count;
    if (cc_major) *cc_major = g_lib.cc_major;
    if (cc_minor) *cc_minor = g_lib.cc_minor;
    if (total_mem) *total_mem = (int64_t)t;
    if (free_mem) *free_mem = (int64_t)f;
    return LVS_OK;
}

// ------------------------------------------------------------------------------------------------------
// collections
// ------------------------------------------------------------------------------------------------------
static void free_arrays(lvs_collection* c) {
    cudaFree(c->d_vec); cudaFree(c->d_live); cudaFree(c->d_epoch); cudaFree(c->d_tie); cudaFree(c->d_inv_norm);
    for (int i = 0; i < kMaxFilterCols; ++i) { cudaFree(c->d_codes[i]); c->d_codes[i] = nullptr; }
    c->d_vec = nullptr; c->d_live = nullptr; c->d_epoch = nullptr; c->d_tie = nullptr; c->d_inv_norm = nullptr;
}

static int grow_to(lvs_collection* c, int64_t new_cap) {
    if (new_cap <= c->capacity) return LVS_OK;
    if (new_cap >= (int64_t)0xFFFFFFF0ll) return fail(LVS_ELIMIT, "capacity %lld rows exceeds the 32-bit local row space", (long long)new_cap);
    uint8_t* nv = nullptr; uint8_t* nl = nullptr; uint64_t* ne = nullptr; uint64_t* nt = nullptr; float* ni = nullptr;
    uint32_t* nc[kMaxFilterCols] = {nullptr};
    const size_t vb = (size_t)new_cap * c->row_bytes;
    cudaError_t e = cudaMalloc(&nv, vb);
    if (e == cudaSuccess) e = cudaMalloc(&nl, (size_t)new_cap);
    if (e == cudaSuccess) e = cudaMalloc(&ne, (size_t)new_cap * 8);
    if (e == cudaSuccess) e = cudaMalloc(&nt, (size_t)new_cap * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ni, (size_t)new_cap * 4);
    for (int i = 0; i < c->n_cols && e == cudaSuccess; ++i) e = cudaMalloc(&nc[i], (size_t)new_cap * 4);
    if (e != cudaSuccess) {
        cudaFree(nv); cudaFree(nl); cudaFree(ne); cudaFree(nt); cudaFree(ni);
        for (int i = 0; i < kMaxFilterCols; ++i) cudaFree(nc[i]);
        cudaGetLastError();
        return fail(LVS_ENOMEM, "cannot allocate %lld rows x %u bytes on the device: %s", (long long)new_cap, c->row_bytes,
                    cudaGetErrorString(e));
    }
    cudaStream_t st = c->stream;
    CU(cudaMemsetAsync(nl, 0, (size_t)new_cap, st));
    for (int i = 0; i < c->n_cols; ++i) CU(cudaMemsetAsync(nc[i], 0, (size_t)new_cap * 4, st));
    if (c->n_rows > 0) {
        const size_t n = (size_t)c->n_rows;
        CU(cudaMemcpyAsync(nv, c->d_vec, n * c->row_bytes, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(nl, c->d_live, n, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(ne, c->d_epoch, n * 8, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(nt, c->d_tie, n * 8, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(ni, c->d_inv_norm, n * 4, cudaMemcpyDeviceToDevice, st));
        for (int i = 0; i < c->n_cols; ++i) CU(cudaMemcpyAsync(nc[i], c->d_codes[i], n * 4, cudaMemcpyDeviceToDevice, st));
    }
    CU(cudaStreamSynchronize(st));
    free_arrays(c);
    c->d_vec = nv; c->d_live = nl; c->d_epoch = ne; c->d_tie = nt; c->d_inv_norm = ni;
    for (int i = 0; i < c->n_cols; ++i) c->d_codes[i] = nc[i];
    c->capacity = new_cap;
    return LVS_OK;
}

extern "C" int lvs_collection_create(const char* name, int dim, int storage, int metric, int n_filter_cols,
                                     int64_t capacity_rows, int64_t row_base, lvs_collection** out) {
    bind_thread();
    if (!g_lib.ready) return fail(LVS_ESTATE, "lvs_init() has not been called (no CUDA device bound)");
    if (!out) return fail(LVS_EINVAL, "out is NULL");
    *out = nullptr;
    if (dim < 1 || dim > 8192) return fail(LVS_ELIMIT, "dim %d outside 1..8192", dim);
    if (storage != LVS_STORAGE_F32 && storage != LVS_STORAGE_BF16) return fail(LVS_EINVAL, "unknown storage %d", storage);
    if (metric != LVS_METRIC_COSINE && metric != LVS_METRIC_DOT) return fail(LVS_EINVAL, "unknown metric %d", metric);
    if (n_filter_cols < 0 || n_filter_cols > kMaxFilterCols) return fail(LVS_ELIMIT, "n_filter_cols %d outside 0..%d", n_filter_cols, kMaxFilterCols);
    if (capacity_rows < 0 || row_base < 0) return fail(LVS_EINVAL, "negative capacity or row_base");
    lvs_collection* c = new (std::nothrow) lvs_collection();
    if (!c) return fail(LVS_ENOMEM, "host allocation failed");
    c->name = name ? name : "";
    c->dim = dim; c->storage = storage; c->metric = metric; c->n_cols = n_filter_cols; c->row_base = row_base;
    const int esz = storage == LVS_STORAGE_F32 ? 4 : 2;
    const int epc = 16 / esz;
    const int ld = (dim + epc - 1) / epc * epc;
    c->row_bytes = (uint32_t)(ld * esz);
    c->chunks_per_row = c->row_bytes / 16;
    c->q_stride = (uint32_t)ld;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 2 * kEventRing && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ring_ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_max_norm, 8);
    if (e == cudaSuccess) e = cudaMemset(c->d_max_norm, 0, 8);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counter, 64);
    if (e == cudaSuccess) e = cudaMemset(c->d_counter, 0, 64);
    PwProgram pg;
    memset(&pg, 0, sizeof(pg));
    if (dim < 8) { pg.n_leaves = 1; pg.leaf_off[0] = 0; pg.leaf_len[0] = (uint16_t)dim; }
    else if (pw_build(pg, 0, dim) < 0) { lvs_collection_destroy(c); return fail(LVS_ELIMIT, "dim %d needs too many pairwise blocks", dim); }
    pw_classify(pg);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_pw, sizeof(PwProgram));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_pw, &pg, sizeof(PwProgram), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        lvs_collection_destroy(c);
        return fail(LVS_ECUDA, "collection setup failed: %s", cudaGetErrorString(e));
    }
    int rc = grow_to(c, std::max<int64_t>(capacity_rows, 1024));
    if (rc != LVS_OK) { lvs_collection_destroy(c); return rc; }
    *out = c;
    return LVS_OK;
}

extern "C" int lvs_collection_destroy(lvs_collection* c) {
    bind_thread();
    if (!c) return LVS_OK;
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_rk_key); cudaFree(c->d_rk_file); cudaFree(c->d_rk_cent); cudaFree(c->d_rk_name); cudaFree(c->d_rk_clen);
    cudaFree(c->d_rk_flags); cudaFree(c->d_name_off); cudaFree(c->d_name_bytes);
    if (c->s_rk_dev.p) cudaFree(c->s_rk_dev.p);
    if (c->h_rk_pin.p) cudaFreeHost(c->h_rk_pin.p);
    for (auto& e : c->rk_ev) if (e) cudaEventDestroy(e);
    free_arrays(c);
    cudaFree(c->d_max_norm); cudaFree(c->d_pw); cudaFree(c->d_counter);
    Scratch* ds[] = {&c->s_qraw, &c->s_q64, &c->s_q32, &c->s_qnorm, &c->s_keys, &c->s_mins, &c->s_flags, &c->s_res, &c->s_stage_dev, &c->s_misc,
                     &c->s_gkeys, &c->s_gtops, &c->s_gdrops, &c->s_qb16, &c->s_tickets, &c->s_dbg, &c->s_geps, &c->s_xlocal, &c->s_qstage, &c->s_dbg_times};
    if (c->order_ev) cudaEventDestroy(c->order_ev);
    for (Scratch* s : ds) if (s->p) cudaFree(s->p);
    if (c->h_pin.p) cudaFreeHost(c->h_pin.p);
    if (c->h_pin2.p) cudaFreeHost(c->h_pin2.p);
    for (int i = 0; i < 6; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2 * kEventRing; ++i) if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
    if (c->s_cand.p) cudaFree(c->s_cand.p);
    for (auto& sl : c->slots) {
        if (sl.h.p) cudaFreeHost(sl.h.p);
        if (sl.d.p) cudaFree(sl.d.p);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    if (c->h_flags.p) cudaFreeHost(c->h_flags.p);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return LVS_OK;
}

extern "C" int lvs_collection_reserve(lvs_collection* c, int64_t capacity_rows) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    return grow_to(c, capacity_rows);
}

extern "C" int64_t lvs_rows(const lvs_collection* c) {
    bind_thread(); return c ? c->n_rows : 0; }
extern "C" int64_t lvs_capacity(const lvs_collection* c) {
    bind_thread(); return c ? c->capacity : 0; }
extern "C" uint64_t lvs_search_counter(const lvs_collection* c) {
    bind_thread(); return c ? c->search_counter : 0; }
extern "C" int lvs_advance_search_counter(lvs_collection* c, uint64_t n) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    for (auto& sl : c->slots) if (sl.in_use) return fail(LVS_ESTATE, "searches are in flight: call lvs_search_wait first");
    c->search_counter += n;
    return LVS_OK;
}

__global__ void count_live_kernel(const uint8_t* live, uint32_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += live[i];
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

extern "C" int64_t lvs_count(const lvs_collection* cc) {
    bind_thread();
    lvs_collection* c = const_cast<lvs_collection*>(cc);
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->n_rows == 0) return 0;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(c->d_counter + 8);
    unsigned long long h = 0;
    if (cudaMemsetAsync(d, 0, 8, c->stream) != cudaSuccess) return -1;
    count_live_kernel<<<std::min<int64_t>(1024, (c->n_rows + 255) / 256), 256, 0, c->stream>>>(c->d_live, (uint32_t)c->n_rows, d);
    if (cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    return (int64_t)h;
}

// ------------------------------------------------------------------------------------------------------
// upsert
// ------------------------------------------------------------------------------------------------------
static size_t dt_size(int dtype) { return dtype == LVS_DT_F64 ? 8 : dtype == LVS_DT_F32 ? 4 : 2; }

static int launch_upsert(lvs_collection* c, const void* d_src, int dtype, int64_t n, const int64_t* d_rows, int64_t row0,
                         const uint32_t* d_codes, const uint64_t* d_ties, cudaStream_t st) {
    UpsertParams p;
    memset(&p, 0, sizeof(p));
    p.src = d_src; p.src_dtype = dtype; p.n = n; p.rows = d_rows; p.row0 = row0;
    p.base = c->d_vec; p.row_bytes = c->row_bytes; p.dim = c->dim; p.storage = c->storage; p.metric = c->metric;
    p.inv_norm = c->d_inv_norm; p.live = c->d_live; p.epoch = c->d_epoch; p.epoch_val = c->search_counter;
    p.tiekey = c->d_tie; p.ties_src = d_ties; p.row_base = c->row_base;
    for (int i = 0; i < c->n_cols; ++i) p.codes[i] = c->d_codes[i];
    p.codes_src = d_codes; p.n_cols = c->n_cols; p.max_norm = c->d_max_norm;
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)g_lib.sm_count * 16);
    upsert_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, st>>>(p);
    CU(cudaGetLastError());
    return LVS_OK;
}

// Shared body of lvs_upsert (host vectors) and lvs_upsert_device_vectors (vectors already in HBM: the encoder's output): row numbers,
// tie keys and filter codes are staged through pinned memory either way.
static int upsert_impl(lvs_collection* c, const void* vecs, const void* d_vecs, int dtype, int64_t n, const int64_t* rows,
                       const uint32_t* codes, const uint64_t* ties) {
    std::lock_guard<std::mutex> lk(c->mu);
    int64_t hi = c->n_rows;
    int64_t row0 = c->n_rows;
    if (rows) {
        for (int64_t i = 0; i < n; ++i) {
            if (rows[i] < c->row_base) return fail(LVS_EINVAL, "rows[%lld] is below the shard's row_base", (long long)i);
            hi = std::max(hi, rows[i] - c->row_base + 1);
        }
    } else {
        hi = row0 + n;
    }
    if (hi > c->capacity) {
        int rc = grow_to(c, std::max(hi, c->capacity * 2));
        if (rc != LVS_OK) return rc;
    }
    // stage through pinned memory in batches of <= 32 MiB of vector data
    const size_t vrow = (size_t)c->dim * dt_size(dtype);
    const size_t vstage = d_vecs ? 0 : vrow;                       // device vectors are read in place
    const int64_t batch = std::max<int64_t>(1, (int64_t)((32u << 20) / vrow));
    const size_t per_row = vstage + 8 + 8 + (size_t)c->n_cols * 4;
    const int64_t b0 = std::min(batch, n);
    int rc = ensure_pinned(c->h_pin, (size_t)b0 * per_row + 64);
    if (rc != LVS_OK) return rc;
    rc = ensure_dev(c->s_stage_dev, (size_t)b0 * per_row + 64);
    if (rc != LVS_OK) return rc;
    cudaStream_t st = c->stream;
    for (int64_t s = 0; s < n; s += batch) {
        const int64_t m = std::min(batch, n - s);
        uint8_t* hp = (uint8_t*)c->h_pin.p;
        uint8_t* dp = (uint8_t*)c->s_stage_dev.p;
        size_t o_vec = 0;
        size_t o_rows = ((size_t)m * vstage + 15) & ~(size_t)15;
        size_t o_ties = o_rows + (size_t)m * 8;
        size_t o_codes = o_ties + (size_t)m * 8;
        size_t total = o_codes + (size_t)m * c->n_cols * 4;
        if (!d_vecs) memcpy(hp + o_vec, (const uint8_t*)vecs + (size_t)s * vrow, (size_t)m * vrow);
        if (rows) {
            int64_t* lr = (int64_t*)(hp + o_rows);
            for (int64_t j = 0; j < m; ++j) lr[j] = rows[s + j] - c->row_base;   // global -> local
        }
        if (ties) memcpy(hp + o_ties, ties + s, (size_t)m * 8);
        if (codes && c->n_cols) memcpy(hp + o_codes, codes + (size_t)s * c->n_cols, (size_t)m * c->n_cols * 4);
        CU(cudaMemcpyAsync(dp, hp, total, cudaMemcpyHostToDevice, st));
        const void* src = d_vecs ? (const void*)((const uint8_t*)d_vecs + (size_t)s * vrow) : (const void*)(dp + o_vec);
        rc = launch_upsert(c, src, dtype, m, rows ? (const int64_t*)(dp + o_rows) : nullptr, row0 + s,
                           (codes && c->n_cols) ? (const uint32_t*)(dp + o_codes) : nullptr,
                           ties ? (const uint64_t*)(dp + o_ties) : nullptr, st);
        if (rc != LVS_OK) return rc;
        CU(cudaStreamSynchronize(st));   // the pinned buffer is reused by the next batch
    }
    c->n_rows = hi;
    CU(cudaMemcpy(c->h_norm_stats, c->d_max_norm, 8, cudaMemcpyDeviceToHost));
    return LVS_OK;
}

extern "C" int lvs_upsert(lvs_collection* c, const void* vecs, int dtype, int64_t n, const int64_t* rows,
                          const uint32_t* codes, const uint64_t* ties) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n < 0 || (n > 0 && !vecs)) return fail(LVS_EINVAL, "bad vecs / n");
    if (dtype != LVS_DT_F32 && dtype != LVS_DT_F64 && dtype != LVS_DT_BF16) return fail(LVS_EINVAL, "unknown dtype %d", dtype);
    if (n == 0) return LVS_OK;
    return upsert_impl(c, vecs, nullptr, dtype, n, rows, codes, ties);
}

int lvs_upsert_device_vectors(lvs_collection* c, const void* d_vecs, int dtype, int64_t n, const int64_t* rows, const uint32_t* codes,
                              const uint64_t* ties) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n < 0 || (n > 0 && !d_vecs)) return fail(LVS_EINVAL, "bad d_vecs / n");
    if (n == 0) return LVS_OK;
    // the vectors were produced on another stream: the caller has synchronised it (encoder_api.cu does)
    return upsert_impl(c, nullptr, d_vecs, dtype, n, rows, codes, ties);
}

extern "C" int lvs_upsert_device(lvs_collection* c, const void* d_vecs, int dtype, int64_t n, int64_t row0,
                                 const uint32_t* d_codes, const uint64_t* d_ties, void* stream) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n < 0 || (n > 0 && !d_vecs)) return fail(LVS_EINVAL, "bad d_vecs / n");
    if (dtype != LVS_DT_F32 && dtype != LVS_DT_F64 && dtype != LVS_DT_BF16) return fail(LVS_EINVAL, "unknown dtype %d", dtype);
    if (n == 0) return LVS_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    row0 -= c->row_base;   // global -> local
    if (row0 < 0) return fail(LVS_EINVAL, "row0 is below the shard's row_base");
    const int64_t hi = std::max(c->n_rows, row0 + n);
    if (hi > c->capacity) {
        int rc = grow_to(c, std::max(hi, c->capacity * 2));
        if (rc != LVS_OK) return rc;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = launch_upsert(c, d_vecs, dtype, n, nullptr, row0, d_codes, d_ties, st);
    if (rc != LVS_OK) return rc;
    CU(cudaStreamSynchronize(st));
    c->n_rows = hi;
    CU(cudaMemcpy(c->h_norm_stats, c->d_max_norm, 8, cudaMemcpyDeviceToHost));
    return LVS_OK;
}

extern "C" int lvs_set_codes(lvs_collection* c, int col, const int64_t* rows, int64_t row0, int64_t n, const uint32_t* codes) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (col < 0 || col >= c->n_cols) return fail(LVS_EINVAL, "filter column %d outside 0..%d", col, c->n_cols - 1);
    if (n <= 0) return LVS_OK;
    if (!codes) return fail(LVS_EINVAL, "codes is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    std::vector<int64_t> local;
    if (rows) {
        local.assign(rows, rows + n);
        for (int64_t i = 0; i < n; ++i) { local[i] -= c->row_base; if (local[i] < 0 || local[i] >= c->n_rows) return fail(LVS_EINVAL, "rows[%lld] out of range", (long long)i); }
        rows = local.data();
    } else {
        row0 -= c->row_base;
        if (row0 < 0 || row0 + n > c->n_rows) return fail(LVS_EINVAL, "row range out of bounds");
    }
    const size_t bytes = (size_t)n * 12 + 16;
    int rc = ensure_dev(c->s_stage_dev, bytes);
    if (rc != LVS_OK) return rc;
    uint8_t* dp = (uint8_t*)c->s_stage_dev.p;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(dp, codes, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const size_t o_rows = ((size_t)n * 4 + 15) & ~(size_t)15;
    if (rows) CU(cudaMemcpyAsync(dp + o_rows, rows, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    set_codes_kernel<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, st>>>(
        c->d_codes[col], rows ? (const int64_t*)(dp + o_rows) : nullptr, row0, (const uint32_t*)dp, n);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return LVS_OK;
}

// ------------------------------------------------------------------------------------------------------
// delete / match
// ------------------------------------------------------------------------------------------------------
extern "C" int lvs_delete_rows(lvs_collection* c, const int64_t* rows, int64_t n, int64_t* n_deleted) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n_deleted) *n_deleted = 0;
    if (n <= 0) return LVS_OK;
    if (!rows) return fail(LVS_EINVAL, "rows is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = ensure_dev(c->s_stage_dev, (size_t)n * 8);
    if (rc != LVS_OK) return rc;
    cudaStream_t st = c->stream;
    std::vector<int64_t> local(rows, rows + n);
    for (auto& r : local) r -= c->row_base;
    CU(cudaMemcpyAsync(c->s_stage_dev.p, local.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->d_counter, 0, 4, st));
    set_live_kernel<<<(unsigned)std::min<int64_t>(1024, (n + 255) / 256), 256, 0, st>>>(c->d_live, (const int64_t*)c->s_stage_dev.p, n, 0,
                                                                                    (uint32_t)c->n_rows, c->d_counter);
    CU(cudaGetLastError());
    uint32_t h = 0;
    CU(cudaMemcpyAsync(&h, c->d_counter, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n_deleted) *n_deleted = h;
    return LVS_OK;
}

// Compaction: after a mass delete (projects/cleanup.py:38-73 removes a whole project) the tombstones still cost
// scan bandwidth; the host moves the live rows of the tail into the holes and truncates.
extern "C" int lvs_move_rows(lvs_collection* c, const int64_t* src, const int64_t* dst, int64_t n) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n <= 0) return LVS_OK;
    if (!src || !dst) return fail(LVS_EINVAL, "NULL row list");
    std::lock_guard<std::mutex> lk(c->mu);
    std::vector<int64_t> both((size_t)2 * n);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t s_ = src[i] - c->row_base, d_ = dst[i] - c->row_base;
        if (s_ < 0 || s_ >= c->n_rows || d_ < 0 || d_ >= c->n_rows || s_ == d_) return fail(LVS_EINVAL, "move %lld -> %lld is outside the shard", (long long)src[i], (long long)dst[i]);
        both[i] = s_; both[n + i] = d_;
    }
    {   // the two sets must be disjoint and free of repeats (a row moved twice would depend on the order)
        std::vector<int64_t> chk(both);
        std::sort(chk.begin(), chk.end());
        if (std::adjacent_find(chk.begin(), chk.end()) != chk.end()) return fail(LVS_EINVAL, "source and destination rows must be distinct");
    }
    int rc = ensure_dev(c->s_stage_dev, (size_t)2 * n * 8);
    if (rc != LVS_OK) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->s_stage_dev.p, both.data(), (size_t)2 * n * 8, cudaMemcpyHostToDevice, st));
    MoveRowsParams p;
    memset(&p, 0, sizeof(p));
    p.src = (const int64_t*)c->s_stage_dev.p; p.dst = p.src + n; p.n = n;
    p.vec = c->d_vec; p.row_bytes = c->row_bytes; p.live = c->d_live; p.epoch = c->d_epoch; p.tie = c->d_tie; p.inv_norm = c->d_inv_norm;
    for (int i = 0; i < c->n_cols; ++i) p.codes[i] = c->d_codes[i];
    p.n_cols = c->n_cols;
    p.rk_key = c->d_rk_key; p.rk_file = c->d_rk_file; p.rk_cent = c->d_rk_cent; p.rk_name = c->d_rk_name; p.rk_clen = c->d_rk_clen;
    p.rk_flags = c->d_rk_flags; p.rk_rows = c->rk_cap;
    move_rows_kernel<<<(unsigned)std::min<int64_t>(4096, (n + 7) / 8), 256, 0, st>>>(p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return LVS_OK;
}

extern "C" int lvs_truncate(lvs_collection* c, int64_t n_rows) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    if (n_rows < 0 || n_rows > c->n_rows) return fail(LVS_EINVAL, "cannot truncate %lld rows to %lld", (long long)c->n_rows, (long long)n_rows);
    if (n_rows == c->n_rows) return LVS_OK;
    // every dropped row must be a tombstone
    cudaStream_t st = c->stream;
    const int64_t m = c->n_rows - n_rows;
    std::vector<uint8_t> lv((size_t)m);
    CU(cudaMemcpyAsync(lv.data(), c->d_live + n_rows, (size_t)m, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < m; ++i) if (lv[i]) return fail(LVS_ESTATE, "row %lld is still live: move it before truncating", (long long)(c->row_base + n_rows + i));
    c->n_rows = n_rows;
    return LVS_OK;
}

static int build_filter(const lvs_collection* c, const uint32_t* want, const uint32_t** codes, uint32_t* wantv, uint32_t* nf) {
    *nf = 0;
    if (!want) return LVS_OK;
    for (int i = 0; i < c->n_cols; ++i) {
        if (want[i] == kAnyCode) continue;
        codes[*nf] = c->d_codes[i];
        wantv[*nf] = want[i];
        (*nf)++;
    }
    return LVS_OK;
}

static int match_impl(lvs_collection* c, const uint32_t* want, int64_t* out_rows, int64_t cap, int64_t* n_matched, int tombstone) {
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (cap < 0 || (cap > 0 && !out_rows)) return fail(LVS_EINVAL, "bad out_rows / cap");
    if (cap > 0xFFFFFFF0ll) cap = 0xFFFFFFF0ll;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n_matched) *n_matched = 0;
    if (c->n_rows == 0) return LVS_OK;
    MatchParams p;
    memset(&p, 0, sizeof(p));
    build_filter(c, want, p.codes, p.want, &p.n_filter);
    int rc = ensure_dev(c->s_res, (size_t)std::max<int64_t>(cap, 1) * 8);
    if (rc != LVS_OK) return rc;
    cudaStream_t st = c->stream;
    p.n_rows = (uint32_t)c->n_rows; p.live = c->d_live; p.row_base = c->row_base;
    p.out_rows = (int64_t*)c->s_res.p; p.cap = (uint32_t)cap; p.counter = c->d_counter; p.tombstone = tombstone; p.live_rw = c->d_live;
    CU(cudaMemsetAsync(c->d_counter, 0, 4, st));
    match_rows_kernel<<<(unsigned)std::min<int64_t>((int64_t)g_lib.sm_count * 8, (c->n_rows + 255) / 256), 256, 0, st>>>(p);
    CU(cudaGetLastError());
    uint32_t h = 0;
    CU(cudaMemcpyAsync(&h, c->d_counter, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int64_t ncopy = std::min<int64_t>(h, cap);
    if (ncopy > 0) CU(cudaMemcpy(out_rows, c->s_res.p, (size_t)ncopy * 8, cudaMemcpyDeviceToHost));
    if (n_matched) *n_matched = h;
    return LVS_OK;
}

extern "C" int lvs_delete_where(lvs_collection* c, const uint32_t* want, int64_t* out_rows, int64_t cap, int64_t* n_matched) {
    bind_thread();
    return match_impl(c, want, out_rows, cap, n_matched, 1);
}
extern "C" int lvs_match_rows(lvs_collection* c, const uint32_t* want, int64_t* out_rows, int64_t cap, int64_t* n_matched) {
    bind_thread();
    return match_impl(c, want, out_rows, cap, n_matched, 0);
}

// ------------------------------------------------------------------------------------------------------
// search
// ------------------------------------------------------------------------------------------------------
// The 36 instantiations of the fused scan kernel are compiled in three translation units of their own (scan_f32.cu,
// scan_bf16_cos.cu, scan_bf16_dot.cu: build.py compiles every .cu in parallel), through scan_launch.cuh.
static cudaError_t launch_scan(const lvs_collection* c, int qt, int kpl, bool filter, const ScanParams& p, const FinalizeParams& fp,
                               const ExchangeParams& xp, const InlineQueries& iq, int grid, size_t smem, cudaStream_t st) {
    const size_t optin = g_lib.smem_optin;
    if (c->storage == LVS_STORAGE_F32) return lvs_launch_scan_f32(qt, kpl, filter, p, fp, xp, iq, grid, smem, st, optin);
    if (c->metric == LVS_METRIC_COSINE) return lvs_launch_scan_bf16_cos(qt, kpl, filter, p, fp, xp, iq, grid, smem, st, optin);
    return lvs_launch_scan_bf16_dot(qt, kpl, filter, p, fp, xp, iq, grid, smem, st, optin);
}

template <int KPL>
static cudaError_t launch_finalize(const FinalizeParams& fp, int nq, unsigned ncta, size_t smem, cudaStream_t st) {
    static std::once_flag once;
    static cudaError_t once_err = cudaSuccess;
    std::call_once(once, [] { once_err = cudaFuncSetAttribute(finalize_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_lib.smem_optin); });
    if (once_err != cudaSuccess) return once_err;
    finalize_kernel<KPL><<<dim3((unsigned)nq, ncta), kFinThreads, smem, st>>>(fp);
    return cudaGetLastError();
}

static int max_qt_for_kpl(int kpl) { return kpl >= 8 ? 1 : kpl >= 4 ? 2 : 4; }

struct ScanGeom {
    uint32_t stage_rows, n_stages, stage_bytes;
    size_t smem;
};

static int scan_geometry(const lvs_collection* c, int qt, bool filter, size_t fin_bytes, ScanGeom* g) {
    const size_t budget = g_lib.smem_optin - kScanStaticSmem;   // 227 KB on B200, minus the kernel's static shared memory
    uint32_t target = (uint32_t)std::max(4, c->opt_stage_kb) * 1024u;
    uint32_t R = std::max<uint32_t>(kScanRW, (target / c->row_bytes) / kScanRW * kScanRW);
    for (;;) {
        const uint32_t sb = R * c->row_bytes;
        uint32_t S = kScanMaxStages;
        if (c->opt_stages > 0) S = std::min<uint32_t>(S, (uint32_t)c->opt_stages);
        while (S >= 2 && scan_smem_bytes(S, sb, qt, c->q_stride, R, filter, fin_bytes) > budget) --S;
        if (S >= 2) {
            g->stage_rows = R; g->n_stages = S; g->stage_bytes = sb;
            g->smem = scan_smem_bytes(S, sb, qt, c->q_stride, R, filter, fin_bytes);
            return LVS_OK;
        }
        if (R <= (uint32_t)kScanRW) break;
        R -= kScanRW;
    }
    return fail(LVS_ELIMIT, "dim %d does not fit the scan's shared-memory ring", c->dim);
}

struct lvs_exchange {
    int world = 1, rank = 0, max_q = 0, max_k = 0;
    size_t blk_stride = 0;          // int64 elements per rank block
    size_t slot_elems = 0;          // int64 elements per slot = world * blk_stride
    uint8_t* base = nullptr;        // [2 slots][world][blk_stride] int64, then flags [2][kMaxRanks] uint64
    size_t flags_off = 0;
    uint8_t* peers[kMaxRanks] = {nullptr};
    bool opened[kMaxRanks] = {false};
    uint32_t* d_counter = nullptr;  // [0] done counter, [1] error word
    uint64_t seq = 0;
    bool connected = false;
};

// Parameters of the next exchange of (Q queries x k results) on `ex`: advances the sequence number (every rank must issue the
// same sequence of exchanges) and selects the slot by its parity.
static int exchange_params(lvs_exchange* ex, int Q, int k, ExchangeParams* out) {
    if (!ex->connected && ex->world > 1) return fail(LVS_ESTATE, "lvs_exchange_connect() has not been called");
    if (Q < 1 || k < 1 || (size_t)3 * Q * k + (size_t)Q > ex->blk_stride)
        return fail(LVS_ELIMIT, "Q x k = %d x %d exceeds the exchange buffer (%d x %d)", Q, k, ex->max_q, ex->max_k);
    if ((size_t)ex->world * k * 24 > 200 * 1024) return fail(LVS_ELIMIT, "world * k too large for the merge");
    ex->seq += 1;
    const int slot = (int)(ex->seq & 1);
    ExchangeParams& p = *out;
    p.Q = Q; p.k = k; p.world = ex->world; p.rank = ex->rank;
    for (int r = 0; r < ex->world; ++r) {
        p.peer_data[r] = (int64_t*)ex->peers[r] + (size_t)slot * ex->slot_elems;
        p.peer_flags[r] = (uint64_t*)(ex->peers[r] + ex->flags_off) + (size_t)slot * kMaxRanks;
    }
    p.my_data = (const int64_t*)ex->base + (size_t)slot * ex->slot_elems;
    p.my_flags = (const uint64_t*)(ex->base + ex->flags_off) + (size_t)slot * kMaxRanks;
    p.seq = ex->seq; p.blk_stride = ex->blk_stride; p.done_counter = ex->d_counter; p.err = ex->d_counter + 1;
    return LVS_OK;
}

static int launch_exchange(const ExchangeParams& p, cudaStream_t st) {
    const size_t smem = (size_t)p.world * p.k * 24;
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(p.Q, 2 * g_lib.sm_count));      // a CTA merges one query at a time
    exchange_merge_kernel<<<grid, 256, smem, st>>>(p);
    CU(cudaGetLastError());
    return LVS_OK;
}

// Where a level's results go.  Local form: the four Q x k / Q arrays.  Sharded form (ex != nullptr): every group of queries ends in
// an exchange with the other ranks and the MERGED lists land in xout ([3][Q][k] int64: score bits | global rows | tie keys).
struct LevelOut {
    double* scores; int64_t* rows; uint64_t* ties; uint32_t* counts; int32_t* flags;
    lvs_exchange* ex; int64_t* xout; int Q_total;
    uint32_t* host_ready; uint32_t host_ready_val;   // armed only when the level is ONE launch (see search_core)
};

// Enqueue one level of the search for the query indices in `pending` (ascending): ONE fused kernel (query prep + scan + exact
// rescoring [+ exchange and merge]) per group of up to max_qt consecutive queries.  No synchronisation.
static int enqueue_level(lvs_collection* c, const void* d_queries, const void* h_queries, int dtype, const std::vector<int>& pending, int k, int kpl, bool filter,
                         const uint32_t* const* fcodes, const uint32_t* fwant, uint32_t nf, uint64_t search_base, const LevelOut& out,
                         cudaStream_t st, int* launches, bool time_first) {
    const int sm = g_lib.sm_count;
    uint64_t* keys = (uint64_t*)c->s_keys.p;
    uint64_t* mins = (uint64_t*)c->s_mins.p;
    const int chain = (int)((c->chunks_per_row + 31) / 32) * (c->storage == LVS_STORAGE_F32 ? 4 : 8);
    const float eps_rel = (float)(chain + 12) * 1.1920929e-7f;
    const int max_qt = max_qt_for_kpl(kpl);
    const size_t qrow = (size_t)c->dim * dt_size(dtype);
    int rc;
    size_t i = 0;
    bool first_group = time_first;
    while (i < pending.size()) {
        // a group = up to max_qt CONSECUTIVE query indices; kernels exist for 1, 2 and 4 query slots
        // (a 3-query group runs the 4-slot kernel, whose last slot scores an all-zero query and is ignored)
        const int first = pending[i];
        int cnt = 1;
        while (cnt < max_qt && i + cnt < pending.size() && pending[i + cnt] == first + cnt) ++cnt;
        const int qt_use = cnt == 1 ? 1 : cnt == 2 ? 2 : 4;
        const uint32_t kpw = 32u * kpl;

        FinalizeParams fp;
        memset(&fp, 0, sizeof(fp));
        fp.kp = kpw; fp.k = (uint32_t)k;
        fp.base = c->d_vec; fp.row_bytes = c->row_bytes; fp.dim = c->dim; fp.dim_pad = (int)c->q_stride;
        fp.storage = c->storage; fp.metric = c->metric;
        fp.tiekey = c->d_tie; fp.epoch = c->d_epoch; fp.search_no = search_base + (uint64_t)first;
        fp.pw = c->d_pw; fp.eps = eps_rel; fp.row_base = c->row_base;
        int nrw = kFinWarps;
        const size_t xbytes = out.ex ? (size_t)kMaxRanks * k * 24 : 0;       // merge scratch behind the finalize carve-up
        const size_t fin_budget = g_lib.smem_optin - kScanStaticSmem - xbytes;
        while (nrw > 1 && finalize_smem_bytes(fp.dim_pad, nrw) > fin_budget) --nrw;
        if (finalize_smem_bytes(fp.dim_pad, nrw) > fin_budget) return fail(LVS_ELIMIT, "dim %d does not fit the rescoring buffers", c->dim);
        fp.n_rescore_warps = nrw;
        fp.cand_scores = (double*)c->s_cand.p; fp.tickets = c->d_counter + 12;   // 4 tickets (one per query slot)
        if (!out.ex) {
            fp.out_scores = out.scores + (size_t)first * k; fp.out_rows = out.rows + (size_t)first * k; fp.out_ties = out.ties + (size_t)first * k;
            fp.out_flags = out.flags + first; fp.out_counts = out.counts + first;
        }
        fp.qnorm = nullptr; fp.max_norm = c->d_max_norm;
        const size_t fin_bytes = finalize_smem_bytes(fp.dim_pad, nrw) + xbytes;

        ScanGeom g;
        if ((rc = scan_geometry(c, qt_use, filter, fin_bytes, &g)) != LVS_OK) return rc;
        ScanParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.base = c->d_vec; sp.row_bytes = c->row_bytes; sp.chunks_per_row = c->chunks_per_row;
        sp.n_rows = (uint32_t)c->n_rows; sp.stage_rows = g.stage_rows; sp.n_stages = g.n_stages; sp.stage_bytes = g.stage_bytes;
        sp.n_tiles = (uint32_t)((c->n_rows + g.stage_rows - 1) / g.stage_rows);
        sp.n_blocks32 = (uint32_t)((c->n_rows + 31) / 32);
        sp.q_raw = (const uint8_t*)d_queries + (size_t)first * qrow; sp.q_dtype = dtype; sp.dim = c->dim; sp.metric = c->metric;
        sp.n_queries = (uint32_t)cnt; sp.q_stride = c->q_stride;
        const uint32_t seq = c->launch_seq + 1;
        sp.seq = seq; sp.done_seq = c->d_counter + 10;
        InlineQueries& iq = c->iq;        // only the first cnt * qrow bytes mean anything; the launch copies the block
        // One unfiltered host query on a shard that is scanned in well under a millisecond - where the microseconds count - travels
        // in the kernel's parameter block (scan_kernel.cuh).  Longer scans let CTA 0 stage the query instead: 3 us more in front of
        // a millisecond, and their launches keep the classic sub-4 KB parameter block (pipelined host searches on a 5M-row shard
        // measured the same either way: profiles/r02_inline_overlap_probe.jsonl).
        const bool short_scan = (double)c->n_rows * c->row_bytes <= (double)c->opt_inline_max_mb * 1048576.0;
        if (h_queries != nullptr && c->opt_inline_query && cnt == 1 && !filter && short_scan && qrow <= kInlineQueryBytes) {
            sp.q_inline = 1u; sp.q_raw = nullptr;
            memcpy(iq.bytes, (const uint8_t*)h_queries + (size_t)first * qrow, (size_t)cnt * qrow);
        } else if (h_queries != nullptr) {
            // mapped pinned host memory: CTA 0 stages the group's queries in HBM for the other CTAs (scan_kernel.cuh)
            const size_t slot_bytes = (size_t)4 * c->dim * 8;
            if ((rc = ensure_dev(c->s_qstage, 2 * slot_bytes)) != LVS_OK) return rc;
            sp.q_host = sp.q_raw;
            sp.q_stage = (uint8_t*)c->s_qstage.p + (size_t)(seq & 1u) * slot_bytes;
            sp.q_raw = sp.q_stage; sp.q_flag = c->d_counter + 11; sp.q_bytes = (uint32_t)((size_t)cnt * qrow);
        }
        sp.live = c->d_live;
        for (uint32_t f = 0; f < nf; ++f) { sp.codes[f] = fcodes[f]; sp.want[f] = fwant[f]; }
        sp.n_filter = nf;
        sp.out_keys = keys; sp.out_tops = mins;
        int grid = c->opt_grid > 0 ? c->opt_grid : sm;
        const uint32_t units = filter ? sp.n_blocks32 : sp.n_tiles;
        grid = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)grid, units));
        // keys of query slot s of this group live at keys + s*grid*kpw
        fp.keys = keys; fp.tops = mins; fp.M = (uint32_t)grid * kpw; fp.L = (uint32_t)grid;
        // CTAs per query for the rescoring: as many as it takes to give every candidate its own warp, at most 8 CTAs in all
        const uint32_t cq = std::max<uint32_t>(1u, std::min<uint32_t>(8u / (uint32_t)cnt, (kpw + nrw - 1) / nrw));
        sp.ticket = c->d_counter + 4; sp.n_helpers = (uint32_t)cnt * cq;
        sp.host_ready = out.host_ready; sp.host_ready_val = out.host_ready_val;
        if (c->opt_dbg_times) {
            if ((rc = ensure_dev(c->s_dbg_times, 128)) != LVS_OK) return rc;
            CU(cudaMemsetAsync(c->s_dbg_times.p, 0xFF, 8, st));                       // slot 0 takes a minimum
            CU(cudaMemsetAsync((uint8_t*)c->s_dbg_times.p + 8, 0, 120, st));
            sp.dbg_times = (unsigned long long*)c->s_dbg_times.p; fp.dbg_times = sp.dbg_times;
        }
        const bool dyn = !filter && grid == sm && c->opt_dyn_tiles;       // dynamic tile scheduling needs one CTA per SM (see the kernel)
        if (dyn) { sp.tile_counter = c->d_counter + 6 + (seq & 1u); sp.tile_base = c->tile_base[seq & 1u]; }
        sp.pdl = c->opt_pdl && !c->opt_timing ? 1u : 0u;
        ExchangeParams xp;
        memset(&xp, 0, sizeof(xp));
        if (out.ex) {
            if ((rc = exchange_params(out.ex, cnt, k, &xp)) != LVS_OK) return rc;
            xp.out = out.xout; xp.out_counts = out.counts; xp.out_flags = out.flags; xp.Q_out = out.Q_total; xp.q_out0 = first;
            sp.exchange = 1u;
        }
        cudaEvent_t es = nullptr, ee = nullptr;
        if (c->opt_timing) {
            const int slot = c->ring_pos % kEventRing;
            es = c->ring_ev[2 * slot]; ee = c->ring_ev[2 * slot + 1];
            c->ring_bytes[slot] = (double)c->n_rows * c->row_bytes;
            c->ring_pos++;
            CU(cudaEventRecord(es, st));
        }
        cudaError_t e = launch_scan(c, qt_use, kpl, filter, sp, fp, xp, iq, grid, g.smem, st);
        if (e != cudaSuccess) return fail(LVS_ECUDA, "scan kernel launch failed: %s (qt=%d kpl=%d smem=%zu)", cudaGetErrorString(e), qt_use, kpl, g.smem);
        ++*launches;
        c->launch_seq = seq;
        if (dyn) c->tile_base[seq & 1u] += sp.n_tiles + (uint32_t)grid;     // every producer draws exactly one tile past the end
        if (c->opt_timing) CU(cudaEventRecord(ee, st));
        if (first_group) {
            c->first_scan_start = es; c->first_scan_end = ee;
            if (c->opt_timing) CU(cudaEventRecord(c->ev[4], st));
            first_group = false;
        }
        i += cnt;
    }
    return LVS_OK;
}

// ---- K2: tensor-core path -------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;

static int get_encode_tiled() {
    if (g_encode_tiled) return LVS_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(LVS_ECUDA, "cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
    g_encode_tiled = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    return LVS_OK;
}

int lvs_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
bool lvs_lib_ready() { return g_lib.ready; }
int lvs_lib_sm_count() { return g_lib.sm_count; }
size_t lvs_lib_smem_optin() { return g_lib.smem_optin; }
void lvs_lib_bind_thread() { bind_thread(); }
PFN_cuTensorMapEncodeTiled_v12000 lvs_lib_encode_tiled() { return get_encode_tiled() == LVS_OK ? g_encode_tiled : nullptr; }

static bool gemm_eligible(const lvs_collection* c, int Q, bool filter) {
    (void)filter;                 // K2 evaluates the payload filter next to the tombstone check
    if (c->opt_path == 1) return false;
    if (c->storage != LVS_STORAGE_BF16 && c->opt_gemm_no_tf32) return false;   // fp32 shards: kind::tf32
    const uint32_t kce = c->storage == LVS_STORAGE_BF16 ? kGemmKC : kGemmKC / 2;
    const uint32_t nk = (c->q_stride + kce - 1) / kce;
    if (nk > (uint32_t)kGemmMaxKChunks || c->dim < 64) return false;
    if (c->n_rows < (int64_t)kGemmN * 64) return false;             // too few tiles to fill the machine / the lists
    int min_q = c->storage == LVS_STORAGE_BF16 ? c->opt_gemm_min_q : std::max(c->opt_gemm_min_q, 5);
    // two queries on a large bf16 shard: one tensor-core pass (the price of one single-query scan) beats the two-query scan, which is
    // FMA-issue-bound; on small shards the scan path's single launch wins
    if (c->storage == LVS_STORAGE_BF16 && min_q == 3 && c->n_rows >= (int64_t)2000000) min_q = 2;
    return c->opt_path == 2 || Q >= min_q;
}

// Enqueue K2 + finalize for queries [0, Q) in batches of up to 256.  No synchronisation.
static int enqueue_gemm(lvs_collection* c, int Q, int k, int kpl, uint64_t search_base, double* d_scores, int64_t* d_rows,
                        uint64_t* d_ties, uint32_t* d_counts, int32_t* d_flags, cudaStream_t st, int* launches,
                        const uint32_t* const* fcodes = nullptr, const uint32_t* fwant = nullptr, uint32_t nf = 0) {
    int rc;
    if ((rc = get_encode_tiled()) != LVS_OK) return rc;
    const int sm = g_lib.sm_count;
    const bool tf32 = c->storage == LVS_STORAGE_F32;                 // fp32 shards are multiplied as tf32, straight from their rows
    const uint32_t kce = tf32 ? kGemmKC / 2 : kGemmKC;               // K elements per 128-byte swizzle row
    const uint32_t esz = tf32 ? 4 : 2;
    const uint32_t nk = (c->q_stride + kce - 1) / kce;
    const uint32_t k_pad = nk * kce;
    const size_t keys_bytes = (size_t)256 * sm * 2 * kGemmList * 8;
    if ((rc = ensure_dev(c->s_gkeys, keys_bytes)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_gtops, (size_t)256 * sm * 2 * 8)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_gdrops, (size_t)256 * sm * 2 * 8)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_qb16, (size_t)256 * k_pad * esz)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_geps, (size_t)256 * 4)) != LVS_OK) return rc;
    kpl = std::min(8, 2 * kpl);      // the bf16-query scores are coarser than K1's: rescore a larger candidate set
    if ((rc = ensure_dev(c->s_cand, (size_t)256 * kMaxCand * 8)) != LVS_OK) return rc;
    if (c->s_tickets.bytes < 256 * 4) {
        if ((rc = ensure_dev(c->s_tickets, 256 * 4)) != LVS_OK) return rc;
        CU(cudaMemsetAsync(c->s_tickets.p, 0, c->s_tickets.bytes, st));
    }
    if (c->opt_gemm_dbg & 1) { if ((rc = ensure_dev(c->s_dbg, (size_t)kGemmM * kGemmN * 4)) != LVS_OK) return rc; }
    // tensor maps: B over the shard [n_rows][ld] bf16 (box = 256 rows x 64 elements; 128 rows for the CTA-pair form, where
    // each CTA of the pair loads half a tile) and A over the bf16 queries [256][k_pad] (box = 128 rows x 64 elements);
    // 128-byte swizzle, zero fill out of bounds
    CUtensorMap tmap_b, tmap_b_half, tmap_a;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)c->q_stride, (cuuint64_t)c->n_rows};
        cuuint64_t gstr[1] = {(cuuint64_t)c->row_bytes};
        cuuint32_t box[2] = {(cuuint32_t)kce, (cuuint32_t)kGemmN};
        cuuint32_t estr[2] = {1, 1};
        const CUtensorMapDataType tdt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        CUresult r = g_encode_tiled(&tmap_b, tdt, 2, c->d_vec, gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(LVS_ECUDA, "cuTensorMapEncodeTiled (corpus) failed with CUresult %d", (int)r);
        cuuint32_t hbox[2] = {(cuuint32_t)kce, (cuuint32_t)(kGemmN / 2)};
        r = g_encode_tiled(&tmap_b_half, tdt, 2, c->d_vec, gdim, gstr, hbox, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(LVS_ECUDA, "cuTensorMapEncodeTiled (corpus, half tile) failed with CUresult %d", (int)r);
        cuuint64_t qdim[2] = {(cuuint64_t)k_pad, (cuuint64_t)256};
        cuuint64_t qstr[1] = {(cuuint64_t)k_pad * esz};
        cuuint32_t qbox[2] = {(cuuint32_t)kce, (cuuint32_t)kGemmM};
        r = g_encode_tiled(&tmap_a, tdt, 2, c->s_qb16.p, qdim, qstr, qbox, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(LVS_ECUDA, "cuTensorMapEncodeTiled (queries) failed with CUresult %d", (int)r);
    }
    {   // collections are searched from several threads (one lock per collection): set the attributes exactly once
        static std::once_flag once;
        static cudaError_t once_err = cudaSuccess;
        std::call_once(once, [] {
            const int sm_ = (int)g_lib.smem_optin;
            cudaError_t e_ = cudaFuncSetAttribute(gemm_topk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_);
            if (e_ == cudaSuccess) e_ = cudaFuncSetAttribute(gemm_topk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_);
            if (e_ == cudaSuccess) e_ = cudaFuncSetAttribute(gemm_topk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_);
            if (e_ == cudaSuccess) e_ = cudaFuncSetAttribute(gemm_topk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_);
            once_err = e_;
        });
        if (once_err != cudaSuccess) return fail(LVS_ECUDA, "cudaFuncSetAttribute (tensor-core kernel) failed: %s", cudaGetErrorString(once_err));
    }
    const uint32_t n_tiles = (uint32_t)((c->n_rows + kGemmN - 1) / kGemmN);
    for (int q0 = 0; q0 < Q; q0 += 256) {
        const int qb = std::min(256, Q - q0);
        const uint32_t G = (uint32_t)((qb + kGemmM - 1) / kGemmM);
        // more than 128 queries: clusters of two CTAs share every corpus tile (tcgen05 cta_group::2)
        const bool pair_form = G == 2 && !c->opt_gemm_no_pair;
        const uint32_t P = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)sm / G, n_tiles));
        uint32_t S = pair_form ? 4 : 3;
        if (c->opt_gemm_stages > 0) S = std::min<uint32_t>(S, (uint32_t)c->opt_gemm_stages);
        // pair form, split rings: S query-chunk buffers + SB corpus half-tile buffers (16 KB each) instead of S stages of both
        uint32_t SB = 0;
        if (pair_form) {
            static const int env_b = [] { const char* e = getenv("LATTICE_B200_GEMM_STAGES_B"); return e ? atoi(e) : -1; }();   // A/B switch
            const int want_b = c->opt_gemm_stages_b >= 0 ? c->opt_gemm_stages_b : env_b >= 0 ? env_b : kGemmDefaultStagesB;
            SB = (uint32_t)std::min(want_b, kGemmMaxStagesB);
            if (SB != 0 && c->opt_gemm_stages <= 0) S = kGemmDefaultStagesA;
            while (SB > 2 && gemm_smem_bytes(S, true, SB) > g_lib.smem_optin) --SB;
            if (SB != 0 && (SB < 2 || gemm_smem_bytes(S, true, SB) > g_lib.smem_optin)) SB = 0;
        }
        while (S > 2 && gemm_smem_bytes(S, pair_form, SB) > g_lib.smem_optin) --S;
        const size_t smem = gemm_smem_bytes(S, pair_form, SB);
        if (tf32)
            prep_qtf32_kernel<<<G * kGemmM, 256, 0, st>>>((const double*)c->s_q64.p + (size_t)q0 * c->dim, qb, c->dim,
                                                        (float*)c->s_qb16.p, k_pad, G * kGemmM, (float*)c->s_geps.p);
        else
            prep_qb16_kernel<<<G * kGemmM, 256, 0, st>>>((const double*)c->s_q64.p + (size_t)q0 * c->dim, qb, c->dim,
                                                       (__nv_bfloat16*)c->s_qb16.p, k_pad, G * kGemmM, (float*)c->s_geps.p);
        CU(cudaGetLastError());
        ++*launches;
        GemmParams gp;
        memset(&gp, 0, sizeof(gp));
        gp.n_kchunks = nk; gp.n_queries = (uint32_t)qb; gp.n_rows = (uint32_t)c->n_rows;
        gp.n_tiles = n_tiles; gp.n_groups = G; gp.n_pairs = P; gp.n_stages = S; gp.n_stages_b = SB;
        gp.inv_norm = c->metric == LVS_METRIC_COSINE ? c->d_inv_norm : nullptr; gp.live = c->d_live;
        gp.n_filter = nf;
        for (uint32_t i = 0; i < nf; ++i) { gp.fcodes[i] = fcodes[i]; gp.fwant[i] = fwant[i]; }

        gp.out_keys = (uint64_t*)c->s_gkeys.p; gp.out_tops = (uint64_t*)c->s_gtops.p; gp.out_drops = (uint64_t*)c->s_gdrops.p;
        const bool unit_rows = c->metric == LVS_METRIC_COSINE && c->h_norm_stats[1] <= 0.001953125f && !c->opt_gemm_no_unit;
        gp.unit_rows = unit_rows ? 1u : 0u;
        gp.keep = (uint32_t)std::min(kGemmList, std::max(4, c->opt_gemm_keep));
        gp.dbg = (c->opt_gemm_dbg & 1) ? (float*)c->s_dbg.p : nullptr;
        gp.dbg_mode = (uint32_t)(c->opt_gemm_dbg >> 1);
        cudaEvent_t es = nullptr, ee = nullptr;
        if (c->opt_timing) {
            const int slot = c->ring_pos % kEventRing;
            es = c->ring_ev[2 * slot]; ee = c->ring_ev[2 * slot + 1];
            c->ring_bytes[slot] = (double)c->n_rows * c->row_bytes;
            c->ring_pos++;
            CU(cudaEventRecord(es, st));
        }
        cudaError_t e;
        if (pair_form) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(P * 2); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute la[1];
            la[0].id = cudaLaunchAttributeClusterDimension;
            la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
            cfg.attrs = la; cfg.numAttrs = 1;
            e = tf32 ? cudaLaunchKernelEx(&cfg, gemm_topk_kernel<true, true>, tmap_b_half, tmap_a, gp)
                     : cudaLaunchKernelEx(&cfg, gemm_topk_kernel<true, false>, tmap_b_half, tmap_a, gp);
        } else {
            if (tf32) gemm_topk_kernel<false, true><<<P * G, kGemmThreads, smem, st>>>(tmap_b, tmap_a, gp);
            else gemm_topk_kernel<false, false><<<P * G, kGemmThreads, smem, st>>>(tmap_b, tmap_a, gp);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return fail(LVS_ECUDA, "gemm kernel launch failed: %s (smem=%zu grid=%u)", cudaGetErrorString(e), smem, P * G);
        ++*launches;
        if (c->opt_timing) CU(cudaEventRecord(ee, st));
        if (q0 == 0) { c->first_scan_start = es; c->first_scan_end = ee; }

        FinalizeParams fp;
        memset(&fp, 0, sizeof(fp));
        const uint32_t kpw = 32u * kpl;
        fp.keys = gp.out_keys; fp.tops = gp.out_tops; fp.drops = gp.out_drops;
        fp.M = 2 * P * kGemmList; fp.L = 2 * P; fp.kp = kpw; fp.k = (uint32_t)k;   // two lists per CTA and query
        fp.base = c->d_vec; fp.row_bytes = c->row_bytes; fp.dim = c->dim; fp.dim_pad = (int)c->q_stride;
        fp.storage = c->storage; fp.metric = c->metric;
        fp.q64 = (const double*)c->s_q64.p + (size_t)q0 * c->dim;
        fp.tiekey = c->d_tie; fp.epoch = c->d_epoch; fp.search_no = search_base + (uint64_t)q0;
        fp.pw = c->d_pw; fp.row_base = c->row_base;
        fp.eps = 2.2e-3f;    // fallback; the per-query bound ||q - bf16(q)||_2 + accumulation slack is used
        fp.eps_q = c->metric == LVS_METRIC_COSINE ? (const float*)c->s_geps.p : nullptr;
        fp.eps_add = unit_rows ? c->h_norm_stats[1] * 1.01f : 0.f;
        // tf32: the tensor core drops the low 13 mantissa bits of every stored value: |q.(x - tf32(x))| <= 2^-10 ||q|| ||x||
        if (tf32) fp.eps_add += 0.0009775f + 1.0e-6f;
        int nrw = kFinWarps;
        while (nrw > 1 && finalize_smem_bytes(fp.dim_pad, nrw) > g_lib.smem_optin) --nrw;
        fp.n_rescore_warps = nrw;
        fp.cand_scores = (double*)c->s_cand.p; fp.tickets = (uint32_t*)c->s_tickets.p;
        fp.out_scores = d_scores + (size_t)q0 * k; fp.out_rows = d_rows + (size_t)q0 * k; fp.out_ties = d_ties + (size_t)q0 * k;
        fp.out_flags = d_flags + q0; fp.out_counts = d_counts + q0;
        fp.qnorm = (const float*)c->s_qnorm.p + q0; fp.max_norm = c->d_max_norm;
        const size_t fsm = finalize_smem_bytes(fp.dim_pad, nrw);
        // CTAs per query: every CTA of a query repeats the selection, so only as many as it takes to fill the machine once
        // (two finalize CTAs fit on an SM): 8 for a handful of queries, 1 from ~300 queries on
        const uint32_t fill = (uint32_t)std::max(1, (2 * sm) / std::max(1, qb));
        const unsigned ncta = (unsigned)std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(8u, fill), (kpw + nrw - 1) / nrw));
        cudaError_t fe = kpl == 1 ? launch_finalize<1>(fp, qb, ncta, fsm, st) : kpl == 2 ? launch_finalize<2>(fp, qb, ncta, fsm, st)
                       : kpl == 4 ? launch_finalize<4>(fp, qb, ncta, fsm, st) : launch_finalize<8>(fp, qb, ncta, fsm, st);
        if (fe != cudaSuccess) return fail(LVS_ECUDA, "finalize kernel launch failed: %s", cudaGetErrorString(fe));
        ++*launches;
        if (q0 == 0 && c->opt_timing) CU(cudaEventRecord(c->ev[4], st));
    }
    return LVS_OK;
}

// Core: queries already on the device (raw, `dtype`); outputs are device buffers.  With `async` the work is only
// enqueued (no escalation, flags stay on the device in d_flags_out); otherwise flagged queries are repeated with a
// larger candidate set and the host flags are returned.  Sharded form (ex != nullptr, async only): every rank calls it with
// the same arguments; the merged result of all ranks lands in xout ([3][Q][k] int64), d_counts and the flags.
static int search_core(lvs_collection* c, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                       double* d_scores, int64_t* d_rows, uint64_t* d_ties, uint32_t* d_counts, int32_t* h_flags,
                       int32_t* d_flags_out, bool async, cudaStream_t st, int64_t base_override = -1, int kpl_min = 0,
                       const std::vector<int>* only = nullptr, lvs_exchange* ex = nullptr, int64_t* xout = nullptr,
                       const void* h_queries = nullptr, uint32_t* host_ready = nullptr, uint32_t host_ready_val = 0) {
    c->last_ready_armed = false;
    if (Q <= 0) return LVS_OK;
    if (k < 1 || k > LVS_MAX_K) return fail(LVS_ELIMIT, "limit %d outside 1..%d", k, LVS_MAX_K);
    if (dtype != LVS_DT_F32 && dtype != LVS_DT_F64) return fail(LVS_EINVAL, "query dtype must be f32 or f64");
    if (ex && !async) return fail(LVS_EINVAL, "the sharded search is enqueue-only");
    if (h_queries == nullptr) {
        // callers may hand in mapped pinned HOST memory as a "device" pointer (unified addressing): such queries travel in the
        // kernel's parameter block or are staged by one CTA instead of being pulled over PCIe by every CTA
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, d_queries) == cudaSuccess) { if (pa.type == cudaMemoryTypeHost) h_queries = pa.hostPointer; }
        else cudaGetLastError();
    }
    const int sm = g_lib.sm_count;
    int rc;
    if ((rc = ensure_dev(c->s_keys, (size_t)4 * sm * 256 * 8)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_mins, (size_t)4 * sm * 8)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_flags, (size_t)Q * 4)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_cand, (size_t)4 * kMaxCand * 8)) != LVS_OK) return rc;
    if ((rc = ensure_pinned(c->h_flags, (size_t)Q * 4)) != LVS_OK) return rc;
    int32_t* d_flags = d_flags_out ? d_flags_out : (int32_t*)c->s_flags.p;
    // the per-collection scratch (lists, tickets, candidate scores) is shared by every search on this handle: a search on
    // another stream than the previous one is ordered behind it
    if (c->last_stream_valid && c->last_stream != st) {
        CU(cudaEventRecord(c->order_ev, c->last_stream));
        CU(cudaStreamWaitEvent(st, c->order_ev, 0));
    }
    c->last_stream = st; c->last_stream_valid = true;

    int launches = 0;
    if (c->opt_timing) CU(cudaEventRecord(c->ev[0], st));
    const uint32_t* fcodes[kMaxFilterCols];
    uint32_t fwant[kMaxFilterCols];
    uint32_t nf = 0;
    build_filter(c, want, fcodes, fwant, &nf);
    const bool filter = nf > 0;

    // candidate-set size: smallest list of 32*KPL keys that leaves a margin over k
    int kpl = 1;
    while (kpl < 8 && 32 * kpl < k + std::max(8, k / 4)) kpl <<= 1;
    if (c->opt_force_kpl > 0) kpl = std::min(8, std::max(kpl, c->opt_force_kpl));
    if (kpl_min > 0) kpl = std::min(8, std::max(kpl, kpl_min));

    std::vector<int> pending;
    if (only) pending = *only;
    else { pending.resize(Q); for (int i = 0; i < Q; ++i) pending[i] = i; }
    const uint64_t search_base = base_override >= 0 ? (uint64_t)base_override : c->search_counter + 1;
    int32_t* hf = (int32_t*)c->h_flags.p;
    bool first = true;
    int kind = 1;
    if (!only && gemm_eligible(c, Q, filter)) {
        // K2 for the whole batch; queries whose exactness bound is not met fall through to the K1 levels below
        kind = 2;
        if ((rc = ensure_dev(c->s_q64, (size_t)Q * c->dim * 8)) != LVS_OK) return rc;
        if ((rc = ensure_dev(c->s_q32, (size_t)(Q + 4) * c->q_stride * 4)) != LVS_OK) return rc;
        if ((rc = ensure_dev(c->s_qnorm, (size_t)Q * 4)) != LVS_OK) return rc;
        {
            PrepParams pp;
            pp.src = d_queries; pp.src_dtype = dtype; pp.dim = c->dim; pp.metric = c->metric;
            pp.q64 = (double*)c->s_q64.p; pp.q32 = (float*)c->s_q32.p; pp.q_stride = c->q_stride; pp.qnorm = (float*)c->s_qnorm.p;
            pp.n_zero_rows = 4;
            prep_queries_kernel<<<Q + 1, 256, 0, st>>>(pp);   // CTA Q zeroes the 4 padding query rows
            CU(cudaGetLastError());
            ++launches;
        }
        double* ls = d_scores; int64_t* lr = d_rows; uint64_t* lt = d_ties; uint32_t* lc = d_counts; int32_t* lf = d_flags;
        if (ex) {
            // sharded: the local lists go to scratch, the exchange kernel publishes and merges them
            const size_t n = (size_t)Q * k;
            if ((rc = ensure_dev(c->s_xlocal, n * 24 + (size_t)Q * 8)) != LVS_OK) return rc;
            int64_t* xl = (int64_t*)c->s_xlocal.p;
            ls = (double*)xl; lr = xl + n; lt = (uint64_t*)(xl + 2 * n); lc = (uint32_t*)(xl + 3 * n); lf = (int32_t*)(lc + Q);
        }
        rc = enqueue_gemm(c, Q, k, kpl, search_base, ls, lr, lt, lc, lf, st, &launches, fcodes, fwant, nf);
        if (rc != LVS_OK) return rc;
        if (ex) {
            ExchangeParams xp;
            memset(&xp, 0, sizeof(xp));
            if ((rc = exchange_params(ex, Q, k, &xp)) != LVS_OK) return rc;
            xp.local = (const int64_t*)c->s_xlocal.p; xp.local_flags = lf;
            xp.out = xout; xp.out_counts = d_counts; xp.out_flags = d_flags; xp.Q_out = Q; xp.q_out0 = 0;
            if ((rc = launch_exchange(xp, st)) != LVS_OK) return rc;
            ++launches;
        }
        first = false;
        pending.clear();
        if (!async) {
            CU(cudaMemcpyAsync(hf, d_flags, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (int qi = 0; qi < Q; ++qi) if (hf[qi] & 1) pending.push_back(qi);
        }
    }
    LevelOut lo;
    lo.scores = d_scores; lo.rows = d_rows; lo.ties = d_ties; lo.counts = d_counts; lo.flags = d_flags;
    lo.ex = ex; lo.xout = xout; lo.Q_total = Q;
    lo.host_ready = nullptr; lo.host_ready_val = 0;
    // a search that is exactly ONE fused launch can publish its completion in host memory itself (no event between consecutive
    // kernels of the stream, so they keep overlapping; the host polls a word instead of synchronising)
    if (host_ready && async && kind == 1 && !c->opt_timing && (int)pending.size() == Q && Q <= max_qt_for_kpl(kpl)) {
        lo.host_ready = host_ready; lo.host_ready_val = host_ready_val;
        c->last_ready_armed = true;
    }
    while (!pending.empty()) {
        rc = enqueue_level(c, d_queries, h_queries, dtype, pending, k, kpl, filter, fcodes, fwant, nf, search_base, lo, st, &launches, first);
        if (rc != LVS_OK) return rc;
        first = false;
        if (async) break;
        CU(cudaMemcpyAsync(hf, d_flags, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        std::vector<int> next;
        for (int qi : pending) if ((hf[qi] & 1) && kpl < 8) next.push_back(qi);
        pending.swap(next);
        if (!pending.empty()) kpl <<= 1;
    }
    c->last_launches = launches;
    if (!only) c->last_kind = kind;
    c->last_kpl = kpl;
    if (base_override < 0) c->search_counter += (uint64_t)Q;
    if (!async) {
        if (c->opt_timing) {
            CU(cudaEventRecord(c->ev[5], st));
            CU(cudaEventSynchronize(c->ev[5]));
            c->last_ms[0] = 0.f;
            if (c->first_scan_start) cudaEventElapsedTime(&c->last_ms[0], c->ev[0], c->first_scan_start);
            if (c->first_scan_start) cudaEventElapsedTime(&c->last_ms[1], c->first_scan_start, c->first_scan_end);
            if (c->first_scan_end) cudaEventElapsedTime(&c->last_ms[2], c->first_scan_end, c->ev[4]);
            cudaEventElapsedTime(&c->last_ms[3], c->ev[0], c->ev[5]);
        }
        if (h_flags) memcpy(h_flags, hf, (size_t)Q * 4);
    }
    return LVS_OK;
}

extern "C" int lvs_search_device(lvs_collection* c, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                                 double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                                 int32_t* out_flags, void* stream) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (Q < 0 || (Q > 0 && (!d_queries || !d_out_scores || !d_out_rows || !d_out_ties || !d_out_counts)))
        return fail(LVS_EINVAL, "NULL device buffer");
    if (Q > 65535) return fail(LVS_ELIMIT, "batch of %d queries exceeds 65535", Q);
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return search_core(c, d_queries, dtype, Q, k, want, d_out_scores, d_out_rows, d_out_ties, d_out_counts, out_flags, nullptr, false, st);
}

extern "C" int lvs_search_device_at(lvs_collection* c, uint64_t search_no, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                                    double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                                    int32_t* out_flags, void* stream) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (Q < 0 || (Q > 0 && (!d_queries || !d_out_scores || !d_out_rows || !d_out_ties || !d_out_counts)))
        return fail(LVS_EINVAL, "NULL device buffer");
    if (Q > 65535) return fail(LVS_ELIMIT, "batch of %d queries exceeds 65535", Q);
    std::lock_guard<std::mutex> lk(c->mu);
    if (search_no < 1 || search_no + (uint64_t)Q - 1 > c->search_counter) return fail(LVS_EINVAL, "search numbers outside 1..%llu", (unsigned long long)c->search_counter);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    // a repeat: the exact scan (K1) with a candidate set one size up, numbered as the searches it repeats
    std::vector<int> all(Q);
    for (int i = 0; i < Q; ++i) all[i] = i;
    return search_core(c, d_queries, dtype, Q, k, want, d_out_scores, d_out_rows, d_out_ties, d_out_counts, out_flags, nullptr, false, st,
                       (int64_t)search_no, 0, &all);
}

extern "C" int lvs_search_device_async(lvs_collection* c, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                                       double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                                       int32_t* d_out_flags, void* stream) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (Q < 0 || (Q > 0 && (!d_queries || !d_out_scores || !d_out_rows || !d_out_ties || !d_out_counts || !d_out_flags)))
        return fail(LVS_EINVAL, "NULL device buffer");
    if (Q > 65535) return fail(LVS_ELIMIT, "batch of %d queries exceeds 65535", Q);
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return search_core(c, d_queries, dtype, Q, k, want, d_out_scores, d_out_rows, d_out_ties, d_out_counts, nullptr, d_out_flags, true, st);
}

// Host-buffer searches are zero-copy: the query block and the result block of a slot live in mapped pinned host memory;
// the prep kernel reads the queries over PCIe and the finalize kernel stores scores/rows/ties/counts/flags straight into
// host memory, so a step is kernels only (no copy-engine hops between them).
// Slot layout: [queries, padded to 256 B][scores | rows | ties (Q*k*8 each) | counts (Q*4) | flags (Q*4)]
static size_t res_bytes(int Q, int k) { return (size_t)Q * k * 24 + (size_t)Q * 8; }
static size_t ready_off(int Q, int k) { return (res_bytes(Q, k) + 15) & ~(size_t)15; }   // the slot's completion word sits behind the results

static int submit_locked(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want, int* ticket,
                         lvs_exchange* ex = nullptr) {
    int si = -1;
    for (int i = 0; i < kSubmitSlots; ++i) if (!c->slots[i].in_use) { si = i; break; }
    if (si < 0) return fail(LVS_ELIMIT, "%d searches already in flight: call lvs_search_wait first", kSubmitSlots);
    auto& sl = c->slots[si];
    const size_t qbytes = ((size_t)Q * c->dim * dt_size(dtype) + 255) & ~(size_t)255;
    const size_t rbytes = res_bytes(Q, k);
    int rc;
    if ((rc = ensure_pinned(sl.h, qbytes + ready_off(Q, k) + 16)) != LVS_OK) return rc;
    if (!sl.done) CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    void* dview = nullptr;
    CU(cudaHostGetDevicePointer(&dview, sl.h.p, 0));
    cudaStream_t st = c->stream;
    const size_t qraw = (size_t)Q * c->dim * dt_size(dtype);
    // copy into the slot and look for NaN on the way (the reference's local mode refuses such a query)
    {
        const size_t n = (size_t)Q * c->dim;
        int bad = 0;
        if (dtype == LVS_DT_F64) { const double* s = (const double*)queries; double* d = (double*)sl.h.p; for (size_t i = 0; i < n; ++i) { const double x = s[i]; d[i] = x; bad |= x != x; } }
        else { const float* s = (const float*)queries; float* d = (float*)sl.h.p; for (size_t i = 0; i < n; ++i) { const float x = s[i]; d[i] = x; bad |= x != x; } }
        if (bad) return fail(LVS_ENAN, "Query vector must not contain NaN");
    }
    const size_t nres = (size_t)Q * k;
    sl.Q = Q; sl.k = k; sl.dtype = dtype; sl.base = c->search_counter + 1; sl.qbytes = qbytes;
    sl.has_want = want != nullptr;
    if (want) memcpy(sl.want, want, sizeof(uint32_t) * kMaxFilterCols);
    // A handful of queries: zero-copy (the kernels read the query block from and store the result block to mapped pinned memory:
    // no copy-engine hop on a latency-bound step).  A batch: one H2D copy of the queries and one D2H copy of the results around
    // device-resident work - hundreds of CTAs storing 24 Q k bytes over PCIe from inside the finalize kernel serialise.
    sl.staged = qraw + rbytes >= (size_t)64 * 1024;
    sl.sharded = ex != nullptr;
    sl.ready_val = ++c->ready_seq ? c->ready_seq : ++c->ready_seq;      // never 0
    volatile uint32_t* h_ready = (volatile uint32_t*)((uint8_t*)sl.h.p + qbytes + ready_off(Q, k));
    *h_ready = 0;
    uint8_t* qp = (uint8_t*)dview;
    if (sl.staged) {
        if ((rc = ensure_dev(sl.d, qbytes + rbytes)) != LVS_OK) return rc;
        CU(cudaMemcpyAsync(sl.d.p, sl.h.p, qraw, cudaMemcpyHostToDevice, st));
        qp = (uint8_t*)sl.d.p;
    }
    uint8_t* rp = qp + qbytes;
    // sharded: the merged lists of all ranks land in the same place, in the same layout ([3][Q][k] = scores | rows | ties)
    rc = search_core(c, qp, dtype, Q, k, want, (double*)rp, (int64_t*)(rp + nres * 8), (uint64_t*)(rp + nres * 16),
                     (uint32_t*)(rp + nres * 24), nullptr, (int32_t*)(rp + nres * 24 + (size_t)Q * 4), true, st, -1, 0, nullptr, ex,
                     ex ? (int64_t*)rp : nullptr, sl.staged ? nullptr : sl.h.p,
                     sl.staged ? nullptr : (uint32_t*)((uint8_t*)dview + qbytes + ready_off(Q, k)), sl.ready_val);
    if (rc == LVS_OK && sl.staged) CU(cudaMemcpyAsync((uint8_t*)sl.h.p + qbytes, rp, rbytes, cudaMemcpyDeviceToHost, st));
    sl.poll = rc == LVS_OK && c->last_ready_armed;
    if (rc != LVS_OK) return rc;
    sl.kpl = c->last_kpl;
    sl.kind = c->last_kind;
    if (!sl.poll) CU(cudaEventRecord(sl.done, st));
    sl.in_use = true;
    *ticket = si;
    return LVS_OK;
}

// Blocks until the search in `sl` has finished: polls the completion word its kernel publishes in the slot's pinned memory, or
// synchronises the slot's event (batches, multi-launch searches, timing mode).
static int wait_slot(lvs_collection* c, lvs_collection::Slot& sl) {
    if (!sl.poll) { CU(cudaEventSynchronize(sl.done)); return LVS_OK; }
    volatile uint32_t* w = (volatile uint32_t*)((uint8_t*)sl.h.p + sl.qbytes + ready_off(sl.Q, sl.k));
    for (uint64_t spins = 1;; ++spins) {
        if (*w == sl.ready_val) { std::atomic_thread_fence(std::memory_order_acquire); return LVS_OK; }
        if ((spins & 0x3FFF) == 0) {
            // not there yet: make sure the stream is still alive (a faulted kernel never publishes)
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e == cudaSuccess) {
                if (*w == sl.ready_val) return LVS_OK;
                return fail(LVS_ECUDA, "the search finished without publishing its completion word");
            }
            if (e != cudaErrorNotReady) return fail(LVS_ECUDA, "search failed: %s", cudaGetErrorString(e));
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

static int finish_locked(lvs_collection* c, int ticket, double* out_scores, int64_t* out_rows, uint64_t* out_ties,
                         uint32_t* out_counts, int32_t* out_flags) {
    auto& sl = c->slots[ticket];
    const int Q = sl.Q, k = sl.k;
    const size_t nres = (size_t)Q * k;
    uint8_t* hp = (uint8_t*)sl.h.p + sl.qbytes;
    int32_t* hflags = (int32_t*)(hp + nres * 24 + (size_t)Q * 4);
    std::vector<int> redo;
    // a K2 slot (bf16-rounded queries, coarse bound) always gets the exact-scan fallback for its flagged queries, whatever k is;
    // a K1 slot is repeated with a larger candidate set while one exists
    if (!sl.sharded)     // a sharded slot's flags are the merged ones: the caller repeats those queries on every rank (lvs_search_device_at)
        for (int i = 0; i < Q; ++i) if ((hflags[i] & 1) && (sl.kind == 2 || sl.kpl < 8)) redo.push_back(i);
    if (!redo.empty()) {
        // rare: repeat the flagged queries (K1, larger candidate sets), as the same reference searches (same numbers)
        void* dview = nullptr;
        CU(cudaHostGetDevicePointer(&dview, sl.h.p, 0));
        uint8_t* qp = sl.staged ? (uint8_t*)sl.d.p : (uint8_t*)dview;
        uint8_t* rp = qp + sl.qbytes;
        std::vector<int32_t> f2(Q, 0);
        int rc = search_core(c, qp, sl.dtype, Q, k, sl.has_want ? sl.want : nullptr, (double*)rp, (int64_t*)(rp + nres * 8),
                             (uint64_t*)(rp + nres * 16), (uint32_t*)(rp + nres * 24), f2.data(), nullptr, false, c->stream,
                             (int64_t)sl.base, sl.kind == 2 ? sl.kpl : sl.kpl * 2, &redo, nullptr, nullptr, sl.staged ? nullptr : sl.h.p);
        if (rc == LVS_OK && sl.staged) {
            cudaError_t e = cudaMemcpyAsync(hp, rp, res_bytes(Q, k), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = fail(LVS_ECUDA, "copy of the repeated results failed: %s", cudaGetErrorString(e));
        }
        if (rc != LVS_OK) { sl.in_use = false; return rc; }
        for (int i : redo) hflags[i] = f2[i];
    }
    if (out_scores) memcpy(out_scores, hp, nres * 8);
    if (out_rows) memcpy(out_rows, hp + nres * 8, nres * 8);
    if (out_ties) memcpy(out_ties, hp + nres * 16, nres * 8);
    if (out_counts) memcpy(out_counts, hp + nres * 24, (size_t)Q * 4);
    if (out_flags) memcpy(out_flags, hflags, (size_t)Q * 4);
    sl.in_use = false;
    return LVS_OK;
}

static int check_search_args(const lvs_collection* c, const void* queries, int dtype, int Q, int k) {
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (Q < 0 || (Q > 0 && !queries)) return fail(LVS_EINVAL, "bad queries / Q");
    if (Q > 65535) return fail(LVS_ELIMIT, "batch of %d queries exceeds 65535", Q);
    if (k < 1 || k > LVS_MAX_K) return fail(LVS_ELIMIT, "limit %d outside 1..%d", k, LVS_MAX_K);
    if (dtype != LVS_DT_F32 && dtype != LVS_DT_F64) return fail(LVS_EINVAL, "query dtype must be f32 or f64");
    return LVS_OK;
}

extern "C" int lvs_search_submit(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want, int* ticket) {
    bind_thread();
    if (!ticket) return fail(LVS_EINVAL, "ticket is NULL");
    int rc = check_search_args(c, queries, dtype, Q, k);
    if (rc != LVS_OK) return rc;
    if (Q < 1) return fail(LVS_EINVAL, "Q must be >= 1");
    std::lock_guard<std::mutex> lk(c->mu);
    return submit_locked(c, queries, dtype, Q, k, want, ticket);
}

extern "C" int lvs_search_submit_sharded(lvs_collection* c, lvs_exchange* ex, const void* queries, int dtype, int Q, int k, const uint32_t* want,
                                         int* ticket) {
    bind_thread();
    if (!ticket || !ex) return fail(LVS_EINVAL, "ticket / exchange is NULL");
    int rc = check_search_args(c, queries, dtype, Q, k);
    if (rc != LVS_OK) return rc;
    if (Q < 1) return fail(LVS_EINVAL, "Q must be >= 1");
    std::lock_guard<std::mutex> lk(c->mu);
    return submit_locked(c, queries, dtype, Q, k, want, ticket, ex);
}

extern "C" int lvs_search_wait(lvs_collection* c, int ticket, double* out_scores, int64_t* out_rows, uint64_t* out_ties,
                               uint32_t* out_counts, int32_t* out_flags) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (ticket < 0 || ticket >= kSubmitSlots) return fail(LVS_EINVAL, "bad ticket %d", ticket);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->slots[ticket].in_use) return fail(LVS_EINVAL, "ticket %d is not in flight", ticket);
    }
    int rcw = wait_slot(c, c->slots[ticket]);   // outside the lock: other threads may submit meanwhile (slots are stable storage)
    if (rcw != LVS_OK) { std::lock_guard<std::mutex> lk(c->mu); c->slots[ticket].in_use = false; return rcw; }
    std::lock_guard<std::mutex> lk(c->mu);
    return finish_locked(c, ticket, out_scores, out_rows, out_ties, out_counts, out_flags);
}

extern "C" int lvs_search_poll(lvs_collection* c, int ticket, int* done) {
    bind_thread();
    if (!c || !done) return fail(LVS_EINVAL, "NULL argument");
    if (ticket < 0 || ticket >= kSubmitSlots) return fail(LVS_EINVAL, "bad ticket %d", ticket);
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->slots[ticket].in_use) return fail(LVS_EINVAL, "ticket %d is not in flight", ticket);
    }
    auto& sl = c->slots[ticket];              // stable storage while the ticket is in flight
    if (sl.poll) {
        volatile uint32_t* w = (volatile uint32_t*)((uint8_t*)sl.h.p + sl.qbytes + ready_off(sl.Q, sl.k));
        *done = *w == sl.ready_val ? 1 : 0;
        return LVS_OK;
    }
    cudaError_t e = cudaEventQuery(sl.done);
    if (e == cudaSuccess) { *done = 1; return LVS_OK; }
    if (e == cudaErrorNotReady) { cudaGetLastError(); *done = 0; return LVS_OK; }
    return fail(LVS_ECUDA, "search failed: %s", cudaGetErrorString(e));
}

extern "C" int lvs_search(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want,
                          double* out_scores, int64_t* out_rows, uint64_t* out_ties, uint32_t* out_counts, int32_t* out_flags) {
    bind_thread();
    int rc = check_search_args(c, queries, dtype, Q, k);
    if (rc != LVS_OK) return rc;
    if (Q == 0) return LVS_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    int ticket = -1;
    if ((rc = submit_locked(c, queries, dtype, Q, k, want, &ticket)) != LVS_OK) return rc;
    if ((rc = wait_slot(c, c->slots[ticket])) != LVS_OK) { c->slots[ticket].in_use = false; return rc; }
    rc = finish_locked(c, ticket, out_scores, out_rows, out_ties, out_counts, out_flags);
    if (rc == LVS_OK && c->opt_timing && c->first_scan_start) {
        c->last_ms[0] = 0.f;
        cudaEventElapsedTime(&c->last_ms[0], c->ev[0], c->first_scan_start);
        cudaEventElapsedTime(&c->last_ms[1], c->first_scan_start, c->first_scan_end);
        cudaEventElapsedTime(&c->last_ms[2], c->first_scan_end, c->ev[4]);
        c->last_ms[3] = c->last_ms[0] + c->last_ms[1] + c->last_ms[2];
    }
    return rc;
}

extern "C" int lvs_scan_times(lvs_collection* c, int max_n, float* out_ms, double* out_bytes, int* n) {
    bind_thread();
    if (!c || !out_ms || !n) return fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    const uint64_t have = std::min<uint64_t>(c->ring_pos, (uint64_t)kEventRing);
    const int cnt = (int)std::min<uint64_t>(have, (uint64_t)std::max(0, max_n));
    for (int i = 0; i < cnt; ++i) {
        const int slot = (int)((c->ring_pos - cnt + i) % kEventRing);
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, c->ring_ev[2 * slot], c->ring_ev[2 * slot + 1]);
        if (e != cudaSuccess) return fail(LVS_ECUDA, "scan events not complete (synchronise the stream first): %s", cudaGetErrorString(e));
        out_ms[i] = ms;
        if (out_bytes) out_bytes[i] = c->ring_bytes[slot];
    }
    *n = cnt;
    return LVS_OK;
}

extern "C" int lvs_merge_topk_device(const double* d_scores, const int64_t* d_rows, const uint64_t* d_ties, int64_t shard_stride,
                                     int G, int Q, int k, double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                                     void* stream) {
    bind_thread();
    if (!g_lib.ready) return fail(LVS_ESTATE, "lvs_init() has not been called");
    if (G < 1 || Q < 0 || k < 1) return fail(LVS_EINVAL, "bad G / Q / k");
    if (Q == 0) return LVS_OK;
    const size_t smem = (size_t)G * k * 24;
    if (smem > g_lib.smem_optin - 2048) return fail(LVS_ELIMIT, "G*k = %d too large for the merge kernel", G * k);
    MergeParams p;
    p.in_scores = d_scores; p.in_rows = d_rows; p.in_ties = d_ties; p.G = G; p.Q = Q; p.k = k;
    p.shard_stride = shard_stride > 0 ? shard_stride : (int64_t)Q * k;
    p.out_scores = d_out_scores; p.out_rows = d_out_rows; p.out_ties = d_out_ties; p.out_counts = d_out_counts;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_topk_kernel<<<Q, 256, smem, st>>>(p);
    CU(cudaGetLastError());
    return LVS_OK;
}

// ------------------------------------------------------------------------------------------------------
// K5': fused exchange + merge over peer memory
// ------------------------------------------------------------------------------------------------------
extern "C" int lvs_exchange_create(int world, int rank, int max_q, int max_k, lvs_exchange** out, void* ipc_handle_out) {
    bind_thread();
    if (!g_lib.ready) return fail(LVS_ESTATE, "lvs_init() has not been called (no CUDA device bound)");
    if (!out || !ipc_handle_out) return fail(LVS_EINVAL, "NULL argument");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(LVS_EINVAL, "bad world / rank");
    if (max_q < 1 || max_k < 1) return fail(LVS_EINVAL, "bad max_q / max_k");
    lvs_exchange* ex = new (std::nothrow) lvs_exchange();
    if (!ex) return fail(LVS_ENOMEM, "host allocation failed");
    ex->world = world; ex->rank = rank; ex->max_q = max_q; ex->max_k = max_k;
    ex->blk_stride = (((size_t)3 * max_q * max_k + (size_t)max_q) + 31) & ~(size_t)31;   // [3][Q][k] lists + [Q] flags
    ex->slot_elems = (size_t)world * ex->blk_stride;
    ex->flags_off = 2 * ex->slot_elems * 8;
    const size_t bytes = ex->flags_off + 2 * kMaxRanks * 8;
    cudaError_t e = cudaMalloc(&ex->base, bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->base, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&ex->d_counter, 8);
    if (e == cudaSuccess) e = cudaMemset(ex->d_counter, 0, 8);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ex->base);
    if (e != cudaSuccess) { lvs_exchange_destroy(ex); return fail(LVS_ECUDA, "exchange setup failed: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &h, 64);
    ex->peers[rank] = ex->base;
    *out = ex;
    return LVS_OK;
}

extern "C" int lvs_exchange_connect(lvs_exchange* ex, const void* all_handles) {
    bind_thread();
    if (!ex || !all_handles) return fail(LVS_EINVAL, "NULL argument");
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)all_handles + (size_t)r * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(LVS_ECUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
        ex->peers[r] = (uint8_t*)ptr;
        ex->opened[r] = true;
    }
    ex->connected = true;
    return LVS_OK;
}

extern "C" int lvs_exchange_merge_device(lvs_exchange* ex, const int64_t* d_local, const int32_t* d_local_flags, int Q, int k,
                                         int64_t* d_out, uint32_t* d_out_counts, int32_t* d_out_flags, void* stream) {
    bind_thread();
    if (!ex || !d_local || !d_out || !d_out_counts) return fail(LVS_EINVAL, "NULL argument");
    ExchangeParams p;
    memset(&p, 0, sizeof(p));
    int rc = exchange_params(ex, Q, k, &p);
    if (rc != LVS_OK) return rc;
    p.local = d_local; p.local_flags = d_local_flags;
    p.out = d_out; p.out_counts = d_out_counts; p.out_flags = d_out_flags; p.Q_out = Q; p.q_out0 = 0;
    return launch_exchange(p, (cudaStream_t)stream);
}

extern "C" int lvs_search_sharded_device_async(lvs_collection* c, lvs_exchange* ex, const void* d_queries, int dtype, int Q, int k,
                                               const uint32_t* want, int64_t* d_out, uint32_t* d_out_counts, int32_t* d_out_flags,
                                               void* stream) {
    bind_thread();
    if (!c || !ex) return fail(LVS_EINVAL, "collection / exchange is NULL");
    if (Q < 0 || (Q > 0 && (!d_queries || !d_out || !d_out_counts || !d_out_flags))) return fail(LVS_EINVAL, "NULL device buffer");
    if (Q > 65535) return fail(LVS_ELIMIT, "batch of %d queries exceeds 65535", Q);
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return search_core(c, d_queries, dtype, Q, k, want, nullptr, nullptr, nullptr, d_out_counts, nullptr, d_out_flags, true, st, -1, 0,
                       nullptr, ex, d_out);
}

extern "C" int lvs_exchange_error(lvs_exchange* ex) {
    bind_thread();
    if (!ex) return 0;
    uint32_t h = 0;
    if (cudaMemcpy(&h, ex->d_counter + 1, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return (int)h;
}

extern "C" int lvs_exchange_destroy(lvs_exchange* ex) {
    bind_thread();
    if (!ex) return LVS_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < kMaxRanks; ++r) if (ex->opened[r] && ex->peers[r]) cudaIpcCloseMemHandle(ex->peers[r]);
    cudaFree(ex->base); cudaFree(ex->d_counter);
    delete ex;
    return LVS_OK;
}

// ------------------------------------------------------------------------------------------------------
// K3 hybrid ranking
// ------------------------------------------------------------------------------------------------------
static std::mutex g_rank_mu;
static Scratch g_rank_dev, g_rank_pin;
static cudaStream_t g_rank_stream = nullptr;
static cudaEvent_t g_rank_ev[2] = {nullptr, nullptr};

extern "C" int lvs_rank_fuse(const lvs_rank_batch* in, int mode, int max_per_file, int max_total, double entity_bonus, double rel_bonus,
                             int32_t* out_count, int32_t* out_index, double* out_score, double* out_norm, double* out_signals,
                             uint8_t* out_sigmask, uint8_t* out_source, int32_t* out_leader, float* device_ms) {
    bind_thread();
    if (!g_lib.ready) return fail(LVS_ESTATE, "lvs_init() has not been called (no CUDA device bound)");
    if (!in || !in->offsets) return fail(LVS_EINVAL, "NULL batch");
    if (mode != 0 && mode != 1) return fail(LVS_EINVAL, "mode must be 0 (HybridRanker) or 1 (ResultReranker)");
    const int nq = in->n_queries;
    if (nq < 0 || max_total < 1) return fail(LVS_EINVAL, "bad n_queries / max_total");
    if (device_ms) *device_ms = 0.f;
    if (nq == 0) return LVS_OK;
    if (!out_count || !out_index || !out_score || !out_leader) return fail(LVS_EINVAL, "NULL output");
    const int64_t nc = in->offsets[nq];
    int max_c = 0;
    for (int q = 0; q < nq; ++q) {
        const int c = in->offsets[q + 1] - in->offsets[q];
        if (c < 0) return fail(LVS_EINVAL, "offsets must be non-decreasing");
        max_c = std::max(max_c, c);
    }
    if (max_c > kRankMaxCand) return fail(LVS_ELIMIT, "%d candidates in one query exceed %d", max_c, kRankMaxCand);
    std::lock_guard<std::mutex> lk(g_rank_mu);
    if (!g_rank_stream) {
        CU(cudaStreamCreateWithFlags(&g_rank_stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&g_rank_ev[0]));
        CU(cudaEventCreate(&g_rank_ev[1]));
    }
    // one staging block: inputs then outputs, 16-byte aligned sections
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0;
    const size_t o_off = o; o = al(o + (size_t)(nq + 1) * 4);
    const size_t o_kind = o; o = al(o + (size_t)nc);
    const size_t o_key = o; o = al(o + (size_t)nc * 4);
    const size_t o_file = o; o = al(o + (size_t)nc * 4);
    const size_t o_depth = o; o = al(o + (size_t)nc * 4);
    const size_t o_em = o; o = al(o + (size_t)nc * 8);
    const size_t o_deg = o; o = al(o + (size_t)nc * 4);
    const size_t o_flags = o; o = al(o + (size_t)nc);
    const size_t o_clen = o; o = al(o + (size_t)nc * 4);
    const size_t o_vs = o; o = al(o + (size_t)nc * 8);
    const size_t o_w = o; o = al(o + (size_t)nq * 32);
    const size_t in_bytes = o;
    const size_t rows = (size_t)nq * max_total;
    const size_t r_count = o; o = al(o + (size_t)nq * 4);
    const size_t r_index = o; o = al(o + rows * 4);
    const size_t r_score = o; o = al(o + rows * 8);
    const size_t r_norm = o; o = al(o + rows * 8);
    const size_t r_sig = o; o = al(o + rows * 8 * kRankSignals);
    const size_t r_mask = o; o = al(o + rows);
    const size_t r_src = o; o = al(o + rows);
    const size_t r_lead = o; o = al(o + (size_t)std::max<int64_t>(nc, 1) * 4);
    const size_t total = o;
    int rc;
    if ((rc = ensure_pinned(g_rank_pin, total)) != LVS_OK) return rc;
    if ((rc = ensure_dev(g_rank_dev, total)) != LVS_OK) return rc;
    uint8_t* hp = (uint8_t*)g_rank_pin.p;
    uint8_t* dp = (uint8_t*)g_rank_dev.p;
    memcpy(hp + o_off, in->offsets, (size_t)(nq + 1) * 4);
    if (nc > 0) {
        if (!in->kind || !in->key_id || !in->file_id || !in->depth || !in->entity_match || !in->degree || !in->flags ||
            !in->content_len || !in->vscore) return fail(LVS_EINVAL, "NULL candidate array");
        memcpy(hp + o_kind, in->kind, (size_t)nc);
        memcpy(hp + o_key, in->key_id, (size_t)nc * 4);
        memcpy(hp + o_file, in->file_id, (size_t)nc * 4);
        memcpy(hp + o_depth, in->depth, (size_t)nc * 4);
        memcpy(hp + o_em, in->entity_match, (size_t)nc * 8);
        memcpy(hp + o_deg, in->degree, (size_t)nc * 4);
        memcpy(hp + o_flags, in->flags, (size_t)nc);
        memcpy(hp + o_clen, in->content_len, (size_t)nc * 4);
        memcpy(hp + o_vs, in->vscore, (size_t)nc * 8);
    }
    if (!in->weights) return fail(LVS_EINVAL, "NULL weights");
    memcpy(hp + o_w, in->weights, (size_t)nq * 32);
    cudaStream_t st = g_rank_stream;
    CU(cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, st));
    RankParams p;
    memset(&p, 0, sizeof(p));
    p.n_queries = nq; p.offsets = (const int32_t*)(dp + o_off); p.kind = dp + o_kind; p.key_id = (const uint32_t*)(dp + o_key);
    p.file_id = (const uint32_t*)(dp + o_file); p.depth = (const int32_t*)(dp + o_depth); p.entity_match = (const double*)(dp + o_em);
    p.degree = (const int32_t*)(dp + o_deg); p.flags = dp + o_flags; p.content_len = (const int32_t*)(dp + o_clen);
    p.vscore = (const double*)(dp + o_vs); p.weights = (const double*)(dp + o_w);
    p.mode = mode; p.max_per_file = max_per_file; p.max_total = max_total; p.entity_bonus = entity_bonus; p.rel_bonus = rel_bonus;
    p.out_count = (int32_t*)(dp + r_count); p.out_index = (int32_t*)(dp + r_index); p.out_score = (double*)(dp + r_score);
    p.out_norm = (double*)(dp + r_norm); p.out_signals = (double*)(dp + r_sig); p.out_sigmask = dp + r_mask; p.out_source = dp + r_src;
    p.out_leader = (int32_t*)(dp + r_lead);
    const size_t smem = rank_smem_bytes(std::max(max_c, 1));
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(rank_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaEventRecord(g_rank_ev[0], st));
    rank_fuse_kernel<<<nq, kRankThreads, smem, st>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(g_rank_ev[1], st));
    CU(cudaMemcpyAsync(hp + in_bytes, dp + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (device_ms) cudaEventElapsedTime(device_ms, g_rank_ev[0], g_rank_ev[1]);
    memcpy(out_count, hp + r_count, (size_t)nq * 4);
    memcpy(out_index, hp + r_index, rows * 4);
    memcpy(out_score, hp + r_score, rows * 8);
    if (out_norm) memcpy(out_norm, hp + r_norm, rows * 8);
    if (out_signals) memcpy(out_signals, hp + r_sig, rows * 8 * kRankSignals);
    if (out_sigmask) memcpy(out_sigmask, hp + r_mask, rows);
    if (out_source) memcpy(out_source, hp + r_src, rows);
    if (nc > 0) memcpy(out_leader, hp + r_lead, (size_t)nc * 4);
    return LVS_OK;
}

// ------------------------------------------------------------------------------------------------------
// fused search -> rank
// ------------------------------------------------------------------------------------------------------
template <typename T>
static int regrow(T*& ptr, int64_t old_n, int64_t new_n, cudaStream_t st, int fill = -1) {
    T* np_ = nullptr;
    cudaError_t e = cudaMalloc(&np_, (size_t)new_n * sizeof(T));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LVS_ENOMEM, "cannot allocate %lld ranking-attribute entries: %s", (long long)new_n, cudaGetErrorString(e)); }
    if (fill >= 0) CU(cudaMemsetAsync(np_, fill, (size_t)new_n * sizeof(T), st));
    if (ptr && old_n > 0) CU(cudaMemcpyAsync(np_, ptr, (size_t)old_n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(ptr);
    ptr = np_;
    return LVS_OK;
}

extern "C" int lvs_rank_names_append(lvs_collection* c, const uint8_t* bytes, const uint32_t* lens, int n, uint32_t* first_id) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n < 0 || (n > 0 && !lens)) return fail(LVS_EINVAL, "bad name batch");
    std::lock_guard<std::mutex> lk(c->mu);
    if (first_id) *first_id = (uint32_t)c->n_names;
    if (n == 0) return LVS_OK;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) total += lens[i];
    if (total > 0 && !bytes) return fail(LVS_EINVAL, "NULL name bytes");
    if ((uint64_t)c->name_bytes_used + total >= 0xFFFFFFFFull) return fail(LVS_ELIMIT, "entity-name pool exceeds 4 GiB");
    cudaStream_t st = c->stream;
    int rc;
    if (c->n_names + n + 1 > c->names_cap) {
        const int64_t nc = std::max<int64_t>(c->n_names + n + 1, std::max<int64_t>(1024, c->names_cap * 2));
        if ((rc = regrow(c->d_name_off, c->names_cap ? c->n_names + 1 : 0, nc, st, 0)) != LVS_OK) return rc;
        c->names_cap = nc;
    }
    if (c->name_bytes_used + (int64_t)total > c->name_bytes_cap) {
        const int64_t nb = std::max<int64_t>(c->name_bytes_used + (int64_t)total, std::max<int64_t>(65536, c->name_bytes_cap * 2));
        if ((rc = regrow(c->d_name_bytes, c->name_bytes_used, nb, st)) != LVS_OK) return rc;
        c->name_bytes_cap = nb;
    }
    std::vector<uint32_t> offs((size_t)n + 1);
    uint32_t o = (uint32_t)c->name_bytes_used;
    for (int i = 0; i < n; ++i) { offs[i] = o; o += lens[i]; }
    offs[n] = o;
    CU(cudaMemcpyAsync(c->d_name_off + c->n_names, offs.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, st));
    if (total) CU(cudaMemcpyAsync(c->d_name_bytes + c->name_bytes_used, bytes, (size_t)total, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    c->n_names += n;
    c->name_bytes_used += (int64_t)total;
    return LVS_OK;
}

extern "C" int lvs_rank_attrs_set(lvs_collection* c, const int64_t* rows, int n, const uint32_t* key_id, const uint32_t* file_id,
                                  const uint32_t* cent_id, const uint32_t* name_id, const int32_t* content_len, const uint8_t* flags) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n < 0) return fail(LVS_EINVAL, "bad n");
    if (n == 0) return LVS_OK;
    if (!rows || !key_id || !file_id || !cent_id || !name_id || !content_len || !flags) return fail(LVS_EINVAL, "NULL attribute array");
    std::lock_guard<std::mutex> lk(c->mu);
    int64_t max_row = -1;
    for (int i = 0; i < n; ++i) {
        const int64_t r = rows[i] - c->row_base;
        if (r < 0 || r >= c->n_rows) return fail(LVS_EINVAL, "row %lld is not in the collection", (long long)rows[i]);
        if ((int64_t)name_id[i] >= c->n_names) return fail(LVS_EINVAL, "name id %u has not been appended", name_id[i]);
        max_row = std::max(max_row, r);
    }
    cudaStream_t st = c->stream;
    int rc;
    if (max_row >= c->rk_cap) {
        const int64_t nc = std::max<int64_t>(std::max<int64_t>(max_row + 1, c->capacity), c->rk_cap * 2);
        if ((rc = regrow(c->d_rk_key, c->rk_cap, nc, st, 0xFF)) != LVS_OK) return rc;
        if ((rc = regrow(c->d_rk_file, c->rk_cap, nc, st, 0xFF)) != LVS_OK) return rc;
        if ((rc = regrow(c->d_rk_cent, c->rk_cap, nc, st, 0xFF)) != LVS_OK) return rc;
        if ((rc = regrow(c->d_rk_name, c->rk_cap, nc, st, 0xFF)) != LVS_OK) return rc;
        if ((rc = regrow(c->d_rk_clen, c->rk_cap, nc, st, 0xFF)) != LVS_OK) return rc;
        if ((rc = regrow(c->d_rk_flags, c->rk_cap, nc, st, 0)) != LVS_OK) return rc;
        c->rk_cap = nc;
    }
    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0;
    const size_t o_rows = o; o = al(o + (size_t)n * 8);
    const size_t o_key = o; o = al(o + (size_t)n * 4);
    const size_t o_file = o; o = al(o + (size_t)n * 4);
    const size_t o_cent = o; o = al(o + (size_t)n * 4);
    const size_t o_name = o; o = al(o + (size_t)n * 4);
    const size_t o_clen = o; o = al(o + (size_t)n * 4);
    const size_t o_flags = o; o = al(o + (size_t)n);
    if ((rc = ensure_pinned(c->h_rk_pin, o)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_rk_dev, o)) != LVS_OK) return rc;
    uint8_t* hp = (uint8_t*)c->h_rk_pin.p;
    uint8_t* dp = (uint8_t*)c->s_rk_dev.p;
    int64_t* hr = (int64_t*)(hp + o_rows);
    for (int i = 0; i < n; ++i) hr[i] = rows[i] - c->row_base;
    memcpy(hp + o_key, key_id, (size_t)n * 4); memcpy(hp + o_file, file_id, (size_t)n * 4); memcpy(hp + o_cent, cent_id, (size_t)n * 4);
    memcpy(hp + o_name, name_id, (size_t)n * 4); memcpy(hp + o_clen, content_len, (size_t)n * 4); memcpy(hp + o_flags, flags, (size_t)n);
    CU(cudaMemcpyAsync(dp, hp, o, cudaMemcpyHostToDevice, st));
    RankAttrScatter sp;
    sp.rows = (const int64_t*)(dp + o_rows); sp.n = n;
    sp.key = (const uint32_t*)(dp + o_key); sp.file = (const uint32_t*)(dp + o_file); sp.cent = (const uint32_t*)(dp + o_cent);
    sp.name = (const uint32_t*)(dp + o_name); sp.clen = (const int32_t*)(dp + o_clen); sp.flags = dp + o_flags;
    sp.row_key = c->d_rk_key; sp.row_file = c->d_rk_file; sp.row_cent = c->d_rk_cent; sp.row_name = c->d_rk_name;
    sp.row_clen = c->d_rk_clen; sp.row_flags = c->d_rk_flags;
    rank_attr_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(sp);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return LVS_OK;
}

extern "C" int lvs_search_rank2(lvs_collection* c, lvs_collection* c2, const void* queries, int dtype, int Q, int k, int k2,
                                const uint32_t* want, const uint32_t* want2, const int32_t* sel2, int Q2,
                                const lvs_rank_batch* graph, const lvs_rank_query_ctx* ctx, int max_per_file, int max_total,
                                double entity_bonus, double rel_bonus, const lvs_rank_hits* hits1, const lvs_rank_hits* hits2,
                                int32_t* out_count, int32_t* out_index, double* out_score, double* out_signals, uint8_t* out_sigmask,
                                uint8_t* out_source, int32_t* out_leader, float* device_ms) {
    bind_thread();
    int rc = check_search_args(c, queries, dtype, Q, k);
    if (rc != LVS_OK) return rc;
    if (Q == 0) return LVS_OK;
    if (!graph || !graph->offsets || !graph->weights || graph->n_queries != Q) return fail(LVS_EINVAL, "graph batch must describe the same %d queries", Q);
    if (!ctx || !ctx->ent_off || !ctx->ent_str_off || !ctx->cen_off) return fail(LVS_EINVAL, "NULL query context");
    if (max_total < 1) return fail(LVS_EINVAL, "bad max_total");
    if (!hits1 || !hits1->scores || !hits1->rows || !hits1->counts || !hits1->flags || !out_count || !out_index || !out_score || !out_leader)
        return fail(LVS_EINVAL, "NULL output");
    const bool two = c2 != nullptr && Q2 > 0;
    if (c2 == c) return fail(LVS_EINVAL, "the second collection must be a different one");
    if (two) {
        if (k2 < 1 || k2 > LVS_MAX_K) return fail(LVS_ELIMIT, "limit %d outside 1..%d", k2, LVS_MAX_K);
        if (Q2 > Q || !sel2) return fail(LVS_EINVAL, "bad query subset for the second collection");
        if (c2->dim != c->dim) return fail(LVS_EINVAL, "the two collections have different dimensions (%d, %d)", c->dim, c2->dim);
        if (!hits2 || !hits2->scores || !hits2->rows || !hits2->counts || !hits2->flags) return fail(LVS_EINVAL, "NULL output (second collection)");
        for (int j = 0; j < Q2; ++j) if (sel2[j] < 0 || sel2[j] >= Q || (j > 0 && sel2[j] <= sel2[j - 1])) return fail(LVS_EINVAL, "sel2 must be increasing query indices");
    } else { k2 = 0; Q2 = 0; }
    std::unique_lock<std::mutex> lk(c->mu, std::defer_lock), lk2;
    if (two) { lk2 = std::unique_lock<std::mutex>(c2->mu, std::defer_lock); std::lock(lk, lk2); } else lk.lock();
    if (c->rk_cap < c->n_rows || !c->d_rk_key) return fail(LVS_ESTATE, "ranking attributes are not set for every row (lvs_rank_attrs_set)");
    if (two && c2->n_rows > 0 && (c2->rk_cap < c2->n_rows || !c2->d_rk_key))
        return fail(LVS_ESTATE, "ranking attributes are not set for every row of the second collection");
    const int kk = k + k2;                                   // hit slots per query
    const int32_t* goff = graph->offsets;
    const int64_t ngc = goff[Q];
    int max_c = 0;
    for (int q = 0; q < Q; ++q) {
        const int g = goff[q + 1] - goff[q];
        if (g < 0) return fail(LVS_EINVAL, "offsets must be non-decreasing");
        max_c = std::max(max_c, g + kk);
    }
    if (max_c > kRankMaxCand) return fail(LVS_ELIMIT, "%d candidates in one query exceed %d", max_c, kRankMaxCand);
    if (ngc > 0 && (!graph->kind || !graph->key_id || !graph->file_id || !graph->depth || !graph->entity_match || !graph->degree ||
                    !graph->flags)) return fail(LVS_EINVAL, "NULL graph candidate array");
    const int64_t nc = ngc + (int64_t)Q * kk;                // combined candidates: per query its graph candidates, then the hit slots
    const int n_ent = ctx->ent_off[Q], n_cen = ctx->cen_off[Q];
    const size_t ent_bytes = n_ent > 0 ? ctx->ent_str_off[n_ent] : 0;
    if (n_ent > 0 && ent_bytes > 0 && !ctx->ent_bytes) return fail(LVS_EINVAL, "NULL entity bytes");
    if (n_cen > 0 && (!ctx->cen_id || !ctx->cen_deg)) return fail(LVS_EINVAL, "NULL centrality table");

    auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t esz = dt_size(dtype);
    const size_t qrow = (size_t)c->dim * esz;
    size_t o = 0;
    const size_t o_q = o; o = al(o + (size_t)Q * qrow);
    const size_t o_q2 = o; o = al(o + (size_t)Q2 * qrow);
    const size_t o_sel = o; o = al(o + (size_t)std::max(Q2, 1) * 4);
    const size_t o_off = o; o = al(o + (size_t)(Q + 1) * 4);
    const size_t o_ng = o; o = al(o + (size_t)Q * 4);
    const size_t o_kind = o; o = al(o + (size_t)nc);
    const size_t o_key = o; o = al(o + (size_t)nc * 4);
    const size_t o_file = o; o = al(o + (size_t)nc * 4);
    const size_t o_depth = o; o = al(o + (size_t)nc * 4);
    const size_t o_em = o; o = al(o + (size_t)nc * 8);
    const size_t o_deg = o; o = al(o + (size_t)nc * 4);
    const size_t o_flags = o; o = al(o + (size_t)nc);
    const size_t o_clen = o; o = al(o + (size_t)nc * 4);
    const size_t o_vs = o; o = al(o + (size_t)nc * 8);
    const size_t o_w = o; o = al(o + (size_t)Q * 32);
    const size_t o_eoff = o; o = al(o + (size_t)(Q + 1) * 4);
    const size_t o_esoff = o; o = al(o + (size_t)(n_ent + 1) * 4);
    const size_t o_eb = o; o = al(o + ent_bytes);
    const size_t o_coff = o; o = al(o + (size_t)(Q + 1) * 4);
    const size_t o_cid = o; o = al(o + (size_t)std::max(n_cen, 1) * 4);
    const size_t o_cdeg = o; o = al(o + (size_t)std::max(n_cen, 1) * 4);
    const size_t in_bytes = o;
    const size_t nres = (size_t)Q * k, nres2 = (size_t)Q2 * std::max(k2, 1), rows = (size_t)Q * max_total;
    // outputs (device -> host in one copy)
    const size_t r_hs = o; o = al(o + nres * 8);
    const size_t r_hr = o; o = al(o + nres * 8);
    const size_t r_ht = o; o = al(o + nres * 8);
    const size_t r_hc = o; o = al(o + (size_t)Q * 4);
    const size_t r_hf = o; o = al(o + (size_t)Q * 4);
    const size_t r_hs2 = o; o = al(o + nres2 * 8);
    const size_t r_hr2 = o; o = al(o + nres2 * 8);
    const size_t r_ht2 = o; o = al(o + nres2 * 8);
    const size_t r_hc2 = o; o = al(o + (size_t)std::max(Q2, 1) * 4);
    const size_t r_hf2 = o; o = al(o + (size_t)std::max(Q2, 1) * 4);
    const size_t r_cnt = o; o = al(o + (size_t)Q * 4);     // candidates present per query (gather output)
    const size_t r_err = o; o = al(o + 4);
    const size_t r_count = o; o = al(o + (size_t)Q * 4);
    const size_t r_index = o; o = al(o + rows * 4);
    const size_t r_score = o; o = al(o + rows * 8);
    const size_t r_norm = o; o = al(o + rows * 8);
    const size_t r_sig = o; o = al(o + rows * 8 * kRankSignals);
    const size_t r_mask = o; o = al(o + rows);
    const size_t r_src = o; o = al(o + rows);
    const size_t r_lead = o; o = al(o + (size_t)nc * 4);
    const size_t total = o;
    if ((rc = ensure_pinned(c->h_rk_pin, total)) != LVS_OK) return rc;
    if ((rc = ensure_dev(c->s_rk_dev, total)) != LVS_OK) return rc;
    uint8_t* hp = (uint8_t*)c->h_rk_pin.p;
    uint8_t* dp = (uint8_t*)c->s_rk_dev.p;
    memcpy(hp + o_q, queries, (size_t)Q * qrow);
    for (int j = 0; j < Q2; ++j) {
        memcpy(hp + o_q2 + (size_t)j * qrow, (const uint8_t*)queries + (size_t)sel2[j] * qrow, qrow);
        ((int32_t*)(hp + o_sel))[j] = sel2[j];
    }
    int32_t* off2 = (int32_t*)(hp + o_off);
    int32_t* ngq = (int32_t*)(hp + o_ng);
    for (int q = 0; q <= Q; ++q) off2[q] = goff[q] + q * kk;
    for (int q = 0; q < Q; ++q) {
        const int g = goff[q + 1] - goff[q];
        ngq[q] = g;
        if (g == 0) continue;
        const size_t src = (size_t)goff[q], dst = (size_t)off2[q];
        memcpy(hp + o_kind + dst, graph->kind + src, (size_t)g);
        memcpy(hp + o_key + dst * 4, graph->key_id + src, (size_t)g * 4);
        memcpy(hp + o_file + dst * 4, graph->file_id + src, (size_t)g * 4);
        memcpy(hp + o_depth + dst * 4, graph->depth + src, (size_t)g * 4);
        memcpy(hp + o_em + dst * 8, graph->entity_match + src, (size_t)g * 8);
        memcpy(hp + o_deg + dst * 4, graph->degree + src, (size_t)g * 4);
        memcpy(hp + o_flags + dst, graph->flags + src, (size_t)g);
        for (int i = 0; i < g; ++i) { ((int32_t*)(hp + o_clen))[dst + i] = -1; ((double*)(hp + o_vs))[dst + i] = 0.0; }
    }
    memcpy(hp + o_w, graph->weights, (size_t)Q * 32);
    memcpy(hp + o_eoff, ctx->ent_off, (size_t)(Q + 1) * 4);
    memcpy(hp + o_esoff, ctx->ent_str_off, (size_t)(n_ent + 1) * 4);
    if (ent_bytes) memcpy(hp + o_eb, ctx->ent_bytes, ent_bytes);
    memcpy(hp + o_coff, ctx->cen_off, (size_t)(Q + 1) * 4);
    if (n_cen > 0) { memcpy(hp + o_cid, ctx->cen_id, (size_t)n_cen * 4); memcpy(hp + o_cdeg, ctx->cen_deg, (size_t)n_cen * 4); }

    cudaStream_t st = c->stream;
    for (auto& e : c->rk_ev) if (!e) CU(cudaEventCreate(&e));
    CU(cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dp + r_err, 0, 4, st));
    CU(cudaEventRecord(c->rk_ev[0], st));
    // 1. the search(es): top-k rows and float64 scores stay on the device
    rc = search_core(c, dp + o_q, dtype, Q, k, want, (double*)(dp + r_hs), (int64_t*)(dp + r_hr), (uint64_t*)(dp + r_ht),
                     (uint32_t*)(dp + r_hc), nullptr, (int32_t*)(dp + r_hf), true, st);
    if (rc != LVS_OK) return rc;
    if (two) {
        if (c2->n_rows > 0) {
            rc = search_core(c2, dp + o_q2, dtype, Q2, k2, want2, (double*)(dp + r_hs2), (int64_t*)(dp + r_hr2), (uint64_t*)(dp + r_ht2),
                             (uint32_t*)(dp + r_hc2), nullptr, (int32_t*)(dp + r_hf2), true, st);
            if (rc != LVS_OK) return rc;
        } else {
            CU(cudaMemsetAsync(dp + r_hc2, 0, (size_t)Q2 * 4, st));
            CU(cudaMemsetAsync(dp + r_hf2, 0, (size_t)Q2 * 4, st));
            CU(cudaMemsetAsync(dp + r_hr2, 0xFF, nres2 * 8, st));
            CU(cudaMemsetAsync(dp + r_hs2, 0, nres2 * 8, st));
        }
    }
    CU(cudaEventRecord(c->rk_ev[1], st));
    // 2. hits -> vector candidates, 3. K3
    RankGatherParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.k = k; gp.offsets = (const int32_t*)(dp + o_off); gp.n_graph = (const int32_t*)(dp + o_ng);
    gp.hit_scores = (const double*)(dp + r_hs); gp.hit_rows = (const int64_t*)(dp + r_hr); gp.hit_counts = (const uint32_t*)(dp + r_hc);
    gp.row_base = c->row_base; gp.attr_rows = std::min(c->rk_cap, c->n_rows);
    gp.row_key = c->d_rk_key; gp.row_file = c->d_rk_file; gp.row_cent = c->d_rk_cent; gp.row_name = c->d_rk_name;
    gp.row_clen = c->d_rk_clen; gp.row_flags = c->d_rk_flags;
    gp.name_off = c->d_name_off; gp.name_bytes = c->d_name_bytes; gp.n_names = (uint32_t)c->n_names;
    gp.ent_off = (const int32_t*)(dp + o_eoff); gp.ent_str_off = (const uint32_t*)(dp + o_esoff); gp.ent_bytes = dp + o_eb;
    gp.cen_off = (const int32_t*)(dp + o_coff); gp.cen_id = (const uint32_t*)(dp + o_cid); gp.cen_deg = (const int32_t*)(dp + o_cdeg);
    gp.kind = dp + o_kind; gp.key_id = (uint32_t*)(dp + o_key); gp.file_id = (uint32_t*)(dp + o_file); gp.depth = (int32_t*)(dp + o_depth);
    gp.entity_match = (double*)(dp + o_em); gp.degree = (int32_t*)(dp + o_deg); gp.flags = dp + o_flags;
    gp.content_len = (int32_t*)(dp + o_clen); gp.vscore = (double*)(dp + o_vs);
    gp.counts = (int32_t*)(dp + r_cnt); gp.error = (int32_t*)(dp + r_err);
    gp.sel = nullptr;
    rank_gather_kernel<<<Q, 128, 0, st>>>(gp);
    CU(cudaGetLastError());
    if (two) {
        RankGatherParams g2 = gp;
        g2.k = k2; g2.hit_scores = (const double*)(dp + r_hs2); g2.hit_rows = (const int64_t*)(dp + r_hr2);
        g2.hit_counts = (const uint32_t*)(dp + r_hc2);
        g2.row_base = c2->row_base; g2.attr_rows = std::min(c2->rk_cap, c2->n_rows);
        g2.row_key = c2->d_rk_key; g2.row_file = c2->d_rk_file; g2.row_cent = c2->d_rk_cent; g2.row_name = c2->d_rk_name;
        g2.row_clen = c2->d_rk_clen; g2.row_flags = c2->d_rk_flags;
        g2.name_off = c2->d_name_off; g2.name_bytes = c2->d_name_bytes; g2.n_names = (uint32_t)c2->n_names;
        g2.sel = (const int32_t*)(dp + o_sel);
        rank_gather_kernel<<<Q2, 128, 0, st>>>(g2);
        CU(cudaGetLastError());
    }
    RankParams p;
    memset(&p, 0, sizeof(p));
    p.n_queries = Q; p.offsets = gp.offsets; p.counts = gp.counts; p.kind = gp.kind; p.key_id = gp.key_id; p.file_id = gp.file_id;
    p.depth = gp.depth; p.entity_match = gp.entity_match; p.degree = gp.degree; p.flags = gp.flags; p.content_len = gp.content_len;
    p.vscore = gp.vscore; p.weights = (const double*)(dp + o_w);
    p.mode = 0; p.max_per_file = max_per_file; p.max_total = max_total; p.entity_bonus = entity_bonus; p.rel_bonus = rel_bonus;
    p.out_count = (int32_t*)(dp + r_count); p.out_index = (int32_t*)(dp + r_index); p.out_score = (double*)(dp + r_score);
    p.out_norm = (double*)(dp + r_norm); p.out_signals = (double*)(dp + r_sig); p.out_sigmask = dp + r_mask; p.out_source = dp + r_src;
    p.out_leader = (int32_t*)(dp + r_lead);
    const size_t smem = rank_smem_bytes(std::max(max_c, 1));
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(rank_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rank_fuse_kernel<<<Q, kRankThreads, smem, st>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->rk_ev[2], st));
    CU(cudaMemcpyAsync(hp + in_bytes, dp + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->last_launches += two ? 3 : 2;
    if (device_ms) {
        cudaEventElapsedTime(&device_ms[0], c->rk_ev[0], c->rk_ev[1]);
        cudaEventElapsedTime(&device_ms[1], c->rk_ev[1], c->rk_ev[2]);
    }
    if (*(int32_t*)(hp + r_err)) return fail(LVS_ESTATE, "a hit row has no ranking attributes");
    memcpy(hits1->scores, hp + r_hs, nres * 8);
    memcpy(hits1->rows, hp + r_hr, nres * 8);
    memcpy(hits1->counts, hp + r_hc, (size_t)Q * 4);
    memcpy(hits1->flags, hp + r_hf, (size_t)Q * 4);
    if (two) {
        memcpy(hits2->scores, hp + r_hs2, (size_t)Q2 * k2 * 8);
        memcpy(hits2->rows, hp + r_hr2, (size_t)Q2 * k2 * 8);
        memcpy(hits2->counts, hp + r_hc2, (size_t)Q2 * 4);
        memcpy(hits2->flags, hp + r_hf2, (size_t)Q2 * 4);
    }
    memcpy(out_count, hp + r_count, (size_t)Q * 4);
    memcpy(out_index, hp + r_index, rows * 4);
    memcpy(out_score, hp + r_score, rows * 8);
    if (out_signals) memcpy(out_signals, hp + r_sig, rows * 8 * kRankSignals);
    if (out_sigmask) memcpy(out_sigmask, hp + r_mask, rows);
    if (out_source) memcpy(out_source, hp + r_src, rows);
    memcpy(out_leader, hp + r_lead, (size_t)nc * 4);
    return LVS_OK;
}

extern "C" int lvs_search_rank(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want,
                               const lvs_rank_batch* graph, const lvs_rank_query_ctx* ctx, int max_per_file, int max_total,
                               double entity_bonus, double rel_bonus, double* out_hit_scores, int64_t* out_hit_rows,
                               uint32_t* out_hit_counts, int32_t* out_flags, int32_t* out_count, int32_t* out_index, double* out_score,
                               double* out_signals, uint8_t* out_sigmask, uint8_t* out_source, int32_t* out_leader, float* device_ms) {
    bind_thread();
    lvs_rank_hits h1;
    h1.scores = out_hit_scores; h1.rows = out_hit_rows; h1.counts = out_hit_counts; h1.flags = out_flags;
    return lvs_search_rank2(c, nullptr, queries, dtype, Q, k, 0, want, nullptr, nullptr, 0, graph, ctx, max_per_file, max_total,
                            entity_bonus, rel_bonus, &h1, nullptr, out_count, out_index, out_score, out_signals, out_sigmask, out_source,
                            out_leader, device_ms);
}

// ------------------------------------------------------------------------------------------------------
// snapshots: a shard survives a restart the way the Qdrant volume does
// (reference docker-compose.yml:42-43).  One file per shard: header, then the raw device arrays in row order.
// ------------------------------------------------------------------------------------------------------
struct SnapshotHeader {
    char magic[8];                // "LVSSNAP2" (64-bit write epochs and search counter)
    int32_t dim, storage, metric, n_cols;
    int64_t n_rows, row_base;
    uint32_t reserved0, row_bytes;
    float norm_stats[2];
    int64_t rk_rows;              // rows covered by the ranking attribute columns (0 = none)
    int64_t n_names, name_bytes;
    uint64_t search_counter;
    int64_t reserved[3];
};

static int snap_write(FILE* f, const void* dptr, size_t bytes, Scratch& pin, cudaStream_t st) {
    const size_t chunk = (size_t)32 << 20;
    int rc;
    if ((rc = ensure_pinned(pin, std::min(std::max(bytes, (size_t)4096), chunk))) != LVS_OK) return rc;
    for (size_t o = 0; o < bytes; o += chunk) {
        const size_t n = std::min(chunk, bytes - o);
        CU(cudaMemcpyAsync(pin.p, (const uint8_t*)dptr + o, n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (fwrite(pin.p, 1, n, f) != n) return fail(LVS_EINVAL, "short write to the snapshot file");
    }
    return LVS_OK;
}
static int snap_read(FILE* f, void* dptr, size_t bytes, Scratch& pin, cudaStream_t st) {
    const size_t chunk = (size_t)32 << 20;
    int rc;
    if ((rc = ensure_pinned(pin, std::min(std::max(bytes, (size_t)4096), chunk))) != LVS_OK) return rc;
    for (size_t o = 0; o < bytes; o += chunk) {
        const size_t n = std::min(chunk, bytes - o);
        if (fread(pin.p, 1, n, f) != n) return fail(LVS_EINVAL, "snapshot file is truncated");
        CU(cudaMemcpyAsync((uint8_t*)dptr + o, pin.p, n, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    return LVS_OK;
}

extern "C" int lvs_snapshot_save(lvs_collection* c, const char* path) {
    bind_thread();
    if (!c || !path) return fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    for (auto& sl : c->slots) if (sl.in_use) return fail(LVS_ESTATE, "searches are in flight: call lvs_search_wait first");
    cudaStream_t st = c->stream;
    CU(cudaStreamSynchronize(st));
    FILE* f = fopen(path, "wb");
    if (!f) return fail(LVS_EINVAL, "cannot open %s for writing", path);
    SnapshotHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "LVSSNAP2", 8);
    h.dim = c->dim; h.storage = c->storage; h.metric = c->metric; h.n_cols = c->n_cols;
    h.n_rows = c->n_rows; h.row_base = c->row_base; h.search_counter = c->search_counter; h.row_bytes = c->row_bytes;
    h.norm_stats[0] = c->h_norm_stats[0]; h.norm_stats[1] = c->h_norm_stats[1];
    h.rk_rows = c->d_rk_key ? std::min(c->rk_cap, c->n_rows) : 0;
    h.n_names = c->n_names; h.name_bytes = c->name_bytes_used;
    int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? LVS_OK : fail(LVS_EINVAL, "short write to the snapshot file");
    const size_t n = (size_t)c->n_rows;
    Scratch& pin = c->h_pin2;
    if (rc == LVS_OK && n) rc = snap_write(f, c->d_vec, n * c->row_bytes, pin, st);
    if (rc == LVS_OK && n) rc = snap_write(f, c->d_live, n, pin, st);
    if (rc == LVS_OK && n) rc = snap_write(f, c->d_epoch, n * 8, pin, st);
    if (rc == LVS_OK && n) rc = snap_write(f, c->d_tie, n * 8, pin, st);
    if (rc == LVS_OK && n) rc = snap_write(f, c->d_inv_norm, n * 4, pin, st);
    for (int i = 0; i < c->n_cols && rc == LVS_OK && n; ++i) rc = snap_write(f, c->d_codes[i], n * 4, pin, st);
    if (rc == LVS_OK && h.rk_rows) {
        const size_t r = (size_t)h.rk_rows;
        rc = snap_write(f, c->d_rk_key, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_write(f, c->d_rk_file, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_write(f, c->d_rk_cent, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_write(f, c->d_rk_name, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_write(f, c->d_rk_clen, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_write(f, c->d_rk_flags, r, pin, st);
    }
    if (rc == LVS_OK && h.n_names) {
        rc = snap_write(f, c->d_name_off, ((size_t)h.n_names + 1) * 4, pin, st);
        if (rc == LVS_OK && h.name_bytes) rc = snap_write(f, c->d_name_bytes, (size_t)h.name_bytes, pin, st);
    }
    if (fclose(f) != 0 && rc == LVS_OK) rc = fail(LVS_EINVAL, "closing %s failed", path);
    return rc;
}

extern "C" int lvs_snapshot_load(const char* path, const char* name, int64_t capacity_rows, lvs_collection** out) {
    bind_thread();
    if (!path || !out) return fail(LVS_EINVAL, "NULL argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(LVS_EINVAL, "cannot open %s", path);
    SnapshotHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "LVSSNAP2", 8) != 0) { fclose(f); return fail(LVS_EINVAL, "%s is not a lattice-b200 snapshot", path); }
    // the header must describe a file of exactly this size before anything is allocated from it
    if (h.n_rows < 0 || h.n_rows >= (int64_t)0xFFFFFFF0ll || h.row_base < 0 || h.dim < 1 || h.dim > 8192 || h.n_cols < 0 ||
        h.n_cols > kMaxFilterCols || h.rk_rows < 0 || h.rk_rows > h.n_rows || h.n_names < 0 || h.name_bytes < 0 ||
        h.n_names > (int64_t)0xFFFFFFFFll || h.name_bytes > (int64_t)0xFFFFFFFFll || h.row_bytes == 0 || h.row_bytes > 8192u * 4u) {
        fclose(f);
        return fail(LVS_EINVAL, "%s: corrupt snapshot header", path);
    }
    {
        const int64_t expect = (int64_t)sizeof(h) + h.n_rows * ((int64_t)h.row_bytes + 1 + 8 + 8 + 4 + 4 * (int64_t)h.n_cols) + h.rk_rows * 21 +
                               (h.n_names ? (h.n_names + 1) * 4 + h.name_bytes : 0);
        if (fseek(f, 0, SEEK_END) != 0 || (int64_t)ftell(f) != expect || fseek(f, (long)sizeof(h), SEEK_SET) != 0) {
            fclose(f);
            return fail(LVS_EINVAL, "%s: snapshot size does not match its header (expected %lld bytes)", path, (long long)expect);
        }
    }
    lvs_collection* c = nullptr;
    int rc = lvs_collection_create(name, h.dim, h.storage, h.metric, h.n_cols, std::max(capacity_rows, h.n_rows), h.row_base, &c);
    if (rc != LVS_OK) { fclose(f); return rc; }
    if (c->row_bytes != h.row_bytes) { fclose(f); lvs_collection_destroy(c); return fail(LVS_EINVAL, "snapshot row layout differs (%u vs %u bytes)", h.row_bytes, c->row_bytes); }
    cudaStream_t st = c->stream;
    const size_t n = (size_t)h.n_rows;
    Scratch& pin = c->h_pin2;
    if (n) rc = snap_read(f, c->d_vec, n * c->row_bytes, pin, st);
    if (rc == LVS_OK && n) rc = snap_read(f, c->d_live, n, pin, st);
    if (rc == LVS_OK && n) rc = snap_read(f, c->d_epoch, n * 8, pin, st);
    if (rc == LVS_OK && n) rc = snap_read(f, c->d_tie, n * 8, pin, st);
    if (rc == LVS_OK && n) rc = snap_read(f, c->d_inv_norm, n * 4, pin, st);
    for (int i = 0; i < c->n_cols && rc == LVS_OK && n; ++i) rc = snap_read(f, c->d_codes[i], n * 4, pin, st);
    if (rc == LVS_OK) {
        c->n_rows = h.n_rows; c->search_counter = h.search_counter;
        c->h_norm_stats[0] = h.norm_stats[0]; c->h_norm_stats[1] = h.norm_stats[1];
        cudaError_t e = cudaMemcpy(c->d_max_norm, c->h_norm_stats, 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = fail(LVS_ECUDA, "snapshot load: %s", cudaGetErrorString(e));
    }
    if (rc == LVS_OK && h.rk_rows) {
        const int64_t cap = std::max<int64_t>(h.rk_rows, c->capacity);
        const size_t r = (size_t)h.rk_rows;
        rc = regrow(c->d_rk_key, 0, cap, st, 0xFF);
        if (rc == LVS_OK) rc = regrow(c->d_rk_file, 0, cap, st, 0xFF);
        if (rc == LVS_OK) rc = regrow(c->d_rk_cent, 0, cap, st, 0xFF);
        if (rc == LVS_OK) rc = regrow(c->d_rk_name, 0, cap, st, 0xFF);
        if (rc == LVS_OK) rc = regrow(c->d_rk_clen, 0, cap, st, 0xFF);
        if (rc == LVS_OK) rc = regrow(c->d_rk_flags, 0, cap, st, 0);
        if (rc == LVS_OK) c->rk_cap = cap;
        if (rc == LVS_OK) rc = snap_read(f, c->d_rk_key, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_read(f, c->d_rk_file, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_read(f, c->d_rk_cent, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_read(f, c->d_rk_name, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_read(f, c->d_rk_clen, r * 4, pin, st);
        if (rc == LVS_OK) rc = snap_read(f, c->d_rk_flags, r, pin, st);
    }
    if (rc == LVS_OK && h.n_names) {
        rc = regrow(c->d_name_off, 0, h.n_names + 1, st, 0);
        if (rc == LVS_OK) { c->names_cap = h.n_names + 1; rc = snap_read(f, c->d_name_off, ((size_t)h.n_names + 1) * 4, pin, st); }
        if (rc == LVS_OK && h.name_bytes) {
            rc = regrow(c->d_name_bytes, 0, h.name_bytes, st);
            if (rc == LVS_OK) { c->name_bytes_cap = h.name_bytes; rc = snap_read(f, c->d_name_bytes, (size_t)h.name_bytes, pin, st); }
        }
        if (rc == LVS_OK) { c->n_names = h.n_names; c->name_bytes_used = h.name_bytes; }
    }
    fclose(f);
    if (rc != LVS_OK) { lvs_collection_destroy(c); return rc; }
    *out = c;
    return LVS_OK;
}

// ------------------------------------------------------------------------------------------------------
// instrumentation
// ------------------------------------------------------------------------------------------------------
extern "C" int lvs_last_search_timing(const lvs_collection* c, float* ms4, int* n_launches, int* kernel_kind) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (ms4) memcpy(ms4, c->last_ms, sizeof(float) * 4);
    if (n_launches) *n_launches = c->last_launches;
    if (kernel_kind) *kernel_kind = c->last_kind;
    return LVS_OK;
}

extern "C" int lvs_last_kernel_phases(lvs_collection* c, uint64_t* ns16) {
    bind_thread();
    if (!c || !ns16) return fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->s_dbg_times.p) return fail(LVS_ESTATE, "option dbg_times was not set before the search");
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(ns16, c->s_dbg_times.p, 128, cudaMemcpyDeviceToHost));
    return LVS_OK;
}

extern "C" int lvs_set_option(lvs_collection* c, const char* name, int value) {
    bind_thread();
    if (!c || !name) return fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!strcmp(name, "stage_kb")) c->opt_stage_kb = value;
    else if (!strcmp(name, "stages")) c->opt_stages = value;
    else if (!strcmp(name, "grid")) c->opt_grid = std::max(0, std::min(value, g_lib.sm_count));   // the scratch lists are sized for one CTA per SM
    else if (!strcmp(name, "force_kpl")) c->opt_force_kpl = value;
    else if (!strcmp(name, "timing")) c->opt_timing = value ? 1 : 0;
    else if (!strcmp(name, "pdl")) c->opt_pdl = value ? 1 : 0;
    else if (!strcmp(name, "dyn_tiles")) c->opt_dyn_tiles = value ? 1 : 0;
    else if (!strcmp(name, "dbg_times")) c->opt_dbg_times = value ? 1 : 0;
    else if (!strcmp(name, "inline_query")) c->opt_inline_query = value ? 1 : 0;
    else if (!strcmp(name, "inline_max_mb")) c->opt_inline_max_mb = value < 0 ? 0 : value;
    else if (!strcmp(name, "gemm_min_q")) c->opt_gemm_min_q = value;
    else if (!strcmp(name, "path")) c->opt_path = value;
    else if (!strcmp(name, "gemm_dbg")) c->opt_gemm_dbg = value;
    else if (!strcmp(name, "gemm_stages")) c->opt_gemm_stages = value;
    else if (!strcmp(name, "gemm_stages_b")) c->opt_gemm_stages_b = std::max(-1, std::min(value, kGemmMaxStagesB));
    else if (!strcmp(name, "gemm_no_unit")) c->opt_gemm_no_unit = value;
    else if (!strcmp(name, "gemm_keep")) c->opt_gemm_keep = value;
    else if (!strcmp(name, "gemm_no_pair")) c->opt_gemm_no_pair = value;
    else if (!strcmp(name, "gemm_no_tf32")) c->opt_gemm_no_tf32 = value;
    else return fail(LVS_EINVAL, "unknown option '%s'", name);
    return LVS_OK;
}

__global__ void fetch_rows_kernel(const uint8_t* base, uint32_t row_bytes, int dim, int storage, const int64_t* rows, int64_t n, float* out) {
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint8_t* rp = base + (size_t)rows[i] * row_bytes;
        for (int c = threadIdx.x; c < dim; c += blockDim.x)
            out[(size_t)i * dim + c] = storage == LVS_STORAGE_F32 ? reinterpret_cast<const float*>(rp)[c]
                                                                  : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rp)[c]);
    }
}

extern "C" int lvs_fetch_rows_f32(lvs_collection* c, const int64_t* rows, int64_t n, float* out) {
    bind_thread();
    if (!c) return fail(LVS_EINVAL, "collection is NULL");
    if (n <= 0) return LVS_OK;
    if (!rows || !out) return fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    std::vector<int64_t> local(rows, rows + n);
    for (auto& r : local) { r -= c->row_base; if (r < 0 || r >= c->n_rows) return fail(LVS_EINVAL, "row out of range"); }
    int rc = ensure_dev(c->s_misc, (size_t)n * 8 + (size_t)n * c->dim * 4);
    if (rc != LVS_OK) return rc;
    cudaStream_t st = c->stream;
    int64_t* dr = (int64_t*)c->s_misc.p;
    float* dout = (float*)((uint8_t*)c->s_misc.p + (size_t)n * 8);
    CU(cudaMemcpyAsync(dr, local.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    fetch_rows_kernel<<<(unsigned)std::min<int64_t>(n, 4096), 128, 0, st>>>(c->d_vec, c->row_bytes, c->dim, c->storage, dr, n, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout, (size_t)n * c->dim * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return LVS_OK;
}
