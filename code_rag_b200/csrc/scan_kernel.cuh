// K1 - HBM-streaming exact scan for 1..4 queries (regime 1 of BASELINE.json north_star).
//
// Replaces the O(N*D) loop that the reference delegates to Qdrant behind
// QdrantManager.search (reference src/lattice/embeddings/client.py:132-157): every stored row is scored
// against the query and the best k' rows (k' >= k, see finalize_kernel.cuh) are kept.
//
// Shape of the kernel (one persistent CTA per SM, 9 warps):
//   * warp 8 = producer.  One elected lane streams row tiles global -> shared with cp.async.bulk (TMA,
//     SASS UBLKCP) into an S-stage ring guarded by full/empty mbarriers.  Rows are contiguous in a row-major
//     shard, so a tile of R rows is ONE linear bulk copy of R*row_bytes bytes - no tensor map is needed.
//     With a payload filter the 32 lanes evaluate the conjunction of code columns + the tombstone byte for 32
//     rows at a time and only the surviving rows are copied (one bulk copy per row), so a selective filter
//     reads only the rows it needs ("effective bytes").
//   * warps 0..7 = consumers.  A warp owns RW rows at a time; each lane reads 16-byte chunks of the rows
//     (LDS.128, conflict free) and of the fp32 queries held in shared memory, accumulates dot products (and
//     the row's sum of squares for cosine over bf16 storage - the fused norm) in fp32, then a transposed
//     butterfly leaves one (row, query) total per lane group.  The per-warp running top-k' lives in registers
//     (KPL keys per lane); a row is only looked at again if its key beats the warp's current k'-th key.
//   * epilogue: the 8 warp lists of a CTA are merged (shuffle networks + a three-level merge tree for 32-key lists) and
//     written to global memory.
//
// ONE KERNEL PER SEARCH (round 2).  The launch also does what used to be three more launches:
//   * query preparation: every CTA normalises the raw query itself (float64 norm, the arithmetic of prep_queries_kernel)
//     while its producer warp is already streaming tiles;
//   * candidate selection + exact rescoring + ordering (finalize_kernel.cuh): CTAs take a ticket when their list is
//     written; the last H finishers stay ("helpers"), wait until every list is there, and run the finalize body - the same
//     split of the candidates over H CTAs the standalone finalize kernel uses, without a launch or an idle machine;
//   * sharded collections: the CTA that holds a query's final local result stores it into every peer's gather buffer over
//     NVLink, raises its flag, waits for the peers and merges (exchange_kernel.cuh) - a sharded search is one kernel per GPU.
//   Launched with programmatic stream serialisation: `griddepcontrol.launch_dependents` at the top, `griddepcontrol.wait`
//   only before the first write to memory shared with the previous search (lists, tickets), so the next search's CTAs start
//   streaming as soon as SMs free up while this search's helpers are still rescoring / waiting for peers.
//
// Algorithmic bytes per launch = n_rows * row_bytes (one read of the shard, shared by the QT queries).
#pragma once
#include "common.cuh"
#include "exchange_kernel.cuh"
#include "finalize_kernel.cuh"

namespace lvs {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanThreads = (kScanConsumerWarps + 1) * kWarp;
constexpr int kScanRW = 4;            // rows a consumer warp processes together
constexpr int kScanMaxStages = 12;
constexpr size_t kScanStaticSmem = 1024;   // static shared memory of the kernel (a few words, padded to the ring's alignment)

// A handful of host queries ride in the kernel's parameter block (one copy at launch, readable by every CTA from its first
// instruction): no PCIe read, no staging hop, no flag.  8 KB = one 1024-dimensional float64 query.
constexpr uint32_t kInlineQueryBytes = 8192;
struct InlineQueries { uint8_t bytes[kInlineQueryBytes]; };
// The block is a parameter of the INLINE instantiations only (one query slot, no filter); every other instantiation keeps the
// classic sub-4 KB parameter block.
struct NoInlineQueries { uint8_t bytes[16]; };
template <bool INLINE> struct InlineBlock { using type = NoInlineQueries; };
template <> struct InlineBlock<true> { using type = InlineQueries; };

struct ScanParams {
    const uint8_t* base;        // shard, row-major, row stride = row_bytes
    uint32_t row_bytes;         // multiple of 16
    uint32_t chunks_per_row;    // row_bytes / 16
    uint32_t n_rows;            // rows to scan (high-water mark of the shard)
    uint32_t stage_rows;        // R: rows per ring stage, multiple of kScanRW
    uint32_t n_stages;          // S
    uint32_t stage_bytes;       // R * row_bytes
    uint32_t n_tiles;           // ceil(n_rows / R)                    (unfiltered)
    uint32_t n_blocks32;        // ceil(n_rows / 32)                   (filtered)
    const void* q_raw;          // [n_queries][dim] RAW queries (f32 / f64) as the caller passed them; normalised in the kernel
    int q_dtype;                // LVS_DT_F32 / LVS_DT_F64
    int dim;
    int metric;
    uint32_t n_queries;         // 1..QT (a 3-query group runs the 4-slot kernel; unused slots score a zero query)
    uint32_t q_stride;          // floats per query row = chunks_per_row * elems_per_chunk
    const uint8_t* live;        // [n_rows] 1 = live, 0 = tombstone
    const uint32_t* codes[kMaxFilterCols];
    uint32_t want[kMaxFilterCols];
    uint32_t n_filter;          // number of constrained columns compacted into codes[]/want[]
    uint64_t* out_keys;         // [QT][gridDim.x][32*KPL]
    uint64_t* out_tops;         // [QT][gridDim.x] best key of each CTA list (threshold for the finalize step)
    // ---- fused finalize (+ exchange) ----
    uint32_t seq;               // sequence number of this launch on its collection (1, 2, ...): orders what griddepcontrol alone
    uint32_t* done_seq;         //   does not - the last helper of launch s publishes *done_seq = s when everything is written
    uint32_t* tile_counter;     // dynamic tile scheduling (nullptr = static, tile t -> CTA t mod grid): producers draw tiles with
    uint32_t tile_base;         //   atomicAdd(tile_counter) - tile_base; a launch advances the counter by n_tiles + grid
    const void* q_host;         // the raw queries live in mapped pinned HOST memory: CTA 0 reads them over PCIe once and stages
    void* q_stage;              //   them here (device) for everybody, then publishes *q_flag = seq; q_raw == q_stage
    uint32_t* q_flag;
    uint32_t q_bytes;           // n_queries * dim * sizeof(query element)
    uint32_t q_inline;          // 1: the raw queries are in the InlineQueries parameter of an INLINE instantiation (q_raw / q_host unused)
    unsigned long long* dbg_times;  // optional [8]: globaltimer stamps of the launch's phases (min of the starts, max of the rest); profiling aid
    uint32_t* host_ready;       // optional word in mapped pinned host memory: set to host_ready_val (system-scope release) when every
    uint32_t host_ready_val;    //   result of this launch has been stored - the host polls it instead of synchronising an event
    uint32_t* ticket;           // [2] arrival counters of this launch's CTAs, zero between launches
    uint32_t n_helpers;         // H = n_queries * C: the last H CTAs to finish run the finalize body (C CTAs per query)
    uint32_t pdl;               // 1: launched with programmatic stream serialisation (griddepcontrol.* are executed)
    uint32_t exchange;          // 1: sharded collection - publish / wait / merge through `ex`
};

// Transposed warp reduction: `vals[0..V)` per lane, V a power of two <= 32.  On return vals[0] of lane l
// holds the warp-wide sum of value index  v = l >> (5 - log2 V).
template <int V>
__device__ __forceinline__ void warp_reduce_transposed(float (&vals)[V], int lane) {
    static_assert(V >= 1 && V <= 32 && (V & (V - 1)) == 0, "V must be a power of two");
    int off = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            float keep = upper ? vals[i + n / 2] : vals[i];
            float send = upper ? vals[i] : vals[i + n / 2];
            vals[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
        }
        off >>= 1;
    }
#pragma unroll
    for (; off >= 1; off >>= 1) vals[0] += __shfl_xor_sync(0xFFFFFFFFu, vals[0], off);
}

template <typename T> struct ChunkTraits;
template <> struct ChunkTraits<float> { static constexpr int kElems = 4; };
template <> struct ChunkTraits<__nv_bfloat16> { static constexpr int kElems = 8; };

template <typename T>
__device__ __forceinline__ void unpack_chunk(const uint4& v, float (&x)[ChunkTraits<T>::kElems]);
template <>
__device__ __forceinline__ void unpack_chunk<float>(const uint4& v, float (&x)[4]) {
    x[0] = __uint_as_float(v.x); x[1] = __uint_as_float(v.y); x[2] = __uint_as_float(v.z); x[3] = __uint_as_float(v.w);
}
template <>
__device__ __forceinline__ void unpack_chunk<__nv_bfloat16>(const uint4& v, float (&x)[8]) {
    x[0] = bf16lo(v.x); x[1] = bf16hi(v.x); x[2] = bf16lo(v.y); x[3] = bf16hi(v.y);
    x[4] = bf16lo(v.z); x[5] = bf16hi(v.z); x[6] = bf16lo(v.w); x[7] = bf16hi(v.w);
}

__host__ __device__ constexpr int next_pow2_ce(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// Shared-memory carve-up (host mirrors this in scan_smem_bytes()).
//   [stages: S*stage_bytes][queries: QT*q_stride*4][meta: S*(R+4)*4 (filtered) or S*16 (unfiltered: rows, first row)][full[S], empty[S]]
// fin_bytes: shared memory the fused finalize (+ exchange merge) needs; it aliases the drained ring
__host__ __device__ inline size_t scan_smem_bytes(uint32_t n_stages, uint32_t stage_bytes, uint32_t qt, uint32_t q_stride,
                                                  uint32_t stage_rows, bool filtered, size_t fin_bytes = 0) {
    size_t b = (size_t)n_stages * stage_bytes;
    b += (size_t)qt * q_stride * 4;
    b += filtered ? (size_t)n_stages * (stage_rows + 4) * 4 : (size_t)n_stages * 16;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)2 * n_stages * 8;
    // the epilogue's merge scratch (8 warps * 32*KPL <= 256 keys * 8 B = 16 KB) aliases the stage ring
    const size_t merge = (size_t)kScanConsumerWarps * 32 * 8 * 8;
    b = b > merge ? b : merge;
    return b > fin_bytes ? b : fin_bytes;
}

// ||q||_2 in float64 with the arithmetic of prep_queries_kernel (aux_kernels.cuh), so that the fused kernel, the standalone
// prep kernel and every CTA agree bit for bit.  Called by the 256 consumer threads (named barrier 1).
__device__ __forceinline__ double block_query_norm(const void* src, int dtype, size_t so, int dim, double* red, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    double ss = 0.0;
    for (int c = tid; c < dim; c += 256) { const double x = load_as_f64(src, dtype, so + c); ss = fma(x, x, ss); }
    ss = warp_sum_f64(ss);
    named_bar_sync(1, 256);                               // previous readers of red[] are done
    if (lane == 0) red[warp] = ss;
    named_bar_sync(1, 256);
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    return sqrt(t);
}

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one thread: wait until *word has reached `target` (sequence numbers wrap: compare the signed difference)
__device__ __forceinline__ void wait_seq_reached(const uint32_t* word, uint32_t target) {
    while ((int32_t)(ld_acquire_gpu_u32(word) - target) < 0) __nanosleep(64);
}

template <typename T, int QT, int KPL, bool NORM, bool FILTER, bool INLINE>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const __grid_constant__ ScanParams p, const __grid_constant__ FinalizeParams fp,
                                                                    const __grid_constant__ ExchangeParams xp,
                                                                    const __grid_constant__ typename InlineBlock<INLINE>::type iq) {
    constexpr int E = ChunkTraits<T>::kElems;
    constexpr int RW = kScanRW;
    constexpr int NVAL = RW * (QT + (NORM ? 1 : 0));
    constexpr int VP = next_pow2_ce(NVAL);
    constexpr int LV = (VP == 1) ? 0 : (VP == 2) ? 1 : (VP == 4) ? 2 : (VP == 8) ? 3 : (VP == 16) ? 4 : 5;
    constexpr int KPW = 32 * KPL;

    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ double s_red[8];
    __shared__ double s_div[QT];        // what each raw query is divided by (its norm; EPSILON for a zero query; 1 for dot)
    __shared__ float s_qnorm[QT];
    __shared__ uint32_t s_bcast[2];
    const uint32_t S = p.n_stages;
    uint8_t* stages = smem;
    float* qsm = reinterpret_cast<float*>(smem + (size_t)S * p.stage_bytes);
    size_t off = (size_t)S * p.stage_bytes + (size_t)QT * p.q_stride * 4;
    uint32_t* meta = reinterpret_cast<uint32_t*>(smem + off);          // per stage: [0]=count, filtered: [4..4+R)=rows; else [1]=first row
    const uint32_t meta_stride = FILTER ? p.stage_rows + 4 : 4;
    off += (size_t)S * meta_stride * 4;
    off = (off + 15) & ~(size_t)15;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + off);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* merge_buf = reinterpret_cast<uint64_t*>(smem);            // [8 warps][KPW], aliases the (drained) stage ring

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        stamp(p.dbg_times, 0);
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kScanConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // the next search on this stream may be scheduled as soon as SMs free up: nothing it does before ITS griddepcontrol.wait
    // (streaming the shard, scoring) depends on this search
    if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == kScanConsumerWarps) {
        // =============================== producer ===============================
        if (!FILTER) {
            if (lane == 0) {
                // tiles are drawn from a global counter (late starters - the SMs that just finished the previous search's
                // rescoring - take fewer), or assigned round-robin when the grid does not cover the machine
                const bool dyn = p.tile_counter != nullptr;
                if (dyn) wait_seq_reached(p.done_seq, p.seq - 2u);      // the launch that last used this counter has drawn its last tile
                uint32_t tile = dyn ? atomicAdd(p.tile_counter, 1u) - p.tile_base : blockIdx.x;
                uint32_t it = 0;
                for (;; ++it) {
                    const uint32_t s = it % S;
                    const uint32_t ph = (it / S) & 1u;
                    // the next tile's number is requested before this stage is waited for, so the atomic's latency is hidden
                    uint32_t next = 0;
                    if (tile < p.n_tiles) next = dyn ? atomicAdd(p.tile_counter, 1u) - p.tile_base : tile + gridDim.x;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    if (tile >= p.n_tiles) {                            // terminator stage: count == 0
                        meta[s * meta_stride] = 0;
                        mbar_arrive(&full_bar[s]);
                        break;
                    }
                    const uint32_t row0 = tile * p.stage_rows;
                    const uint32_t rows = min(p.stage_rows, p.n_rows - row0);
                    const uint32_t bytes = rows * p.row_bytes;
                    meta[s * meta_stride] = rows; meta[s * meta_stride + 1] = row0;
                    mbar_arrive_expect_tx(&full_bar[s], bytes);
                    bulk_g2s(stages + (size_t)s * p.stage_bytes, p.base + (size_t)row0 * p.row_bytes, bytes, &full_bar[s]);
                    tile = next;
                }
            }
        } else {
            uint32_t it = 0, slot = 0, s = 0;
            constexpr uint32_t UB = 4;                            // 32-row blocks whose flag loads are in flight together
            for (uint32_t blk0 = blockIdx.x * UB; blk0 < p.n_blocks32; blk0 += gridDim.x * UB) {
                bool pass[UB];
#pragma unroll
                for (uint32_t u = 0; u < UB; ++u) {
                    const uint32_t row = (blk0 + u) * 32u + lane;
                    pass[u] = row < p.n_rows;
                    uint8_t lv = 0;
                    if (pass[u]) lv = p.live[row];
                    uint32_t cv[kMaxFilterCols];
#pragma unroll
                    for (uint32_t f = 0; f < (uint32_t)kMaxFilterCols; ++f)
                        cv[f] = (f < p.n_filter && pass[u]) ? __ldg(p.codes[f] + row) : 0u;
                    pass[u] = pass[u] && lv != 0;
#pragma unroll
                    for (uint32_t f = 0; f < (uint32_t)kMaxFilterCols; ++f)
                        if (f < p.n_filter) pass[u] = pass[u] && (cv[f] == p.want[f]);
                }
#pragma unroll
                for (uint32_t u = 0; u < UB; ++u) {
                    const uint32_t blk = blk0 + u;
                    uint32_t m = __ballot_sync(0xFFFFFFFFu, pass[u]);
                    while (m) {
                        const uint32_t b = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t r = blk * 32u + b;
                        if (slot == 0) {
                            s = it % S;
                            if (lane == 0) mbar_wait(&empty_bar[s], ((it / S) & 1u) ^ 1u);
                            __syncwarp();
                        }
                        if (lane == 0) {
                            mbar_expect_tx(&full_bar[s], p.row_bytes);
                            bulk_g2s(stages + (size_t)s * p.stage_bytes + (size_t)slot * p.row_bytes,
                                     p.base + (size_t)r * p.row_bytes, p.row_bytes, &full_bar[s]);
                            meta[s * meta_stride + 4 + slot] = r;
                        }
                        ++slot;
                        if (slot == p.stage_rows) {
                            if (lane == 0) { meta[s * meta_stride] = slot; mbar_arrive(&full_bar[s]); }
                            slot = 0; ++it;
                        }
                    }
                }
            }
            if (slot > 0) {
                if (lane == 0) { meta[s * meta_stride] = slot; mbar_arrive(&full_bar[s]); }
                slot = 0; ++it;
            }
            // terminator stage: count == 0
            s = it % S;
            if (lane == 0) {
                mbar_wait(&empty_bar[s], ((it / S) & 1u) ^ 1u);
                meta[s * meta_stride] = 0;
                mbar_arrive(&full_bar[s]);
            }
        }
        return;
    }

    // ================================= consumers =================================
    // raw queries -> unit fp32 queries in shared memory, while the producer is already streaming.  For 8-element chunks the two
    // float4 halves of a chunk go to separate planes ([h][chunk] float4) so that a warp's LDS.128 over consecutive chunks is
    // bank-conflict free.
    const void* const q_raw = INLINE ? static_cast<const void*>(iq.bytes) : p.q_raw;
    if (p.q_host != nullptr) {
        // queries in mapped pinned host memory (the host-buffer API): ONE CTA pulls them over PCIe and stages them in HBM,
        // the others wait for its flag and read the staged copy - 6 KB over the bus instead of 148 x 6 KB
        if (blockIdx.x == 0) {
            if (tid == 0) wait_seq_reached(p.done_seq, p.seq - 2u);     // the launch that last read this staging slot is done
            named_bar_sync(1, 256);
            // every load of a thread is in flight before its first store: one PCIe round trip for the block, not one per word
            if ((reinterpret_cast<uintptr_t>(p.q_host) & 15u) == 0) {
                const uint4* src = reinterpret_cast<const uint4*>(p.q_host);      // reads up to 15 bytes past the block: slots are padded
                uint4* dst = reinterpret_cast<uint4*>(p.q_stage);
                const uint32_t n16 = (p.q_bytes + 15u) / 16u;
                for (uint32_t i0 = tid; i0 < n16; i0 += 256 * 4) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i0 + u * 256 < n16) v[u] = src[i0 + u * 256];
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i0 + u * 256 < n16) dst[i0 + u * 256] = v[u];
                }
            } else {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(p.q_host);
                uint32_t* dst = reinterpret_cast<uint32_t*>(p.q_stage);
                const uint32_t n4 = p.q_bytes / 4u;
                for (uint32_t i0 = tid; i0 < n4; i0 += 256 * 8) {
                    uint32_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (i0 + u * 256 < n4) v[u] = src[i0 + u * 256];
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (i0 + u * 256 < n4) dst[i0 + u * 256] = v[u];
                }
            }
            __threadfence();
            named_bar_sync(1, 256);
            if (tid == 0) st_release_gpu_u32(p.q_flag, p.seq);
        } else {
            if (tid == 0) wait_seq_reached(p.q_flag, p.seq);
            named_bar_sync(1, 256);
        }
    }
#pragma unroll 1
    for (uint32_t qi = 0; qi < (uint32_t)QT; ++qi) {
        double div = 1.0;
        if (qi < p.n_queries) {
            const double nrm = block_query_norm(q_raw, p.q_dtype, (size_t)qi * p.dim, p.dim, s_red, tid);
            div = (p.metric == LVS_METRIC_COSINE) ? (nrm != 0.0 ? nrm : 1.1920929e-7) : 1.0;
            if (tid == 0) { s_div[qi] = div; s_qnorm[qi] = (float)nrm; }
        }
        for (uint32_t e = tid; e < p.q_stride; e += 256) {
            float v = 0.f;
            if (qi < p.n_queries && (int)e < p.dim) v = (float)(load_as_f64(q_raw, p.q_dtype, (size_t)qi * p.dim + e) / div);
            const uint32_t c = e / E, w = e % E;
            qsm[(size_t)qi * p.q_stride + (size_t)(w / 4) * (p.chunks_per_row * 4) + c * 4 + (w % 4)] = v;
        }
    }
    named_bar_sync(1, 256);
    if (tid == 0) stamp(p.dbg_times, 1);                  // queries ready

    WarpTopK<KPL> top[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) top[q].init();

    const uint32_t cpr = p.chunks_per_row;
    const int my_val = lane >> (5 - LV);                  // value index this lane holds after the reduction
    const bool holder = (lane & ((1 << (5 - LV)) - 1)) == 0;
    const int my_type = my_val / RW;                      // 0..QT-1 = query, QT = sum of squares
    const int my_j = my_val % RW;                         // row within the group

    uint32_t it = 0;
    uint32_t group_base = 0;                              // running group counter (rotates warps over stages)
    for (;;) {
        uint32_t cnt, row0 = 0;
        const uint32_t s = it % S;
        mbar_wait(&full_bar[s], (it / S) & 1u);
        cnt = meta[s * meta_stride];
        if (cnt == 0) break;
        if (!FILTER) row0 = meta[s * meta_stride + 1];
        const uint4* st = reinterpret_cast<const uint4*>(stages + (size_t)s * p.stage_bytes);
        const uint32_t groups = (cnt + RW - 1) / RW;
        // first group of this stage owned by this warp: (group_base + g) % NW == warp
        uint32_t g = (warp + kScanConsumerWarps - (group_base % kScanConsumerWarps)) % kScanConsumerWarps;
        for (; g < groups; g += kScanConsumerWarps) {
            float vals[VP];
#pragma unroll
            for (int i = 0; i < VP; ++i) vals[i] = 0.f;
            const uint4* rowp = st + (size_t)g * RW * cpr;
            for (uint32_t c = lane; c < cpr; c += 32) {
                float q[QT][E];
#pragma unroll
                for (int qi = 0; qi < QT; ++qi) {
                    const float4* qp = reinterpret_cast<const float4*>(qsm + (size_t)qi * p.q_stride) + c;
#pragma unroll
                    for (int h = 0; h < E / 4; ++h) {
                        float4 t = qp[(size_t)h * cpr];
                        q[qi][4 * h + 0] = t.x; q[qi][4 * h + 1] = t.y; q[qi][4 * h + 2] = t.z; q[qi][4 * h + 3] = t.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < RW; ++j) {
                    const uint4 v = rowp[(size_t)j * cpr + c];
                    float x[E];
                    unpack_chunk<T>(v, x);
#pragma unroll
                    for (int e = 0; e < E; ++e) {
#pragma unroll
                        for (int qi = 0; qi < QT; ++qi) vals[qi * RW + j] = fmaf(x[e], q[qi][e], vals[qi * RW + j]);
                        if (NORM) vals[QT * RW + j] = fmaf(x[e], x[e], vals[QT * RW + j]);
                    }
                }
            }
            warp_reduce_transposed<VP>(vals, lane);
            float tot = vals[0];
            if (NORM) {
                // the sum of squares of row my_j sits on the holder lane of value (QT*RW + my_j)
                const float ss = __shfl_sync(0xFFFFFFFFu, tot, (QT * RW + my_j) << (5 - LV));
                tot = ss > 0.f ? tot * rsqrtf(ss) : 0.f;
            }
            const uint32_t slot = g * RW + my_j;
            uint32_t row;
            if (!FILTER) row = row0 + slot;
            else row = meta[s * meta_stride + 4 + min(slot, p.stage_rows - 1)];
            const bool valid = holder && my_type < QT && slot < cnt;
            const uint64_t key = valid ? make_key(tot, row) : 0ull;
            bool pred = false;
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) pred |= (my_type == qi) && (key > top[qi].thr);
            uint32_t ball = __ballot_sync(0xFFFFFFFFu, pred);
            while (ball) {
                const int src = __ffs(ball) - 1;
                ball &= ball - 1;
                const uint64_t k = shfl_u64(key, src);
                const int t = __shfl_sync(0xFFFFFFFFu, my_type, src);
                bool alive = true;
                if (!FILTER) alive = p.live[key_row(k)] != 0;   // tombstones are checked lazily (rare path)
                if (alive) {
#pragma unroll
                    for (int qi = 0; qi < QT; ++qi) {
                        if (t == qi && k > top[qi].thr) top[qi].insert(k, lane);
                    }
                }
            }
        }
        group_base += groups;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        ++it;
    }

    if (tid == 0) stamp(p.dbg_times, 2);                  // this CTA's share of the shard is scanned
    // ============== CTA epilogue: merge the 8 warp lists of each query, write one list per CTA ==============
    // From here on the kernel writes memory that the previous search on this stream also used (lists, tickets, candidate
    // scores): wait until that search has completed (it has, long ago, unless this shard is tiny).
    if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr int NCT = kScanConsumerWarps * kWarp;
    if (tid == 0) wait_seq_reached(p.done_seq, p.seq - 1u);   // explicit, whatever the launch chain looks like (already true in practice)
    named_bar_sync(1, NCT);
    if constexpr (KPL == 1) {
        // 32-key lists: sort each warp's list on shuffles, then a three-level merge tree (top 32 of two sorted lists =
        // elementwise max of one with the other reversed, a bitonic sequence) - ~1 us instead of 36 barrier-separated steps
#pragma unroll 1
        for (int qi = 0; qi < QT; ++qi) {
            uint64_t x = 0;
#pragma unroll
            for (int q2 = 0; q2 < QT; ++q2) if (q2 == qi) x = top[q2].key[0];
            x = warp_sort_desc(x, lane);
#pragma unroll 1
            for (int half = kScanConsumerWarps / 2; half >= 1; half >>= 1) {
                named_bar_sync(1, NCT);                   // the previous level's readers are done
                if (warp >= half && warp < 2 * half) merge_buf[(size_t)warp * 32 + lane] = x;
                named_bar_sync(1, NCT);
                if (warp < half) {
                    const uint64_t o = merge_buf[(size_t)(warp + half) * 32 + (31 - lane)];
                    x = warp_bitonic_merge_desc(o > x ? o : x, lane);
                }
            }
            if (warp == 0) {
                p.out_keys[((size_t)qi * gridDim.x + blockIdx.x) * KPW + lane] = x;
                if (lane == 0) p.out_tops[(size_t)qi * gridDim.x + blockIdx.x] = x;
            }
        }
    } else {
        // The 8 * KPW keys of a query are sorted in shared memory by a bitonic network run by all 256 consumer threads
        // (named barrier between stages); the first KPW keys are the CTA's list.
        constexpr int NM = kScanConsumerWarps * KPW;      // keys to merge per query (a power of two)
#pragma unroll 1
        for (int qi = 0; qi < QT; ++qi) {
            named_bar_sync(1, NCT);                       // previous round's readers are done with merge_buf
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                uint64_t kv = 0;
#pragma unroll
                for (int q2 = 0; q2 < QT; ++q2) if (q2 == qi) kv = top[q2].key[j];
                merge_buf[(size_t)warp * KPW + j * 32 + lane] = kv;
            }
            named_bar_sync(1, NCT);
            for (int k = 2; k <= NM; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < NM; i += NCT) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const uint64_t a = merge_buf[i], b2 = merge_buf[ixj];
                            const bool desc = (i & k) == 0;
                            if (desc ? (a < b2) : (a > b2)) { merge_buf[i] = b2; merge_buf[ixj] = a; }
                        }
                    }
                    named_bar_sync(1, NCT);
                }
            }
            uint64_t* dst = p.out_keys + ((size_t)qi * gridDim.x + blockIdx.x) * KPW;
            for (int i = tid; i < KPW; i += NCT) dst[i] = merge_buf[i];
            if (tid == 0) p.out_tops[(size_t)qi * gridDim.x + blockIdx.x] = merge_buf[0];
        }
    }

    // ============== ticket: the last H CTAs to finish stay and finalize ==============
    __threadfence();
    named_bar_sync(1, NCT);
    if (tid == 0) { s_bcast[0] = atomicAdd(p.ticket, 1u); stamp(p.dbg_times, 3); }      // list written, ticket taken
    named_bar_sync(1, NCT);
    const uint32_t G = gridDim.x;
    const uint32_t n_items = p.n_helpers;                 // (query, CTA-of-the-query) work items: n_queries * C
    const uint32_t H = min(n_items, G);
    const uint32_t my_ticket = s_bcast[0];
    if (my_ticket < G - H) return;
    const uint32_t h = my_ticket - (G - H);               // 0..H-1 in finishing order; helper H-1 finished last
    if (h + 1 < H) {
        if (tid == 0) { while (ld_acquire_gpu_u32(p.ticket) < G) __nanosleep(128); }
        named_bar_sync(1, NCT);
    }
    __threadfence();
    if (tid == 0) stamp(p.dbg_times, 4);                  // helper: every list is there

    const uint32_t nq = p.n_queries;
    const uint32_t C = n_items / nq;
    uint32_t owned = 0;                                   // queries whose final local result this CTA produced
    const bool early_ready = p.host_ready != nullptr && nq == 1 && !p.exchange;
#pragma unroll 1
    for (uint32_t w = h; w < n_items; w += H) {
        const uint32_t qi = w % nq, cy = w / nq;
        named_bar_sync(1, NCT);                           // the previous item's shared memory is no longer read
        const auto stage_q = [&](double* qdst) {
            for (int e = tid; e < p.dim; e += NCT) qdst[e] = load_as_f64(q_raw, p.q_dtype, (size_t)qi * p.dim + e) / s_div[qi];
        };
        FinResult r;
        const bool mine = finalize_body<KPL>(fp, qi, cy, C, smem, s_qnorm[qi], r, stage_q);
        if (!mine) continue;
        owned |= 1u << qi;
        if (!p.exchange) {
            finalize_store_local(fp, qi, r, tid);
            if (early_ready) {
                // the only result of this launch is stored: tell the polling host now, the housekeeping below is device business.
                // One system-scope release by one thread: the barrier orders every thread's stores before it, the fence is cumulative
                named_bar_sync(1, NCT);
                if (tid == 0) {
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.host_ready), "r"(p.host_ready_val) : "memory");
                    stamp(p.dbg_times, 13);
                }
            }
            continue;
        }
        // sharded collection: this query's local top-k (+ its flag) -> every rank's gather buffer
        const size_t qk = (size_t)nq * fp.k;
        for (uint32_t c = tid; c < r.ncand + fp.k; c += NCT) {
            uint32_t slot; int64_t v0, v1, v2;
            if (c < r.ncand) {
                slot = r.rank[c];
                if (slot >= fp.k) continue;
                v0 = __double_as_longlong(r.score[c]); v1 = fp.row_base + (int64_t)r.row[c]; v2 = (int64_t)r.tie[c];
            } else {
                slot = c - r.ncand;
                if (slot < r.nout) continue;
                v0 = 0; v1 = -1; v2 = 0;
            }
            const size_t o = (size_t)xp.rank * xp.blk_stride + (size_t)qi * fp.k + slot;
            for (int g = 0; g < xp.world; ++g) { int64_t* d = xp.peer_data[g]; d[o] = v0; d[o + qk] = v1; d[o + 2 * qk] = v2; }
        }
        if (tid == 0) {
            const size_t o = (size_t)xp.rank * xp.blk_stride + 3 * qk + qi;
            for (int g = 0; g < xp.world; ++g) xp.peer_data[g][o] = (int64_t)r.flag;
        }
        __threadfence_system();
        named_bar_sync(1, NCT);
        if (tid == 0 && atomicAdd(xp.done_counter, 1u) == nq - 1) { *xp.done_counter = 0; exchange_signal(xp); }
    }
    if (p.exchange && owned) {
        if (tid == 0) s_bcast[1] = 0;
        named_bar_sync(1, NCT);
        if (tid < xp.world && !exchange_wait_one(xp, tid)) s_bcast[1] = 1;
        named_bar_sync(1, NCT);
        const bool timed_out = s_bcast[1] != 0;
        uint8_t* xsm = smem + finalize_smem_bytes(fp.dim_pad, fp.n_rescore_warps);
        for (uint32_t qi = 0; qi < nq; ++qi)
            if (owned & (1u << qi)) exchange_merge_query(xp, (int)qi, xsm, &s_bcast[0], tid, NCT, timed_out, [] { named_bar_sync(1, NCT); });
    }
    // the helper that leaves last re-arms the tickets for the next launch, publishes the launch's completion on the device and,
    // for host-polled searches, in host memory (every result store above is fenced at system scope before its CTA checks out)
    named_bar_sync(1, NCT);                               // every thread's result stores are ordered before thread 0's fence below
    if (tid == 0) {
        if (owned) stamp(p.dbg_times, 7);                 // results stored (and merged)
        if (p.host_ready != nullptr && !early_ready) __threadfence_system();   // cumulative over this CTA's stores to host memory
        if (atomicAdd(p.ticket + 1, 1u) == H - 1) {
            p.ticket[0] = 0; p.ticket[1] = 0;
            __threadfence();
            st_release_gpu_u32(p.done_seq, p.seq);
            if (p.host_ready != nullptr && !early_ready)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.host_ready), "r"(p.host_ready_val) : "memory");
        }
    }
}

}  // namespace lvs
