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
//   * epilogue: the 8 warp lists of a CTA are merged by one warp per query and written to global memory.
//
// Algorithmic bytes per launch = n_rows * row_bytes (one read of the shard, shared by the QT queries).
#pragma once
#include "common.cuh"

namespace lvs {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanThreads = (kScanConsumerWarps + 1) * kWarp;
constexpr int kScanRW = 4;            // rows a consumer warp processes together
constexpr int kScanMaxStages = 12;

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
    const float* queries;       // [QT][q_stride] fp32, unit norm for cosine; unused slots are zero
    uint32_t q_stride;          // floats per query row = chunks_per_row * elems_per_chunk
    const uint8_t* live;        // [n_rows] 1 = live, 0 = tombstone
    const uint32_t* codes[kMaxFilterCols];
    uint32_t want[kMaxFilterCols];
    uint32_t n_filter;          // number of constrained columns compacted into codes[]/want[]
    uint64_t* out_keys;         // [QT][gridDim.x][32*KPL]
    uint64_t* out_tops;         // [QT][gridDim.x] best key of each CTA list (threshold for the finalize step)
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
//   [stages: S*stage_bytes][queries: QT*q_stride*4][meta: S*(R+4)*4 (filtered only)][full[S], empty[S]]
__host__ __device__ inline size_t scan_smem_bytes(uint32_t n_stages, uint32_t stage_bytes, uint32_t qt, uint32_t q_stride,
                                                  uint32_t stage_rows, bool filtered) {
    size_t b = (size_t)n_stages * stage_bytes;
    b += (size_t)qt * q_stride * 4;
    if (filtered) b += (size_t)n_stages * (stage_rows + 4) * 4;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)2 * n_stages * 8;
    // the epilogue's merge scratch (8 warps * 32*KPL <= 256 keys * 8 B = 16 KB) aliases the stage ring
    const size_t merge = (size_t)kScanConsumerWarps * 32 * 8 * 8;
    return b > merge ? b : merge;
}

template <typename T, int QT, int KPL, bool NORM, bool FILTER>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int E = ChunkTraits<T>::kElems;
    constexpr int RW = kScanRW;
    constexpr int NVAL = RW * (QT + (NORM ? 1 : 0));
    constexpr int VP = next_pow2_ce(NVAL);
    constexpr int LV = (VP == 1) ? 0 : (VP == 2) ? 1 : (VP == 4) ? 2 : (VP == 8) ? 3 : (VP == 16) ? 4 : 5;
    constexpr int KPW = 32 * KPL;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t S = p.n_stages;
    uint8_t* stages = smem;
    float* qsm = reinterpret_cast<float*>(smem + (size_t)S * p.stage_bytes);
    size_t off = (size_t)S * p.stage_bytes + (size_t)QT * p.q_stride * 4;
    uint32_t* meta = reinterpret_cast<uint32_t*>(smem + off);          // per stage: [0]=count, [4..4+R)=rows
    const uint32_t meta_stride = p.stage_rows + 4;
    if (FILTER) off += (size_t)S * meta_stride * 4;
    off = (off + 15) & ~(size_t)15;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + off);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* merge_buf = reinterpret_cast<uint64_t*>(smem);            // [8 warps][KPW], aliases the (drained) stage ring

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kScanConsumerWarps);
        }
        mbar_fence_init();
    }
    // queries -> shared (fp32).  For 8-element chunks the two float4 halves of a chunk go to separate planes
    // ([h][chunk] float4) so that a warp's LDS.128 over consecutive chunks is bank-conflict free.
    for (uint32_t i = tid; i < QT * p.q_stride; i += kScanThreads) {
        const uint32_t qi = i / p.q_stride, e = i % p.q_stride;
        const uint32_t c = e / E, w = e % E;
        qsm[(size_t)qi * p.q_stride + (size_t)(w / 4) * (p.chunks_per_row * 4) + c * 4 + (w % 4)] = p.queries[i];
    }
    __syncthreads();

    if (warp == kScanConsumerWarps) {
        // =============================== producer ===============================
        if (!FILTER) {
            if (lane == 0) {
                uint32_t it = 0;
                for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                    const uint32_t s = it % S;
                    const uint32_t ph = (it / S) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const uint32_t row0 = tile * p.stage_rows;
                    const uint32_t rows = min(p.stage_rows, p.n_rows - row0);
                    const uint32_t bytes = rows * p.row_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], bytes);
                    bulk_g2s(stages + (size_t)s * p.stage_bytes, p.base + (size_t)row0 * p.row_bytes, bytes, &full_bar[s]);
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
    WarpTopK<KPL> top[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) top[q].init();

    const uint32_t cpr = p.chunks_per_row;
    const int my_val = lane >> (5 - LV);                  // value index this lane holds after the reduction
    const bool holder = (lane & ((1 << (5 - LV)) - 1)) == 0;
    const int my_type = my_val / RW;                      // 0..QT-1 = query, QT = sum of squares
    const int my_j = my_val % RW;                         // row within the group

    uint32_t it = 0;
    uint32_t tile = blockIdx.x;
    uint32_t group_base = 0;                              // running group counter (rotates warps over stages)
    for (;;) {
        uint32_t cnt, row0 = 0;
        const uint32_t s = it % S;
        if (!FILTER) {
            if (tile >= p.n_tiles) break;
            row0 = tile * p.stage_rows;
            cnt = min(p.stage_rows, p.n_rows - row0);
            mbar_wait(&full_bar[s], (it / S) & 1u);
        } else {
            mbar_wait(&full_bar[s], (it / S) & 1u);
            cnt = meta[s * meta_stride];
            if (cnt == 0) break;
        }
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
        tile += gridDim.x;
    }

    // ============== CTA epilogue: merge the 8 warp lists of each query, write one list per CTA ==============
    // The 8 * KPW keys of a query are sorted in shared memory by a bitonic network run by all 256 consumer threads
    // (named barrier between stages); the first KPW keys are the CTA's list.  This replaced a serial fold by one warp
    // (dozens of dependent warp-level inserts), which cost ~9 us at the tail of every CTA - visible on small shards.
    constexpr int NCT = kScanConsumerWarps * kWarp;
    constexpr int NM = kScanConsumerWarps * KPW;          // keys to merge per query (a power of two)
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

}  // namespace lvs
