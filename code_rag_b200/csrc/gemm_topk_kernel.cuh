// K2 - tensor-core path for batched queries (regime 2 of BASELINE.json north_star): tcgen05.mma with the accumulator
// in TMEM, corpus tiles fed by TMA, top-k selection fused into the epilogue.
//
// Replaces, for Q >= 8 queries at a time, the O(Q*N*D) loop that the reference delegates to Qdrant behind
// QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) - there one gRPC call per query.
//
// Orientation: D[128 queries x 64 corpus rows] += A[128 x K] * B[64 x K]^T, bf16 inputs, fp32 accumulate.
//   * A (the queries of this CTA, unit-norm, rounded to bf16) is loaded ONCE into tensor memory: 128 lanes (one per
//     query) x K/2 32-bit columns (two bf16 per column) - 384 of the 512 TMEM columns for K = 768.  The MMA reads A
//     from TMEM ("TS" form), so shared memory is left to the corpus pipeline.
//   * B = 64 corpus rows x 64 K-elements per stage (8 KB, K-major, 128-byte swizzle), streamed with
//     cp.async.bulk.tensor.2d (TMA, SASS UTMALDG) into a 14-stage mbarrier ring: ~112 KB in flight per SM.
//   * D is double buffered in the remaining 2 x 64 TMEM columns: the MMA of tile t+1 overlaps the epilogue of tile t.
//   * Epilogue (8 warps, thread = query = TMEM lane, 32 columns each): tcgen05.ld the scores of the tile, release the accumulator,
//     scale by 1/||row|| (NaN for tombstones and rows past the end, so they never pass), compare with the query's
//     current 32nd-best score.  Passing (query, row, score) items go to a per-warp shared-memory queue by ballot; the
//     queue is drained WARP-COOPERATIVELY into sorted 32-key lists in shared memory (one list per query: lane j holds
//     key j, the insertion position is a ballot/popc, the shift is one shuffle), so the rare inserts cost ~15
//     full-warp instructions instead of a 32-step single-lane chain.
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocation,
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4; two warps per quarter split the tile's 64 columns).
// More than 128 queries: G = ceil(Q/128) CTAs ("a pair") walk the same tile sequence for different query groups; the
// second reader of a tile hits the 126 MB L2, so HBM still sees every row once.
//
// Exactness: lists hold 32 keys per (CTA, query); the finalize kernel proves the result against the bound
// max(k'-th kept fast score, largest dropped score) + eps_q and the host repeats flagged queries on the K1 path.
// Algorithmic bytes per launch = rows x row_bytes; FLOPs = 2 * Q * rows * K.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lvs {

constexpr int kGemmEpiWarps = 8;       // two per TMEM lane quarter: each takes 32 of the tile's 64 columns
constexpr int kGemmThreads = (2 + kGemmEpiWarps) * 32;
constexpr int kGemmM = 128;            // queries per CTA (TMEM lanes)
constexpr int kGemmN = 64;             // corpus rows per tile (accumulator columns per buffer)
constexpr int kGemmKC = 64;            // K elements per pipeline stage (= one 128-byte swizzle row)
constexpr int kGemmStageBytes = kGemmN * kGemmKC * 2;   // 8 KB
constexpr int kGemmMaxStages = 14;
constexpr int kGemmList = 32;          // keys kept per (CTA, query)
constexpr int kGemmMaxKChunks = 12;    // A occupies 32 columns per chunk: 12 * 32 + 2 * 64 (D) = 512 TMEM columns
constexpr uint32_t kGemmDCol = 384;    // first accumulator column
constexpr int kGemmQueue = 256;        // pending items per epilogue warp

struct GemmParams {
    const __nv_bfloat16* qb16;   // [n_groups * 128][k_pad] unit queries rounded to bf16, zero padded
    uint32_t k_pad;              // n_kchunks * 64
    uint32_t n_kchunks;
    uint32_t n_rows;
    uint32_t n_tiles;            // ceil(n_rows / 64)
    uint32_t n_groups;           // G: CTAs per pair
    uint32_t n_pairs;            // P: lists per query
    uint32_t n_stages;
    const float* inv_norm;       // [rows] 1/||row||, or nullptr (dot metric)
    const uint8_t* live;
    const uint8_t* base;         // shard base (for the linear L2 prefetch)
    uint32_t row_bytes;
    uint32_t prefetch_tiles;     // how many tiles ahead the L2 prefetch runs (0 = off)
    uint64_t* out_keys;          // [n_groups * 128][2P][32]   (two lists per CTA and query: one per column half)
    uint64_t* out_tops;          // [n_groups * 128][2P]  best key of the list
    uint64_t* out_drops;         // [n_groups * 128][2P]  32nd key when the list is full (bound on what was dropped), else 0
    float* dbg;                  // optional: CTA 0 dumps its first accumulator tile [128 x 64] (already scaled)
    uint32_t dbg_mode;           // profiling aid: bit0 skip the epilogue's scoring, bit1 skip the MMA issue, bit2 skip tcgen05.ld
};

// shared memory: [stages][lists 8*32*32*8][queue keys 8*256*8][queue lanes 8*256][inv 2*64*4][thr 256*4][barriers]
__host__ __device__ inline size_t gemm_smem_bytes(uint32_t n_stages) {
    return 1024 /* alignment slack */ + (size_t)n_stages * kGemmStageBytes + (size_t)kGemmEpiWarps * 32 * kGemmList * 8 +
           kGemmEpiWarps * kGemmQueue * 8 + kGemmEpiWarps * kGemmQueue + 2 * kGemmN * 4 + kGemmEpiWarps * 32 * 4 +
           (2 * kGemmMaxStages + 8) * 8 + 16;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// contiguous L2 prefetch: the tile's rows are adjacent in HBM, so ONE linear request keeps DRAM pages open; the 12
// swizzled 128-byte-wide tensor loads of the tile then hit L2 instead of touching each DRAM page twelve times
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem_d] (+)= A[tmem_a] * B[smem desc]^T, bf16 x bf16 -> fp32, M = 128, N = 128, K = 16
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: 8-row groups are 1024 B apart (SBO), version 1 (sm_100), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Warp-cooperative drain of one epilogue warp's queue into its sorted per-query lists (lane j holds key j of a list).
// Kept out of line: it is called from every column of the unrolled score loop and must not be replicated 33 times.
__device__ __noinline__ float gemm_drain_queue(const uint64_t* q_key, const uint8_t* q_lane, uint64_t* warp_lists, float* my_thr,
                                               uint32_t qcnt, int lane) {
    __syncwarp();
    for (uint32_t i = 0; i < qcnt; ++i) {
        const uint64_t key = q_key[i];
        const uint32_t ql = q_lane[i];
        uint64_t* L = warp_lists + ql * kGemmList;
        const uint64_t cur = L[lane];
        const uint32_t pos = __popc(__ballot_sync(0xFFFFFFFFu, cur > key));
        if (pos < (uint32_t)kGemmList) {
            const uint64_t up = shfl_up_u64(cur, 1);
            const uint64_t nv = (uint32_t)lane < pos ? cur : ((uint32_t)lane == pos ? key : up);
            L[lane] = nv;
            if (lane == kGemmList - 1) my_thr[ql] = nv != 0ull ? key_score(nv) : -INFINITY;
        }
        __syncwarp();
    }
    return my_thr[lane];
}

__global__ void __launch_bounds__(kGemmThreads, 1) gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    // 1024-byte alignment for the 128-byte swizzle; pointer arithmetic (not an integer round trip) keeps the accesses LDS/STS
    uint8_t* gsm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
    const uint32_t S = p.n_stages;
    uint8_t* stages = gsm;                                                   // S x 8 KB, 1024-byte aligned
    uint64_t* lists = reinterpret_cast<uint64_t*>(gsm + (size_t)S * kGemmStageBytes);     // [8 warps][32 queries][32] sorted desc
    uint64_t* wq_key = lists + kGemmEpiWarps * 32 * kGemmList;               // [8][kGemmQueue]
    uint8_t* wq_lane = reinterpret_cast<uint8_t*>(wq_key + kGemmEpiWarps * kGemmQueue);   // [8][kGemmQueue]
    float* inv_sm = reinterpret_cast<float*>(wq_lane + kGemmEpiWarps * kGemmQueue);       // [2][64]
    float* thr_sm = inv_sm + 2 * kGemmN;                                     // [8][32]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(thr_sm + kGemmEpiWarps * 32);
    uint64_t* empty_bar = full_bar + kGemmMaxStages;
    uint64_t* tmem_full = empty_bar + kGemmMaxStages;                        // [2]
    uint64_t* tmem_empty = tmem_full + 2;                                    // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t group = blockIdx.x % p.n_groups;
    const uint32_t pair = blockIdx.x / p.n_groups;
    const uint32_t nk = p.n_kchunks;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kGemmEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < kGemmEpiWarps * 32 * kGemmList; i += kGemmThreads) lists[i] = 0ull;
    for (int i = tid; i < kGemmEpiWarps * 32; i += kGemmThreads) thr_sm[i] = -INFINITY;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // ---- queries of this CTA -> TMEM (A operand): thread = query (lane of its quarter), 8 columns per store ----
    if (warp >= 2 && warp < 6) {
        const uint32_t lq = warp & 3;
        const uint32_t qrow = group * kGemmM + lq * 32 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(p.qb16 + (size_t)qrow * p.k_pad);
        for (uint32_t c = 0; c < nk * 4; ++c) {            // 16 bf16 = 8 columns per step
            const uint4 a = __ldg(src + 2 * c), b = __ldg(src + 2 * c + 1);
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            tc_st8(tmem_base + ((lq * 32u) << 16) + c * 8u, v);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t my_tiles = p.n_tiles > pair ? (p.n_tiles - pair + p.n_pairs - 1) / p.n_pairs : 0;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            auto prefetch = [&](uint32_t lt) {
                if (lt >= my_tiles) return;
                const uint32_t row0 = (pair + lt * p.n_pairs) * kGemmN;
                const uint32_t rows = min((uint32_t)kGemmN, p.n_rows - row0);
                bulk_prefetch_l2(p.base + (size_t)row0 * p.row_bytes, rows * p.row_bytes);
            };
            if (p.prefetch_tiles && group == 0)
                for (uint32_t lt = 0; lt < p.prefetch_tiles; ++lt) prefetch(lt);
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t tile = pair + lt * p.n_pairs;
                if (p.prefetch_tiles && group == 0) prefetch(lt + p.prefetch_tiles);
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&empty_bar[s], ((it / S) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], kGemmStageBytes);
                    tma_load_2d(stages + (size_t)s * kGemmStageBytes, &tmap, (int)(kc * kGemmKC), (int)(tile * kGemmN), &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, N = 64, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGemmN >> 3) << 17) | ((uint32_t)(kGemmM >> 4) << 24);
            uint32_t it = 0;
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t buf = lt & 1u;
                const uint32_t tmem_d = tmem_base + kGemmDCol + buf * kGemmN;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
                tc_fence_after();
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&full_bar[s], (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(stages + (size_t)s * kGemmStageBytes);
                    if (!(p.dbg_mode & 2u)) {
#pragma unroll
                        for (uint32_t k = 0; k < kGemmKC / 16; ++k) {
                            const uint64_t desc_b = make_b_desc(b_addr + k * 32u);
                            tc_mma_ts(tmem_d, tmem_base + kc * 32u + k * 8u, desc_b, idesc, (kc | k) != 0u ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty_bar[s]);                    // frees the stage when these MMAs have read it
                }
                tc_commit(&tmem_full[buf]);                      // accumulator of this tile is complete
            }
        }
    } else {
        // ================================ epilogue: thread = query ================================
        const uint32_t ew = warp - 2;                            // 0..7
        const uint32_t lq = warp & 3;                            // TMEM lane quarter this warp may read
        const uint32_t half = ew >> 2;                           // which 32 of the tile's 64 columns
        const uint32_t et = lq * 32 + lane;                      // 0..127: TMEM lane == query within the group
        uint64_t* my_q_key = wq_key + ew * kGemmQueue;
        uint8_t* my_q_lane = wq_lane + ew * kGemmQueue;
        uint64_t* warp_lists = lists + (size_t)ew * 32 * kGemmList;
        float* my_thr = thr_sm + ew * 32;
        float thr = -INFINITY;
        uint32_t qcnt = 0;
        const uint32_t lt_mask = (1u << lane) - 1u;

        auto drain = [&]() {
            thr = gemm_drain_queue(my_q_key, my_q_lane, warp_lists, my_thr, qcnt, lane);
            qcnt = 0;
        };

        // lane c holds 1/||row|| of column c of this warp's half of the tile (NaN for tombstones and rows past the end:
        // a NaN score never passes a comparison); the value for the NEXT tile is fetched while this one is processed
        auto fetch_inv = [&](uint32_t lt) -> float {
            float f = __int_as_float(0x7FC00000);
            if (lt < my_tiles) {
                const uint32_t r = (pair + lt * p.n_pairs) * kGemmN + half * 32 + lane;
                if (r < p.n_rows && p.live[r] != 0) f = p.inv_norm ? p.inv_norm[r] : 1.0f;
            }
            return f;
        };
        float inv_next = fetch_inv(0);

        for (uint32_t lt = 0; lt < my_tiles; ++lt) {
            const uint32_t tile = pair + lt * p.n_pairs;
            const uint32_t row0 = tile * kGemmN + half * 32;     // first row of this warp's 32 columns
            const uint32_t buf = lt & 1u;
            const float inv_reg = inv_next;
            inv_next = fetch_inv(lt + 1);
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((lq * 32u) << 16) + kGemmDCol + buf * kGemmN + half * 32u;
            uint32_t v[32];
            if (!(p.dbg_mode & 4u)) {
                tc_ld32(taddr, v);
                tc_wait_ld();
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0u;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);        // the MMA warp may overwrite this accumulator now
            if (p.dbg_mode & 1u) continue;
            if (p.dbg != nullptr && blockIdx.x == 0 && lt == 0) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    p.dbg[(size_t)et * kGemmN + half * 32 + c] = __uint_as_float(v[c]) * __shfl_sync(0xFFFFFFFFu, inv_reg, c);
            }
            if (lt == 0) {
                // first tile: the 32 scores of every query ARE its list; transpose through shared memory and sort each
                // list with a warp bitonic network instead of 1024 one-by-one inserts
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float sc = __uint_as_float(v[c]) * __shfl_sync(0xFFFFFFFFu, inv_reg, c);
                    warp_lists[lane * kGemmList + c] = (sc == sc) ? make_key(sc, row0 + c) : 0ull;
                }
                __syncwarp();
                for (uint32_t ql = 0; ql < 32; ++ql) {
                    uint64_t x = warp_lists[ql * kGemmList + lane];
#pragma unroll
                    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            const uint64_t o = shfl_xor_u64(x, j);
                            const bool asc = (lane & k) != 0, lower = (lane & j) == 0;
                            const bool take_min = (lower == asc);
                            x = take_min ? (o < x ? o : x) : (o > x ? o : x);
                        }
                    }
                    warp_lists[ql * kGemmList + lane] = x;               // descending: lane 0 holds the best key
                    if (lane == kGemmList - 1) my_thr[ql] = x != 0ull ? key_score(x) : -INFINITY;
                }
                __syncwarp();
                thr = my_thr[lane];
                continue;
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float sc = __uint_as_float(v[c]) * __shfl_sync(0xFFFFFFFFu, inv_reg, c);
                const bool pass = sc > thr;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, pass);
                if (m) {
                    if (pass) {
                        const uint32_t slot = qcnt + __popc(m & lt_mask);
                        my_q_key[slot] = make_key(sc, row0 + c);
                        my_q_lane[slot] = (uint8_t)lane;
                    }
                    qcnt += __popc(m);
                    if (qcnt > (uint32_t)(kGemmQueue - 32)) drain();
                }
            }
            if (qcnt) drain();
        }
        // ---- write this warp's 32 (CTA, column half, query) lists ----
        __syncwarp();
        const uint32_t L2 = 2 * p.n_pairs;
        for (uint32_t ql = 0; ql < 32; ++ql) {
            const size_t q = (size_t)group * kGemmM + lq * 32 + ql;
            const size_t li = q * L2 + pair * 2 + half;
            const uint64_t kv = warp_lists[ql * kGemmList + lane];
            p.out_keys[li * kGemmList + lane] = kv;
            if (lane == 0) p.out_tops[li] = kv;
            if (lane == kGemmList - 1) p.out_drops[li] = kv;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace lvs
