// K2 - tensor-core path for batched queries (regime 2 of BASELINE.json north_star): tcgen05.mma with the accumulators
// in TMEM, operands fed by TMA, top-k selection fused into the epilogue.
//
// Replaces, for Q >= 8 queries at a time, the O(Q*N*D) loop that the reference delegates to Qdrant behind
// QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) - there one gRPC call per query.
//
// Orientation: D[128 queries x 256 corpus rows] += A[128 x K] * B[256 x K]^T, bf16 inputs, fp32 accumulate, one
// tcgen05.mma (M = 128, N = 256, K = 16) per 16 K-elements: the full-rate shape of a single-CTA MMA (128 cycles).
//   * A pipeline stage = one 64-wide K chunk of BOTH operands, K-major with 128-byte swizzle: the corpus tile
//     B[256 rows x 64] (32 KB, from HBM) and the matching chunk of the CTA's 128 unit-norm bf16 queries A[128 x 64]
//     (16 KB; the queries total 192 KB and stay L2-resident, so re-streaming them costs L2 bandwidth, not HBM).
//     Both arrive by cp.async.bulk.tensor.2d (TMA, SASS UTMALDG) on one mbarrier; 3 stages = 144 KB in flight.
//     (A first version kept A in TMEM - the "TS" form.  Ablation showed that form pays ~160 cycles per MMA to read
//     the A slice out of TMEM whatever N is, 5x the N = 64 floor; see profiles/r01_k2_notes.md.)
//   * D is double buffered: 2 x 256 TMEM columns, so the MMAs of tile t+1 overlap the epilogue of tile t.
//   * Epilogue (8 warps; thread = query = TMEM lane; the two warps of a lane quarter split the 256 columns):
//     tcgen05.ld the scores, release the accumulator, scale by 1/||row|| (NaN for tombstones and rows past the end,
//     so they never pass), compare with the query's current 32nd-best score.  Passing (query, row, score) items go to
//     a per-warp shared-memory queue by ballot; the queue is drained WARP-COOPERATIVELY into sorted 32-key lists in
//     shared memory (lane j holds key j; the insertion position is a ballot + popc, the shift one shuffle).  The first
//     32 columns a warp ever sees are sorted into the lists directly (bitonic network) instead of 1024 inserts.
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocation,
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4).
// More than 128 queries: G = ceil(Q/128) CTAs ("a pair") walk the same tile sequence for different query groups; the
// second reader of a tile hits the 126 MB L2, so HBM still sees every row once.
//
// Exactness: lists hold 32 keys per (CTA, column half, query); the finalize kernel proves the result against the bound
// max(k'-th kept fast score, largest dropped score) + eps_q and the host repeats flagged queries on the K1 path.
// Algorithmic bytes per launch = rows x row_bytes; FLOPs = 2 * Q * rows * K.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lvs {

constexpr int kGemmEpiWarps = 8;       // two per TMEM lane quarter: each takes 128 of the tile's 256 columns
constexpr int kGemmThreads = (2 + kGemmEpiWarps) * 32;
constexpr int kGemmM = 128;            // queries per CTA (TMEM lanes)
constexpr int kGemmN = 256;            // corpus rows per tile (accumulator columns per buffer)
constexpr int kGemmKC = 64;            // K elements per pipeline stage (= one 128-byte swizzle row)
constexpr int kGemmABytes = kGemmM * kGemmKC * 2;       // 16 KB
constexpr int kGemmBBytes = kGemmN * kGemmKC * 2;       // 32 KB
constexpr int kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr int kGemmMaxStages = 3;
constexpr int kGemmList = 32;          // keys kept per (CTA, column half, query)
constexpr int kGemmMaxKChunks = 64;    // dim <= 4096
constexpr int kGemmQueue = 128;        // pending items per epilogue warp

struct GemmParams {
    uint32_t n_kchunks;
    uint32_t n_rows;
    uint32_t n_tiles;            // ceil(n_rows / 256)
    uint32_t n_groups;           // G: CTAs per pair
    uint32_t n_pairs;            // P
    uint32_t n_stages;
    const float* inv_norm;       // [rows] 1/||row||, or nullptr (dot metric)
    const uint8_t* live;
    uint64_t* out_keys;          // [n_groups * 128][2P][32]   (two lists per CTA and query: one per column half)
    uint64_t* out_tops;          // [n_groups * 128][2P]  best key of the list
    uint64_t* out_drops;         // [n_groups * 128][2P]  32nd key when the list is full (bound on what was dropped), else 0
    float* dbg;                  // optional: CTA 0 dumps its first accumulator tile [128 x 256] (already scaled)
    uint32_t keep;               // keys kept per list (<= 32): fewer keys = fewer inserts but a weaker drop bound
    uint32_t unit_rows;          // every stored row has | ||row|| - 1 | <= 2^-9: clean tiles skip the 1/||row|| scaling
    uint32_t dbg_mode;           // profiling aid: bit0 skip the epilogue's scoring, bit1 skip the MMA issue, bit2 skip tcgen05.ld
};

// shared memory: [stages (A 16 KB | B 32 KB)][lists 8*32*32*8][queue keys 8*128*8][queue lanes 8*128][thr 256*4][barriers]
__host__ __device__ inline size_t gemm_smem_bytes(uint32_t n_stages) {
    return 1024 /* alignment slack */ + (size_t)n_stages * kGemmStageBytes + (size_t)kGemmEpiWarps * 32 * kGemmList * 8 +
           kGemmEpiWarps * kGemmQueue * 8 + kGemmEpiWarps * kGemmQueue + kGemmEpiWarps * 32 * 4 + (2 * kGemmMaxStages + 8) * 8 + 16;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem_d] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32, K = 16
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
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

// K-major, 128-byte-swizzled operand tile: 8-row groups are 1024 B apart (SBO), version 1 (sm_100), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Warp-cooperative drain of one epilogue warp's queue into its sorted per-query lists (lane j holds key j of a list).
// Kept out of line: it is called from every column block of the unrolled score loop.  The next item's list row is
// fetched while the current one is inserted.  thr_q is the CTA-wide per-query threshold (orderable u32 of the score),
// shared by the two warps that serve a query: raising it to the 32nd key of EITHER list is safe because that key is
// recorded as the list's drop bound.
__device__ __noinline__ float gemm_drain_queue(const uint64_t* q_key, const uint8_t* q_lane, uint64_t* warp_lists, uint32_t* thr_q,
                                               uint32_t qcnt, int lane, uint32_t keep) {
    __syncwarp();
    uint64_t key = 0, cur = 0;
    uint32_t ql = 0;
    if (qcnt) { key = q_key[0]; ql = q_lane[0]; cur = warp_lists[ql * kGemmList + lane]; }
    for (uint32_t i = 0; i < qcnt; ++i) {
        uint64_t nkey = 0, ncur = 0;
        uint32_t nql = 0;
        if (i + 1 < qcnt) { nkey = q_key[i + 1]; nql = q_lane[i + 1]; ncur = warp_lists[nql * kGemmList + lane]; }
        const uint32_t pos = __popc(__ballot_sync(0xFFFFFFFFu, cur > key));
        if (pos < keep) {
            const uint64_t up = shfl_up_u64(cur, 1);
            uint64_t nv = (uint32_t)lane < pos ? cur : ((uint32_t)lane == pos ? key : up);
            if ((uint32_t)lane >= keep) nv = 0ull;
            warp_lists[ql * kGemmList + lane] = nv;
            if ((uint32_t)lane == keep - 1 && nv != 0ull) atomicMax(thr_q + ql, (uint32_t)(nv >> 32));
            if (nql == ql) ncur = nv;                    // the prefetched row of the same list is stale
        }
        key = nkey; ql = nql; cur = ncur;
    }
    __syncwarp();
    const uint32_t t = thr_q[lane];
    return t ? f32_from_orderable(t) : -INFINITY;
}

__global__ void __launch_bounds__(kGemmThreads, 1) gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_b,
                                                                    const __grid_constant__ CUtensorMap tmap_a, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    // 1024-byte alignment for the 128-byte swizzle; pointer arithmetic (not an integer round trip) keeps the accesses LDS/STS
    uint8_t* gsm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
    const uint32_t S = p.n_stages;
    uint8_t* stages = gsm;                                                   // S x (A 16 KB | B 32 KB), 1024-byte aligned
    uint64_t* lists = reinterpret_cast<uint64_t*>(gsm + (size_t)S * kGemmStageBytes);     // [8 warps][32 queries][32] sorted desc
    uint64_t* wq_key = lists + kGemmEpiWarps * 32 * kGemmList;               // [8][kGemmQueue]
    uint8_t* wq_lane = reinterpret_cast<uint8_t*>(wq_key + kGemmEpiWarps * kGemmQueue);   // [8][kGemmQueue]
    uint32_t* thr_sm = reinterpret_cast<uint32_t*>(wq_lane + kGemmEpiWarps * kGemmQueue); // [128] orderable threshold per query
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
    for (int i = tid; i < kGemmEpiWarps * 32; i += kGemmThreads) thr_sm[i] = 0u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const uint32_t my_tiles = p.n_tiles > pair ? (p.n_tiles - pair + p.n_pairs - 1) / p.n_pairs : 0;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t tile = pair + lt * p.n_pairs;
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&empty_bar[s], ((it / S) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], kGemmStageBytes);
                    uint8_t* st = stages + (size_t)s * kGemmStageBytes;
                    tma_load_2d(st, &tmap_a, (int)(kc * kGemmKC), (int)(group * kGemmM), &full_bar[s]);
                    tma_load_2d(st + kGemmABytes, &tmap_b, (int)(kc * kGemmKC), (int)(tile * kGemmN), &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGemmN >> 3) << 17) | ((uint32_t)(kGemmM >> 4) << 24);
            uint32_t it = 0;
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t buf = lt & 1u;
                const uint32_t tmem_d = tmem_base + buf * kGemmN;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
                tc_fence_after();
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&full_bar[s], (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stages + (size_t)s * kGemmStageBytes);
                    const uint32_t b_addr = a_addr + kGemmABytes;
                    if (!(p.dbg_mode & 2u)) {
#pragma unroll
                        for (uint32_t k = 0; k < kGemmKC / 16; ++k)
                            tc_mma_ss(tmem_d, make_kmajor_desc(a_addr + k * 32u), make_kmajor_desc(b_addr + k * 32u), idesc,
                                      (kc | k) != 0u ? 1u : 0u);
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
        const uint32_t half = ew >> 2;                           // which 128 of the tile's 256 columns
        const uint32_t et = lq * 32 + lane;                      // 0..127: TMEM lane == query within the group
        uint64_t* my_q_key = wq_key + ew * kGemmQueue;
        uint8_t* my_q_lane = wq_lane + ew * kGemmQueue;
        uint64_t* warp_lists = lists + (size_t)ew * 32 * kGemmList;
        uint32_t* my_thr = thr_sm + lq * 32;                     // shared by the two warps of this lane quarter
        float thr = -INFINITY;
        uint32_t qcnt = 0;
        const uint32_t lt_mask = (1u << lane) - 1u;
        constexpr int NB = kGemmN / 2 / 32;                      // 4 blocks of 32 columns per warp and tile

        auto drain = [&]() {
            thr = gemm_drain_queue(my_q_key, my_q_lane, warp_lists, my_thr, qcnt, lane, p.keep);
            qcnt = 0;
        };
        // lane c holds 1/||row|| of column (32 j + c) of this warp's half of the tile for j = 0..3 (NaN for tombstones and
        // rows past the end: a NaN score never passes a comparison); the NEXT tile's values are fetched during this one
        auto fetch_inv = [&](uint32_t lt, float (&f)[NB]) -> bool {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                f[j] = __int_as_float(0x7FC00000);
                if (lt < my_tiles) {
                    const uint32_t r = (pair + lt * p.n_pairs) * kGemmN + half * (kGemmN / 2) + j * 32 + lane;
                    if (r < p.n_rows && p.live[r] != 0) f[j] = p.inv_norm ? p.inv_norm[r] : 1.0f;
                    else ok = false;
                }
            }
            return ok;
        };
        float inv_next[NB];
        bool ok_next = fetch_inv(0, inv_next);

        for (uint32_t lt = 0; lt < my_tiles; ++lt) {
            const uint32_t tile = pair + lt * p.n_pairs;
            const uint32_t row_base = tile * kGemmN + half * (kGemmN / 2);   // first row of this warp's 128 columns
            const uint32_t buf = lt & 1u;
            float inv_reg[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) inv_reg[j] = inv_next[j];
            // unit-norm shards: when all 128 rows of this warp's half are live and in range the raw dot IS the score
            // (to within the norm deviation that the host adds to the error bound), so the scaling can be skipped
            const bool raw = p.unit_rows != 0u && __all_sync(0xFFFFFFFFu, ok_next);
            ok_next = fetch_inv(lt + 1, inv_next);
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((lq * 32u) << 16) + buf * kGemmN + half * (kGemmN / 2);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                uint32_t v[32];
                if (!(p.dbg_mode & 4u)) {
                    tc_ld32(taddr + j * 32u, v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = 0u;
                }
                if (j == NB - 1) {
                    // everything this warp needs from the accumulator is in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                if (p.dbg_mode & 1u) continue;
                const uint32_t row0 = row_base + j * 32;
                if (p.dbg != nullptr && blockIdx.x == 0 && lt == 0) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        p.dbg[(size_t)et * kGemmN + half * (kGemmN / 2) + j * 32 + c] = __uint_as_float(v[c]) * __shfl_sync(0xFFFFFFFFu, inv_reg[j], c);
                }
                if (lt == 0 && j == 0) {
                    // the first 32 scores of every query ARE its list: transpose through shared memory and sort each list
                    // with a warp bitonic network instead of 1024 one-by-one inserts
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float sc = __uint_as_float(v[c]) * __shfl_sync(0xFFFFFFFFu, inv_reg[j], c);
                        warp_lists[lane * kGemmList + c] = (sc == sc) ? make_key(sc, row0 + c) : 0ull;
                    }
                    __syncwarp();
                    for (uint32_t ql = 0; ql < 32; ++ql) {
                        uint64_t x = warp_lists[ql * kGemmList + lane];
#pragma unroll
                        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                                const uint64_t o = shfl_xor_u64(x, jj);
                                const bool asc = (lane & k) != 0, lower = (lane & jj) == 0;
                                const bool take_min = (lower == asc);
                                x = take_min ? (o < x ? o : x) : (o > x ? o : x);
                            }
                        }
                        if ((uint32_t)lane >= p.keep) x = 0ull;
                        warp_lists[ql * kGemmList + lane] = x;           // descending: lane 0 holds the best key
                        if ((uint32_t)lane == p.keep - 1 && x != 0ull) atomicMax(my_thr + ql, (uint32_t)(x >> 32));
                    }
                    __syncwarp();
                    { const uint32_t t = my_thr[lane]; thr = t ? f32_from_orderable(t) : -INFINITY; }
                    continue;
                }
                if (raw) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float sc = __uint_as_float(v[c]);
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
                    continue;
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float sc = __uint_as_float(v[c]) * __shfl_sync(0xFFFFFFFFu, inv_reg[j], c);
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
            if ((uint32_t)lane == p.keep - 1) p.out_drops[li] = kv;
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
