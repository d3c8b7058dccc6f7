// K2 - tensor-core path for batched queries (regime 2 of BASELINE.json north_star): tcgen05.mma with the accumulators
// in TMEM, operands fed by TMA, top-k selection fused into the epilogue.
//
// Replaces, for batches of queries (from 3 on bf16 shards, from 5 on fp32 ones), the O(Q*N*D) loop that the reference delegates to Qdrant behind
// QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) - there one gRPC call per query.
//
// Orientation: D[queries x 256 corpus rows] += A[queries x K] * B[256 x K]^T, fp32 accumulate.  bf16 shards use kind::f16
// (64 K-elements per 128-byte swizzle row, K = 16 per MMA); fp32 shards are read AS THEY ARE with kind::tf32 (template parameter
// TF32: 32 K-elements per swizzle row, K = 8 per MMA - the same bytes per stage and per instruction; the tensor core uses the
// upper 19 bits of every fp32 value, which the finalize kernel's bound accounts for with 2^-10 ||q|| ||x||).
// Two forms of the same kernel (template parameter PAIR):
//   * PAIR = false (Q <= 128 per CTA): tcgen05.mma.cta_group::1, M = 128, N = 256, K = 16 - the full-rate shape of a
//     single-CTA MMA (128 cycles).  A pipeline stage = one 64-wide K chunk of BOTH operands, K-major with 128-byte
//     swizzle: the corpus tile B[256 rows x 64] (32 KB, from HBM) and the matching chunk of the CTA's 128 unit-norm
//     bf16 queries A[128 x 64] (16 KB, L2-resident).  3 stages.  For 128 < Q <= 256 two CTAs would walk the same
//     tiles and every CTA would pull 576 KB per tile through L2 -> SM, so:
//   * PAIR = true (128 < Q <= 256): a CLUSTER OF TWO CTAs (one TPC) issues tcgen05.mma.cta_group::2 with M = 256
//     (128 queries in each CTA's TMEM), N = 256.  Each CTA loads only ITS half of the corpus tile (128 rows x 64,
//     16 KB) plus its own 128 queries' chunk (16 KB); the tensor cores read the other half from the peer's shared
//     memory.  L2 -> SM traffic per tile and CTA drops from 576 KB to 384 KB.  The leader CTA
//     (cluster rank 0) issues all MMAs; both CTAs' TMA loads signal the leader's mbarrier; tcgen05.commit is
//     multicast to both CTAs' barriers; the peer's epilogue warps arrive remotely on the leader's tmem_empty barrier.
//     SPLIT RINGS (default of the pair form): the two operands have very different latencies - the query chunks come out of L2,
//     the corpus out of HBM - so they do not share pipeline stages.  4 query-chunk buffers and 5 corpus half-tile buffers
//     (16 KB each, own full / empty barriers) are requested independently by the one producer thread, which polls both
//     rings (mbarrier.test_wait) and serves the corpus ring first: 80 KB of HBM bytes in flight per SM instead of 64 KB in
//     the same shared memory, and a late query chunk no longer holds a corpus request back.  The MMA thread waits for one
//     buffer of each ring and commits to both.  6-10 % shorter kernel on every same-box comparison (profiles/r02_k2_split_rings_notes.md);
//     n_stages_b = 0 gives the former single ring of 4 x 32 KB stages.
//   Both arrive by cp.async.bulk.tensor.2d (TMA, SASS UTMALDG).
//   * D is double buffered: 2 x 256 TMEM columns, so the MMAs of tile t+1 overlap the epilogue of tile t.
//   * Epilogue (8 warps; thread = query = TMEM lane; the two warps of a lane quarter split the 256 columns):
//     tcgen05.ld 32 columns, (scale by 1/||row|| unless the shard is unit-norm and the tile clean; NaN for tombstones,
//     rows that fail the payload filter and rows past the end, and NaN never passes), then a NaN-ignoring MAX TREE over
//     the 32 scores and ONE vote:
//     if no query of the warp has a score above its current threshold (the common case once the lists have warmed
//     up) the block costs ~40 instructions.  Otherwise every lane builds the bit mask of its passing columns, the
//     passing lanes park their 32 scores in a shared-memory column (so that they can be indexed) and ALL LANES INSERT
//     IN PARALLEL, one item per round: a lane owns its query's list (16 unsorted keys in a bank-conflict-free shared
//     memory column, the minimum tracked in registers), an insert replaces the minimum and rescans.  No queue, no
//     atomics, no warp-wide serialisation per item; the lists are sorted once, when the kernel ends.
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocation,
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4).
//
// Exactness: lists hold up to 16 keys per (CTA, column half, query) and record the threshold below which they dropped
// scores; the finalize kernel proves the result against the bound max(k'-th kept fast score, largest drop bound) + eps_q and
// the host repeats flagged queries on the K1 path.
// Algorithmic bytes per launch = rows x row_bytes; FLOPs = 2 * Q * rows * K.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lvs {

constexpr int kGemmEpiWarps = 8;       // two per TMEM lane quarter: each takes 128 of the tile's 256 columns
constexpr int kGemmThreads = (2 + kGemmEpiWarps) * 32;
constexpr int kGemmM = 128;            // queries per CTA (TMEM lanes)
constexpr int kGemmN = 256;            // corpus rows per tile (accumulator columns per buffer)
constexpr int kGemmKC = 64;            // K elements per pipeline stage (= one 128-byte swizzle row) for bf16; 32 for fp32 / tf32
constexpr int kGemmABytes = kGemmM * kGemmKC * 2;       // 16 KB
constexpr int kGemmBBytes = kGemmN * kGemmKC * 2;       // 32 KB (PAIR: each CTA holds half, 16 KB)
constexpr int kGemmMaxStages = 4;
constexpr int kGemmMaxStagesB = 8;     // split rings (pair form): corpus half-tiles in flight
constexpr int kGemmDefaultStagesA = 4; // defaults of the pair form: 0 corpus buffers = one combined ring of kGemmDefaultStagesA stages
constexpr int kGemmDefaultStagesB = 5;
constexpr int kGemmList = 16;          // keys kept per (CTA, column half, query)
constexpr int kGemmMaxKChunks = 128;   // dim <= 8192 (bf16) / 4096 (fp32)
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA of a pair

__host__ __device__ constexpr int gemm_stage_bytes(bool pair) { return kGemmABytes + (pair ? kGemmBBytes / 2 : kGemmBBytes); }

struct GemmParams {
    uint32_t n_kchunks;
    uint32_t n_queries;          // queries of this launch (<= n_groups * 128); the TMEM lanes behind them score a zero query and are
                                 // masked out of the selection (they would never raise a threshold and drag every block through the insert path)
    uint32_t n_rows;
    uint32_t n_tiles;            // ceil(n_rows / 256)
    uint32_t n_groups;           // G: CTAs that walk the same tiles (PAIR: 2 = the cluster)
    uint32_t n_pairs;            // P: tile walkers
    uint32_t n_stages;
    uint32_t n_stages_b;         // pair form only, 0 = one ring of (query chunk | corpus half-tile) stages.  > 0: SPLIT RINGS - n_stages query-chunk
                                 // buffers (L2-resident, short latency) and n_stages_b corpus half-tile buffers (HBM, long latency) of 16 KB
                                 // each, requested independently, so that more of the shared memory holds bytes that are in flight from HBM
    const float* inv_norm;       // [rows] 1/||row||, or nullptr (dot metric)
    const uint8_t* live;
    const uint32_t* fcodes[kMaxFilterCols];   // payload filter: dictionary-code columns that must equal fwant[] (conjunction)
    uint32_t fwant[kMaxFilterCols];
    uint32_t n_filter;
    uint64_t* out_keys;          // [n_groups * 128][2P][16]   (two lists per CTA and query: one per column half), sorted descending
    uint64_t* out_tops;          // [n_groups * 128][2P]  best key of the list
    uint64_t* out_drops;         // [n_groups * 128][2P]  key-shaped bound on every score this list dropped (its final threshold), 0 = nothing dropped
    float* dbg;                  // optional: CTA 0 dumps its first accumulator tile [128 x 256] (already scaled)
    uint32_t keep;               // keys kept per list (<= 16): fewer keys = fewer inserts but a weaker drop bound
    uint32_t unit_rows;          // every stored row has | ||row|| - 1 | <= 2^-9: clean tiles skip the 1/||row|| scaling
    uint32_t dbg_mode;           // profiling aid: bit0 skip the epilogue's scoring, bit1 skip the MMA issue, bit2 skip tcgen05.ld,
                                 // bit3 (pair form) stop re-loading the query chunks after the first pipeline fill, bit4 no loads at all
                                 // after the first fill (MMA + epilogue hand-off alone); bits 3 and 4 act on the combined ring only
                                 // (n_stages_b = 0)
};

// shared memory: [stages][lists 8 warps * 16 keys * 32 lanes * 8][score columns 8 * 32 * 32 * 4][thr 128*4][inv 8*128*4][barriers]
__host__ __device__ inline size_t gemm_smem_bytes(uint32_t n_stages, bool pair, uint32_t n_stages_b = 0) {
    const size_t ring = n_stages_b ? (size_t)n_stages * kGemmABytes + (size_t)n_stages_b * (kGemmBBytes / 2)
                                   : (size_t)n_stages * gemm_stage_bytes(pair);
    return 1024 /* alignment slack */ + ring + (size_t)kGemmEpiWarps * 32 * kGemmList * 8 +
           (size_t)kGemmEpiWarps * 32 * 32 * 4 + kGemmM * 4 + kGemmEpiWarps * 128 * 4 + 256 /* barriers */;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// CTA-pair form: the data lands in THIS CTA's shared memory, the bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <bool PAIR>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    if constexpr (PAIR) {
        // arrives on the barrier at this offset in BOTH CTAs of the pair
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
}
// D[tmem_d] (+)= A[smem desc] * B[smem desc]^T: bf16 x bf16 -> fp32 with K = 16, or tf32 x tf32 -> fp32 with K = 8 (32 bytes either way)
template <bool PAIR, bool TF32>
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR && TF32) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else if constexpr (PAIR) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else if constexpr (TF32) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
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
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at this offset in the leader CTA of the pair (works from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

// K-major, 128-byte-swizzled operand tile: 8-row groups are 1024 B apart (SBO), version 1 (sm_100), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <bool PAIR, bool TF32>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_b,
                                                                    const __grid_constant__ CUtensorMap tmap_a, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    // 1024-byte alignment for the 128-byte swizzle; pointer arithmetic (not an integer round trip) keeps the accesses LDS/STS
    uint8_t* gsm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
    constexpr int kStageBytes = gemm_stage_bytes(PAIR);
    constexpr int kBRows = PAIR ? kGemmN / 2 : kGemmN;       // corpus rows this CTA loads per tile
    constexpr int kKcElems = TF32 ? kGemmKC / 2 : kGemmKC;  // K elements per stage (128 bytes of a row)
    const uint32_t S = p.n_stages;
    const uint32_t SB = PAIR ? p.n_stages_b : 0u;                            // split rings: S query-chunk buffers, then SB corpus buffers
    uint8_t* stages = gsm;                                                   // S x (A 16 KB | B), 1024-byte aligned
    uint8_t* b_bufs = gsm + (size_t)S * kGemmABytes;                         // split rings only
    const size_t ring_bytes = SB ? (size_t)S * kGemmABytes + (size_t)SB * (kGemmBBytes / 2) : (size_t)S * kStageBytes;
    uint64_t* lists = reinterpret_cast<uint64_t*>(gsm + ring_bytes);         // [8 warps][16 keys][32 lanes], unsorted
    float* sbuf = reinterpret_cast<float*>(lists + kGemmEpiWarps * 32 * kGemmList);       // [8 warps][32 columns][32 lanes]
    uint32_t* thr_sm = reinterpret_cast<uint32_t*>(sbuf + kGemmEpiWarps * 32 * 32);       // [128] orderable threshold per query
    float* inv_sm = reinterpret_cast<float*>(thr_sm + kGemmM);               // [8][128] 1/||row|| of a warp's 128 columns
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(inv_sm + kGemmEpiWarps * 128);
    uint64_t* empty_bar = full_bar + kGemmMaxStages;
    uint64_t* tmem_full = empty_bar + kGemmMaxStages;                        // [2]
    uint64_t* tmem_empty = tmem_full + 2;                                    // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint64_t* full_b = tmem_empty + 3;                                       // split rings: [kGemmMaxStagesB]
    uint64_t* empty_b = full_b + kGemmMaxStagesB;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t group = PAIR ? rank : blockIdx.x % p.n_groups;
    const uint32_t pair = PAIR ? blockIdx.x / 2 : blockIdx.x / p.n_groups;
    const uint32_t nk = p.n_kchunks;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (uint32_t s = 0; s < SB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], (PAIR ? 2 : 1) * kGemmEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    for (int i = tid; i < kGemmEpiWarps * 32 * kGemmList; i += kGemmThreads) lists[i] = 0ull;
    for (int i = tid; i < kGemmM; i += kGemmThreads) thr_sm[i] = 0u;
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const uint32_t my_tiles = p.n_tiles > pair ? (p.n_tiles - pair + p.n_pairs - 1) / p.n_pairs : 0;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0 && SB != 0u) {
            if constexpr (PAIR) {
                // split rings: one thread polls both rings and requests whatever has a free buffer, the corpus first - its ring is
                // the deep one, so its requests run ahead of the query chunks' by the difference in depth
                const uint32_t total = my_tiles * nk;
                uint32_t ia = 0, sa = 0, pa = 1, kca = 0;                    // next query chunk: index, buffer, parity to wait for, K chunk
                uint32_t ib = 0, sb = 0, pb = 1, kcb = 0, ltb = 0;           // next corpus half-tile
                while (ia < total || ib < total) {
                    if (ib < total && mbar_test(&empty_b[sb], pb)) {
                        if (rank == 0) mbar_arrive_expect_tx(&full_b[sb], (uint32_t)kGemmBBytes);      // both CTAs' halves
                        tma_load_2d_pair(b_bufs + (size_t)sb * (kGemmBBytes / 2), &tmap_b, (int)(kcb * kKcElems),
                                         (int)((pair + ltb * p.n_pairs) * kGemmN + rank * kBRows), &full_b[sb]);
                        ++ib;
                        if (++sb == SB) { sb = 0; pb ^= 1u; }
                        if (++kcb == nk) { kcb = 0; ++ltb; }
                    }
                    if (ia < total && mbar_test(&empty_bar[sa], pa)) {
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[sa], 2u * kGemmABytes);         // both CTAs' query chunks
                        tma_load_2d_pair(stages + (size_t)sa * kGemmABytes, &tmap_a, (int)(kca * kKcElems), (int)(group * kGemmM), &full_bar[sa]);
                        ++ia;
                        if (++sa == S) { sa = 0; pa ^= 1u; }
                        if (++kca == nk) kca = 0;
                    }
                }
            }
        } else if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t tile = pair + lt * p.n_pairs;
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&empty_bar[s], ((it / S) & 1u) ^ 1u);
                    uint8_t* st = stages + (size_t)s * kStageBytes;
                    if ((p.dbg_mode & 16u) != 0u && it >= S) {                       // profiling aid: no loads at all after the first fill
                        if (rank == 0) mbar_arrive(&full_bar[s]);                   // (stale operands, timing only)
                        continue;
                    }
                    if constexpr (PAIR) {
                        // the leader's barrier counts the bytes of both CTAs
                        const bool skip_a = (p.dbg_mode & 8u) != 0u && it >= S;      // profiling aid: stale query chunks, timing only
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], skip_a ? 2u * (kStageBytes - kGemmABytes) : 2u * kStageBytes);
                        if (!skip_a)
                        tma_load_2d_pair(st, &tmap_a, (int)(kc * kKcElems), (int)(group * kGemmM), &full_bar[s]);
                        tma_load_2d_pair(st + kGemmABytes, &tmap_b, (int)(kc * kKcElems), (int)(tile * kGemmN + rank * kBRows), &full_bar[s]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                        tma_load_2d(st, &tmap_a, (int)(kc * kKcElems), (int)(group * kGemmM), &full_bar[s]);
                        tma_load_2d(st + kGemmABytes, &tmap_b, (int)(kc * kKcElems), (int)(tile * kGemmN), &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (PAIR: the leader CTA only) ================================
        if (lane == 0 && rank == 0) {
            // instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, N = 256, M = 128 (PAIR: 256 over both CTAs)
            // (formats: D = F32 (1 at bit 4); A, B = BF16 (1) or TF32 (2) at bits 7 and 10)
            constexpr uint32_t kFmt = TF32 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(kGemmN >> 3) << 17) |
                                   ((uint32_t)((PAIR ? 2 * kGemmM : kGemmM) >> 4) << 24);
            uint32_t it = 0;
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;                  // split rings: buffer and parity of the next chunk of each ring
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t buf = lt & 1u;
                const uint32_t tmem_d = tmem_base + buf * kGemmN;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);   // the epilogue(s) have drained this accumulator
                tc_fence_after();
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    uint32_t a_addr, b_addr;
                    if (SB != 0u) {
                        mbar_wait(&full_b[sb], pb);
                        mbar_wait(&full_bar[sa], pa);
                        a_addr = smem_u32(stages + (size_t)sa * kGemmABytes);
                        b_addr = smem_u32(b_bufs + (size_t)sb * (kGemmBBytes / 2));
                    } else {
                        const uint32_t s = it % S;
                        mbar_wait(&full_bar[s], (it / S) & 1u);
                        a_addr = smem_u32(stages + (size_t)s * kStageBytes);
                        b_addr = a_addr + kGemmABytes;
                    }
                    tc_fence_after();
                    if (!(p.dbg_mode & 2u)) {
#pragma unroll
                        for (uint32_t k = 0; k < kGemmKC / 16; ++k)
                            tc_mma_ss<PAIR, TF32>(tmem_d, make_kmajor_desc(a_addr + k * 32u), make_kmajor_desc(b_addr + k * 32u), idesc,
                                            (kc | k) != 0u ? 1u : 0u);
                    }
                    if (SB != 0u) {                                  // each ring's buffer is freed (in both CTAs) when these MMAs have read it
                        tc_commit<PAIR>(&empty_bar[sa]);
                        tc_commit<PAIR>(&empty_b[sb]);
                        if (++sa == S) { sa = 0; pa ^= 1u; }
                        if (++sb == SB) { sb = 0; pb ^= 1u; }
                    } else {
                        tc_commit<PAIR>(&empty_bar[it % S]);         // frees the stage (in both CTAs) when these MMAs have read it
                    }
                }
                tc_commit<PAIR>(&tmem_full[buf]);                    // accumulator of this tile is complete (in both CTAs)
            }
        }
    } else {
        // ================================ epilogue: thread = query ================================
        const uint32_t ew = warp - 2;                            // 0..7
        const uint32_t lq = warp & 3;                            // TMEM lane quarter this warp may read
        const uint32_t half = ew >> 2;                           // which 128 of the tile's 256 columns
        const uint32_t et = lq * 32 + lane;                      // 0..127: TMEM lane == query within the group
        uint64_t* my_list = lists + (size_t)ew * 32 * kGemmList + lane;     // key j of this lane's list at my_list[j * 32]
        float* my_sbuf = sbuf + (size_t)ew * 32 * 32 + lane;                 // score of column c at my_sbuf[c * 32]
        uint32_t* my_thr = thr_sm + lq * 32;                     // shared by the two warps of this lane quarter
        float* my_inv = inv_sm + ew * 128;
        const uint32_t keep = p.keep;
        const bool active = group * kGemmM + et < p.n_queries;   // a padding lane never selects anything
        float thr = active ? -INFINITY : INFINITY;                // scores at or below thr are dropped
        uint64_t minkey = 0ull;                                  // smallest key of the list (0 while it has empty slots)
        uint32_t minpos = 0;
        constexpr int NB = kGemmN / 2 / 32;                      // 4 blocks of 32 columns per warp and tile

        // new minimum of the lane's list; a full list raises the query's threshold (published for the partner warp)
        auto rescan = [&]() {
            uint64_t mn = ~0ull;
            uint32_t mp = 0;
#pragma unroll 4
            for (uint32_t j = 0; j < keep; ++j) {
                const uint64_t kk = my_list[j * 32];
                if (kk < mn) { mn = kk; mp = j; }
            }
            minkey = mn; minpos = mp;
            if (mn != 0ull) {
                const float t = key_score(mn);
                if (t > thr) { thr = t; atomicMax(my_thr + lane, (uint32_t)(mn >> 32)); }
            }
        };
        // the NEXT tile's 1/||row|| and live bytes are requested during this tile and only looked at when it starts
        // (lane c: column 32 j + c of this warp's half of the tile, j = 0..3)
        auto fetch_inv = [&](uint32_t lt, float (&f)[NB], uint32_t (&lv)[NB]) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                f[j] = 1.0f; lv[j] = 0u;
                if (lt < my_tiles) {
                    uint32_t r = (pair + lt * p.n_pairs) * kGemmN + half * (kGemmN / 2) + j * 32 + lane;
                    r = r < p.n_rows ? r : p.n_rows - 1;
                    uint32_t ok = p.live[r];
                    // payload filter (reference client.py:171-176): rows that fail it are treated like tombstones
                    for (uint32_t c = 0; c < p.n_filter; ++c) ok = (p.fcodes[c][r] == p.fwant[c]) ? ok : 0u;
                    lv[j] = ok;
                    if (p.inv_norm) f[j] = p.inv_norm[r];
                }
            }
        };
        // One block of 32 scores per lane (v = fp32 bit patterns, already scaled; NaN = not a candidate).
        auto select_block = [&](uint32_t (&v)[32], uint32_t row0, bool first) {
            uint32_t mask = 0;
            if (first) {
                // the first columns a list sees are its first keys
                if (active) {
#pragma unroll
                    for (int c = 0; c < kGemmList; ++c) {
                        const float sc = __uint_as_float(v[c]);
                        if ((uint32_t)c < keep) my_list[c * 32] = (sc == sc) ? make_key(sc, row0 + c) : 0ull;
                    }
                    rescan();
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) mask |= ((uint32_t)c >= keep && __uint_as_float(v[c]) > thr) ? (1u << c) : 0u;
            } else {
                float g[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    g[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                                 fmaxf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
                const float mx = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                if (!__any_sync(0xFFFFFFFFu, mx > thr)) return;                  // fast exit: nothing above any threshold
#pragma unroll
                for (int c = 0; c < 32; ++c) mask |= (__uint_as_float(v[c]) > thr) ? (1u << c) : 0u;
            }
            if (mask) {
#pragma unroll
                for (int c = 0; c < 32; ++c) my_sbuf[c * 32] = __uint_as_float(v[c]);
                // every lane inserts its own items; the threshold rises as it goes, so later items are re-checked
                while (mask) {
                    const uint32_t c = (uint32_t)__ffs((int)mask) - 1u;
                    mask &= mask - 1u;
                    const float sc = my_sbuf[c * 32];
                    if (sc > thr) {
                        my_list[minpos * 32] = make_key(sc, row0 + c);
                        rescan();
                    }
                }
            }
            __syncwarp();
        };
        float inv_next[NB];
        uint32_t live_next[NB];
        fetch_inv(0, inv_next, live_next);

        for (uint32_t lt = 0; lt < my_tiles; ++lt) {
            const uint32_t tile = pair + lt * p.n_pairs;
            const uint32_t row_base = tile * kGemmN + half * (kGemmN / 2);   // first row of this warp's 128 columns
            const uint32_t buf = lt & 1u;
            bool ok = true;
            float inv_cur[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const bool good = row_base + j * 32 + lane < p.n_rows && live_next[j] != 0u;
                inv_cur[j] = good ? inv_next[j] : __int_as_float(0x7FC00000);      // NaN: tombstone or past the end
                ok = ok && good;
            }
            // unit-norm shards: when all 128 rows of this warp's half are live and in range the raw dot IS the score
            // (to within the norm deviation that the host adds to the error bound), so the scaling can be skipped
            const bool raw = p.unit_rows != 0u && p.dbg == nullptr && __all_sync(0xFFFFFFFFu, ok);
            if (!raw) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NB; ++j) my_inv[j * 32 + lane] = inv_cur[j];
                __syncwarp();
            }
            fetch_inv(lt + 1, inv_next, live_next);
            { const uint32_t t = my_thr[lane]; if (t) thr = fmaxf(thr, f32_from_orderable(t)); }   // the partner warp's progress
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((lq * 32u) << 16) + buf * kGemmN + half * (kGemmN / 2);
#pragma unroll 1
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
                    if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&tmem_empty[buf]); else mbar_arrive(&tmem_empty[buf]); }
                }
                if (p.dbg_mode & 1u) continue;
                const uint32_t row0 = row_base + j * 32;
                if (!raw) {
                    const float4* iv = reinterpret_cast<const float4*>(my_inv + j * 32);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 f = iv[c4];
                        v[4 * c4 + 0] = __float_as_uint(__uint_as_float(v[4 * c4 + 0]) * f.x);
                        v[4 * c4 + 1] = __float_as_uint(__uint_as_float(v[4 * c4 + 1]) * f.y);
                        v[4 * c4 + 2] = __float_as_uint(__uint_as_float(v[4 * c4 + 2]) * f.z);
                        v[4 * c4 + 3] = __float_as_uint(__uint_as_float(v[4 * c4 + 3]) * f.w);
                    }
                }
                if (p.dbg != nullptr && blockIdx.x == 0 && lt == 0) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) p.dbg[(size_t)et * kGemmN + half * (kGemmN / 2) + j * 32 + c] = __uint_as_float(v[c]);
                }
                select_block(v, row0, lt == 0 && j == 0);
            }
        }
        // ---- sort (warp bitonic network, descending) and write this warp's 32 (CTA, column half, query) lists ----
        __syncwarp();
        const uint32_t L2 = 2 * p.n_pairs;
        uint64_t* warp_lists = lists + (size_t)ew * 32 * kGemmList;
        for (uint32_t ql = 0; ql < 32; ++ql) {
            uint64_t x = (uint32_t)lane < keep ? warp_lists[lane * 32 + ql] : 0ull;
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
            const size_t q = (size_t)group * kGemmM + lq * 32 + ql;
            const size_t li = q * L2 + pair * 2 + half;
            if (lane < kGemmList) p.out_keys[li * kGemmList + lane] = x;             // lane 0 holds the best key
            if (lane == 0) p.out_tops[li] = x;
        }
        {
            const size_t q = (size_t)group * kGemmM + lq * 32 + lane;
            const size_t li = q * L2 + pair * 2 + half;
            p.out_drops[li] = (active && thr > -INFINITY) ? (((uint64_t)f32_orderable(thr) << 32) | 0xFFFFFFFFull) : 0ull;
        }
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace lvs
