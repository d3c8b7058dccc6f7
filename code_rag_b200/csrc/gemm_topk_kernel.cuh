// K2 - tensor-core path for batched queries (regime 2 of BASELINE.json north_star): tcgen05.mma with the accumulators
// in TMEM, operands fed by TMA, top-k selection fused into the epilogue.
//
// Replaces, for Q >= 8 queries at a time, the O(Q*N*D) loop that the reference delegates to Qdrant behind
// QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) - there one gRPC call per query.
//
// Orientation: D[queries x 256 corpus rows] += A[queries x K] * B[256 x K]^T, bf16 inputs, fp32 accumulate.
// Two forms of the same kernel (template parameter PAIR):
//   * PAIR = false (Q <= 128 per CTA): tcgen05.mma.cta_group::1, M = 128, N = 256, K = 16 - the full-rate shape of a
//     single-CTA MMA (128 cycles).  A pipeline stage = one 64-wide K chunk of BOTH operands, K-major with 128-byte
//     swizzle: the corpus tile B[256 rows x 64] (32 KB, from HBM) and the matching chunk of the CTA's 128 unit-norm
//     bf16 queries A[128 x 64] (16 KB, L2-resident).  3 stages.  For 128 < Q <= 256 two CTAs would walk the same
//     tiles, and every CTA pulls 576 KB per tile through L2 -> SM (above the ~42 B/clk/SM the L2 can deliver), so:
//   * PAIR = true (128 < Q <= 256): a CLUSTER OF TWO CTAs (one TPC) issues tcgen05.mma.cta_group::2 with M = 256
//     (128 queries in each CTA's TMEM), N = 256.  Each CTA loads only ITS half of the corpus tile (128 rows x 64,
//     16 KB) plus its own 128 queries' chunk (16 KB); the tensor cores read the other half from the peer's shared
//     memory.  L2 -> SM traffic per tile and CTA drops from 576 KB to 384 KB; 4 stages of 32 KB.  The leader CTA
//     (cluster rank 0) issues all MMAs; both CTAs' TMA loads signal the leader's mbarrier; tcgen05.commit is
//     multicast to both CTAs' barriers; the peer's epilogue warps arrive remotely on the leader's tmem_empty barrier.
//   Both arrive by cp.async.bulk.tensor.2d (TMA, SASS UTMALDG).
//   * D is double buffered: 2 x 256 TMEM columns, so the MMAs of tile t+1 overlap the epilogue of tile t.
//   * Epilogue (8 warps; thread = query = TMEM lane; the two warps of a lane quarter split the 256 columns):
//     tcgen05.ld 32 columns, (scale by 1/||row|| unless the shard is unit-norm and the tile clean; NaN for tombstones
//     and rows past the end, and NaN never passes), then a NaN-ignoring MAX TREE over the 32 scores and ONE vote:
//     if no query of the warp has a score above its current threshold (the common case once the lists have warmed
//     up) the block costs ~40 instructions.  Otherwise each lane counts its passing scores, a warp reduction sizes
//     the batch, a shared-memory atomic hands out queue slots and the lanes write (key, query) items; the queue is
//     drained WARP-COOPERATIVELY into sorted per-query lists in shared memory (lane j holds key j; the insertion
//     position is a ballot + popc, the shift one shuffle).  The first 32 columns a warp ever sees are sorted into
//     the lists directly (bitonic network) instead of 1024 inserts.
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocation,
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4).
//
// Exactness: lists hold up to 32 keys per (CTA, column half, query); the finalize kernel proves the result against the
// bound max(k'-th kept fast score, largest dropped score) + eps_q and the host repeats flagged queries on the K1 path.
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
constexpr int kGemmBBytes = kGemmN * kGemmKC * 2;       // 32 KB (PAIR: each CTA holds half, 16 KB)
constexpr int kGemmMaxStages = 4;
constexpr int kGemmList = 32;          // keys kept per (CTA, column half, query)
constexpr int kGemmMaxKChunks = 64;    // dim <= 4096
constexpr int kGemmQueue = 128;        // pending items per epilogue warp
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA of a pair

__host__ __device__ constexpr int gemm_stage_bytes(bool pair) { return kGemmABytes + (pair ? kGemmBBytes / 2 : kGemmBBytes); }

struct GemmParams {
    uint32_t n_kchunks;
    uint32_t n_rows;
    uint32_t n_tiles;            // ceil(n_rows / 256)
    uint32_t n_groups;           // G: CTAs that walk the same tiles (PAIR: 2 = the cluster)
    uint32_t n_pairs;            // P: tile walkers
    uint32_t n_stages;
    const float* inv_norm;       // [rows] 1/||row||, or nullptr (dot metric)
    const uint8_t* live;
    uint64_t* out_keys;          // [n_groups * 128][2P][32]   (two lists per CTA and query: one per column half)
    uint64_t* out_tops;          // [n_groups * 128][2P]  best key of the list
    uint64_t* out_drops;         // [n_groups * 128][2P]  last kept key when the list is full (bound on what was dropped), else 0
    float* dbg;                  // optional: CTA 0 dumps its first accumulator tile [128 x 256] (already scaled)
    uint32_t keep;               // keys kept per list (<= 32): fewer keys = fewer inserts but a weaker drop bound
    uint32_t unit_rows;          // every stored row has | ||row|| - 1 | <= 2^-9: clean tiles skip the 1/||row|| scaling
    uint32_t dbg_mode;           // profiling aid: bit0 skip the epilogue's scoring, bit1 skip the MMA issue, bit2 skip tcgen05.ld
};

// shared memory: [stages][lists 8*32*32*8][queue keys 8*128*8][queue lanes 8*128][thr 128*4][inv 8*128*4][qcnt 8*4][barriers]
__host__ __device__ inline size_t gemm_smem_bytes(uint32_t n_stages, bool pair) {
    return 1024 /* alignment slack */ + (size_t)n_stages * gemm_stage_bytes(pair) + (size_t)kGemmEpiWarps * 32 * kGemmList * 8 +
           kGemmEpiWarps * kGemmQueue * 8 + kGemmEpiWarps * kGemmQueue + kGemmM * 4 + kGemmEpiWarps * 128 * 4 + kGemmEpiWarps * 4 +
           (2 * kGemmMaxStages + 8) * 8 + 16;
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
// D[tmem_d] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32, K = 16
template <bool PAIR>
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

// K-major, 128-byte-swizzled operand tile: 8-row groups are 1024 B apart (SBO), version 1 (sm_100), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Warp-cooperative drain of one epilogue warp's queue into its sorted per-query lists (lane j holds key j of a list).
// Kept out of line.  The next item's list row is fetched while the current one is inserted.  thr_q is the CTA-wide
// per-query threshold (orderable u32 of the score), shared by the two warps that serve a query: raising it to the
// last kept key of EITHER list is safe because that key is recorded as the list's drop bound.
__device__ __noinline__ float gemm_drain_queue(const uint64_t* q_key, const uint8_t* q_lane, uint64_t* warp_lists, uint32_t* thr_q,
                                               uint32_t* q_count, uint32_t qcnt, int lane, uint32_t keep) {
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
    if (lane == 0) *q_count = 0u;
    __syncwarp();
    const uint32_t t = thr_q[lane];
    return t ? f32_from_orderable(t) : -INFINITY;
}

template <bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_b,
                                                                    const __grid_constant__ CUtensorMap tmap_a, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t gsm_raw[];
    // 1024-byte alignment for the 128-byte swizzle; pointer arithmetic (not an integer round trip) keeps the accesses LDS/STS
    uint8_t* gsm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
    constexpr int kStageBytes = gemm_stage_bytes(PAIR);
    constexpr int kBRows = PAIR ? kGemmN / 2 : kGemmN;       // corpus rows this CTA loads per tile
    const uint32_t S = p.n_stages;
    uint8_t* stages = gsm;                                                   // S x (A 16 KB | B), 1024-byte aligned
    uint64_t* lists = reinterpret_cast<uint64_t*>(gsm + (size_t)S * kStageBytes);         // [8 warps][32 queries][32] sorted desc
    uint64_t* wq_key = lists + kGemmEpiWarps * 32 * kGemmList;               // [8][kGemmQueue]
    uint8_t* wq_lane = reinterpret_cast<uint8_t*>(wq_key + kGemmEpiWarps * kGemmQueue);   // [8][kGemmQueue]
    uint32_t* thr_sm = reinterpret_cast<uint32_t*>(wq_lane + kGemmEpiWarps * kGemmQueue); // [128] orderable threshold per query
    float* inv_sm = reinterpret_cast<float*>(thr_sm + kGemmM);               // [8][128] 1/||row|| of a warp's 128 columns
    uint32_t* qcnt_sm = reinterpret_cast<uint32_t*>(inv_sm + kGemmEpiWarps * 128);        // [8] queue fill
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(qcnt_sm + kGemmEpiWarps);
    uint64_t* empty_bar = full_bar + kGemmMaxStages;
    uint64_t* tmem_full = empty_bar + kGemmMaxStages;                        // [2]
    uint64_t* tmem_empty = tmem_full + 2;                                    // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t group = PAIR ? rank : blockIdx.x % p.n_groups;
    const uint32_t pair = PAIR ? blockIdx.x / 2 : blockIdx.x / p.n_groups;
    const uint32_t nk = p.n_kchunks;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
    if (tid < kGemmEpiWarps) qcnt_sm[tid] = 0u;
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the peer's barriers must exist before anything signals them
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
                    uint8_t* st = stages + (size_t)s * kStageBytes;
                    if constexpr (PAIR) {
                        // the leader's barrier counts the bytes of both CTAs
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * kStageBytes);
                        tma_load_2d_pair(st, &tmap_a, (int)(kc * kGemmKC), (int)(group * kGemmM), &full_bar[s]);
                        tma_load_2d_pair(st + kGemmABytes, &tmap_b, (int)(kc * kGemmKC), (int)(tile * kGemmN + rank * kBRows), &full_bar[s]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                        tma_load_2d(st, &tmap_a, (int)(kc * kGemmKC), (int)(group * kGemmM), &full_bar[s]);
                        tma_load_2d(st + kGemmABytes, &tmap_b, (int)(kc * kGemmKC), (int)(tile * kGemmN), &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (PAIR: the leader CTA only) ================================
        if (lane == 0 && rank == 0) {
            // instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, N = 256, M = 128 (PAIR: 256 over both CTAs)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGemmN >> 3) << 17) |
                                   ((uint32_t)((PAIR ? 2 * kGemmM : kGemmM) >> 4) << 24);
            uint32_t it = 0;
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t buf = lt & 1u;
                const uint32_t tmem_d = tmem_base + buf * kGemmN;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);   // the epilogue(s) have drained this accumulator
                tc_fence_after();
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&full_bar[s], (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stages + (size_t)s * kStageBytes);
                    const uint32_t b_addr = a_addr + kGemmABytes;
                    if (!(p.dbg_mode & 2u)) {
#pragma unroll
                        for (uint32_t k = 0; k < kGemmKC / 16; ++k)
                            tc_mma_ss<PAIR>(tmem_d, make_kmajor_desc(a_addr + k * 32u), make_kmajor_desc(b_addr + k * 32u), idesc,
                                            (kc | k) != 0u ? 1u : 0u);
                    }
                    tc_commit<PAIR>(&empty_bar[s]);                  // frees the stage (in both CTAs) when these MMAs have read it
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
        uint64_t* my_q_key = wq_key + ew * kGemmQueue;
        uint8_t* my_q_lane = wq_lane + ew * kGemmQueue;
        uint64_t* warp_lists = lists + (size_t)ew * 32 * kGemmList;
        uint32_t* my_thr = thr_sm + lq * 32;                     // shared by the two warps of this lane quarter
        float* my_inv = inv_sm + ew * 128;
        uint32_t* my_qcnt = qcnt_sm + ew;
        float thr = -INFINITY;
        uint32_t qcnt = 0;                                       // warp-uniform mirror of *my_qcnt
        const uint32_t lt_mask = (1u << lane) - 1u;
        constexpr int NB = kGemmN / 2 / 32;                      // 4 blocks of 32 columns per warp and tile

        auto drain = [&]() {
            thr = gemm_drain_queue(my_q_key, my_q_lane, warp_lists, my_thr, my_qcnt, qcnt, lane, p.keep);
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
        // One block of 32 scores per lane (v = fp32 bit patterns, already scaled).  Fast exit: nothing above any threshold.
        auto select_block = [&](uint32_t (&v)[32], uint32_t row0) {
            float g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                g[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                             fmaxf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
            const float mx = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
            if (!__any_sync(0xFFFFFFFFu, mx > thr)) return;
            uint32_t cnt = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c) cnt += (__uint_as_float(v[c]) > thr) ? 1u : 0u;
            uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (qcnt + tot > (uint32_t)kGemmQueue) {
                drain();                                         // raises thr: count again
                cnt = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) cnt += (__uint_as_float(v[c]) > thr) ? 1u : 0u;
                tot = __reduce_add_sync(0xFFFFFFFFu, cnt);
            }
            if (tot <= (uint32_t)kGemmQueue) {
                if (tot == 0u) return;
                uint32_t slot = 0;
                if (cnt) slot = atomicAdd(my_qcnt, cnt);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (g[i] > thr) {
#pragma unroll
                        for (int c = 4 * i; c < 4 * i + 4; ++c) {
                            const float sc = __uint_as_float(v[c]);
                            if (sc > thr) {
                                my_q_key[slot] = make_key(sc, row0 + c);
                                my_q_lane[slot] = (uint8_t)lane;
                                ++slot;
                            }
                        }
                    }
                }
                qcnt += tot;
                __syncwarp();
            } else {
                // early in the run more scores pass than the queue holds: column by column, draining in between
#pragma unroll 1
                for (int c4 = 0; c4 < 8; ++c4) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        float sc = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (i == c4) sc = __uint_as_float(v[4 * i + cc]);   // register select, no local memory
                        const bool pass = sc > thr;
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, pass);
                        if (m) {
                            if (pass) {
                                const uint32_t slot = qcnt + __popc(m & lt_mask);
                                my_q_key[slot] = make_key(sc, row0 + 4 * c4 + cc);
                                my_q_lane[slot] = (uint8_t)lane;
                            }
                            qcnt += __popc(m);
                            if (qcnt > (uint32_t)(kGemmQueue - 32)) drain();
                        }
                    }
                }
                if (lane == 0) *my_qcnt = qcnt;
                __syncwarp();
            }
        };
        float inv_next[NB];
        bool ok_next = fetch_inv(0, inv_next);

        for (uint32_t lt = 0; lt < my_tiles; ++lt) {
            const uint32_t tile = pair + lt * p.n_pairs;
            const uint32_t row_base = tile * kGemmN + half * (kGemmN / 2);   // first row of this warp's 128 columns
            const uint32_t buf = lt & 1u;
            // unit-norm shards: when all 128 rows of this warp's half are live and in range the raw dot IS the score
            // (to within the norm deviation that the host adds to the error bound), so the scaling can be skipped
            const bool raw = p.unit_rows != 0u && p.dbg == nullptr && __all_sync(0xFFFFFFFFu, ok_next);
            if (!raw) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NB; ++j) my_inv[j * 32 + lane] = inv_next[j];
                __syncwarp();
            }
            ok_next = fetch_inv(lt + 1, inv_next);
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
                if (lt == 0 && j == 0) {
                    // the first 32 scores of every query ARE its list: transpose through shared memory and sort each list
                    // with a warp bitonic network instead of 1024 one-by-one inserts
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float sc = __uint_as_float(v[c]);
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
                select_block(v, row0);
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
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace lvs
