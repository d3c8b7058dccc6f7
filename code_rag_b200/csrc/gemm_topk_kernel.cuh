// K2 - tensor-core path for batched queries (regime 2 of BASELINE.json north_star): tcgen05.mma with the accumulator
// in TMEM, corpus tiles fed by TMA, top-k selection fused into the epilogue.
//
// Replaces, for Q >= 8 queries at a time, the O(Q*N*D) loop that the reference delegates to Qdrant behind
// QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) - there one gRPC call per query.
//
// Orientation: D[128 queries x 128 corpus rows] += A[128 x K] * B[128 x K]^T, bf16 inputs, fp32 accumulate.
//   * A (the queries of this CTA, unit-norm, rounded to bf16) is loaded ONCE into tensor memory: 128 lanes (one per
//     query) x K/2 32-bit columns (two bf16 per column) - 384 of the 512 TMEM columns for K = 768.  The MMA reads A
//     from TMEM ("TS" form), so shared memory is left entirely to the corpus pipeline.
//   * B = 128 corpus rows x 64 K-elements per stage (16 KB, K-major, 128-byte swizzle), streamed with
//     cp.async.bulk.tensor.2d (TMA, SASS UTMALDG) into a 12-stage mbarrier ring: ~192 KB in flight per SM.
//   * D lives in the remaining 128 TMEM columns.  The epilogue warps read it with tcgen05.ld; thread = query, so each
//     thread walks the 128 scores of "its" query, multiplies by 1/||row|| (cosine) and keeps a sorted top-32 in
//     registers (a score is looked at again only if it beats the thread's 32nd best).
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocation,
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
// More than 128 queries: G = ceil(Q/128) CTAs ("a pair") walk the same tile sequence for different query groups; the
// second reader of a tile hits the 126 MB L2, so HBM still sees every row once.
//
// Exactness: lists hold 32 keys per (CTA, query); the finalize kernel rescoring proves the result against the bound
// max(k'-th kept fast score, largest dropped score) + eps and the host repeats flagged queries on the K1 path.
// Algorithmic bytes per launch = rows x row_bytes; FLOPs = 2 * Q * rows * K.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lvs {

constexpr int kGemmThreads = 192;
constexpr int kGemmM = 128;            // queries per CTA (TMEM lanes)
constexpr int kGemmN = 128;            // corpus rows per tile (accumulator columns)
constexpr int kGemmKC = 64;            // K elements per pipeline stage (= one 128-byte swizzle row)
constexpr int kGemmStageBytes = kGemmN * kGemmKC * 2;   // 16 KB
constexpr int kGemmMaxStages = 12;
constexpr int kGemmList = 32;          // keys kept per (CTA, query)
constexpr int kGemmMaxKChunks = 12;    // A occupies 32 columns per chunk: 12 * 32 + 128 (D) = 512 TMEM columns
constexpr uint32_t kGemmDCol = 384;    // first accumulator column

struct GemmParams {
    const __nv_bfloat16* qb16;   // [n_groups * 128][k_pad] unit queries rounded to bf16, zero padded
    uint32_t k_pad;              // n_kchunks * 64
    uint32_t n_kchunks;
    uint32_t n_rows;
    uint32_t n_tiles;            // ceil(n_rows / 128)
    uint32_t n_groups;           // G: CTAs per pair
    uint32_t n_pairs;            // P: lists per query
    uint32_t n_stages;
    const float* inv_norm;       // [rows] 1/||row||, or nullptr (dot metric)
    const uint8_t* live;
    uint64_t* out_keys;          // [n_groups * 128][P][32]
    uint64_t* out_tops;          // [n_groups * 128][P]  best key of the list
    uint64_t* out_drops;         // [n_groups * 128][P]  32nd key when the list is full (bound on what was dropped), else 0
    float* dbg;                  // optional: CTA 0 dumps its first accumulator tile [128 x 128] (already scaled)
};

__host__ __device__ inline size_t gemm_smem_bytes(uint32_t n_stages) {
    return 1024 /* alignment slack */ + (size_t)n_stages * kGemmStageBytes + 2 * kGemmN * 4 + (2 * kGemmMaxStages + 4) * 8 + 16;
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

__global__ void __launch_bounds__(kGemmThreads, 1) gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap, const GemmParams p) {
    extern __shared__ uint8_t gsm_raw[];
    uint8_t* gsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t S = p.n_stages;
    uint8_t* stages = gsm;                                                   // S x 16 KB, 1024-byte aligned
    float* inv_sm = reinterpret_cast<float*>(gsm + (size_t)S * kGemmStageBytes);          // [2][128]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(inv_sm + 2 * kGemmN);
    uint64_t* empty_bar = full_bar + kGemmMaxStages;
    uint64_t* tmem_full = empty_bar + kGemmMaxStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t group = blockIdx.x % p.n_groups;
    const uint32_t pair = blockIdx.x / p.n_groups;
    const uint32_t nk = p.n_kchunks;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // ---- queries of this CTA -> TMEM (A operand): thread = query (lane of its quarter), 8 columns per store ----
    if (warp >= 2) {
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
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                const uint32_t tile = pair + lt * p.n_pairs;
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
            // instruction descriptor: fp32 accumulate, bf16 x bf16, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGemmN >> 3) << 17) | ((uint32_t)(kGemmM >> 4) << 24);
            const uint32_t tmem_d = tmem_base + kGemmDCol;
            uint32_t it = 0;
            for (uint32_t lt = 0; lt < my_tiles; ++lt) {
                mbar_wait(tmem_empty, (lt & 1u) ^ 1u);          // the epilogue has drained the accumulator
                tc_fence_after();
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % S;
                    mbar_wait(&full_bar[s], (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(stages + (size_t)s * kGemmStageBytes);
#pragma unroll
                    for (uint32_t k = 0; k < kGemmKC / 16; ++k) {
                        const uint64_t desc_b = make_b_desc(b_addr + k * 32u);
                        tc_mma_ts(tmem_d, tmem_base + kc * 32u + k * 8u, desc_b, idesc, (kc | k) != 0u ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[s]);                    // frees the stage when these MMAs have read it
                }
                tc_commit(tmem_full);                            // accumulator of this tile is complete
            }
        }
    } else {
        // ================================ epilogue: thread = query ================================
        const uint32_t lq = warp & 3;
        const uint32_t et = lq * 32 + lane;                      // 0..127: TMEM lane == query within the group
        float ls[kGemmList];
        uint32_t lr[kGemmList];
#pragma unroll
        for (int j = 0; j < kGemmList; ++j) { ls[j] = -INFINITY; lr[j] = 0xFFFFFFFFu; }
        float thr = -INFINITY;
        const uint32_t taddr = tmem_base + ((lq * 32u) << 16) + kGemmDCol;
        for (uint32_t lt = 0; lt < my_tiles; ++lt) {
            const uint32_t tile = pair + lt * p.n_pairs;
            const uint32_t row0 = tile * kGemmN;
            float* inv = inv_sm + (lt & 1u) * kGemmN;
            {
                const uint32_t r = row0 + et;
                inv[et] = (r < p.n_rows) ? (p.inv_norm ? p.inv_norm[r] : 1.0f) : 0.0f;
            }
            named_bar_sync(2, 128);
            mbar_wait(tmem_full, lt & 1u);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t cb = 0; cb < kGemmN / 32; ++cb) {
                uint32_t v[32];
                tc_ld32(taddr + cb * 32u, v);
                tc_wait_ld();
                if (cb == kGemmN / 32 - 1) {
                    // all of this tile's accumulator is in registers: let the MMA warp start the next tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const uint32_t col = cb * 32u + c;
                    const float sc = __uint_as_float(v[c]) * inv[col];
                    if (p.dbg != nullptr && blockIdx.x == 0 && lt == 0) p.dbg[(size_t)et * kGemmN + col] = sc;
                    const uint32_t row = row0 + col;
                    bool pass = (row < p.n_rows) && (sc > thr);
                    if (__any_sync(0xFFFFFFFFu, pass)) {
                        if (pass) pass = p.live[row] != 0;
                        if (pass) {
                            float cs = sc; uint32_t cr = row;
#pragma unroll
                            for (int j = 0; j < kGemmList; ++j) {
                                if (cs > ls[j]) { const float ts = ls[j]; const uint32_t tr = lr[j]; ls[j] = cs; lr[j] = cr; cs = ts; cr = tr; }
                            }
                            thr = ls[kGemmList - 1];
                        }
                    }
                }
            }
        }
        // ---- write this (CTA, query) list ----
        const size_t q = (size_t)group * kGemmM + et;
        uint64_t* dst = p.out_keys + (q * p.n_pairs + pair) * kGemmList;
#pragma unroll
        for (int j = 0; j < kGemmList; ++j) dst[j] = lr[j] != 0xFFFFFFFFu ? make_key(ls[j], lr[j]) : 0ull;
        p.out_tops[q * p.n_pairs + pair] = lr[0] != 0xFFFFFFFFu ? make_key(ls[0], lr[0]) : 0ull;
        p.out_drops[q * p.n_pairs + pair] = lr[kGemmList - 1] != 0xFFFFFFFFu ? make_key(ls[kGemmList - 1], lr[kGemmList - 1]) : 0ull;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace lvs
