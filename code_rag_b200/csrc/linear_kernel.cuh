// Dense layers of the code encoder (embedding the chunks on the GPUs that will search them):
//   Y[M x N] = epilogue(X[M x K] * W[N x K]^T + bias)
// as a tcgen05 GEMM with the epilogue fused: X = bf16 activations (tokens x features), W = a torch.nn.Linear weight ([out, in], so both
// operands are K-major), fp32 accumulators double-buffered in TMEM.  Same machinery as the batched search kernel
// (gemm_topk_kernel.cuh): TMA producer warp (cp.async.bulk.tensor.2d, 128-byte swizzle), one MMA-issuing lane
// (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 256, K = 16 per instruction, tcgen05.commit -> mbarriers), 8 epilogue warps
// (thread = output row = TMEM lane; two warps per lane quarter split the 256 columns).  Two forms (template parameter PAIR):
//   PAIR = false: a persistent CTA walks 128 x 256 output tiles (cta_group::1); every tile pulls 48 KB of operands per 64-wide K chunk.
//   PAIR = true:  a CLUSTER OF TWO CTAs walks 256 x 256 tiles (tcgen05.mma.cta_group::2, M = 256: 128 rows in each CTA's TMEM).  Each
//                 CTA loads its own 128 rows of X and only HALF of the weight tile (128 of the 256 output features); the tensor cores
//                 read the other half from the peer's shared memory.  Operand traffic per CTA and K chunk drops from 48 to 32 KB (the
//                 single-CTA form is bound by the L2 -> SM fabric at ~50-60 % tensor-pipe activity), and 6 stages fit instead of 4.
//                 The leader CTA issues all MMAs; both CTAs' TMA loads count on the leader's barrier; tcgen05.commit is multicast.
// Epilogues (replace what the reference leaves to torch inside transformers' RobertaLayer, reached from
// reference src/lattice/providers/unixcoder_provider.py:137-155):
//   EPI_BIAS       y = acc + b                      -> bf16   (fused Q|K|V projection)
//   EPI_BIAS_GELU  y = gelu_erf(acc + b)            -> bf16   (intermediate dense)
//   EPI_BIAS_RESID y = acc + b + residual(bf16)     -> fp32   (attention / FFN output dense; LayerNorm follows in add_ln_kernel)
// FLOPs per launch = 2 M N K; bytes = (M K + N K) * 2 in, M N * 2 (or 4) out - tensor-bound for every layer shape of RoBERTa-base.
#pragma once
#include <cuda.h>

#include "gemm_topk_kernel.cuh"   // tma_load_2d, tc_* wrappers, make_kmajor_desc

namespace lvs {

constexpr int kLinM = 128;             // output rows (tokens) per tile = TMEM lanes
constexpr int kLinN = 256;             // output columns per tile = accumulator columns per buffer
constexpr int kLinKC = 64;             // K elements per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int kLinStages = 4;          // single-CTA form; the pair form has 32 KB stages and takes 6
constexpr int kLinEpiWarps = 8;
constexpr int kLinThreads = (2 + kLinEpiWarps) * 32;
constexpr int kLinABytes = kLinM * kLinKC * 2;   // 16 KB
constexpr int kLinBBytes = kLinN * kLinKC * 2;   // 32 KB
constexpr int kLinStageBytes = kLinABytes + kLinBBytes;
constexpr int kLinPairStages = 6;
__host__ __device__ constexpr int lin_stage_bytes(bool pair) { return kLinABytes + (pair ? kLinBBytes / 2 : kLinBBytes); }
__host__ __device__ constexpr int lin_stages(bool pair) { return pair ? kLinPairStages : kLinStages; }

enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2 };

struct LinearParams {
    uint32_t M, N, K;              // K a multiple of 64
    uint32_t tiles_m, tiles_n;
    const float* bias;             // [N]
    const __nv_bfloat16* resid;    // [M][N] (EPI_BIAS_RESID)
    void* out;                     // [M][N] bf16, or fp32 for EPI_BIAS_RESID
};

__host__ __device__ constexpr size_t linear_smem_bytes(bool pair = false) {
    return 1024 /* alignment slack */ + (size_t)lin_stages(pair) * lin_stage_bytes(pair) + kLinN * 4 * 2 /* bias, two tiles in flight */ +
           (2 * kLinPairStages + 4) * 8 + 16;
}

// x * Phi(x) with erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 the result is rounded to): one rcp, one
// ex2 and a degree-5 polynomial instead of erff's two branches - the intermediate layer's epilogue is instruction-bound
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float e = fmaf(-poly, ex2_approx(-1.4426950408889634f * z * z), 1.0f);   // erf(|x| / sqrt 2)
    return 0.5f * x * (1.0f + copysignf(e, x));
}

template <int EPI, bool PAIR>
__global__ void __launch_bounds__(kLinThreads, 1) linear_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                                                                const LinearParams p) {
    extern __shared__ __align__(1024) uint8_t lsm_raw[];
    uint8_t* lsm = lsm_raw + ((1024u - (smem_u32(lsm_raw) & 1023u)) & 1023u);
    constexpr int NS = lin_stages(PAIR);
    constexpr int kStageBytes = lin_stage_bytes(PAIR);
    constexpr int kBRows = PAIR ? kLinN / 2 : kLinN;          // rows of the weight tile this CTA loads
    uint8_t* stages = lsm;
    float* bias_sm = reinterpret_cast<float*>(lsm + (size_t)NS * kStageBytes);      // [2][256]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_sm + 2 * kLinN);
    uint64_t* empty_bar = full_bar + kLinPairStages;
    uint64_t* tmem_full = empty_bar + kLinPairStages;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;                 // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t walker = PAIR ? blockIdx.x / 2 : blockIdx.x;          // tile walker: a CTA, or a pair of CTAs
    const uint32_t n_walkers = PAIR ? gridDim.x / 2 : gridDim.x;
    const uint32_t nk = p.K / kLinKC;
    const uint32_t n_tiles = p.tiles_m * p.tiles_n;                      // PAIR: tiles_m counts 256-row blocks

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], (PAIR ? 2 : 1) * kLinEpiWarps); }
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
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // tile t -> (m, n): n fastest, so that the CTAs running at any moment share the same few row blocks of X through L2
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = walker; t < n_tiles; t += n_walkers) {
                const uint32_t tm = t / p.tiles_n, tn = t % p.tiles_n;
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % NS;
                    mbar_wait(&empty_bar[s], ((it / NS) & 1u) ^ 1u);
                    uint8_t* st = stages + (size_t)s * kStageBytes;
                    if constexpr (PAIR) {
                        // the leader's barrier counts the bytes of both CTAs; the data lands in each CTA's own shared memory
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * kStageBytes);
                        tma_load_2d_pair(st, &tmap_x, (int)(kc * kLinKC), (int)(tm * 2 * kLinM + rank * kLinM), &full_bar[s]);
                        tma_load_2d_pair(st + kLinABytes, &tmap_w, (int)(kc * kLinKC), (int)(tn * kLinN + rank * kBRows), &full_bar[s]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                        tma_load_2d(st, &tmap_x, (int)(kc * kLinKC), (int)(tm * kLinM), &full_bar[s]);
                        tma_load_2d(st + kLinABytes, &tmap_w, (int)(kc * kLinKC), (int)(tn * kLinN), &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // instruction descriptor: D = F32 (bit 4), A, B = BF16 (1 at bits 7 and 10), both K-major, N = 256, M = 128 (PAIR: 256 over both CTAs)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kLinN >> 3) << 17) |
                                   ((uint32_t)((PAIR ? 2 * kLinM : kLinM) >> 4) << 24);
            uint32_t it = 0, lt = 0;
            for (uint32_t t = walker; t < n_tiles; t += n_walkers, ++lt) {
                const uint32_t buf = lt & 1u;
                const uint32_t tmem_d = tmem_base + buf * kLinN;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);
                tc_fence_after();
                for (uint32_t kc = 0; kc < nk; ++kc, ++it) {
                    const uint32_t s = it % NS;
                    mbar_wait(&full_bar[s], (it / NS) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stages + (size_t)s * kStageBytes);
                    const uint32_t b_addr = a_addr + kLinABytes;
#pragma unroll
                    for (uint32_t k = 0; k < kLinKC / 16; ++k)
                        tc_mma_ss<PAIR, false>(tmem_d, make_kmajor_desc(a_addr + k * 32u), make_kmajor_desc(b_addr + k * 32u), idesc,
                                               (kc | k) != 0u ? 1u : 0u);
                    tc_commit<PAIR>(&empty_bar[s]);                  // frees the stage (in both CTAs) when these MMAs have read it
                }
                tc_commit<PAIR>(&tmem_full[buf]);                    // the accumulator of this tile is complete (in both CTAs)
            }
        }
    } else {
        // ================================ epilogue: thread = output row ================================
        const uint32_t ew = warp - 2, lq = warp & 3, half = ew >> 2;
        uint32_t lt = 0;
        for (uint32_t t = walker; t < n_tiles; t += n_walkers, ++lt) {
            const uint32_t tm = t / p.tiles_n, tn = t % p.tiles_n;
            const uint32_t buf = lt & 1u;
            // this tile's 256 bias values (the slot of tile lt - 2 is free: its readers passed the barrier below two tiles ago)
            float* bs = bias_sm + buf * kLinN;
            {
                const uint32_t c = (uint32_t)(tid - 64);                                  // 0..255
                const uint32_t col = tn * kLinN + c;
                bs[c] = col < p.N ? p.bias[col] : 0.f;
            }
            named_bar_sync(2, kLinEpiWarps * 32);
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t row = (PAIR ? tm * 2 * kLinM + rank * kLinM : tm * kLinM) + lq * 32 + lane;
            const uint32_t taddr = tmem_base + ((lq * 32u) << 16) + buf * kLinN + half * (kLinN / 2);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                uint32_t v[32];
                tc_ld32(taddr + j * 32u, v);
                tc_wait_ld();
                if (j == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&tmem_empty[buf]); else mbar_arrive(&tmem_empty[buf]); }   // hand the accumulator back
                }
                const uint32_t c0 = half * (kLinN / 2) + j * 32;             // first column of this block within the tile
                const uint32_t col0 = tn * kLinN + c0;
                if (row >= p.M || col0 >= p.N) continue;
                float f[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(v[c]) + bs[c0 + c];
                if (EPI == EPI_BIAS_GELU) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = gelu_erf(f[c]);
                }
                const size_t o = (size_t)row * p.N + col0;
                const uint32_t nvalid = min(32u, p.N - col0);                // N is a multiple of 8 (checked on the host)
                if (EPI == EPI_BIAS_RESID) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.resid + o);
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        if ((uint32_t)c8 * 8 >= nvalid) break;
                        const uint4 r = rp[c8];
                        f[c8 * 8 + 0] += bf16lo(r.x); f[c8 * 8 + 1] += bf16hi(r.x); f[c8 * 8 + 2] += bf16lo(r.y); f[c8 * 8 + 3] += bf16hi(r.y);
                        f[c8 * 8 + 4] += bf16lo(r.z); f[c8 * 8 + 5] += bf16hi(r.z); f[c8 * 8 + 6] += bf16lo(r.w); f[c8 * 8 + 7] += bf16hi(r.w);
                        op[c8 * 2] = make_float4(f[c8 * 8], f[c8 * 8 + 1], f[c8 * 8 + 2], f[c8 * 8 + 3]);
                        op[c8 * 2 + 1] = make_float4(f[c8 * 8 + 4], f[c8 * 8 + 5], f[c8 * 8 + 6], f[c8 * 8 + 7]);
                    }
                } else {
                    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        if ((uint32_t)c8 * 8 >= nvalid) break;
                        uint4 w;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[c8 * 8], f[c8 * 8 + 1]), h1 = __floats2bfloat162_rn(f[c8 * 8 + 2], f[c8 * 8 + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[c8 * 8 + 4], f[c8 * 8 + 5]), h3 = __floats2bfloat162_rn(f[c8 * 8 + 6], f[c8 * 8 + 7]);
                        w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
                        w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
                        op[c8] = w;
                    }
                }
            }
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
