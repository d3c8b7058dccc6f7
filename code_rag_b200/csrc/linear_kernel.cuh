// Dense layers of the code encoder (embedding the chunks on the GPUs that will search them):
//   Y[M x N] = epilogue(X[M x K] * W[N x K]^T + bias)
// as a tcgen05 GEMM with the epilogue fused: X = bf16 activations (tokens x features), W = a torch.nn.Linear weight ([out, in], so both
// operands are K-major), fp32 accumulators double-buffered in TMEM.  Same machinery as the batched search kernel
// (gemm_topk_kernel.cuh): TMA producer warp (cp.async.bulk.tensor.2d, 128-byte swizzle), one MMA-issuing lane
// (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 256, K = 16 per instruction, tcgen05.commit -> mbarriers), 8 epilogue warps
// (TMEM lane = output row; two warps per lane quarter split the 256 columns; every 32 x 32 block goes through a per-warp
// shared-memory transpose so that global stores and residual loads are coalesced).  Two forms (template parameter PAIR):
//   PAIR = false: a persistent CTA walks 128 x 256 output tiles (cta_group::1); every tile pulls 48 KB of operands per 64-wide K chunk.
//   PAIR = true:  a CLUSTER OF TWO CTAs walks 256 x 256 tiles (tcgen05.mma.cta_group::2, M = 256: 128 rows in each CTA's TMEM).  Each
//                 CTA loads its own 128 rows of X and only HALF of the weight tile (128 of the 256 output features); the tensor cores
//                 read the other half from the peer's shared memory.  Operand traffic per CTA and K chunk drops from 48 to 32 KB (the
//                 single-CTA form is bound by the L2 -> SM fabric at ~50-60 % tensor-pipe activity), and 5 stages fit instead of 3.
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
constexpr int kLinABytes = kLinM * kLinKC * 2;   // 16 KB
constexpr int kLinBBytes = kLinN * kLinKC * 2;   // 32 KB
constexpr int kLinStageBytes = kLinABytes + kLinBBytes;
constexpr int kLinMaxStages = 5;
constexpr int kLinXposeLd = 36;        // floats per row of a warp's 32 x 32 transpose buffer: 16-byte rows, conflict-free 128-bit accesses

enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2 };

// Epilogue warps: 8 (two per TMEM lane quarter, 128 columns each) - or 16 (four per quarter, 64 columns each) for the GELU epilogue,
// whose ~20 instructions per value are latency-bound with two warps per scheduler (ncu: epilogue warps stalled 70 % of the time, 53 %
// tensor activity).  The transpose buffers of 16 warps take 72 KB, so that form runs one pipeline stage less.
__host__ __device__ constexpr int lin_epi_warps(int epi) { return 8 + 0 * epi; }     // 16 for EPI_BIAS_GELU measured the same (156 vs 158 us)
__host__ __device__ constexpr int lin_threads(int epi) { return (2 + lin_epi_warps(epi)) * 32; }
__host__ __device__ constexpr int lin_xpose_bytes(int epi) { return lin_epi_warps(epi) * 32 * kLinXposeLd * 4; }   // 36 / 72 KB
__host__ __device__ constexpr int lin_stage_bytes(bool pair) { return kLinABytes + (pair ? kLinBBytes / 2 : kLinBBytes); }
// pair form: 32 KB stages, 5 of them (4 under the GELU epilogue); single-CTA form: 48 KB stages, 3 (2)
__host__ __device__ constexpr int lin_stages(bool pair, int epi) { return (pair ? 5 : 3) - (lin_epi_warps(epi) > 8 ? 1 : 0); }

struct LinearParams {
    uint32_t M, N, K;              // K a multiple of 64
    uint32_t tiles_m, tiles_n;
    const float* bias;             // [N]
    const __nv_bfloat16* resid;    // [M][N] (EPI_BIAS_RESID)
    void* out;                     // [M][N] bf16, or fp32 for EPI_BIAS_RESID
};

__host__ __device__ constexpr size_t linear_smem_bytes(bool pair, int epi) {
    return 1024 /* alignment slack */ + (size_t)lin_stages(pair, epi) * lin_stage_bytes(pair) + kLinN * 4 * 2 /* bias, two tiles in flight */ +
           lin_xpose_bytes(epi) + (2 * kLinMaxStages + 4) * 8 + 16;
}

// x * Phi(x) with erfc by Abramowitz & Stegun 7.1.28, erfc(z) = (1 + a1 z + ... + a6 z^6)^-16 (|error| <= 3e-7; 7e-7 on the result in
// float32, far below the bf16 it is rounded to): six FMAs, four squarings and ONE reciprocal - erff costs two branches, 7.1.26 a
// reciprocal AND an exponential, and the intermediate layer's epilogue has to fit under the 6144 tensor-core cycles of its tile
// (128 values per thread; two MUFU operations per value alone were 4096 cycles of the pipe per scheduler).
//   gelu(x) = x/2 (1 + sign(x) (1 - erfc(|x|/sqrt 2))) = max(x, 0) - |x|/2 erfc(|x|/sqrt 2)
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float p = fmaf(z, fmaf(z, fmaf(z, fmaf(z, fmaf(z, fmaf(z, 0.0000430638f, 0.0002765672f), 0.0001520143f), 0.0092705272f), 0.0422820123f),
                           0.0705230784f), 1.0f);
    p *= p; p *= p; p *= p; p *= p;                       // overflows to +inf from |x| ~ 25 on: erfc = 1/inf = 0, as it should be
    return fmaf(-fabsf(0.5f * x), rcp_approx(p), fmaxf(x, 0.f));
}

template <int EPI, bool PAIR>
__global__ void __launch_bounds__(lin_threads(EPI), 1) linear_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                                                                const LinearParams p) {
    extern __shared__ __align__(1024) uint8_t lsm_raw[];
    uint8_t* lsm = lsm_raw + ((1024u - (smem_u32(lsm_raw) & 1023u)) & 1023u);
    constexpr int NS = lin_stages(PAIR, EPI);
    constexpr int EW = lin_epi_warps(EPI);                  // epilogue warps
    constexpr int NPART = EW / 4;                           // column parts of the tile (one warp per lane quarter and part)
    constexpr int PCOLS = kLinN / NPART;                    // 128 or 64 columns per warp
    constexpr int kStageBytes = lin_stage_bytes(PAIR);
    constexpr int kBRows = PAIR ? kLinN / 2 : kLinN;          // rows of the weight tile this CTA loads
    uint8_t* stages = lsm;
    float* bias_sm = reinterpret_cast<float*>(lsm + (size_t)NS * kStageBytes);      // [2][256]
    float* xpose = bias_sm + 2 * kLinN;                                             // [8 epilogue warps][32][36]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(xpose + lin_xpose_bytes(EPI) / 4);
    uint64_t* empty_bar = full_bar + kLinMaxStages;
    uint64_t* tmem_full = empty_bar + kLinMaxStages;      // [2]
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
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], (PAIR ? 2 : 1) * EW); }
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
        const uint32_t ew = warp - 2, lq = warp & 3, part = ew >> 2;            // TMEM lane quarter = warp % 4 (hardware rule), column part
        uint32_t lt = 0;
        for (uint32_t t = walker; t < n_tiles; t += n_walkers, ++lt) {
            const uint32_t tm = t / p.tiles_n, tn = t % p.tiles_n;
            const uint32_t buf = lt & 1u;
            // this tile's 256 bias values (the slot of tile lt - 2 is free: its readers passed the barrier below two tiles ago)
            float* bs = bias_sm + buf * kLinN;
            {
                const uint32_t c = (uint32_t)(tid - 64);                                  // 0..255 (0..511 with 16 warps)
                const uint32_t col = tn * kLinN + c;
                if (c < (uint32_t)kLinN) bs[c] = col < p.N ? p.bias[col] : 0.f;
            }
            // the residual of this warp's part of the tile (bf16, in the layout the blocks are stored in - see below) is requested
            // NOW, while the tensor cores still work on the tile: it comes from DRAM (the layer input left L2 two GEMMs ago), and
            // four dependent round trips per tile were the whole epilogue (ncu: 67 % of the samples on the first use of the values)
            uint2 rres[PCOLS / 32][8];
            if (EPI == EPI_BIAS_RESID) {
                const uint32_t row_p = (PAIR ? tm * 2 * kLinM + rank * kLinM : tm * kLinM) + lq * 32 + (lane >> 3);
                const uint32_t col_p = tn * kLinN + part * PCOLS + (lane & 7) * 4;
#pragma unroll
                for (int j = 0; j < PCOLS / 32; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t grow = row_p + 4 * i, gcol = col_p + j * 32;
                        rres[j][i] = (grow < p.M && gcol < p.N) ? __ldg(reinterpret_cast<const uint2*>(p.resid + (size_t)grow * p.N + gcol)) : make_uint2(0u, 0u);
                    }
            }
            named_bar_sync(2, EW * 32);
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((lq * 32u) << 16) + buf * kLinN + part * PCOLS;
            // One 32-column block.  A thread holds 32 columns of ITS row (TMEM lane = row): storing them as they are would make every
            // store instruction touch 32 different rows, 16 bytes each - the memory pipe, not the tensor cores, then sets the pace
            // (ncu, attention-output layer: 20 % tensor activity).  So the block is transposed through the warp's 32 x 32 buffer:
            // raw accumulators in; bias (a thread's columns are the same for every row it stores: one read per block), GELU,
            // residual and conversion on the way out; every global access is 64 or 128 contiguous bytes per row.
            float* xb = xpose + (size_t)ew * 32 * kLinXposeLd;
            const uint32_t row_w0 = (PAIR ? tm * 2 * kLinM + rank * kLinM : tm * kLinM) + lq * 32;     // first row of this warp
            auto finish_block = [&](const uint32_t (&v)[32], int j) {
                const uint32_t c0 = part * PCOLS + j * 32;                   // first column of this block within the tile
                const uint32_t col0 = tn * kLinN + c0;
                if (row_w0 >= p.M || col0 >= p.N) return;                    // warp-uniform
                __syncwarp();                                                // the previous block's readers are done with xb
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)                               // the raw accumulators, as they come out of TMEM
                    *reinterpret_cast<uint4*>(xb + lane * kLinXposeLd + c4 * 4) = make_uint4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
                __syncwarp();
                const uint32_t nvalid = min(32u, p.N - col0);                // N is a multiple of 8 (checked on the host)
                if (EPI == EPI_BIAS_RESID) {
                    // fp32 out: 8 lanes x 16 bytes = one row's 128 bytes, 4 rows per instruction; the bf16 residual the same way
                    const uint32_t cq = (lane & 7) * 4;
                    const float4 bq = *reinterpret_cast<const float4*>(bs + c0 + cq);       // this thread's four columns, every row
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t r = (lane >> 3) + 4 * i, grow = row_w0 + r;
                        if (grow >= p.M || cq >= nvalid) continue;
                        float4 f = *reinterpret_cast<const float4*>(xb + r * kLinXposeLd + cq);
                        const size_t o = (size_t)grow * p.N + col0 + cq;
                        const uint2 rr = rres[j][i];
                        f.x += bq.x + bf16lo(rr.x); f.y += bq.y + bf16hi(rr.x); f.z += bq.z + bf16lo(rr.y); f.w += bq.w + bf16hi(rr.y);
                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) = f;
                    }
                } else {
                    // bf16 out: 4 lanes x 16 bytes = one row's 64 bytes, 8 rows per instruction
                    const uint32_t c8 = (lane & 3) * 8;
                    const float4 ba = *reinterpret_cast<const float4*>(bs + c0 + c8), bb = *reinterpret_cast<const float4*>(bs + c0 + c8 + 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t r = (lane >> 2) + 8 * i, grow = row_w0 + r;
                        if (grow >= p.M || c8 >= nvalid) continue;
                        float4 a = *reinterpret_cast<const float4*>(xb + r * kLinXposeLd + c8);
                        float4 b = *reinterpret_cast<const float4*>(xb + r * kLinXposeLd + c8 + 4);
                        a.x += ba.x; a.y += ba.y; a.z += ba.z; a.w += ba.w; b.x += bb.x; b.y += bb.y; b.z += bb.z; b.w += bb.w;
                        if (EPI == EPI_BIAS_GELU) {
                            a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w);
                            b.x = gelu_erf(b.x); b.y = gelu_erf(b.y); b.z = gelu_erf(b.z); b.w = gelu_erf(b.w);
                        }
                        uint4 w;
                        w.x = pack_bf16(a.x, a.y); w.y = pack_bf16(a.z, a.w); w.z = pack_bf16(b.x, b.y); w.w = pack_bf16(b.z, b.w);
                        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)grow * p.N + col0 + c8) = w;
                    }
                }
            };
            // two blocks per TMEM round trip (one wait for both loads); the accumulator goes back to the MMA warp as soon as this
            // warp's last load has landed, before the arithmetic on it
#pragma unroll
            for (int jj = 0; jj < PCOLS / 64; ++jj) {
                uint32_t v0[32], v1[32];
                tc_ld32(taddr + (2 * jj) * 32u, v0);
                tc_ld32(taddr + (2 * jj + 1) * 32u, v1);
                tc_wait_ld();
                if (jj == PCOLS / 64 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&tmem_empty[buf]); else mbar_arrive(&tmem_empty[buf]); }   // hand the accumulator back
                }
                finish_block(v0, 2 * jj);
                finish_block(v1, 2 * jj + 1);
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
