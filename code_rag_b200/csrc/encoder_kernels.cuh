// The non-GEMM steps of the code encoder: what transformers' RobertaModel does around its dense layers when
// the reference embeds a chunk (reference src/lattice/providers/unixcoder_provider.py:137-155), as hand-written kernels.
//   embed_ln_kernel   word + position (pad + running count of non-pad tokens) + token-type embeddings -> LayerNorm -> bf16
//   attention_kernel  softmax(Q K^T / sqrt(d) + key mask) V per (sequence, head): flash-style online softmax over key blocks, both
//                     products on mma.sync m16n8k16 (bf16 in, fp32 accumulate); head size 64
//   add_ln_kernel     LayerNorm of the fp32 (dense + bias + residual) rows -> bf16
//   pool_kernel       masked mean over the non-pad tokens -> fp32 sentence embedding (what UniXcoder.forward returns second)
#pragma once
#include "common.cuh"

namespace lvs {

// ---------------------------------------------------------------------------------------------------------
// embeddings + LayerNorm.  grid = (sequences, blocks of 64 tokens), 256 threads: positions by a block scan over the sequence's mask
// (repeated by every CTA of the sequence: a few hundred integers), then a warp per token of the CTA's block.
// ---------------------------------------------------------------------------------------------------------
struct EmbedParams {
    const int32_t* ids;        // [B][L]
    int B, L, H, pad_id, max_pos, vocab;
    const float* word; const float* pos; const float* type0;   // [vocab][H], [max_pos][H], [H]
    const float* ln_w; const float* ln_b; float eps;
    __nv_bfloat16* out;        // [B*L][H]
    int32_t* error;            // set to 1 on an id / position out of range
};

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

constexpr int kLnMaxPerLane = 32;     // hidden <= 1024

__global__ void __launch_bounds__(256) embed_ln_kernel(const EmbedParams p) {
    extern __shared__ int32_t esm[];   // [L] position ids
    __shared__ int32_t s_carry;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* ids = p.ids + (size_t)b * p.L;
    // positions: pad_id + (number of non-pad tokens up to and including this one), pad_id for pad tokens
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < p.L; base += 256) {
        const int i = base + tid;
        const int m = (i < p.L && ids[i] != p.pad_id) ? 1 : 0;
        // inclusive scan of m over the 256 threads
        int v = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += u; }
        __shared__ int32_t wsum[8];
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        int add = s_carry;
        for (int w = 0; w < warp; ++w) add += wsum[w];
        if (i < p.L) esm[i] = m ? p.pad_id + add + v : p.pad_id;
        __syncthreads();
        if (tid == 255) s_carry = add + v;
        __syncthreads();
    }
    const int nper = (p.H + 31) / 32;
    for (int t = blockIdx.y * 64 + warp; t < min(p.L, (int)(blockIdx.y + 1) * 64); t += 8) {
        const int id = ids[t], pos = esm[t];
        if (id < 0 || id >= p.vocab || pos < 0 || pos >= p.max_pos) { if (lane == 0) *p.error = 1; continue; }
        const float* w = p.word + (size_t)id * p.H;
        const float* q = p.pos + (size_t)pos * p.H;
        float x[kLnMaxPerLane];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kLnMaxPerLane; ++j) {
            const int c = j * 32 + lane;
            x[j] = (j < nper && c < p.H) ? w[c] + q[c] + p.type0[c] : 0.f;
            s += x[j];
        }
        const float mu = warp_sum_f32(s) / (float)p.H;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < kLnMaxPerLane; ++j) { const int c = j * 32 + lane; if (j < nper && c < p.H) { const float d = x[j] - mu; var += d * d; } }
        const float rstd = rsqrtf(warp_sum_f32(var) / (float)p.H + p.eps);
        __nv_bfloat16* o = p.out + ((size_t)b * p.L + t) * p.H;
#pragma unroll
        for (int j = 0; j < kLnMaxPerLane; ++j) {
            const int c = j * 32 + lane;
            if (j < nper && c < p.H) o[c] = __float2bfloat16((x[j] - mu) * rstd * p.ln_w[c] + p.ln_b[c]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over fp32 rows (the dense layer's epilogue has already added bias and residual) -> bf16.  One warp per row.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_ln_kernel(const float* in, int M, int H, const float* ln_w, const float* ln_b, float eps,
                                                     __nv_bfloat16* out) {
    // H is a multiple of 64 (checked at create): a lane owns float4 chunks lane, lane + 32, ... of the row (16-byte loads, 8-byte stores)
    const int lane = threadIdx.x & 31;
    const int nvec = H / 4;
    constexpr int kMaxVec = 8;                 // hidden <= 1024
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < M; r += gridDim.x * 8) {
        const float4* x_ = reinterpret_cast<const float4*>(in + (size_t)r * H);
        float4 x[kMaxVec];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int c = j * 32 + lane;
            x[j] = c < nvec ? x_[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
        }
        const float mu = warp_sum_f32(s) / (float)H;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            if (j * 32 + lane < nvec) {
                const float a = x[j].x - mu, b = x[j].y - mu, c = x[j].z - mu, d = x[j].w - mu;
                var += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum_f32(var) / (float)H + eps);
        uint2* o = reinterpret_cast<uint2*>(out + (size_t)r * H);
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int c = j * 32 + lane;
            if (c < nvec) {
                const float4 w = reinterpret_cast<const float4*>(ln_w)[c], bb = reinterpret_cast<const float4*>(ln_b)[c];
                __nv_bfloat162 lo = __floats2bfloat162_rn((x[j].x - mu) * rstd * w.x + bb.x, (x[j].y - mu) * rstd * w.y + bb.y);
                __nv_bfloat162 hi = __floats2bfloat162_rn((x[j].z - mu) * rstd * w.z + bb.z, (x[j].w - mu) * rstd * w.w + bb.w);
                o[c] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// masked mean pooling: out[b] = sum_{t: ids[b][t] != pad} x[b][t] / count.  grid = (sequences, column blocks of 256): a thread owns
// one column and walks the tokens eight at a time (the mask in shared memory, eight independent loads in flight).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_kernel(const __nv_bfloat16* x, const int32_t* ids, int L, int H, int pad_id, float* out) {
    const int b = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x;
    __shared__ uint8_t s_m[1024];
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int cnt = 0;
    for (int t = threadIdx.x; t < L; t += 256) { const int m = ids[(size_t)b * L + t] != pad_id ? 1 : 0; if (t < 1024) s_m[t] = (uint8_t)m; cnt += m; }
    cnt = (int)warp_sum_f32((float)cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (c >= H) return;
    const float inv = 1.0f / (float)max(s_cnt, 1);
    const __nv_bfloat16* xp = x + (size_t)b * L * H + c;
    float acc = 0.f;
    int t = 0;
    for (; t + 8 <= L; t += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __bfloat162float(xp[(size_t)(t + u) * H]);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += s_m[t + u] ? v[u] : 0.f;
    }
    for (; t < L; ++t) acc += s_m[t] ? __bfloat162float(xp[(size_t)t * H]) : 0.f;
    out[(size_t)b * H + c] = acc * inv;
}

// ---------------------------------------------------------------------------------------------------------
// attention.  qkv [B*L][3H] bf16 (q | k | v, head h at columns h*64 of each third); ctx [B*L][H] bf16.
// One CTA = (sequence, head, block of 128 queries): 8 warps x 16 queries.  K and V stream through a three-stage ring of 64-key blocks
// (row-major, rows padded to 72 bf16) filled by cp.async two blocks ahead of the products - 57 KB of shared memory, so two CTAs share an
// SM and one's softmax overlaps the other's MMAs; each warp walks the key blocks with an online softmax:
//   S = Q K^T        8 n-tiles x 4 k-steps of mma.sync.m16n8k16 (A = Q fragment held in registers, B = K rows by ldmatrix.x4)
//   P = exp(S - m)   in registers; the accumulator layout of two adjacent n-tiles IS the A-fragment layout of one k-step
//   O += P V         4 k-steps x 8 n-tiles (B = V by ldmatrix.x4.trans: no transposed copy of V is ever made)
// Pad keys get -inf before the softmax (the reference masks them: mask.unsqueeze(1) * mask.unsqueeze(2)); pad query rows are
// dropped by the pooling.  Chunks are ragged: key blocks behind a sequence's last real token are skipped (their probabilities are
// exactly zero) and query blocks that hold only padding write zeros instead of attending.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttnD = 64;
constexpr int kAttnQB = 128;
constexpr int kAttnKPad = 72;          // K / V row stride in bf16 (36 words: conflict-free fragment and ldmatrix loads)

constexpr int kAttnStages = 3;
constexpr int kAttnMaxBlocks = 64;        // 64-key blocks per sequence: L <= 4096
constexpr int kAttnStageElems = 2 * 64 * kAttnKPad;      // K block then V block, bf16 elements
__host__ __device__ inline size_t attention_smem_bytes(int Lp) { return (size_t)kAttnStages * kAttnStageElems * 2 + (size_t)Lp * 4; }

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of this thread's committed groups are still in flight (pending is 0..7)
__device__ __forceinline__ void cp_async_wait_pending(int pending) {
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}

__global__ void __launch_bounds__(256, 2) attention_kernel(const __nv_bfloat16* qkv, const int32_t* ids, int L, int Lp, int H, int n_heads, int pad_id,
                                                        __nv_bfloat16* ctx) {
    extern __shared__ __align__(16) uint8_t asm_[];
    __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(asm_);                     // [3 stages][K 64 x 72 | V 64 x 72]
    float* kmask = reinterpret_cast<float*>(ring + (size_t)kAttnStages * kAttnStageElems);   // [Lp] 0 or -inf
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const size_t row0 = (size_t)b * L;
    const int ld = 3 * H;
    const __nv_bfloat16* Qg = qkv + row0 * ld + (size_t)h * kAttnD;
    const __nv_bfloat16* Kg = Qg + H;
    const __nv_bfloat16* Vg = Qg + 2 * H;
    // the sequence's extent: index of its last non-pad token + 1 (pads may sit anywhere, but everything behind `len` is padding)
    __shared__ int s_len;
    if (tid == 0) s_len = 0;
    __syncthreads();
    {
        int last = 0;
        for (int i = tid; i < L; i += 256) if (ids[row0 + i] != pad_id) last = i + 1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, o));
        if (lane == 0 && last) atomicMax(&s_len, last);
    }
    __syncthreads();
    const int len = s_len;
    if (qb * kAttnQB >= len) {
        // nothing but padding in this query block: finite, deterministic rows for the layers that follow (0 x NaN would poison the
        // next layer's P V product through these rows' V values)
        for (int i = tid; i < kAttnQB * 8; i += 256) {
            const int r = qb * kAttnQB + (i >> 3);
            if (r < L) *reinterpret_cast<uint4*>(ctx + (row0 + r) * H + (size_t)h * kAttnD + (i & 7) * 8) = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    const int n_blocks = (len + 63) / 64;                      // key blocks that hold at least one real token
    // one cp.async group per block of 64 keys (keys beyond L: zero rows, mask -inf); a group is committed even when it is empty so
    // that "all but the newest group have landed" always means "block blk is there"
    auto issue_block = [&](int blk) {
        if (blk < n_blocks) {
            __nv_bfloat16* st = ring + (size_t)(blk % kAttnStages) * kAttnStageElems;
            for (int i = tid; i < 64 * 8; i += 256) {          // 8 x 16-byte chunks per key row
                const int kr = i >> 3, ch = i & 7, key = blk * 64 + kr;
                __nv_bfloat16* kd = st + (size_t)kr * kAttnKPad + ch * 8;
                __nv_bfloat16* vd = kd + 64 * kAttnKPad;
                if (key < L) {
                    cp_async_16(kd, Kg + (size_t)key * ld + ch * 8);
                    cp_async_16(vd, Vg + (size_t)key * ld + ch * 8);
                } else {
                    *reinterpret_cast<uint4*>(kd) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(vd) = make_uint4(0, 0, 0, 0);
                }
            }
        }
        cp_async_commit();
    };
    issue_block(0);
    issue_block(1);
    // 0 / -inf per key, and per 64-key block whether it holds a pad at all (most blocks do not: they skip the mask)
    __shared__ int s_blk_pad[kAttnMaxBlocks];
    for (int i = tid; i < kAttnMaxBlocks; i += 256) s_blk_pad[i] = 0;
    __syncthreads();
    for (int i = tid; i < Lp; i += 256) {
        const bool real = i < L && ids[row0 + i] != pad_id;
        kmask[i] = real ? 0.f : -INFINITY;
        if (!real) s_blk_pad[i >> 6] = 1;
    }

    const int q0 = qb * kAttnQB + warp * 16;                   // this warp's 16 query rows
    const bool active = q0 < L;
    const int g = lane >> 2, tq = lane & 3;                    // fragment coordinates: row g (and g + 8), column pair tq
    // ---- Q fragments: 4 k-steps (d = 16 each) x 4 registers, straight from global ----
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int r0 = q0 + g, r1 = q0 + g + 8;
        const int c = ks * 16 + tq * 2;
        qa[ks][0] = (active && r0 < L) ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r0 * ld + c) : 0u;
        qa[ks][1] = (active && r1 < L) ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r1 * ld + c) : 0u;
        qa[ks][2] = (active && r0 < L) ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r0 * ld + c + 8) : 0u;
        qa[ks][3] = (active && r1 < L) ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r1 * ld + c + 8) : 0u;
    }
    const float scale = 0.125f * 1.4426950408889634f;          // 1 / sqrt(64), times log2(e): the softmax runs on ex2
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY;                      // running maxima of rows g and g + 8 (scaled, log2 domain)
    float lsum[4] = {0.f, 0.f, 0.f, 0.f};                      // [0] / [2] of the lanes with tq == 0: running sums of rows g / g + 8
    const uint32_t ones_b = g == 0 ? 0x3F803F80u : 0u;         // B fragment of the ones column: B[k][n] = (n == 0)
    // ldmatrix lane roles: lane supplies the address of row (lane & 7) of matrix (lane >> 3)
    const int lm_row = lane & 7, lm_mat = lane >> 3;

    for (int blk = 0; blk < n_blocks; ++blk) {
        const int kb = blk * 64;
        cp_async_wait_pending(1);                              // this thread's copies of block blk have landed (blk + 1 may be in flight) ...
        __syncthreads();                                       // ... and everybody else's; and every warp is done with block blk - 1
        issue_block(blk + 2);                                  // into the stage block blk - 1 has just vacated
        if (!active) continue;
        const __nv_bfloat16* Ks = ring + (size_t)(blk % kAttnStages) * kAttnStageElems;   // this block's 64 keys
        const __nv_bfloat16* Vs = Ks + 64 * kAttnKPad;
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
        // S = Q K^T.  B fragment of (n-tile, k-step): K[key = kb + 8n + g][d = 16 ks + 2 tq (+8)].  One ldmatrix.x4 = the two
        // registers of k-steps ks and ks + 1 for one n-tile: matrices (d 16ks..+7 | +8..15 | 16(ks+1)..+7 | +8..15) of keys 8n..8n+7
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
                uint32_t kf[4];
                ldmatrix_x4(kf, Ks + (size_t)(n * 8 + lm_row) * kAttnKPad + kp * 32 + lm_mat * 8);
                mma_bf16_16816(s[n], qa[kp * 2], kf[0], kf[1]);
                mma_bf16_16816(s[n], qa[kp * 2 + 1], kf[2], kf[3]);
            }
        }
        // mask (only in blocks that hold a pad key) and row maxima, on the RAW scores: the scale is positive, so it commutes with max.
        // Accumulator (n, i): row g (i < 2) or g + 8, key kb + n*8 + tq*2 + (i & 1)
        if (s_blk_pad[blk]) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float k0 = kmask[kb + n * 8 + tq * 2], k1 = kmask[kb + n * 8 + tq * 2 + 1];
                s[n][0] += k0; s[n][1] += k1; s[n][2] += k0; s[n][3] += k1;
            }
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) { bm0 = fmaxf(bm0, fmaxf(s[n][0], s[n][1])); bm1 = fmaxf(bm1, fmaxf(s[n][2], s[n][3])); }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xFFFFFFFFu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xFFFFFFFFu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xFFFFFFFFu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xFFFFFFFFu, bm1, 2));
        const float nm0 = fmaxf(m0, bm0 * scale), nm1 = fmaxf(m1, bm1 * scale);      // running maxima live in the scaled (log2) domain
        // a block of pad keys only leaves the maximum at -inf: keep exp() away from (-inf) - (-inf)
        const float r0 = nm0 == -INFINITY ? 1.f : ex2_approx(m0 - nm0), r1 = nm1 == -INFINITY ? 1.f : ex2_approx(m1 - nm1);
        const float e0 = nm0 == -INFINITY ? 0.f : -nm0, e1 = nm1 == -INFINITY ? 0.f : -nm1;
        uint32_t pa[4][4];                                     // P as A fragments: k-step j covers keys kb + 16 j .. + 15
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            // p = 2^(s * scale - max): one FMA and one MUFU per score (a masked score is -inf: p = 0)
            const float p0 = ex2_approx(fmaf(s[n][0], scale, e0)), p1 = ex2_approx(fmaf(s[n][1], scale, e0));
            const float p2 = ex2_approx(fmaf(s[n][2], scale, e1)), p3 = ex2_approx(fmaf(s[n][3], scale, e1));
            pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);   // rows g:     a0a1 (keys +0..7) / a4a5 (keys +8..15)
            pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);   // rows g + 8: a2a3 / a6a7
        }
        m0 = nm0; m1 = nm1;
#pragma unroll
        for (int n = 0; n < 8; ++n) { o[n][0] *= r0; o[n][1] *= r0; o[n][2] *= r1; o[n][3] *= r1; }
        lsum[0] *= r0; lsum[2] *= r1;
        // the row sums ride on the tensor cores: P times a column of ones (a B fragment that is 1.0 in output column 0, held in
        // registers) accumulates sum_k P[row][k] of the bf16 P the numerator uses, in column 0 of a ninth n-tile - four HMMAs instead
        // of 32 FADDs and the shuffles
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_16816(lsum, pa[j], ones_b, ones_b);
        // O += P V.  B fragment of (k-step j, n-tile): V[key = kb + 16 j + 2 tq (+8)][d = 8 n + g] = the TRANSPOSE of the 8 x 8 tiles of
        // the row-major V block.  One ldmatrix.x4.trans = both registers for n-tiles n and n + 1: matrices
        // (keys +0..7, d 8n | keys +8..15, d 8n | keys +0..7, d 8n+8 | keys +8..15, d 8n+8)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, Vs + (size_t)(j * 16 + (lm_mat & 1) * 8 + lm_row) * kAttnKPad + np * 16 + (lm_mat >> 1) * 8);
                mma_bf16_16816(o[np * 2], pa[j], vf[0], vf[1]);
                mma_bf16_16816(o[np * 2 + 1], pa[j], vf[2], vf[3]);
            }
        }
    }
    if (!active) return;
    const float l0 = __shfl_sync(0xFFFFFFFFu, lsum[0], lane & ~3), l1 = __shfl_sync(0xFFFFFFFFu, lsum[2], lane & ~3);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    __nv_bfloat16* C0 = ctx + (row0 + r0) * H + (size_t)h * kAttnD + tq * 2;
    __nv_bfloat16* C1 = ctx + (row0 + r1) * H + (size_t)h * kAttnD + tq * 2;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (r0 < L) *reinterpret_cast<uint32_t*>(C0 + n * 8) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
        if (r1 < L) *reinterpret_cast<uint32_t*>(C1 + n * 8) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
    }
}

}  // namespace lvs
