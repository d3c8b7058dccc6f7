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
// embeddings + LayerNorm.  One CTA per sequence (256 threads): positions by a block scan over the mask, then a warp per token.
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
    for (int t = warp; t < p.L; t += 8) {
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
    const int lane = threadIdx.x & 31;
    const int nper = (H + 31) / 32;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < M; r += gridDim.x * 8) {
        const float* x_ = in + (size_t)r * H;
        float x[kLnMaxPerLane];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kLnMaxPerLane; ++j) { const int c = j * 32 + lane; x[j] = (j < nper && c < H) ? x_[c] : 0.f; s += x[j]; }
        const float mu = warp_sum_f32(s) / (float)H;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < kLnMaxPerLane; ++j) { const int c = j * 32 + lane; if (j < nper && c < H) { const float d = x[j] - mu; var += d * d; } }
        const float rstd = rsqrtf(warp_sum_f32(var) / (float)H + eps);
        __nv_bfloat16* o = out + (size_t)r * H;
#pragma unroll
        for (int j = 0; j < kLnMaxPerLane; ++j) { const int c = j * 32 + lane; if (j < nper && c < H) o[c] = __float2bfloat16((x[j] - mu) * rstd * ln_w[c] + ln_b[c]); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// masked mean pooling: out[b] = sum_{t: ids[b][t] != pad} x[b][t] / count.  One CTA per sequence.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_kernel(const __nv_bfloat16* x, const int32_t* ids, int L, int H, int pad_id, float* out) {
    const int b = blockIdx.x;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int cnt = 0;
    for (int t = threadIdx.x; t < L; t += 256) cnt += ids[(size_t)b * L + t] != pad_id ? 1 : 0;
    cnt = (int)warp_sum_f32((float)cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    const float inv = 1.0f / (float)max(s_cnt, 1);
    for (int c = threadIdx.x; c < H; c += 256) {
        float acc = 0.f;
        for (int t = 0; t < L; ++t)
            if (ids[(size_t)b * L + t] != pad_id) acc += __bfloat162float(x[((size_t)b * L + t) * H + c]);
        out[(size_t)b * H + c] = acc * inv;
    }
}

// ---------------------------------------------------------------------------------------------------------
// attention.  qkv [B*L][3H] bf16 (q | k | v, head h at columns h*64 of each third); ctx [B*L][H] bf16.
// One CTA = (sequence, head, block of 128 queries): 8 warps x 16 queries.  K (row-major, padded rows) and V^T (padded rows) of the
// whole sequence sit in shared memory; each warp walks the keys in blocks of 64 with an online softmax:
//   S = Q K^T        8 n-tiles x 4 k-steps of mma.sync.m16n8k16 (A = Q fragment held in registers, B = K rows)
//   P = exp(S - m)   in registers; the accumulator layout of two adjacent n-tiles IS the A-fragment layout of one k-step
//   O += P V         4 k-steps x 8 n-tiles (B = V^T rows)
// Pad keys get -inf before the softmax (the reference masks them: mask.unsqueeze(1) * mask.unsqueeze(2)); pad query rows are
// computed like the others and dropped by the pooling.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttnD = 64;
constexpr int kAttnQB = 128;
constexpr int kAttnKPad = 72;          // K row stride in bf16 (36 words: conflict-free fragment loads)

__host__ __device__ inline int attn_vt_stride(int Lp) { return Lp + 8; }
__host__ __device__ inline size_t attention_smem_bytes(int Lp) {
    return (size_t)Lp * kAttnKPad * 2 + (size_t)kAttnD * attn_vt_stride(Lp) * 2 + (size_t)Lp * 4;
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(256) attention_kernel(const __nv_bfloat16* qkv, const int32_t* ids, int L, int Lp, int H, int n_heads, int pad_id,
                                                        __nv_bfloat16* ctx) {
    extern __shared__ __align__(16) uint8_t asm_[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(asm_);                       // [Lp][72]
    const int vts = attn_vt_stride(Lp);
    __nv_bfloat16* Vt = Ks + (size_t)Lp * kAttnKPad;                                  // [64][Lp + 8]
    float* kmask = reinterpret_cast<float*>(Vt + (size_t)kAttnD * vts);               // [Lp] 0 or -inf
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const size_t row0 = (size_t)b * L;
    const int ld = 3 * H;
    const __nv_bfloat16* Qg = qkv + row0 * ld + (size_t)h * kAttnD;
    const __nv_bfloat16* Kg = Qg + H;
    const __nv_bfloat16* Vg = Qg + 2 * H;
    // ---- K, V^T, mask -> shared (keys beyond L are zero rows with mask -inf) ----
    for (int i = tid; i < Lp * 8; i += 256) {                  // 8 x 16-byte chunks per key row
        const int key = i >> 3, ch = i & 7;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (key < L) {
            kv = *reinterpret_cast<const uint4*>(Kg + (size_t)key * ld + ch * 8);
            vv = *reinterpret_cast<const uint4*>(Vg + (size_t)key * ld + ch * 8);
        }
        *reinterpret_cast<uint4*>(Ks + (size_t)key * kAttnKPad + ch * 8) = kv;
        const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
#pragma unroll
        for (int e = 0; e < 8; ++e) Vt[(size_t)(ch * 8 + e) * vts + key] = ve[e];
    }
    for (int i = tid; i < Lp; i += 256) kmask[i] = (i < L && ids[row0 + i] != pad_id) ? 0.f : -INFINITY;
    __syncthreads();

    const int q0 = qb * kAttnQB + warp * 16;                   // this warp's 16 query rows
    if (q0 >= L) return;
    const int g = lane >> 2, tq = lane & 3;                    // fragment coordinates: row g (and g + 8), column pair tq
    // ---- Q fragments: 4 k-steps (d = 16 each) x 4 registers ----
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int r0 = q0 + g, r1 = q0 + g + 8;
        const int c = ks * 16 + tq * 2;
        qa[ks][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r0 * ld + c) : 0u;
        qa[ks][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r1 * ld + c) : 0u;
        qa[ks][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r0 * ld + c + 8) : 0u;
        qa[ks][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(Qg + (size_t)r1 * ld + c + 8) : 0u;
    }
    const float scale = 0.125f;                                // 1 / sqrt(64)
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max / sum of rows g and g + 8

    for (int kb = 0; kb < Lp; kb += 64) {
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            const __nv_bfloat16* kr = Ks + (size_t)(kb + n * 8 + g) * kAttnKPad + tq * 2;      // B fragment: key = n*8 + g, d pair tq
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
                mma_bf16_16816(s[n], qa[ks], b0, b1);
            }
        }
        // scale + mask; accumulator (n, i): row g (i < 2) or g + 8, key kb + n*8 + tq*2 + (i & 1)
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float k0 = kmask[kb + n * 8 + tq * 2], k1 = kmask[kb + n * 8 + tq * 2 + 1];
            s[n][0] = s[n][0] * scale + k0; s[n][1] = s[n][1] * scale + k1;
            s[n][2] = s[n][2] * scale + k0; s[n][3] = s[n][3] * scale + k1;
            bm0 = fmaxf(bm0, fmaxf(s[n][0], s[n][1])); bm1 = fmaxf(bm1, fmaxf(s[n][2], s[n][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xFFFFFFFFu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xFFFFFFFFu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xFFFFFFFFu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xFFFFFFFFu, bm1, 2));
        const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
        // a block of pad keys only leaves the maximum at -inf: keep exp() away from (-inf) - (-inf)
        const float r0 = nm0 == -INFINITY ? 1.f : __expf(m0 - nm0), r1 = nm1 == -INFINITY ? 1.f : __expf(m1 - nm1);
        const float e0 = nm0 == -INFINITY ? 0.f : nm0, e1 = nm1 == -INFINITY ? 0.f : nm1;
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t pa[4][4];                                     // P as A fragments: k-step j covers keys kb + 16 j .. + 15
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = __expf(s[n][0] - e0), p1 = __expf(s[n][1] - e0), p2 = __expf(s[n][2] - e1), p3 = __expf(s[n][3] - e1);
            ps0 += p0 + p1; ps1 += p2 + p3;
            pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);   // rows g:     a0a1 (keys +0..7) / a4a5 (keys +8..15)
            pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);   // rows g + 8: a2a3 / a6a7
        }
        ps0 += __shfl_xor_sync(0xFFFFFFFFu, ps0, 1); ps0 += __shfl_xor_sync(0xFFFFFFFFu, ps0, 2);
        ps1 += __shfl_xor_sync(0xFFFFFFFFu, ps1, 1); ps1 += __shfl_xor_sync(0xFFFFFFFFu, ps1, 2);
        l0 = l0 * r0 + ps0; l1 = l1 * r1 + ps1; m0 = nm0; m1 = nm1;
#pragma unroll
        for (int n = 0; n < 8; ++n) { o[n][0] *= r0; o[n][1] *= r0; o[n][2] *= r1; o[n][3] *= r1; }
        // O += P V: n-tile = 8 head dims, k-step = 16 keys; B fragment: d = n*8 + g, key pair kb + 16 j + tq*2 (+8)
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const __nv_bfloat16* vr = Vt + (size_t)(n * 8 + g) * vts + kb + tq * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vr + j * 16);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vr + j * 16 + 8);
                mma_bf16_16816(o[n], pa[j], b0, b1);
            }
        }
    }
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
