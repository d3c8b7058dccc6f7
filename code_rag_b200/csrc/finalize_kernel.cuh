// Candidate selection + EXACT rescoring + final ordering (one CTA per query).
//
// The scan (K1) and the tensor-core path (K2) rank rows by an fp32 score.  qdrant-client local mode - the
// engine behind the reference's QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) -
// scores in float64 on a float32 matrix that it re-normalises IN PLACE on every cosine search
// (oracle/qdrant_local.py, point 2).  To return bit-identical id lists this kernel
//   1. selects the k' best fast keys out of the per-CTA lists (threshold = k'-th largest list maximum; when too many keys
//      survive it, an exact MSB-first radix select over the keys), then ranks the survivors in shared memory,
//   2. recomputes each candidate's score the way local mode does: float32 row (for bf16 storage:
//      float32(row / ||row||_f64), i.e. what local mode would have stored), the per-search float32 in-place
//      re-normalisation replayed `searches since the row was written` times (numpy's pair-wise float32 row
//      norm is reproduced add for add; the chain is cut at its fixed point or 2-cycle), then a float64 dot
//      with the float64 unit query,
//   3. orders the candidates by (score desc, tie key asc, row asc) and emits the first k,
//   4. PROVES exactness: every row outside the candidate set has fast score <= t (the k'-th fast score) and
//      hence exact score <= t + eps; if the k-th exact score is not above that bound the query is flagged and
//      the host repeats it with a larger k'.
#pragma once
#include "common.cuh"

namespace lvs {

constexpr int kFinThreads = 256;
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kMaxCand = 256;          // k' upper bound
constexpr int kSortCap = 2048;         // survivors that can be sorted in shared memory
constexpr int kPwMaxLeaves = 128;      // numpy pair-wise blocks for dim <= 8192


// numpy's pairwise_sum(a, n) unrolled into leaves (n <= 128: 8 interleaved accumulators) and in-place combines.
struct PwProgram {
    int n_leaves;
    int n_steps;
    int regular;               // every leaf is a non-zero multiple of 8 long, at most 8 leaves: the register form of np_norm_f32 applies
    int balanced;              // ... and the combine steps form a perfect binary tree over the leaves (a power of two of them)
    uint16_t leaf_off[kPwMaxLeaves];
    uint16_t leaf_len[kPwMaxLeaves];
    uint8_t step_dst[kPwMaxLeaves];
    uint8_t step_src[kPwMaxLeaves];
};

struct FinalizeParams {
    const uint64_t* keys;      // [Q][M]   fast keys (0 = empty)
    const uint64_t* tops;      // [Q][L]   best key of each list
    const uint64_t* drops;     // [Q][L]   K2 only: bound on what a (full) list dropped, 0 = nothing; nullptr for K1 lists
    uint32_t M;
    uint32_t L;
    uint32_t kp;               // k' candidates to rescore (<= kMaxCand)
    uint32_t k;                // results per query (<= kp)
    const uint8_t* base;       // shard
    uint32_t row_bytes;
    int dim;
    int dim_pad;               // floats per chain buffer (multiple of 4)
    int storage;
    int metric;
    const double* q64;         // [Q][dim] unit (cosine) or raw (dot) queries
    const uint64_t* tiekey;    // [rows]
    const uint64_t* epoch;     // [rows] value of the collection's search counter when the row was written
    uint64_t search_no;        // counter value of query 0 of this launch (query qi is search_no + qi); 64-bit: never wraps
    const PwProgram* pw;
    float eps;                 // bound on |fast score - exact score| for unit vectors
    const float* eps_q;        // [Q] per-query bound (K2: bf16 rounding of the query), nullptr => eps
    float eps_add;             // added to eps_q (K2 on unit-norm shards: the norm deviation it ignores)
    const float* qnorm;        // [Q] ||q||      } dot metric: the bound scales with ||q|| * max ||row||   (nullptr: the fused scan kernel passes its own)
    const float* max_norm;     // [1] max ||row|| }
    int64_t row_base;          // global row of local row 0
    int n_rescore_warps;       // warps per CTA that own chain buffers
    double* cand_scores;       // [Q][kMaxCand] scratch: exact scores exchanged between the CTAs of a query
    uint32_t* tickets;         // [Q] zero-initialised arrival counters (reset by the last CTA)
    double* out_scores;        // [Q][k]
    int64_t* out_rows;         // [Q][k]   -1 padded
    uint64_t* out_ties;        // [Q][k]
    int32_t* out_flags;        // [Q]      bit0: exactness not proven
    uint32_t* out_counts;      // [Q]      number of valid results
    unsigned long long* dbg_times;  // optional phase stamps (see stamp() in common.cuh); nullptr in normal operation
};

__host__ __device__ inline size_t finalize_smem_bytes(int dim_pad, int n_rescore_warps) {
    size_t b = 0;
    b += (size_t)kSortCap * 8;                    // merged warp lists / sort buffer
    b += 64 + 1024 + ((sizeof(PwProgram) + 15) & ~(size_t)15);   // scalars + radix histogram + pair-wise program
    b += (size_t)kMaxCand * (8 + 8 + 8 + 4 + 4);  // cand key, score, tie, row, rank
    b = (b + 15) & ~(size_t)15;
    b += (size_t)dim_pad * 8;                     // float64 query
    b += (size_t)n_rescore_warps * ((size_t)3 * dim_pad * 4 + kPwMaxLeaves * 4);
    return b;
}

// byte offset of the float64 query inside the finalize carve-up (the fused scan kernel stages the query there itself)
__host__ __device__ inline size_t finalize_query_offset() {
    size_t b = (size_t)kSortCap * 8 + 64 + 1024 + ((sizeof(PwProgram) + 15) & ~(size_t)15) + (size_t)kMaxCand * (8 + 8 + 8 + 4 + 4);
    return (b + 15) & ~(size_t)15;
}

// ||v||_2 in float32 exactly as numpy computes np.linalg.norm(M, axis=-1) for a float32 row:
// s_i = fl(v_i*v_i); pairwise_sum(s); sqrt.  Warp-cooperative; all lanes return the same value.
__device__ __forceinline__ float np_norm_f32(const float* v, const PwProgram* pw, float* leaf_out, int lane) {
    const int nl = pw->n_leaves;
    if (pw->regular) {
        // Register form (dim 384 / 768 / 1024 ...): 4 lanes per leaf, each with two ADJACENT accumulators of numpy's eight
        // (r[2j], r[2j+1]; one 64-bit load per step), all <= 8 leaves at once.  numpy's combine ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
        // is the in-lane add followed by two xor-shuffles; a perfect tree over the leaves is three more.  Starting the
        // accumulators at +0 instead of at the first square changes nothing (squares are never -0).
        const int li = lane >> 2;
        const bool ok = li < nl;
        const int off = ok ? pw->leaf_off[li] : 0;
        const int len = ok ? pw->leaf_len[li] : 0;
        const float2* vp = reinterpret_cast<const float2*>(v + off + (lane & 3) * 2);
        float r0 = 0.f, r1 = 0.f;
#pragma unroll 4
        for (int i = 0; i < len; i += 8) {
            const float2 x = vp[i >> 1];
            r0 = __fadd_rn(r0, __fmul_rn(x.x, x.x));
            r1 = __fadd_rn(r1, __fmul_rn(x.y, x.y));
        }
        float r = __fadd_rn(r0, r1);
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 2));
        if (pw->balanced) {
            for (int m = 4; m < 4 * nl; m <<= 1) r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, m));
            return __fsqrt_rn(__shfl_sync(0xFFFFFFFFu, r, 0));
        }
        if (ok && (lane & 3) == 0) leaf_out[li] = r;
    } else {
        for (int base = 0; base < nl; base += 4) {
            const int li = base + (lane >> 3);
            const int j = lane & 7;
            const bool ok = li < nl;
            const int off = ok ? pw->leaf_off[li] : 0;
            const int len = ok ? pw->leaf_len[li] : 0;
            float r = 0.f;
            if (len >= 8) {
                const int lim = len - (len & 7);
                float x = v[off + j];
                r = __fmul_rn(x, x);
                for (int i = 8; i < lim; i += 8) { x = v[off + i + j]; r = __fadd_rn(r, __fmul_rn(x, x)); }
            }
            r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
            r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 2));
            r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 4));
            if (len >= 8) {
                for (int i = len - (len & 7); i < len; ++i) { const float x = v[off + i]; r = __fadd_rn(r, __fmul_rn(x, x)); }
            } else {
                r = 0.f;
                for (int i = 0; i < len; ++i) { const float x = v[off + i]; r = __fadd_rn(r, __fmul_rn(x, x)); }
            }
            if (ok && j == 0) leaf_out[li] = r;
        }
    }
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < pw->n_steps; ++s) {
            const int d = pw->step_dst[s];
            leaf_out[d] = __fadd_rn(leaf_out[d], leaf_out[pw->step_src[s]]);
        }
    }
    __syncwarp();
    const float tot = leaf_out[0];
    __syncwarp();
    return __fsqrt_rn(tot);
}

// this lane's share of sum_i v[i] * q[i] in float64: four columns per step (one 128-bit and two 128-bit loads), two accumulators
__device__ __forceinline__ double lane_dot_f64(const float* v, const double* q, int D, int lane) {
    double a0 = 0.0, a1 = 0.0;
    const int D4 = D >> 2;
#pragma unroll 2
    for (int i = lane; i < D4; i += 32) {
        const float4 x = reinterpret_cast<const float4*>(v)[i];
        const double2 qa = reinterpret_cast<const double2*>(q)[2 * i], qb = reinterpret_cast<const double2*>(q)[2 * i + 1];
        a0 = fma((double)x.x, qa.x, a0); a1 = fma((double)x.y, qa.y, a1);
        a0 = fma((double)x.z, qb.x, a0); a1 = fma((double)x.w, qb.y, a1);
    }
    const int i = 4 * D4 + lane;
    if (i < D) a0 = fma((double)v[i], q[i], a0);
    return a0 + a1;
}

// Exact score of one candidate row (warp-cooperative).  buf = 3 * dim_pad floats of shared memory; q = the float64
// query staged in shared memory.
__device__ __forceinline__ double exact_score(const FinalizeParams& p, const PwProgram* pw, uint32_t row, const double* q,
                                              uint64_t search_no, float* buf, float* leaf_out, int lane) {
    const int D = p.dim;
    const uint8_t* rp = p.base + (size_t)row * p.row_bytes;
    const uint32_t cpr = p.row_bytes / 16;
    float* cur = buf;
    float* prev = buf + p.dim_pad;
    float* nxt = buf + 2 * p.dim_pad;
    const uint64_t row_epoch = __ldg(p.epoch + row);    // issued together with the row's loads
    // stage the stored row as float32 (128-bit loads, all chunks of a lane in flight together); the padding
    // columns of a row are zero, so they may be staged and summed as well
    double ss = 0.0;
    if (p.storage == LVS_STORAGE_F32) {
        const uint4* src = reinterpret_cast<const uint4*>(rp);
        for (uint32_t c = lane; c < cpr; c += 32) {
            const uint4 v = __ldg(src + c);
            const float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y), x2 = __uint_as_float(v.z), x3 = __uint_as_float(v.w);
            *reinterpret_cast<float4*>(cur + 4 * c) = make_float4(x0, x1, x2, x3);
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(rp);
        for (uint32_t c = lane; c < cpr; c += 32) {
            const uint4 v = __ldg(src + c);
            const float x[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y), bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
#pragma unroll
            for (int e = 0; e < 8; ++e) { const double t = (double)x[e]; ss = fma(t, t, ss); }
            *reinterpret_cast<float4*>(cur + 8 * c) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(cur + 8 * c + 4) = make_float4(x[4], x[5], x[6], x[7]);
        }
    }
    __syncwarp();
    if (p.metric == LVS_METRIC_DOT) {
        const double acc = lane_dot_f64(cur, q, D, lane);
        __syncwarp();
        return warp_sum_f64(acc);
    }
    if (p.storage == LVS_STORAGE_BF16) {
        // what local mode would have stored for this (bf16-representable) input: float32(x / ||x||_f64)
        ss = warp_sum_f64(ss);
        const double nrm = sqrt(ss);
        if (nrm > 0.0) {
            for (int i = lane; i < D; i += 32) cur[i] = (float)((double)cur[i] / nrm);
        }
        __syncwarp();
    }
    if (lane == 0) stamp(p.dbg_times, 14);                // row staged
    // replay of the per-search in-place re-normalisation
    const uint64_t age = search_no - row_epoch;         // searches run since the row was written, this one included
    const uint32_t count = age > 64ull ? 64u + (uint32_t)((age - 64ull) & 1ull) : (uint32_t)age;
    bool have_prev = false;
    const int n4 = p.dim_pad >> 2;
    for (uint32_t j = 1; j <= count; ++j) {
        const float n = np_norm_f32(cur, pw, leaf_out, lane);
        if (n == 1.0f) break;                            // fixed point
        const float d = (n != 0.0f) ? n : 1.1920929e-7f;
        bool same_cur = true, same_prev = true;
        // four columns per lane and step; the zero padding columns ride along (0 / d = 0, or NaN when d is: "same" either way)
        for (int i = lane; i < n4; i += 32) {
            const float4 x = reinterpret_cast<const float4*>(cur)[i];
            float4 y;
            y.x = __fdiv_rn(x.x, d); y.y = __fdiv_rn(x.y, d); y.z = __fdiv_rn(x.z, d); y.w = __fdiv_rn(x.w, d);
            reinterpret_cast<float4*>(nxt)[i] = y;
            same_cur &= ((y.x == x.x) || (y.x != y.x)) && ((y.y == x.y) || (y.y != y.y)) && ((y.z == x.z) || (y.z != y.z)) && ((y.w == x.w) || (y.w != y.w));
            if (have_prev) {
                const float4 o = reinterpret_cast<const float4*>(prev)[i];
                same_prev &= ((y.x == o.x) || (y.x != y.x)) && ((y.y == o.y) || (y.y != y.y)) && ((y.z == o.z) || (y.z != y.z)) && ((y.w == o.w) || (y.w != y.w));
            }
        }
        same_cur = __all_sync(0xFFFFFFFFu, same_cur);
        same_prev = have_prev && __all_sync(0xFFFFFFFFu, same_prev);
        __syncwarp();
        if (same_cur) break;                             // fixed point with n != 1
        if (same_prev) {                                 // 2-cycle: v_j == v_{j-2}
            if (((count - j) & 1u) == 0u) { float* t = cur; cur = nxt; nxt = t; }
            break;
        }
        float* t = prev; prev = cur; cur = nxt; nxt = t;
        have_prev = true;
    }
    if (lane == 0) stamp(p.dbg_times, 15);                // re-normalisations replayed
    const double acc = lane_dot_f64(cur, q, D, lane);
    __syncwarp();
    return warp_sum_f64(acc);
}

__device__ __forceinline__ void bitonic_sort_desc(uint64_t* buf, int n, int tid) {   // 256 threads, named barrier 1
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += kFinThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = buf[i], b = buf[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { buf[i] = b; buf[ixj] = a; }
                }
            }
            named_bar_sync(1, kFinThreads);
        }
    }
}

// The body runs in two places: as the standalone finalize_kernel<KPL> behind the tensor-core path (grid = (queries, C)), and inside
// the LAST CTAs of the fused scan kernel (scan_kernel.cuh), which then needs no second launch.  (qi, cy, C) = query, this CTA's
// index among the C CTAs that work on the query, C.  Every CTA of a query repeats the (cheap, deterministic) selection, rescoring
// is split over the C CTAs (one candidate per warp), and the last CTA to finish (atomic ticket) orders the candidates.
// 256 threads take part (named barrier 1): in the scan kernel these are the 8 consumer warps, the producer warp has left.
// Returns true in the CTA that holds the final result, left in shared memory (FinResult) for the caller to store or exchange.
__device__ __forceinline__ void fin_sync() { named_bar_sync(1, kFinThreads); }

struct FinResult {
    const double* score;   // [nout]   (shared memory)
    const uint32_t* row;   // [nout]   local rows
    const uint64_t* tie;   // [nout]
    const uint32_t* rank;  // rank[c] of candidate c; candidates with rank < k are the result, in rank order
    uint32_t ncand;
    uint32_t nout;
    int32_t flag;
};

// stage_q(qs): the caller's way of putting the float64 unit query into shared memory (dim values; called by all 256 threads once,
// after the first loads of the selection have been issued, so that its divisions hide their latency)
template <int KPL, class StageQ>
__device__ __forceinline__ bool finalize_body(const FinalizeParams& p, const uint32_t qi, const uint32_t cy, const uint32_t C,
                                              uint8_t* fsm, const float qnorm_in, FinResult& res, StageQ stage_q) {
    constexpr uint32_t KPW = 32u * KPL;
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(fsm);
    uint32_t* scal = reinterpret_cast<uint32_t*>(fsm + (size_t)kSortCap * 8);   // [3]=is_last
    uint32_t* hist = scal + 16;                                                  // [256] radix-select histogram
    PwProgram* pws = reinterpret_cast<PwProgram*>(hist + 256);
    uint8_t* cp = reinterpret_cast<uint8_t*>(pws) + ((sizeof(PwProgram) + 15) & ~(size_t)15);
    uint64_t* ckey = reinterpret_cast<uint64_t*>(cp);  cp += (size_t)kMaxCand * 8;
    double* cscore = reinterpret_cast<double*>(cp);    cp += (size_t)kMaxCand * 8;
    uint64_t* ctie = reinterpret_cast<uint64_t*>(cp);  cp += (size_t)kMaxCand * 8;
    uint32_t* crow = reinterpret_cast<uint32_t*>(cp);  cp += (size_t)kMaxCand * 4;
    uint32_t* crank = reinterpret_cast<uint32_t*>(cp); cp += (size_t)kMaxCand * 4;
    size_t off = (size_t)(cp - fsm);
    off = (off + 15) & ~(size_t)15;
    double* qs = reinterpret_cast<double*>(fsm + off);
    off += (size_t)p.dim_pad * 8;
    float* chain = reinterpret_cast<float*>(fsm + off);
    const size_t per_warp = (size_t)3 * p.dim_pad + kPwMaxLeaves;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t* keys = p.keys + (size_t)qi * p.M;
    const uint32_t kp = p.kp;                            // == KPW

    // stage the pair-wise program (and, below, the float64 query) while the lists stream in
    for (int i = tid; i < (int)(sizeof(PwProgram) / 4); i += kFinThreads)
        reinterpret_cast<uint32_t*>(pws)[i] = reinterpret_cast<const uint32_t*>(p.pw)[i];
    bool q_staged = false;

    // ---- 1a. fast path (L >= kp lists): the kp-th largest LIST MAXIMUM is a threshold with at least kp keys at or
    //          above it (the kp maxima themselves) and, for well-mixed shards, few more; gather those keys.
    bool folded = true;
    uint32_t nsurv = 0;                                  // non-empty keys that survived the threshold
    uint32_t nmerged = 32;                               // keys handed to the ranking step (a power of two, zero padded)
    if (p.L >= kp && p.L < (uint32_t)kSortCap) {
        const uint64_t* tops = p.tops + (size_t)qi * p.L;
        unsigned long long* Tp = reinterpret_cast<unsigned long long*>(scal + 4);
        if (tid == 0) { scal[0] = 0; scal[1] = 0; *Tp = 0ull; }
        // one thread per list: its maximum and its first four keys are requested together (one trip to L2 instead of two);
        // lists beyond the first 256 take their loads in the gather loop below
        const uint32_t llen = p.M / p.L;
        const uint32_t l0 = (uint32_t)tid;
        ulonglong2 ha = make_ulonglong2(0ull, 0ull), hb = ha;
        uint64_t top0 = 0ull;
        if (l0 < p.L) {
            top0 = __ldcg(tops + l0);
            ha = __ldcg(reinterpret_cast<const ulonglong2*>(keys + (size_t)l0 * llen));
            hb = __ldcg(reinterpret_cast<const ulonglong2*>(keys + (size_t)l0 * llen + 2));
        }
        stage_q(qs); q_staged = true;                     // the loads above are in flight meanwhile
        if (p.L <= (uint32_t)kFinThreads) {
            // up to 256 lists (one scan CTA per SM): every warp sorts its 32 maxima on shuffles; a key's rank is its place in its
            // own warp's list plus, by binary search, the number of greater keys in the others - ~40 shared-memory reads per
            // thread instead of L.  Empty maxima (0) tie with each other; no rank may then equal kp - 1 and T stays 0 (keep all).
            const uint64_t x = warp_sort_desc(top0, lane);
            sortbuf[tid] = x;
            fin_sync();
            if (tid == 0) stamp(p.dbg_times, 8);
            const uint32_t nw = (p.L + 31u) / 32u;
            if (x != 0ull) {
                uint32_t rank = (uint32_t)lane;
                for (uint32_t w = 0; w < nw; ++w) if (w != (uint32_t)warp) rank += count_greater_desc32(sortbuf + 32u * w, x);
                if (rank == kp - 1) *Tp = x;
            }
        } else {
            sortbuf[l0] = top0;
            for (uint32_t i = tid + kFinThreads; i < p.L; i += kFinThreads) sortbuf[i] = __ldcg(tops + i);
            if (tid == 0) sortbuf[p.L] = 0ull;            // pad to an even count: the rank loop reads two keys per load (an empty key outranks nothing)
            fin_sync();
            if (tid == 0) stamp(p.dbg_times, 8);
            for (uint32_t i = tid; i < p.L; i += kFinThreads) {
                const uint64_t v = sortbuf[i];
                uint32_t rank = 0;
                for (uint32_t o = 0; o < p.L; o += 2) {
                    const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(sortbuf + o);
                    rank += (w.x > v || (w.x == v && o < i)) ? 1u : 0u;
                    rank += (w.y > v || (w.y == v && o + 1 < i)) ? 1u : 0u;
                }
                if (rank == kp - 1) *Tp = v;
            }
        }
        fin_sync();
        if (tid == 0) stamp(p.dbg_times, 9);
        uint64_t T = *Tp;
        if (T == 0ull) T = 1ull;                          // fewer than kp non-empty lists: keep every key
        // the lists are sorted descending (bitonic merge in K1, sorted insertion in K2), so the keys >= T of a list are a
        // prefix of it: one thread per list takes 4 keys (32 bytes) at a time and stops at the first key below T
        for (uint32_t l = tid; l < p.L; l += kFinThreads) {
            const uint64_t* lk = keys + (size_t)l * llen;
            for (uint32_t i = 0; i < llen; i += 4) {
                ulonglong2 a = ha, b = hb;
                if (i > 0 || l != l0) {
                    a = __ldcg(reinterpret_cast<const ulonglong2*>(lk + i));
                    b = __ldcg(reinterpret_cast<const ulonglong2*>(lk + i + 2));
                }
                const uint64_t v[4] = {a.x, a.y, b.x, b.y};
                bool more = true;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (v[u] >= T) {
                        const uint32_t pos = atomicAdd(&scal[0], 1u);
                        if (pos < kFinWarps * KPW) sortbuf[pos] = v[u];
                    } else {
                        more = false;
                    }
                }
                if (!more) break;
            }
        }
        fin_sync();
        if (tid == 0) stamp(p.dbg_times, 10);
        const uint32_t got = scal[0];
        if (got <= kFinWarps * KPW) {
            while (nmerged < got) nmerged <<= 1;          // sort no more than the next power of two above the survivors
            for (uint32_t i = got + tid; i < nmerged; i += kFinThreads) sortbuf[i] = 0ull;
            folded = false;
            nsurv = got;
        }
        fin_sync();
    }
    // ---- 1b. general path: exact radix select of the kp-th largest key (MSB first, 8 bits per pass over the keys in L2);
    //          stops as soon as at most 2 k' keys lie at or above the current bucket
    if (!q_staged) stage_q(qs);
    if (folded) {
        uint64_t prefix = 0, mask = 0;
        uint32_t remaining = kp;                         // rank still to be located inside the current prefix bucket
        // refine until at most 2 kp keys survive: a small sort beats another pass over the keys only below that
        const uint32_t stop_at = min(kFinWarps * KPW, max(2u * kp, 64u));
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            hist[tid] = 0;                               // kFinThreads == 256
            fin_sync();
            for (uint32_t i = tid; i < p.M; i += kFinThreads) {
                const uint64_t v = __ldcg(keys + i);
                if (v != 0ull && (v & mask) == prefix) atomicAdd(&hist[(uint32_t)(v >> shift) & 255u], 1u);
            }
            fin_sync();
            if (tid == 0) {
                uint32_t cum = 0, digit = 0, rem = 0, stop = 1;
                for (int b = 255; b >= 0; --b) {
                    if (cum + hist[b] >= remaining) {
                        digit = (uint32_t)b; rem = remaining - cum;
                        // keys above the bucket: (kp - remaining) + cum; with the bucket itself they must fit the buffer
                        stop = ((kp - remaining) + cum + hist[b] <= stop_at) ? 1u : 0u;
                        break;
                    }
                    cum += hist[b];
                    if (b == 0) { digit = 0; rem = 0; stop = 1; }   // fewer than kp keys in this bucket chain: keep everything
                }
                scal[0] = digit; scal[1] = rem; scal[2] = stop;
            }
            fin_sync();
            const uint32_t digit = scal[0], rem = scal[1], stop = scal[2];
            fin_sync();
            if (rem == 0) { mask = ~0ull; prefix = 1ull; break; }   // every non-empty key qualifies
            prefix |= (uint64_t)digit << shift;
            mask |= 0xFFull << shift;
            remaining = rem;
            if (stop) break;
        }
        // gather the keys >= prefix (lower bits zero): at most 8*KPW of them by construction
        if (tid == 0) scal[0] = 0;
        fin_sync();
        for (uint32_t i = tid; i < p.M; i += kFinThreads) {
            const uint64_t v = __ldcg(keys + i);
            if (v != 0ull && v >= prefix) {
                const uint32_t pos = atomicAdd(&scal[0], 1u);
                if (pos < kFinWarps * KPW) sortbuf[pos] = v;
            }
        }
        fin_sync();
        const uint32_t got = min(scal[0], kFinWarps * KPW);
        nmerged = 32;
        while (nmerged < got) nmerged <<= 1;
        for (uint32_t i = got + tid; i < nmerged; i += kFinThreads) sortbuf[i] = 0ull;
        nsurv = got;
        fin_sync();
    }
    // ---- 2. rank the merged keys, keep the kp best (non-empty keys are distinct: the row is part of the key).  Every gathered
    //         key is non-empty (>= T >= 1, or tested), so the number of survivors is the number gathered. ----
    const uint32_t ncand = min(nsurv, kp);
    if (nmerged <= 256u) {
        // 256 / nmerged threads per key (adjacent lanes), each counting the greater keys in its share of the buffer
        const uint32_t tpe = (uint32_t)kFinThreads / nmerged, share = nmerged / tpe;
        const uint32_t i = (uint32_t)tid / tpe, part = (uint32_t)tid % tpe;
        const uint64_t v = sortbuf[i];
        uint32_t rank = 0;
        for (uint32_t o = part * share; o < (part + 1) * share; o += 2) {
            const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(sortbuf + o);
            rank += (w.x > v ? 1u : 0u) + (w.y > v ? 1u : 0u);
        }
        for (uint32_t m = 1; m < tpe; m <<= 1) rank += __shfl_xor_sync(0xFFFFFFFFu, rank, m);
        if (part == 0 && v != 0ull && rank < kp) ckey[rank] = v;
        fin_sync();
    } else {
        bitonic_sort_desc(sortbuf, (int)nmerged, tid);
        for (uint32_t c = tid; c < ncand; c += kFinThreads) ckey[c] = sortbuf[c];
        fin_sync();
    }
    // the fast-score bound for every row that is NOT a candidate
    // K2 lists are shorter than k': what a full list dropped is bounded by its last key (drops[])
    uint64_t drop_key = 0ull;
    if (p.drops != nullptr) {
        unsigned long long* Dp = reinterpret_cast<unsigned long long*>(scal + 6);
        if (tid == 0) *Dp = 0ull;
        fin_sync();
        uint64_t t = 0;
        for (uint32_t i = tid; i < p.L; i += kFinThreads) { const uint64_t v = __ldcg(p.drops + (size_t)qi * p.L + i); t = v > t ? v : t; }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { const uint64_t v = shfl_xor_u64(t, o); t = v > t ? v : t; }
        if (lane == 0 && t) atomicMax(Dp, (unsigned long long)t);
        fin_sync();
        drop_key = *Dp;
    }
    const bool lists_dropped_nothing = nsurv < kp && drop_key == 0ull;   // no list ever overflowed
    float t_fast = ncand > 0 ? key_score(ckey[ncand - 1]) : 0.f;
    if (nsurv < kp) t_fast = -INFINITY;                 // every kept key is a candidate
    if (drop_key != 0ull) t_fast = fmaxf(t_fast, key_score(drop_key));
    // rows and tie keys of the candidates: requested now, needed by the ordering step (their latency hides behind the rescoring)
    // (ncand <= kMaxCand = 256 = threads: one candidate per thread, kept in registers until the rescoring is done)
    uint32_t my_row = 0; uint64_t my_tie = 0ull;
    if ((uint32_t)tid < ncand) { my_row = key_row(ckey[tid]); my_tie = p.tiekey[my_row]; }
    if (tid == 0) stamp(p.dbg_times, 5);                 // candidates selected
    // ---- 4. exact rescoring: candidate c belongs to CTA c / nrw, warp c % nrw ----
    const uint32_t nrw = (uint32_t)p.n_rescore_warps;
    double* gscore = p.cand_scores + (size_t)qi * kMaxCand;
    if ((uint32_t)warp < nrw) {
        float* buf = chain + (size_t)warp * per_warp;
        float* leaf_out = buf + (size_t)3 * p.dim_pad;
        for (uint32_t c = cy * nrw + warp; c < ncand; c += C * nrw) {
            const double s = exact_score(p, pws, key_row(ckey[c]), qs, p.search_no + qi, buf, leaf_out, lane);
            if (lane == 0) gscore[c] = s;
        }
    }
    if ((uint32_t)tid < ncand) { crow[tid] = my_row; ctie[tid] = my_tie; }
    fin_sync();
    if (tid == 0) {
        stamp(p.dbg_times, 6);                           // this CTA's share of the rescoring is done
        uint32_t last = 1;
        if (C > 1) {
            __threadfence();
            last = (atomicAdd(p.tickets + qi, 1u) == C - 1) ? 1u : 0u;
        }
        scal[3] = last;
        scal[8] = 0u;                                    // the "exactness not proven" mark of step 6
    }
    fin_sync();
    if (scal[3] == 0) return false;
    __threadfence();
    for (uint32_t c = tid; c < ncand; c += kFinThreads) cscore[c] = __ldcg(gscore + c);
    if (tid == 0 && C > 1) p.tickets[qi] = 0;            // ready for the next launch
    fin_sync();
    if (tid == 0) stamp(p.dbg_times, 11);
    // ---- 5. final order: (score desc, tie asc, row asc) by rank counting, and
    //      6. the exactness proof for the weakest returned result ----
    const uint32_t nout = min(ncand, p.k);
    {
        // 256 / next_pow2(ncand) threads per candidate (adjacent lanes), each comparing with its share of the others
        uint32_t np2 = 32;
        while (np2 < ncand) np2 <<= 1;                   // <= kMaxCand = 256
        const uint32_t tpe = (uint32_t)kFinThreads / np2, share = np2 / tpe;
        const uint32_t c = (uint32_t)tid / tpe, part = (uint32_t)tid % tpe;
        const bool real = c < ncand;
        const double s = real ? cscore[c] : 0.0; const uint64_t t = real ? ctie[c] : 0ull; const uint32_t r = real ? crow[c] : 0u;
        uint32_t rank = 0;
        const uint32_t o_end = min((part + 1) * share, ncand);
        for (uint32_t o = part * share; o < o_end; ++o) {
            const double so = cscore[o]; const uint64_t to = ctie[o]; const uint32_t ro = crow[o];
            const bool better = (so > s) || (so == s && (to < t || (to == t && ro < r)));
            rank += better ? 1u : 0u;
        }
        for (uint32_t m = 1; m < tpe; m <<= 1) rank += __shfl_xor_sync(0xFFFFFFFFu, rank, m);
        if (real && part == 0) {
            crank[c] = rank;
            if (rank == p.k - 1 || (rank == ncand - 1 && ncand < p.k)) {
                // this candidate is the weakest returned result
                if (!lists_dropped_nothing && ncand >= p.k) {
                    // rows outside the candidate set have exact score <= t_fast + eps
                    double eps = p.eps_q != nullptr ? (double)p.eps_q[qi] + (double)p.eps_add : (double)p.eps;
                    if (p.metric == LVS_METRIC_DOT) eps *= (double)(p.qnorm != nullptr ? p.qnorm[qi] : qnorm_in) * (double)(*p.max_norm);
                    if (!(s > (double)t_fast + eps)) scal[8] = 1u;
                    if (ncand == kp && kp == p.k) scal[8] = 1u;  // no margin at all
                }
            }
        }
    }
    fin_sync();
    if (tid == 0) stamp(p.dbg_times, 12);
    res.score = cscore; res.row = crow; res.tie = ctie; res.rank = crank; res.ncand = ncand; res.nout = nout;
    res.flag = ncand == 0 ? 0 : (int32_t)scal[8];
    return true;
}

// Stores a finished query's result (local form: this shard IS the collection, or the exchange runs as its own kernel).
__device__ __forceinline__ void finalize_store_local(const FinalizeParams& p, const uint32_t qi, const FinResult& r, const int tid) {
    for (uint32_t c = tid; c < r.ncand; c += kFinThreads) {
        const uint32_t rk = r.rank[c];
        if (rk < p.k) {
            p.out_scores[(size_t)qi * p.k + rk] = r.score[c];
            p.out_rows[(size_t)qi * p.k + rk] = p.row_base + (int64_t)r.row[c];
            p.out_ties[(size_t)qi * p.k + rk] = r.tie[c];
        }
    }
    for (uint32_t c = r.nout + tid; c < p.k; c += kFinThreads) {
        p.out_scores[(size_t)qi * p.k + c] = 0.0;
        p.out_rows[(size_t)qi * p.k + c] = -1;
        p.out_ties[(size_t)qi * p.k + c] = 0ull;
    }
    if (tid == 0) { p.out_counts[qi] = r.nout; p.out_flags[qi] = r.flag; }
}

// grid = (queries, C): the standalone form behind the tensor-core path (K2 lists)
template <int KPL>
__global__ void __launch_bounds__(kFinThreads, 1) finalize_kernel(const FinalizeParams p) {
    extern __shared__ __align__(16) uint8_t fsm_dyn[];
    FinResult r;
    const uint32_t qi = blockIdx.x;
    const auto stage_q = [&](double* qs) {
        const double* qg = p.q64 + (size_t)qi * p.dim;
        for (int i = threadIdx.x; i < p.dim; i += kFinThreads) qs[i] = qg[i];
    };
    if (finalize_body<KPL>(p, qi, blockIdx.y, gridDim.y, fsm_dyn, 0.f, r, stage_q)) finalize_store_local(p, qi, r, threadIdx.x);
}

}  // namespace lvs
