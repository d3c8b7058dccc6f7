// Candidate selection + EXACT rescoring + final ordering (one CTA per query).
//
// The scan (K1) and the tensor-core path (K2) rank rows by an fp32 score.  qdrant-client local mode - the
// engine behind the reference's QdrantManager.search (reference src/lattice/embeddings/client.py:132-157) -
// scores in float64 on a float32 matrix that it re-normalises IN PLACE on every cosine search
// (oracle/qdrant_local.py, point 2).  To return bit-identical id lists this kernel
//   1. selects the k' best fast keys out of the per-CTA lists (threshold filter + bitonic sort, with an exact
//      8-pass radix select as the fallback when too many keys survive the threshold),
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
    uint16_t leaf_off[kPwMaxLeaves];
    uint16_t leaf_len[kPwMaxLeaves];
    uint8_t step_dst[kPwMaxLeaves];
    uint8_t step_src[kPwMaxLeaves];
};

struct FinalizeParams {
    const uint64_t* keys;      // [Q][M]   fast keys (0 = empty)
    const uint64_t* mins;      // [Q][L]   minimum of each list (L lists of M/L keys), or nullptr
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
    const uint32_t* epoch;     // [rows] value of the collection's search counter when the row was written
    uint32_t search_no;        // counter value of query 0 of this launch (query qi is search_no + qi)
    const PwProgram* pw;
    float eps;                 // bound on |fast score - exact score| for unit vectors
    const float* qnorm;        // [Q] ||q||      } dot metric: the bound scales with ||q|| * max ||row||
    const float* max_norm;     // [1] max ||row|| }
    int64_t row_base;          // global row of local row 0
    int n_rescore_warps;       // warps that own chain buffers
    double* out_scores;        // [Q][k]
    int64_t* out_rows;         // [Q][k]   -1 padded
    uint64_t* out_ties;        // [Q][k]
    int32_t* out_flags;        // [Q]      bit0: exactness not proven
    uint32_t* out_counts;      // [Q]      number of valid results
};

__host__ __device__ inline size_t finalize_smem_bytes(int dim_pad, int n_rescore_warps) {
    size_t b = 0;
    b += (size_t)kSortCap * 8;                    // sort buffer
    b += 256 * 4 + 64;                            // histogram + scalars
    b += (size_t)kMaxCand * (8 + 8 + 8 + 4 + 4);  // cand key, score, tie, row, rank
    b = (b + 15) & ~(size_t)15;
    b += (size_t)n_rescore_warps * ((size_t)3 * dim_pad * 4 + kPwMaxLeaves * 4);
    return b;
}

// ||v||_2 in float32 exactly as numpy computes np.linalg.norm(M, axis=-1) for a float32 row:
// s_i = fl(v_i*v_i); pairwise_sum(s); sqrt.  Warp-cooperative; all lanes return the same value.
__device__ __forceinline__ float np_norm_f32(const float* v, const PwProgram* pw, float* leaf_out, int lane) {
    const int nl = pw->n_leaves;
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

// Exact score of one candidate row (warp-cooperative).  buf = 3 * dim_pad floats of shared memory.
__device__ __forceinline__ double exact_score(const FinalizeParams& p, uint32_t row, const double* q, uint32_t search_no,
                                              float* buf, float* leaf_out, int lane) {
    const int D = p.dim;
    const uint8_t* rp = p.base + (size_t)row * p.row_bytes;
    if (p.metric == LVS_METRIC_DOT) {
        double acc = 0.0;
        if (p.storage == LVS_STORAGE_F32) {
            const float* x = reinterpret_cast<const float*>(rp);
            for (int i = lane; i < D; i += 32) acc = fma((double)x[i], q[i], acc);
        } else {
            const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(rp);
            for (int i = lane; i < D; i += 32) acc = fma((double)__bfloat162float(x[i]), q[i], acc);
        }
        return warp_sum_f64(acc);
    }
    float* cur = buf;
    float* prev = buf + p.dim_pad;
    float* nxt = buf + 2 * p.dim_pad;
    if (p.storage == LVS_STORAGE_F32) {
        const float* x = reinterpret_cast<const float*>(rp);
        for (int i = lane; i < D; i += 32) cur[i] = x[i];
    } else {
        // what local mode would have stored for this (bf16-representable) input: float32(x / ||x||_f64)
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(rp);
        double ss = 0.0;
        for (int i = lane; i < D; i += 32) { const double t = (double)__bfloat162float(x[i]); ss = fma(t, t, ss); }
        ss = warp_sum_f64(ss);
        const double nrm = sqrt(ss);
        for (int i = lane; i < D; i += 32) {
            const double t = (double)__bfloat162float(x[i]);
            cur[i] = (float)(nrm > 0.0 ? t / nrm : t);
        }
    }
    __syncwarp();
    // replay of the per-search in-place re-normalisation
    uint32_t count = search_no - p.epoch[row];          // searches run since the row was written, this one included
    if (count > 64u) count = 64u + ((count - 64u) & 1u);
    bool have_prev = false;
    for (uint32_t j = 1; j <= count; ++j) {
        const float n = np_norm_f32(cur, p.pw, leaf_out, lane);
        if (n == 1.0f) break;                            // fixed point
        const float d = (n != 0.0f) ? n : 1.1920929e-7f;
        bool same_cur = true, same_prev = true;
        for (int i = lane; i < D; i += 32) {
            const float y = __fdiv_rn(cur[i], d);
            nxt[i] = y;
            same_cur &= (y == cur[i]) || (y != y);
            if (have_prev) same_prev &= (y == prev[i]) || (y != y);
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
    double acc = 0.0;
    for (int i = lane; i < D; i += 32) acc = fma((double)cur[i], q[i], acc);
    __syncwarp();
    return warp_sum_f64(acc);
}

__device__ __forceinline__ void bitonic_sort_desc(uint64_t* buf, int n, int tid) {
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
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kFinThreads, 1) finalize_kernel(const FinalizeParams p) {
    extern __shared__ __align__(16) uint8_t fsm[];
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(fsm);
    uint32_t* hist = reinterpret_cast<uint32_t*>(fsm + (size_t)kSortCap * 8);
    uint32_t* scal = hist + 256;                       // [0]=survivor count [1]=digit [2]=remaining [4..5]=T
    uint8_t* cp = reinterpret_cast<uint8_t*>(scal + 16);
    uint64_t* ckey = reinterpret_cast<uint64_t*>(cp);  cp += (size_t)kMaxCand * 8;
    double* cscore = reinterpret_cast<double*>(cp);    cp += (size_t)kMaxCand * 8;
    uint64_t* ctie = reinterpret_cast<uint64_t*>(cp);  cp += (size_t)kMaxCand * 8;
    uint32_t* crow = reinterpret_cast<uint32_t*>(cp);  cp += (size_t)kMaxCand * 4;
    uint32_t* crank = reinterpret_cast<uint32_t*>(cp); cp += (size_t)kMaxCand * 4;
    size_t off = (size_t)(cp - fsm);
    off = (off + 15) & ~(size_t)15;
    float* chain = reinterpret_cast<float*>(fsm + off);
    const size_t per_warp = (size_t)3 * p.dim_pad + kPwMaxLeaves;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t qi = blockIdx.x;
    const uint64_t* keys = p.keys + (size_t)qi * p.M;
    const uint32_t kp = p.kp;

    // ---- 1. threshold T = max over lists of the list minimum (>= kp keys are >= T when lists are full) ----
    unsigned long long* Tp = reinterpret_cast<unsigned long long*>(scal + 4);
    if (tid == 0) { scal[0] = 0; *Tp = 0ull; }
    __syncthreads();
    if (p.mins != nullptr && (p.M / p.L) >= kp) {
        uint64_t t = 0;
        for (uint32_t i = tid; i < p.L; i += kFinThreads) { const uint64_t v = p.mins[(size_t)qi * p.L + i]; t = v > t ? v : t; }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { const uint64_t v = shfl_xor_u64(t, o); t = v > t ? v : t; }
        if (lane == 0) atomicMax(Tp, (unsigned long long)t);
    }
    __syncthreads();
    uint64_t T = *Tp;
    if (T == 0) T = 1;                                  // keep every non-empty key
    // ---- 2. gather survivors ----
    for (uint32_t i = tid; i < p.M; i += kFinThreads) {
        const uint64_t v = keys[i];
        if (v >= T) {
            const uint32_t pos = atomicAdd(&scal[0], 1u);
            if (pos < (uint32_t)kSortCap) sortbuf[pos] = v;
        }
    }
    __syncthreads();
    uint32_t nsurv = scal[0];
    __syncthreads();
    if (nsurv > (uint32_t)kSortCap) {
        // ---- fallback: exact radix select of the kp-th largest key, then gather keys >= it ----
        uint64_t prefix = 0, mask = 0;
        uint32_t remaining = kp;
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            hist[tid] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < p.M; i += kFinThreads) {
                const uint64_t v = keys[i];
                if (v != 0 && (v & mask) == prefix) atomicAdd(&hist[(uint32_t)(v >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t cum = 0; int digit = 0; uint32_t rem = remaining;
                for (int b = 255; b >= 0; --b) {
                    if (cum + hist[b] >= rem) { digit = b; rem -= cum; break; }
                    cum += hist[b];
                    if (b == 0) { digit = 0; rem = 0; }   // fewer than kp keys in total
                }
                scal[1] = (uint32_t)digit; scal[2] = rem;
            }
            __syncthreads();
            prefix |= (uint64_t)scal[1] << shift;
            mask |= 0xFFull << shift;
            remaining = scal[2];
            __syncthreads();
            if (remaining == 0) { prefix = 1; break; }
        }
        if (tid == 0) scal[0] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < p.M; i += kFinThreads) {
            const uint64_t v = keys[i];
            if (v != 0 && v >= prefix) {
                const uint32_t pos = atomicAdd(&scal[0], 1u);
                if (pos < (uint32_t)kSortCap) sortbuf[pos] = v;
            }
        }
        __syncthreads();
        nsurv = min(scal[0], (uint32_t)kSortCap);
        __syncthreads();
    }
    // ---- 3. sort survivors, keep kp ----
    int n2 = 32;
    while (n2 < (int)nsurv) n2 <<= 1;
    for (int i = nsurv + tid; i < n2; i += kFinThreads) sortbuf[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(sortbuf, n2, tid);
    const uint32_t ncand = min(nsurv, kp);
    // the fast-score bound for every row that is NOT a candidate
    const bool lists_dropped_nothing = nsurv < kp;      // fewer keys than k' in total => no list ever overflowed
    const float t_fast = ncand > 0 ? key_score(sortbuf[ncand - 1]) : 0.f;
    for (uint32_t c = tid; c < ncand; c += kFinThreads) {
        const uint64_t kk = sortbuf[c];
        ckey[c] = kk;
        const uint32_t r = key_row(kk);
        crow[c] = r;
        ctie[c] = p.tiekey[r];
    }
    __syncthreads();
    // ---- 4. exact rescoring ----
    const double* q = p.q64 + (size_t)qi * p.dim;
    if (warp < p.n_rescore_warps) {
        float* buf = chain + (size_t)warp * per_warp;
        float* leaf_out = buf + (size_t)3 * p.dim_pad;
        for (uint32_t c = warp; c < ncand; c += p.n_rescore_warps) {
            const double s = exact_score(p, crow[c], q, p.search_no + qi, buf, leaf_out, lane);
            if (lane == 0) cscore[c] = s;
        }
    }
    __syncthreads();
    // ---- 5. final order: (score desc, tie asc, row asc) by rank counting ----
    for (uint32_t c = tid; c < ncand; c += kFinThreads) {
        const double s = cscore[c]; const uint64_t t = ctie[c]; const uint32_t r = crow[c];
        uint32_t rank = 0;
        for (uint32_t o = 0; o < ncand; ++o) {
            const double so = cscore[o]; const uint64_t to = ctie[o]; const uint32_t ro = crow[o];
            const bool better = (so > s) || (so == s && (to < t || (to == t && ro < r)));
            rank += better ? 1u : 0u;
        }
        crank[c] = rank;
    }
    __syncthreads();
    const uint32_t nout = min(ncand, p.k);
    for (uint32_t c = tid; c < ncand; c += kFinThreads) {
        const uint32_t rk = crank[c];
        if (rk < p.k) {
            p.out_scores[(size_t)qi * p.k + rk] = cscore[c];
            p.out_rows[(size_t)qi * p.k + rk] = p.row_base + (int64_t)crow[c];
            p.out_ties[(size_t)qi * p.k + rk] = ctie[c];
        }
        if (rk == p.k - 1 || (rk == ncand - 1 && ncand < p.k)) {
            // this candidate is the weakest returned result
            int32_t flag = 0;
            if (!lists_dropped_nothing && ncand >= p.k) {
                // rows outside the candidate set have exact score <= t_fast + eps
                double eps = (double)p.eps;
                if (p.metric == LVS_METRIC_DOT) eps *= (double)p.qnorm[qi] * (double)(*p.max_norm);
                if (!(cscore[c] > (double)t_fast + eps)) flag = 1;
                if (ncand == kp && kp == p.k) flag = 1;  // no margin at all
            }
            p.out_flags[qi] = flag;
        }
    }
    for (uint32_t c = nout + tid; c < p.k; c += kFinThreads) {
        p.out_scores[(size_t)qi * p.k + c] = 0.0;
        p.out_rows[(size_t)qi * p.k + c] = -1;
        p.out_ties[(size_t)qi * p.k + c] = 0ull;
    }
    if (tid == 0) {
        p.out_counts[qi] = nout;
        if (ncand == 0) p.out_flags[qi] = 0;
    }
}

}  // namespace lvs
