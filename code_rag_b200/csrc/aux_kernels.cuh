// K4 (upsert: normalise / convert / scatter), query preparation, filter-only row matching, tombstones and
// K5 (merge of per-shard top-k lists after the all-gather).
#pragma once
#include "common.cuh"
#include "exchange_kernel.cuh"
#include "finalize_kernel.cuh"

namespace lvs {



// ---------------------------------------------------------------------------------------------------------
// K4 upsert.  Replaces the point marshalling + server-side insert behind QdrantManager.upsert
// (reference src/lattice/embeddings/client.py:115-130).  One warp per point:
//   fp32 storage, cosine : row = float32(x / ||x||_f64)      (what qdrant local mode stores)
//   fp32 storage, dot    : row = float32(x)
//   bf16 storage         : row = bf16(x) (exact when x is bf16-representable) and inv_norm = 1/||row||
// plus tie key, epoch (search counter at write time), tombstone clear and filter codes; padding is zeroed.
// ---------------------------------------------------------------------------------------------------------
struct UpsertParams {
    const void* src;            // [n][dim] device
    int src_dtype;
    int64_t n;
    const int64_t* rows;        // [n] local destination rows, or nullptr => row0 + i
    int64_t row0;
    uint8_t* base;
    uint32_t row_bytes;
    int dim;
    int storage;
    int metric;
    float* inv_norm;
    uint8_t* live;
    uint64_t* epoch;
    uint64_t epoch_val;
    uint64_t* tiekey;
    const uint64_t* ties_src;   // [n] or nullptr => tie key = global row
    int64_t row_base;
    uint32_t* codes[kMaxFilterCols];
    const uint32_t* codes_src;  // [n][n_cols] row-major, or nullptr => kNullCode
    int n_cols;
    float* max_norm;            // [0] running max ||row|| (dot metric error bound), [1] running max | ||row|| - 1 | (bf16 storage)
};

__global__ void __launch_bounds__(256) upsert_kernel(const UpsertParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int D = p.dim;
    for (int64_t i = wid; i < p.n; i += nw) {
        const int64_t row = p.rows ? p.rows[i] : p.row0 + i;
        const size_t so = (size_t)i * D;
        double ss = 0.0;
        for (int c = lane; c < D; c += 32) { const double x = load_as_f64(p.src, p.src_dtype, so + c); ss = fma(x, x, ss); }
        ss = warp_sum_f64(ss);
        const double nrm = sqrt(ss);
        uint8_t* rp = p.base + (size_t)row * p.row_bytes;
        if (p.storage == LVS_STORAGE_F32) {
            float* out = reinterpret_cast<float*>(rp);
            const int ld = p.row_bytes / 4;
            const bool norm = (p.metric == LVS_METRIC_COSINE) && nrm > 0.0;
            for (int c = lane; c < ld; c += 32) {
                float v = 0.f;
                if (c < D) { const double x = load_as_f64(p.src, p.src_dtype, so + c); v = (float)(norm ? x / nrm : x); }
                out[c] = v;
            }
            if (lane == 0) p.inv_norm[row] = 1.0f;
            if (lane == 0 && p.metric == LVS_METRIC_DOT) atomicMax(reinterpret_cast<int*>(p.max_norm), __float_as_int((float)nrm * 1.0000002f));
        } else {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(rp);
            const int ld = p.row_bytes / 2;
            double ss2 = 0.0;
            for (int c = lane; c < ld; c += 32) {
                __nv_bfloat16 b = __float2bfloat16(0.f);
                if (c < D) { b = __double2bfloat16(load_as_f64(p.src, p.src_dtype, so + c)); const double t = (double)__bfloat162float(b); ss2 = fma(t, t, ss2); }
                out[c] = b;
            }
            ss2 = warp_sum_f64(ss2);
            if (lane == 0) {
                const double n2 = sqrt(ss2);
                p.inv_norm[row] = n2 > 0.0 ? (float)(1.0 / n2) : 0.f;
                atomicMax(reinterpret_cast<int*>(p.max_norm + 1), __float_as_int((float)fabs(n2 - 1.0) * 1.0000002f));
                if (p.metric == LVS_METRIC_DOT) atomicMax(reinterpret_cast<int*>(p.max_norm), __float_as_int((float)n2 * 1.0000002f));
            }
        }
        if (lane == 0) {
            p.live[row] = 1;
            p.epoch[row] = p.epoch_val;
            p.tiekey[row] = p.ties_src ? p.ties_src[i] : (uint64_t)(p.row_base + row);
        }
        if (lane < p.n_cols) p.codes[lane][row] = p.codes_src ? p.codes_src[(size_t)i * p.n_cols + lane] : kNullCode;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Query preparation: q64 = q / ||q||_f64 (cosine; local mode divides by EPSILON when the norm is 0) or q (dot);
// q32 = float32(q64) laid out with the scan's row stride; qb16 is filled by the tensor-core path.
// One CTA per query.
// ---------------------------------------------------------------------------------------------------------
struct PrepParams {
    const void* src; int src_dtype; int dim; int metric;
    double* q64;          // [Q][dim]
    float* q32;           // [Q][q_stride]
    uint32_t q_stride;
    float* qnorm;         // [Q] ||q|| (dot metric error bound)
    int n_zero_rows;      // the CTA after the last query zeroes this many fp32 query rows (unused scan slots)
};

__global__ void __launch_bounds__(256) prep_queries_kernel(const PrepParams p) {
    __shared__ double red[8];
    __shared__ double s_norm;
    const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (qi == (int)gridDim.x - 1) {
        for (uint32_t c = tid; c < (uint32_t)p.n_zero_rows * p.q_stride; c += 256) p.q32[(size_t)qi * p.q_stride + c] = 0.f;
        return;
    }
    const size_t so = (size_t)qi * p.dim;
    double ss = 0.0;
    for (int c = tid; c < p.dim; c += 256) { const double x = load_as_f64(p.src, p.src_dtype, so + c); ss = fma(x, x, ss); }
    ss = warp_sum_f64(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int w = 0; w < 8; ++w) t += red[w]; s_norm = sqrt(t); }
    __syncthreads();
    const double nrm = s_norm;
    const double div = (p.metric == LVS_METRIC_COSINE) ? (nrm != 0.0 ? nrm : 1.1920929e-7) : 1.0;
    for (uint32_t c = tid; c < p.q_stride; c += 256) {
        double v = 0.0;
        if ((int)c < p.dim) { v = load_as_f64(p.src, p.src_dtype, so + c) / div; p.q64[so + c] = v; }
        p.q32[(size_t)qi * p.q_stride + c] = (float)v;
    }
    if (tid == 0) p.qnorm[qi] = (float)nrm;
}

// bf16 copy of the unit queries for the tensor-core path: [rows_out][k_pad], zero padded in both directions, plus the
// per-query bound on |bf16(q).x - q.x| for unit rows x:  ||q - bf16(q)||_2  (Cauchy-Schwarz) + fp32 accumulation slack.
__global__ void __launch_bounds__(256) prep_qb16_kernel(const double* q64, int Q, int dim, __nv_bfloat16* out, uint32_t k_pad, uint32_t rows_out,
                                                        float* eps_out) {
    __shared__ double red[8];
    const uint32_t r = blockIdx.x;
    if (r >= rows_out) return;
    double e2 = 0.0;
    for (uint32_t c = threadIdx.x; c < k_pad; c += 256) {
        double v = 0.0;
        if ((int)r < Q && (int)c < dim) v = q64[(size_t)r * dim + c];
        const __nv_bfloat16 b = __double2bfloat16(v);
        out[(size_t)r * k_pad + c] = b;
        const double d = v - (double)__bfloat162float(b);
        e2 = fma(d, d, e2);
    }
    e2 = warp_sum_f64(e2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e2;
    __syncthreads();
    if (threadIdx.x == 0 && (int)r < Q) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        eps_out[r] = (float)(sqrt(t) * 1.0001 + 1.0e-4);
    }
}

// fp32 shards (K2 with kind::tf32): the unit queries rounded to tf32 (round to nearest, low 13 mantissa bits cleared) so that the
// tensor core's own truncation leaves the query side exact; eps_out = ||q - tf32(q)||_2 + slack
__global__ void __launch_bounds__(256) prep_qtf32_kernel(const double* q64, int Q, int dim, float* out, uint32_t k_pad, uint32_t rows_out,
                                                         float* eps_out) {
    __shared__ double red[8];
    const uint32_t r = blockIdx.x;
    if (r >= rows_out) return;
    double e2 = 0.0;
    for (uint32_t c = threadIdx.x; c < k_pad; c += 256) {
        double v = 0.0;
        if ((int)r < Q && (int)c < dim) v = q64[(size_t)r * dim + c];
        uint32_t t;
        const float f = (float)v;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(f));
        t &= 0xFFFFE000u;
        const float b = __uint_as_float(t);
        out[(size_t)r * k_pad + c] = b;
        const double d = v - (double)b;
        e2 = fma(d, d, e2);
    }
    e2 = warp_sum_f64(e2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e2;
    __syncthreads();
    if (threadIdx.x == 0 && (int)r < Q) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        eps_out[r] = (float)(sqrt(t) * 1.0001 + 1.0e-4);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Filter-only matching (query_vector=None searches, delete-by-filter, count, scroll):
// appends every live row whose codes satisfy the conjunction to out_rows (unordered), counts all matches.
// ---------------------------------------------------------------------------------------------------------
struct MatchParams {
    uint32_t n_rows;
    const uint8_t* live;
    const uint32_t* codes[kMaxFilterCols];
    uint32_t want[kMaxFilterCols];
    uint32_t n_filter;
    int64_t row_base;
    int64_t* out_rows; uint32_t cap;
    uint32_t* counter;       // [0] = matches
    int tombstone;           // 1 => also clear the live byte of every match (delete by filter)
    uint8_t* live_rw;
};

__global__ void __launch_bounds__(256) match_rows_kernel(const MatchParams p) {
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < p.n_rows; row += gridDim.x * blockDim.x) {
        bool pass = p.live[row] != 0;
        for (uint32_t f = 0; f < p.n_filter && pass; ++f) pass = p.codes[f][row] == p.want[f];
        if (pass) {
            const uint32_t pos = atomicAdd(p.counter, 1u);
            if (pos < p.cap) p.out_rows[pos] = p.row_base + (int64_t)row;
            if (p.tombstone) p.live_rw[row] = 0;
        }
    }
}

__global__ void set_live_kernel(uint8_t* live, const int64_t* rows, int64_t n, uint8_t value, uint32_t n_rows, uint32_t* changed) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rows[i];
        if (r >= 0 && r < (int64_t)n_rows) {
            if (live[r] != value) { live[r] = value; atomicAdd(changed, 1u); }
        }
    }
}

// Compaction: row src[i] (vector, tombstone, codes, tie key, write epoch, norm, ranking attributes) replaces row dst[i]; the source
// row becomes a tombstone.  One warp per pair; the two row sets are disjoint, so the order of the pairs does not matter.
struct MoveRowsParams {
    const int64_t* src; const int64_t* dst; int64_t n;
    uint8_t* vec; uint32_t row_bytes;
    uint8_t* live; uint64_t* epoch; uint64_t* tie; float* inv_norm;
    uint32_t* codes[kMaxFilterCols]; int n_cols;
    uint32_t* rk_key; uint32_t* rk_file; uint32_t* rk_cent; uint32_t* rk_name; int32_t* rk_clen; uint8_t* rk_flags; int64_t rk_rows;
};
__global__ void __launch_bounds__(256) move_rows_kernel(const MoveRowsParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < p.n; i += nw) {
        const int64_t s = p.src[i], d = p.dst[i];
        const uint4* sp = reinterpret_cast<const uint4*>(p.vec + (size_t)s * p.row_bytes);
        uint4* dp = reinterpret_cast<uint4*>(p.vec + (size_t)d * p.row_bytes);
        for (uint32_t c = lane; c < p.row_bytes / 16; c += 32) dp[c] = sp[c];
        if (lane < p.n_cols) p.codes[lane][d] = p.codes[lane][s];
        if (lane == 8) { p.epoch[d] = p.epoch[s]; p.tie[d] = p.tie[s]; p.inv_norm[d] = p.inv_norm[s]; }
        if (lane == 9 && p.rk_key != nullptr && s < p.rk_rows && d < p.rk_rows) {
            p.rk_key[d] = p.rk_key[s]; p.rk_file[d] = p.rk_file[s]; p.rk_cent[d] = p.rk_cent[s]; p.rk_name[d] = p.rk_name[s];
            p.rk_clen[d] = p.rk_clen[s]; p.rk_flags[d] = p.rk_flags[s];
        }
        __syncwarp();
        if (lane == 0) { p.live[d] = p.live[s]; p.live[s] = 0; }
    }
}

__global__ void set_codes_kernel(uint32_t* col, const int64_t* rows, int64_t row0, const uint32_t* src, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rows ? rows[i] : row0 + i;
        col[r] = src[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5: merge of G per-shard result lists (after the NCCL all-gather) into the global top-k.
// in_*: G blocks of [Q][k], `shard_stride` elements apart; rows < 0 are padding.  Order: (score desc, tie asc, row asc).  One CTA per query.
// ---------------------------------------------------------------------------------------------------------
struct MergeParams {
    const double* in_scores; const int64_t* in_rows; const uint64_t* in_ties;
    int G; int Q; int k;
    int64_t shard_stride;   // elements between consecutive shards' [Q][k] blocks (Q*k when dense)
    double* out_scores; int64_t* out_rows; uint64_t* out_ties; uint32_t* out_counts;
};

__global__ void __launch_bounds__(256) merge_topk_kernel(const MergeParams p) {
    extern __shared__ __align__(16) uint8_t msm[];
    const int n = p.G * p.k;
    double* s = reinterpret_cast<double*>(msm);
    int64_t* r = reinterpret_cast<int64_t*>(msm + (size_t)n * 8);
    uint64_t* t = reinterpret_cast<uint64_t*>(msm + (size_t)n * 16);
    __shared__ uint32_t nvalid;
    const int qi = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) nvalid = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const int g = i / p.k, j = i % p.k;
        const size_t src = (size_t)g * p.shard_stride + (size_t)qi * p.k + j;
        s[i] = p.in_scores[src]; r[i] = p.in_rows[src]; t[i] = p.in_ties[src];
        if (r[i] >= 0) atomicAdd(&nvalid, 1u);
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        if (r[i] < 0) continue;
        uint32_t rank = 0;
        for (int o = 0; o < n; ++o) {
            if (r[o] < 0) continue;
            const bool better = (s[o] > s[i]) || (s[o] == s[i] && (t[o] < t[i] || (t[o] == t[i] && r[o] < r[i])));
            rank += better ? 1u : 0u;
        }
        if (rank < (uint32_t)p.k) {
            p.out_scores[(size_t)qi * p.k + rank] = s[i];
            p.out_rows[(size_t)qi * p.k + rank] = r[i];
            p.out_ties[(size_t)qi * p.k + rank] = t[i];
        }
    }
    const uint32_t nout = min(nvalid, (uint32_t)p.k);
    for (int j = nout + tid; j < p.k; j += blockDim.x) {
        p.out_scores[(size_t)qi * p.k + j] = 0.0; p.out_rows[(size_t)qi * p.k + j] = -1; p.out_ties[(size_t)qi * p.k + j] = 0ull;
    }
    if (tid == 0) p.out_counts[qi] = nout;
}

// ---------------------------------------------------------------------------------------------------------
// K5' as its own launch (behind the tensor-core path; the scan path runs the same steps inside its own kernel)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeParams p) {
    extern __shared__ __align__(16) uint8_t xsm[];
    __shared__ uint32_t s_last, s_nvalid, s_timeout;
    const int tid = threadIdx.x;
    const size_t n = (size_t)3 * p.Q * p.k;
    if (tid == 0) s_timeout = 0;
    // ---- 1. publish: this rank's block (+ flags) -> every rank's gather buffer (its own included) ----
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n + (size_t)p.Q; i += (size_t)gridDim.x * blockDim.x) {
        const int64_t v = i < n ? p.local[i] : (p.local_flags != nullptr ? (int64_t)p.local_flags[i - n] : 0);
        for (int r = 0; r < p.world; ++r) p.peer_data[r][(size_t)p.rank * p.blk_stride + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        s_last = (atomicAdd(p.done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            *p.done_counter = 0;
            exchange_signal(p);
        }
    }
    // ---- 2. wait for every rank's flag of this sequence number ----
    if (tid < p.world && !exchange_wait_one(p, tid)) s_timeout = 1;
    __syncthreads();
    // ---- 3. merge, one query at a time ----
    const bool timed_out = s_timeout != 0;
    for (int qi = blockIdx.x; qi < p.Q; qi += gridDim.x)
        exchange_merge_query(p, qi, xsm, &s_nvalid, tid, (int)blockDim.x, timed_out, [] { __syncthreads(); });
}

}  // namespace lvs
