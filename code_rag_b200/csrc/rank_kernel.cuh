// K3 - fused hybrid ranking (regime 3 of BASELINE.json north_star): per-candidate signal scoring, intent-weighted
// blend, order-dependent merge by key, stable sort, per-file cap and total cap, batched over queries (one CTA each).
//
// Replaces the pure-Python loops of the reference's HybridRanker.rank_results (src/lattice/query/ranking/ranker.py:18-226)
// + ResultScorer (scorer.py:9-126) [mode 0] and of ResultReranker.fuse_results / deduplicate / normalize_scores
// (src/lattice/query/reranker.py:29-145) [mode 1].  All arithmetic is float64 in the reference's operation order, with
// round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an fma: results are bit-identical to CPython's.
// String work (keys, file paths, entity-name matching) is done once on the host and arrives as dense integer ids.
#pragma once
#include "common.cuh"

namespace lvs {

constexpr int kRankThreads = 256;
constexpr int kRankMaxCand = 2048;     // candidates per query
constexpr int kRankSignals = 7;        // graph_match, vector_similarity, centrality, query_entity_match,
                                       // relationship_relevance, code_quality, context_richness (RankingSignal order)
enum : int { SIG_GRAPH = 0, SIG_VECTOR = 1, SIG_CENTRALITY = 2, SIG_ENTITY = 3, SIG_REL = 4, SIG_QUALITY = 5, SIG_CONTEXT = 6 };

struct RankParams {
    int n_queries;
    const int32_t* offsets;      // [n_queries + 1]
    const int32_t* counts;       // optional [n_queries]: candidates actually present (fused path: graph + hits found), else offsets diff
    const uint8_t* kind;         // 0 primary, 1 caller, 2 callee, 3 other graph, 4 vector
    const uint32_t* key_id;
    const uint32_t* file_id;
    const int32_t* depth;
    const double* entity_match;
    const int32_t* degree;       // total_degree, < 0 = not in the centrality dict
    const uint8_t* flags;        // bit0 summary, bit1 docstring, bit2 signature, bit3 content
    const int32_t* content_len;  // < 0 = no content
    const double* vscore;
    const double* weights;       // [n_queries][4] graph, vector, centrality, context
    int mode;                    // 0 = HybridRanker, 1 = ResultReranker
    int max_per_file;            // <= 0: no cap
    int max_total;               // rows of the per-query output
    double entity_bonus, rel_bonus;
    int32_t* out_count;          // [n_queries]
    int32_t* out_index;          // [n_queries][max_total] leader candidate (index within the query)
    double* out_score;           // [n_queries][max_total]
    double* out_norm;            // [n_queries][max_total] min-max normalised scores of the emitted list (mode 1)
    double* out_signals;         // [n_queries][max_total][7]
    uint8_t* out_sigmask;        // [n_queries][max_total] bit s = signal s present
    uint8_t* out_source;         // [n_queries][max_total] 0 graph, 1 vector, 2 hybrid
    int32_t* out_leader;         // [total candidates] leader (index within the query) of every candidate
};

__device__ __forceinline__ double rank_candidate(const RankParams& p, int g, const double* w, double* sig, uint32_t* mask) {
    const int kind = p.kind[g];
    const double em = p.entity_match[g];
    if (p.mode == 1) {                                     // reranker.py:147-172: constant graph weight, scaled vector score
        *mask = 0;
        return kind < 4 ? w[0] : __dmul_rn(p.vscore[g], w[1]);
    }
    double cen = 0.0;
    if (p.degree[g] >= 0) cen = fmin(1.0, __ddiv_rn((double)p.degree[g], 50.0));
#pragma unroll
    for (int s = 0; s < kRankSignals; ++s) sig[s] = 0.0;
    if (kind == 4) {                                       // scorer.py:79-126
        double q = 0.0;
        const int n = p.content_len[g];
        if (n > 0) q = (n > 100 && n < 2000) ? 0.8 : (n > 50 && n < 3000) ? 0.5 : 0.3;
        const double vs = p.vscore[g];
        sig[SIG_VECTOR] = vs; sig[SIG_ENTITY] = em; sig[SIG_CENTRALITY] = cen; sig[SIG_QUALITY] = q;
        *mask = (1u << SIG_VECTOR) | (1u << SIG_ENTITY) | (1u << SIG_CENTRALITY) | (1u << SIG_QUALITY);
        double t = __dmul_rn(vs, w[1]);
        t = __dadd_rn(t, __dmul_rn(em, p.entity_bonus));
        t = __dadd_rn(t, __dmul_rn(cen, w[2]));
        t = __dadd_rn(t, __dmul_rn(q, 0.1));
        return t;
    }
    double base = 1.0;                                     // scorer.py:9-77
    if (kind == 1 || kind == 2) {
        int d = p.depth[g];
        if (d == 0) d = 1;                                 // `depth_from_query or 1`
        base = fmax(0.3, __dsub_rn(1.0, __dmul_rn((double)(d - 1), 0.2)));
    }
    const double rel = kind == 0 ? 1.0 : kind == 1 ? 0.8 : kind == 2 ? 0.7 : 0.5;
    const uint8_t f = p.flags[g];
    double ctx = 0.0;
    if (f & 1) ctx = __dadd_rn(ctx, 0.3);
    if (f & 2) ctx = __dadd_rn(ctx, 0.2);
    if (f & 4) ctx = __dadd_rn(ctx, 0.2);
    if (f & 8) ctx = __dadd_rn(ctx, 0.3);
    sig[SIG_GRAPH] = base; sig[SIG_ENTITY] = em; sig[SIG_REL] = rel; sig[SIG_CENTRALITY] = cen; sig[SIG_CONTEXT] = ctx;
    *mask = (1u << SIG_GRAPH) | (1u << SIG_ENTITY) | (1u << SIG_REL) | (1u << SIG_CENTRALITY) | (1u << SIG_CONTEXT);
    double t = __dmul_rn(base, w[0]);
    t = __dadd_rn(t, __dmul_rn(em, p.entity_bonus));
    t = __dadd_rn(t, __dmul_rn(rel, p.rel_bonus));
    t = __dadd_rn(t, __dmul_rn(cen, w[2]));
    t = __dadd_rn(t, __dmul_rn(ctx, w[3]));
    return t;
}

__global__ void __launch_bounds__(kRankThreads) rank_fuse_kernel(const RankParams p) {
    extern __shared__ __align__(16) uint8_t rsm[];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int base = p.offsets[q];
    const int C = p.counts ? p.counts[q] : p.offsets[q + 1] - base;
    double* sc = reinterpret_cast<double*>(rsm);                       // [C] candidate score, then merged score (leaders)
    double* fin = sc + C;                                              // [C] merged score of leaders
    uint32_t* key = reinterpret_cast<uint32_t*>(fin + C);              // [C]
    uint32_t* fil = key + C;                                           // [C]
    int32_t* lead = reinterpret_cast<int32_t*>(fil + C);               // [C]
    int32_t* rnk = lead + C;                                           // [C] rank among leaders (-1 for non-leaders)
    int32_t* pos = rnk + C;                                            // [C] output position or -1
    __shared__ double s_min, s_max;
    __shared__ int s_nout;
    const double* w = p.weights + (size_t)q * 4;

    // 1. per-candidate scores
    for (int i = tid; i < C; i += kRankThreads) {
        double sig[kRankSignals]; uint32_t m;
        sc[i] = rank_candidate(p, base + i, w, sig, &m);
        key[i] = p.key_id[base + i];
        fil[i] = p.file_id[base + i];
    }
    __syncthreads();
    // 2. leader = first candidate with the same key (dict insertion order)
    for (int i = tid; i < C; i += kRankThreads) {
        int l = i;
        const uint32_t k = key[i];
        for (int j = 0; j < i; ++j) if (key[j] == k) { l = j; break; }
        lead[i] = l;
        p.out_leader[base + i] = l;
    }
    __syncthreads();
    // 3. order-dependent fold of every group, by its leader
    for (int i = tid; i < C; i += kRankThreads) {
        if (lead[i] != i) { fin[i] = 0.0; continue; }
        double f = sc[i];
        for (int j = i + 1; j < C; ++j) {
            if (lead[j] != i) continue;
            if (p.mode == 0) f = __dmul_rn(__ddiv_rn(__dadd_rn(f, sc[j]), 2.0), 1.1);      // ranker.py:183-184
            else f = p.kind[base + j] < 4 ? sc[j] : __dadd_rn(f, sc[j]);                   // reranker.py:94-115
        }
        fin[i] = f;
    }
    __syncthreads();
    // 4. stable rank of the leaders by merged score, descending
    for (int i = tid; i < C; i += kRankThreads) {
        if (lead[i] != i) { rnk[i] = -1; continue; }
        const double f = fin[i];
        int r = 0;
        for (int j = 0; j < C; ++j) {
            if (lead[j] != j) continue;
            const double g = fin[j];
            r += (g > f || (g == f && j < i)) ? 1 : 0;
        }
        rnk[i] = r;
    }
    __syncthreads();
    // 5. per-file cap in sorted order, then the total cap
    for (int i = tid; i < C; i += kRankThreads) {
        int kept = 0;
        if (rnk[i] >= 0) {
            int same = 0;
            for (int j = 0; j < C; ++j) same += (rnk[j] >= 0 && fil[j] == fil[i] && rnk[j] < rnk[i]) ? 1 : 0;
            kept = (p.max_per_file <= 0 || same < p.max_per_file) ? 1 : 0;
        }
        pos[i] = kept ? 0 : -1;
    }
    __syncthreads();
    if (tid == 0) { s_nout = 0; s_min = INFINITY; s_max = -INFINITY; }
    __syncthreads();
    for (int i = tid; i < C; i += kRankThreads) {
        if (pos[i] < 0) continue;
        int before = 0;
        for (int j = 0; j < C; ++j) before += (pos[j] >= 0 && rnk[j] < rnk[i]) ? 1 : 0;
        // pos[] entries are only ever 0 or -1 until this loop finishes for everyone, so reading pos[j] >= 0 is stable
        lead[i] = -1 - before;   // stash (lead is no longer needed for leaders): output slot = before
    }
    __syncthreads();
    const size_t ob = (size_t)q * p.max_total;
    for (int i = tid; i < C; i += kRankThreads) {
        if (pos[i] < 0) continue;
        const int slot = -1 - lead[i];
        if (slot >= p.max_total) continue;
        atomicAdd(&s_nout, 1);
        // merged signals of the group: per-signal max over the members that carry the signal (ranker.py:195-199)
        double sig[kRankSignals]; uint32_t m;
        rank_candidate(p, base + i, w, sig, &m);
        int members = 1;
        bool any_vec = p.kind[base + i] == 4, any_graph = !any_vec;
        for (int j = i + 1; j < C; ++j) {
            if (p.out_leader[base + j] != i) continue;
            ++members;
            double s2[kRankSignals]; uint32_t m2;
            rank_candidate(p, base + j, w, s2, &m2);
#pragma unroll
            for (int s = 0; s < kRankSignals; ++s) {
                if (!(m2 & (1u << s))) continue;
                sig[s] = (m & (1u << s)) ? fmax(sig[s], s2[s]) : s2[s];
            }
            m |= m2;
            if (p.kind[base + j] == 4) any_vec = true; else any_graph = true;
        }
        p.out_index[ob + slot] = i;
        p.out_score[ob + slot] = fin[i];
        p.out_sigmask[ob + slot] = (uint8_t)m;
        uint8_t src;
        if (p.mode == 0) src = members > 1 ? 2 : (p.kind[base + i] == 4 ? 1 : 0);
        else src = (any_vec && members > 1) ? 2 : (any_vec ? 1 : 0);   // graph-on-graph replacement stays "graph"
        p.out_source[ob + slot] = src;
#pragma unroll
        for (int s = 0; s < kRankSignals; ++s) p.out_signals[(ob + slot) * kRankSignals + s] = (m & (1u << s)) ? sig[s] : 0.0;
    }
    __syncthreads();
    const int nout = s_nout;
    if (tid == 0) p.out_count[q] = nout;
    // 6. min-max normalisation of the emitted list (reranker.py:29-70): all-equal -> 1.0
    if (tid == 0) {
        double lo = INFINITY, hi = -INFINITY;
        for (int s = 0; s < nout; ++s) { const double v = p.out_score[ob + s]; lo = fmin(lo, v); hi = fmax(hi, v); }
        s_min = lo; s_max = hi;
    }
    __syncthreads();
    const double range = __dsub_rn(s_max, s_min);
    for (int s = tid; s < nout; s += kRankThreads)
        p.out_norm[ob + s] = range == 0.0 ? 1.0 : __ddiv_rn(__dsub_rn(p.out_score[ob + s], s_min), range);
}


// ---------------------------------------------------------------------------------------------------------
// Fused search -> rank (QueryEngine._execute_vector_search + rank_results glue,
// reference query/engine.py:315-346,176-181 and ranking/ranker.py:150-169).  The search leaves its top-k rows and
// float64 scores on the device; this kernel turns them into vector-hit candidates of K3 WITHOUT a host hop:
// key / file / centrality-key ids, len(content) and the presence flags are per-row columns written at upsert, the
// entity-name match (scorer.py:91-96: exact -> 1.0, any query entity contained in the name -> 0.5) is evaluated here on
// the lower-cased UTF-8 names of a device-side string pool, the centrality degree is looked up in the query's short
// (id, total_degree) table.  Graph candidates were packed by the host into the same arrays, ahead of the hits.
// ---------------------------------------------------------------------------------------------------------
struct RankGatherParams {
    int k;                          // hit slots per query
    const int32_t* offsets;         // [Q + 1] combined candidate offsets
    const int32_t* n_graph;         // [Q] graph candidates of the query; its hits start right after them
    const double* hit_scores;       // [Q][k]
    const int64_t* hit_rows;        // [Q][k] global rows
    const uint32_t* hit_counts;     // [Q]
    int64_t row_base;
    int64_t attr_rows;              // rows covered by the attribute columns
    const uint32_t* row_key; const uint32_t* row_file; const uint32_t* row_cent; const uint32_t* row_name;
    const int32_t* row_clen; const uint8_t* row_flags;
    const uint32_t* name_off;       // [n_names + 1] byte offsets into name_bytes
    const uint8_t* name_bytes;
    uint32_t n_names;
    const int32_t* ent_off;         // [Q + 1] -> entity indices of the query
    const uint32_t* ent_str_off;    // [n_entities + 1] byte offsets into ent_bytes
    const uint8_t* ent_bytes;
    const int32_t* cen_off;         // [Q + 1]
    const uint32_t* cen_id; const int32_t* cen_deg;
    uint8_t* kind; uint32_t* key_id; uint32_t* file_id; int32_t* depth; double* entity_match; int32_t* degree;
    uint8_t* flags; int32_t* content_len; double* vscore;
    int32_t* counts;                // [Q] out: n_graph + hits (second segment: += its hits)
    int32_t* error;                 // set to 1 when a hit row has no attributes
    // second segment (hits of another collection, e.g. summaries behind the code hits, query/engine.py:331-344): the searched
    // queries are a subset, sel[j] = batch query of searched query j; its candidates go right behind what counts[] already holds
    const int32_t* sel;             // nullptr = first segment (searched query j is batch query j)
};

__device__ __forceinline__ bool bytes_equal(const uint8_t* a, const uint8_t* b, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) if (a[i] != b[i]) return false;
    return true;
}

__global__ void __launch_bounds__(128) rank_gather_kernel(const RankGatherParams p) {
    const int j = blockIdx.x;                               // searched query
    const int q = p.sel ? p.sel[j] : j;                     // batch query
    const int hits = (int)min(p.hit_counts[j], (uint32_t)p.k);
    const int before = p.sel ? p.counts[q] : p.n_graph[q];  // candidates already in place
    __syncthreads();                                        // everyone has read counts[q] before it is updated
    if (threadIdx.x == 0) p.counts[q] = before + hits;
    const int e0 = p.ent_off[q], e1 = p.ent_off[q + 1];
    const int c0 = p.cen_off[q], c1 = p.cen_off[q + 1];
    for (int s = threadIdx.x; s < hits; s += blockDim.x) {
        const int g = p.offsets[q] + before + s;
        const int64_t row = p.hit_rows[(size_t)j * p.k + s] - p.row_base;
        if (row < 0 || row >= p.attr_rows) { *p.error = 1; p.kind[g] = 4; p.key_id[g] = 0xFFFFFFFFu; p.file_id[g] = 0xFFFFFFFFu;
            p.depth[g] = 0; p.entity_match[g] = 0.0; p.degree[g] = -1; p.flags[g] = 0; p.content_len[g] = -1; p.vscore[g] = 0.0; continue; }
        const uint32_t nid = p.row_name[row];
        double em = 0.0;
        if (nid < p.n_names) {
            const uint8_t* name = p.name_bytes + p.name_off[nid];
            const uint32_t nlen = p.name_off[nid + 1] - p.name_off[nid];
            bool sub = false;
            for (int e = e0; e < e1; ++e) {
                const uint8_t* es = p.ent_bytes + p.ent_str_off[e];
                const uint32_t el = p.ent_str_off[e + 1] - p.ent_str_off[e];
                if (el == nlen && bytes_equal(name, es, el)) { em = 1.0; break; }
                if (!sub && el <= nlen) {
                    for (uint32_t pos = 0; pos + el <= nlen; ++pos)
                        if (bytes_equal(name + pos, es, el)) { sub = true; break; }
                }
            }
            if (em == 0.0 && sub) em = 0.5;
        }
        const uint32_t cid = p.row_cent[row];
        int deg = -1;
        for (int c = c0; c < c1; ++c) if (p.cen_id[c] == cid) { deg = p.cen_deg[c] < 0 ? 0 : p.cen_deg[c]; break; }
        p.kind[g] = 4;
        p.key_id[g] = p.row_key[row];
        p.file_id[g] = p.row_file[row];
        p.depth[g] = 0;
        p.entity_match[g] = em;
        p.degree[g] = deg;
        p.flags[g] = p.row_flags[row];
        p.content_len[g] = p.row_clen[row];
        p.vscore[g] = p.hit_scores[(size_t)j * p.k + s];
    }
}

// row attribute scatter (upsert of ranking attributes)
struct RankAttrScatter {
    const int64_t* rows; int n;
    const uint32_t* key; const uint32_t* file; const uint32_t* cent; const uint32_t* name; const int32_t* clen; const uint8_t* flags;
    uint32_t* row_key; uint32_t* row_file; uint32_t* row_cent; uint32_t* row_name; int32_t* row_clen; uint8_t* row_flags;
};
__global__ void rank_attr_scatter_kernel(const RankAttrScatter p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int64_t r = p.rows[i];
    p.row_key[r] = p.key[i]; p.row_file[r] = p.file[i]; p.row_cent[r] = p.cent[i]; p.row_name[r] = p.name[i];
    p.row_clen[r] = p.clen[i]; p.row_flags[r] = p.flags[i];
}

__host__ __device__ inline size_t rank_smem_bytes(int max_c) { return (size_t)max_c * (8 + 8 + 4 + 4 + 4 + 4 + 4); }

}  // namespace lvs
