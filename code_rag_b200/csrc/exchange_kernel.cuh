// Exchange + merge over NVLink peer memory (K5').
//
// The sharded path has one exchange step: every rank contributes its Q x k (score, row, tie) lists and needs the union
// of all ranks' lists to merge the global top-k.  Instead of an NCCL all-gather followed by a merge kernel, the rank's
// block is stored straight into every peer's gather buffer (P2P stores through NVSwitch; the buffers are CUDA-IPC mapped
// once at start-up), a sequence-numbered flag is published with system-scope release semantics, the flags of all peers are
// awaited (acquire) and the lists are merged - all inside one kernel.  Two users:
//   * exchange_merge_kernel: its own launch behind the tensor-core path (batches of queries);
//   * the fused scan kernel (scan_kernel.cuh): the CTA that finishes a query publishes, waits and merges in place, so a
//     sharded single-query search is ONE kernel per GPU (scan + exact rescoring + exchange + merge).
// Messages are 24*Q*k + 8*Q bytes per rank (248 B for Q = 1, k = 10): latency-bound.  The per-query flags (exactness not
// proven on some shard) travel with the block and are OR-ed, so every rank sees the same merged flags.  Buffers are double-
// buffered by sequence parity: a rank can only be one exchange ahead of a peer, because finishing exchange s requires the
// peer's block for s, which the peer writes after it has merged s-1.
#pragma once
#include "common.cuh"

namespace lvs {

constexpr int kMaxRanks = 8;
constexpr int32_t kFlagExchangeTimeout = 2;   // LVS_FLAG_EXCHANGE: a peer's block did not arrive in time (result not trustworthy)

struct ExchangeParams {
    const int64_t* local;            // [3][Q][k] packed block of this rank (float64 score bits | global rows | tie keys); standalone kernel only
    const int32_t* local_flags;      // [Q] this rank's per-query flags (nullptr = none); standalone kernel only
    int Q, k, world, rank;
    int64_t* peer_data[kMaxRanks];   // peer p's gather area of this slot; this rank writes at + rank * blk_stride
    uint64_t* peer_flags[kMaxRanks]; // peer p's flag array of this slot; this rank writes element [rank]
    const int64_t* my_data;          // this rank's gather area of this slot: [world][blk_stride]; a block is [3][Q][k] then [Q] flags
    const uint64_t* my_flags;        // this rank's flags of this slot: [world]
    uint64_t seq;
    size_t blk_stride;               // elements between the blocks of consecutive ranks
    uint32_t* done_counter;          // zero between launches
    int64_t* out;                    // [3][Q_out][k] merged result
    uint32_t* out_counts;            // [Q_out]
    int32_t* out_flags;              // [Q_out] OR of every rank's flags (| kFlagExchangeTimeout), or nullptr
    int Q_out, q_out0;               // the merged result of block query q goes to query q_out0 + q of a [Q_out] layout
    uint32_t* err;                   // set to 1 if a peer's flag did not arrive in time
};

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Raise this rank's flag in every rank's flag array (one thread).
__device__ __forceinline__ void exchange_signal(const ExchangeParams& p) {
    __threadfence_system();
    for (int r = 0; r < p.world; ++r) st_release_sys(p.peer_flags[r] + p.rank, p.seq);
}

// Threads tid < world wait for rank tid's flag of this sequence number.  Returns false (and sets *err) on a timeout.
__device__ __forceinline__ bool exchange_wait_one(const ExchangeParams& p, int r) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p.my_flags + r) < p.seq) {
        if (clock64() - t0 > 20000000000ll) { *p.err = 1u; return false; }   // ~10 s: a peer died; do not hang the GPU
        __nanosleep(64);
    }
    return true;
}

// Merge block query q of all ranks into the output (score desc, tie asc, row asc) by rank counting.  nthreads threads of one
// CTA call it together; bar() synchronises exactly those threads.  xsm: world * k * 24 bytes of shared memory.
template <typename Bar>
__device__ __forceinline__ void exchange_merge_query(const ExchangeParams& p, int q, uint8_t* xsm, uint32_t* s_nvalid, int tid, int nthreads,
                                                     bool timed_out, Bar bar) {
    const int m = p.world * p.k;
    double* s = reinterpret_cast<double*>(xsm);
    int64_t* r = reinterpret_cast<int64_t*>(xsm + (size_t)m * 8);
    uint64_t* t = reinterpret_cast<uint64_t*>(xsm + (size_t)m * 16);
    const size_t qk = (size_t)p.Q * p.k;
    const size_t oqk = (size_t)p.Q_out * p.k;
    const int qo = p.q_out0 + q;
    if (tid == 0) *s_nvalid = 0;
    bar();
    for (int i = tid; i < m; i += nthreads) {
        const int g = i / p.k, j = i % p.k;
        const int64_t* blk = p.my_data + (size_t)g * p.blk_stride;
        const size_t o = (size_t)q * p.k + j;
        s[i] = __longlong_as_double(__ldcg(blk + o));
        r[i] = __ldcg(blk + qk + o);
        t[i] = (uint64_t)__ldcg(blk + 2 * qk + o);
        if (r[i] >= 0) atomicAdd(s_nvalid, 1u);
    }
    bar();
    // every rank's list arrives sorted best-first (padding last), so the global rank of an entry is the sum over the lists of how
    // many of their entries beat it: a binary search per list (world * log2 k steps) instead of a scan of all world * k entries
    for (int i = tid; i < m; i += nthreads) {
        if (r[i] < 0) continue;
        const double si = s[i]; const uint64_t ti = t[i]; const int64_t ri = r[i];
        uint32_t rank = 0;
        for (int g = 0; g < p.world; ++g) {
            int lo = 0, hi = p.k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1, o = g * p.k + mid;
                const bool better = r[o] >= 0 && ((s[o] > si) || (s[o] == si && (t[o] < ti || (t[o] == ti && r[o] < ri))));
                if (better) lo = mid + 1; else hi = mid;
            }
            rank += (uint32_t)lo;
        }
        if (rank < (uint32_t)p.k) {
            const size_t o = (size_t)qo * p.k + rank;
            p.out[o] = __double_as_longlong(si);
            p.out[oqk + o] = ri;
            p.out[2 * oqk + o] = (int64_t)ti;
        }
    }
    const uint32_t nout = min(*s_nvalid, (uint32_t)p.k);
    for (int j = nout + tid; j < p.k; j += nthreads) {
        const size_t o = (size_t)qo * p.k + j;
        p.out[o] = 0; p.out[oqk + o] = -1; p.out[2 * oqk + o] = 0;
    }
    if (tid == 0) {
        p.out_counts[qo] = nout;
        if (p.out_flags != nullptr) {
            int32_t f = timed_out ? kFlagExchangeTimeout : 0;
            for (int g = 0; g < p.world; ++g) f |= (int32_t)__ldcg(p.my_data + (size_t)g * p.blk_stride + 3 * qk + q);
            p.out_flags[qo] = f;
        }
    }
    bar();
}

}  // namespace lvs
