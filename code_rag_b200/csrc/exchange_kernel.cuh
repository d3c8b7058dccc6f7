// Fused exchange + merge over NVLink peer memory (K5').
//
// The sharded path has one exchange step: every rank contributes its Q x k (score, row, tie) lists and needs the union
// of all ranks' lists to merge the global top-k.  Instead of an NCCL all-gather followed by a merge kernel, ONE kernel
// per rank does both: it stores this rank's packed block straight into every peer's gather buffer (P2P stores through
// NVSwitch; the buffers are CUDA-IPC mapped once at start-up), publishes a sequence-numbered flag with system-scope
// release semantics, waits (acquire) for the flags of all peers, and merges.  Messages are 24*Q*k bytes per rank
// (240 B for Q = 1, k = 10), so the step is latency-bound: what this removes is a second launch and the collective's
// protocol overhead.  Buffers are double-buffered by sequence parity: a rank can only be one exchange ahead of a peer,
// because finishing exchange s requires the peer's block for s, which the peer writes after it has merged s-1.
#pragma once
#include "common.cuh"

namespace lvs {

constexpr int kMaxRanks = 8;

struct ExchangeParams {
    const int64_t* local;            // [3][Q][k] packed block of this rank (float64 score bits | global rows | tie keys)
    int Q, k, world, rank;
    int64_t* peer_data[kMaxRanks];   // peer p's gather area of this slot; this rank writes at + rank * blk_stride
    uint64_t* peer_flags[kMaxRanks]; // peer p's flag array of this slot; this rank writes element [rank]
    const int64_t* my_data;          // this rank's gather area of this slot: [world][blk_stride]
    const uint64_t* my_flags;        // this rank's flags of this slot: [world]
    uint64_t seq;
    size_t blk_stride;               // elements between the blocks of consecutive ranks
    uint32_t* done_counter;          // zero between launches
    int64_t* out;                    // [3][Q][k] merged result
    uint32_t* out_counts;            // [Q]
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

__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeParams p) {
    extern __shared__ __align__(16) uint8_t xsm[];
    __shared__ uint32_t s_last, s_nvalid;
    const int tid = threadIdx.x;
    const size_t n = (size_t)3 * p.Q * p.k;
    // ---- 1. publish: this rank's block -> every rank's gather buffer (its own included) ----
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int64_t v = p.local[i];
        for (int r = 0; r < p.world; ++r) p.peer_data[r][(size_t)p.rank * p.blk_stride + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        s_last = (atomicAdd(p.done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            *p.done_counter = 0;
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) st_release_sys(p.peer_flags[r] + p.rank, p.seq);
        }
    }
    // ---- 2. wait for every rank's flag of this sequence number ----
    if (tid < p.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(p.my_flags + tid) < p.seq) {
            if (clock64() - t0 > 20000000000ll) { *p.err = 1u; break; }   // ~10 s: a peer died; do not hang the GPU
        }
    }
    __syncthreads();
    // ---- 3. merge: (score desc, tie asc, row asc) by rank counting, one query at a time ----
    const int m = p.world * p.k;
    double* s = reinterpret_cast<double*>(xsm);
    int64_t* r = reinterpret_cast<int64_t*>(xsm + (size_t)m * 8);
    uint64_t* t = reinterpret_cast<uint64_t*>(xsm + (size_t)m * 16);
    const size_t qk = (size_t)p.Q * p.k;
    for (int qi = blockIdx.x; qi < p.Q; qi += gridDim.x) {
        if (tid == 0) s_nvalid = 0;
        __syncthreads();
        for (int i = tid; i < m; i += blockDim.x) {
            const int g = i / p.k, j = i % p.k;
            const int64_t* blk = p.my_data + (size_t)g * p.blk_stride;
            const size_t o = (size_t)qi * p.k + j;
            s[i] = __longlong_as_double(__ldcg(blk + o));
            r[i] = __ldcg(blk + qk + o);
            t[i] = (uint64_t)__ldcg(blk + 2 * qk + o);
            if (r[i] >= 0) atomicAdd(&s_nvalid, 1u);
        }
        __syncthreads();
        for (int i = tid; i < m; i += blockDim.x) {
            if (r[i] < 0) continue;
            uint32_t rank = 0;
            for (int o = 0; o < m; ++o) {
                if (r[o] < 0) continue;
                const bool better = (s[o] > s[i]) || (s[o] == s[i] && (t[o] < t[i] || (t[o] == t[i] && r[o] < r[i])));
                rank += better ? 1u : 0u;
            }
            if (rank < (uint32_t)p.k) {
                const size_t o = (size_t)qi * p.k + rank;
                p.out[o] = __double_as_longlong(s[i]);
                p.out[qk + o] = r[i];
                p.out[2 * qk + o] = (int64_t)t[i];
            }
        }
        const uint32_t nout = min(s_nvalid, (uint32_t)p.k);
        for (int j = nout + tid; j < p.k; j += blockDim.x) {
            const size_t o = (size_t)qi * p.k + j;
            p.out[o] = 0; p.out[qk + o] = -1; p.out[2 * qk + o] = 0;
        }
        if (tid == 0) p.out_counts[qi] = nout;
        __syncthreads();
    }
}

}  // namespace lvs
