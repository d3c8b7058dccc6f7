// Shared device helpers for the lattice-b200 vector-search kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/lvs.h"   // LVS_STORAGE_*, LVS_METRIC_*, LVS_DT_* shared with the C ABI

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "lattice-b200 kernels are written for sm_100a (B200) only"
#endif

namespace lvs {

constexpr int kWarp = 32;
constexpr int kMaxFilterCols = 8;
constexpr uint32_t kAnyCode = 0xFFFFFFFFu;   // filter column not constrained
constexpr uint32_t kNullCode = 0u;           // payload key missing / null: never matches a MatchValue

// ---------------------------------------------------------------------------------------------
// Ordering keys.  The scan ranks by a 64-bit key whose unsigned order is (score desc, row asc):
//   key = orderable(float score) << 32 | (0xFFFFFFFF - local_row).   key 0 == "empty slot".
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_orderable(float s) {
    if (!(s == s)) s = -INFINITY;                       // NaN ranks last
    uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return ((uint64_t)f32_orderable(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu); }
__device__ __forceinline__ float key_score(uint64_t key) { return f32_from_orderable((uint32_t)(key >> 32)); }

// ---------------------------------------------------------------------------------------------
// Shared-memory address, mbarrier and bulk-copy (TMA, SASS UBLKCP) wrappers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking form (try_wait may suspend the thread for a while): for a thread that watches several barriers
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// global -> shared bulk copy, completion counted in bytes on `bar`.  src, dst and bytes are 16 B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Small warp utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(0xFFFFFFFFu, lo, m);
    hi = __shfl_xor_sync(0xFFFFFFFFu, hi, m);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_sync(0xFFFFFFFFu, lo, src);
    hi = __shfl_sync(0xFFFFFFFFu, hi, src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_up_sync(0xFFFFFFFFu, lo, d);
    hi = __shfl_up_sync(0xFFFFFFFFu, hi, d);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    return __longlong_as_double((long long)shfl_xor_u64((uint64_t)__double_as_longlong(v), m));
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

// Running top-(32*KPL) of 64-bit keys held in registers by one warp (KPL keys per lane).
template <int KPL>
struct WarpTopK {
    uint64_t key[KPL];
    uint64_t thr;   // current minimum over the 32*KPL slots (warp-uniform)

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < KPL; ++j) key[j] = 0ull;
        thr = 0ull;
    }
    // warp-uniform call: replace the current minimum by k (k > thr) and recompute the minimum
    __device__ __forceinline__ void insert(uint64_t k, int lane) {
        bool has = false;
#pragma unroll
        for (int j = 0; j < KPL; ++j) has |= (key[j] == thr);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, has);
        const int owner = __ffs(m) - 1;
        if (lane == owner) {
            bool done = false;
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                if (!done && key[j] == thr) { key[j] = k; done = true; }
            }
        }
        uint64_t lm = key[0];
#pragma unroll
        for (int j = 1; j < KPL; ++j) lm = key[j] < lm ? key[j] : lm;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            uint64_t other = shfl_xor_u64(lm, o);
            lm = other < lm ? other : lm;
        }
        thr = lm;
    }
};

// element i of a host-typed array (LVS_DT_*) as float64
__device__ __forceinline__ double load_as_f64(const void* src, int dtype, size_t i) {
    if (dtype == LVS_DT_F64) return reinterpret_cast<const double*>(src)[i];
    if (dtype == LVS_DT_F32) return (double)reinterpret_cast<const float*>(src)[i];
    return (double)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
}

// single-instruction approximations (MUFU): 2^x and 1/x
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// two floats -> a packed bf16 pair (round to nearest even), element 0 in the low half
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// bf16 pair (packed in a 32-bit word, little endian: element 0 in the low half) -> two floats
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// 32 keys, one per lane -> sorted descending across the lanes (bitonic network on shuffles)
__device__ __forceinline__ uint64_t warp_sort_desc(uint64_t x, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint64_t o = shfl_xor_u64(x, j);
            const bool keep_max = (((lane & k) == 0) == ((lane & j) == 0));
            x = keep_max ? (o > x ? o : x) : (o < x ? o : x);
        }
    }
    return x;
}
// a bitonic sequence of 32 keys -> sorted descending
__device__ __forceinline__ uint64_t warp_bitonic_merge_desc(uint64_t x, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint64_t o = shfl_xor_u64(x, j);
        x = ((lane & j) == 0) ? (o > x ? o : x) : (o < x ? o : x);
    }
    return x;
}

// number of keys greater than v in a list of 32 keys sorted descending (binary search: six probes)
__device__ __forceinline__ uint32_t count_greater_desc32(const uint64_t* list, uint64_t v) {
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t step = 16; step >= 1; step >>= 1) if (list[pos + step - 1] > v) pos += step;
    return pos + (list[pos] > v ? 1u : 0u);
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase stamp of the profiling aid (option "dbg_times"): slot 0 keeps the earliest, the others the latest time any CTA got there
__device__ __forceinline__ void stamp(unsigned long long* times, int slot) {
    if (times == nullptr) return;
    if (slot == 0) atomicMin(times, global_ns()); else atomicMax(times + slot, global_ns());
}

}  // namespace lvs
