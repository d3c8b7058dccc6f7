// Launch plumbing of the fused scan kernel, shared by the translation units that instantiate it (one per storage / metric, so
// that build.py can compile them in parallel) and by lvs_api.cu, which only sees the three entry points declared at the end.
#pragma once
#include <cstring>
#include <mutex>

#include "scan_kernel.cuh"

namespace lvs {

template <typename T, int QT, int KPL, bool NORM, bool FILTER, bool INLINE>
static cudaError_t launch_scan_inst(const ScanParams& p, const FinalizeParams& fp, const ExchangeParams& xp, const InlineQueries& iq, int grid,
                                    size_t smem, cudaStream_t st, size_t smem_optin) {
    static std::once_flag once;
    static cudaError_t once_err = cudaSuccess;
    auto kfn = scan_topk_kernel<T, QT, KPL, NORM, FILTER, INLINE>;
    std::call_once(once, [&] {
        // the kernel also has a little static shared memory (padded to the ring's 1024-byte alignment): the opt-in limit covers both
        cudaFuncAttributes fa;
        once_err = cudaFuncGetAttributes(&fa, kfn);
        if (once_err == cudaSuccess)
            once_err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_optin - fa.sharedSizeBytes));
    });
    if (once_err != cudaSuccess) return once_err;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kScanThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // PDL: see the kernel's header
    la[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = la; cfg.numAttrs = p.pdl ? 1 : 0;
    if constexpr (INLINE) {
        return cudaLaunchKernelEx(&cfg, kfn, p, fp, xp, iq);
    } else {
        NoInlineQueries none;
        memset(&none, 0, sizeof(none));
        return cudaLaunchKernelEx(&cfg, kfn, p, fp, xp, none);
    }
}

template <typename T, bool NORM, bool FILTER>
static cudaError_t launch_scan_tnf(int qt, int kpl, const ScanParams& p, const FinalizeParams& fp, const ExchangeParams& xp, const InlineQueries& iq,
                                   int grid, size_t smem, cudaStream_t st, size_t smem_optin) {
    if constexpr (!FILTER) {
        // one query slot, no filter: the instantiations that carry the query in their parameter block (lvs_scan_inline_available)
#define LVS_CASE(K_) if (p.q_inline && qt == 1 && kpl == K_) return launch_scan_inst<T, 1, K_, NORM, false, true>(p, fp, xp, iq, grid, smem, st, smem_optin);
        LVS_CASE(1) LVS_CASE(2) LVS_CASE(4) LVS_CASE(8)
#undef LVS_CASE
    }
    if (p.q_inline) return cudaErrorInvalidValue;
#define LVS_CASE(Q_, K_) if (qt == Q_ && kpl == K_) return launch_scan_inst<T, Q_, K_, NORM, FILTER, false>(p, fp, xp, iq, grid, smem, st, smem_optin);
    LVS_CASE(1, 1) LVS_CASE(1, 2) LVS_CASE(1, 4) LVS_CASE(1, 8)
    LVS_CASE(2, 1) LVS_CASE(2, 2) LVS_CASE(2, 4)
    LVS_CASE(4, 1) LVS_CASE(4, 2)
#undef LVS_CASE
    return cudaErrorInvalidValue;
}

}  // namespace lvs

#define LVS_SCAN_ENTRY(name) \
    cudaError_t name(int qt, int kpl, bool filter, const lvs::ScanParams& p, const lvs::FinalizeParams& fp, const lvs::ExchangeParams& xp, \
                     const lvs::InlineQueries& iq, int grid, size_t smem, cudaStream_t st, size_t smem_optin)
LVS_SCAN_ENTRY(lvs_launch_scan_f32);        // fp32 shards (cosine rows are stored unit-norm; dot uses the same kernel)
LVS_SCAN_ENTRY(lvs_launch_scan_bf16_cos);   // bf16 shards, cosine: the row norm is fused into the scan
LVS_SCAN_ENTRY(lvs_launch_scan_bf16_dot);   // bf16 shards, dot product
