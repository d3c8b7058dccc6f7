// Instantiations of the fused scan kernel (scan_kernel.cuh) for one storage / metric; see scan_launch.cuh.
#include "scan_launch.cuh"

LVS_SCAN_ENTRY(lvs_launch_scan_bf16_dot) {
    return filter ? lvs::launch_scan_tnf<__nv_bfloat16, false, true>(qt, kpl, p, fp, xp, iq, grid, smem, st, smem_optin)
                  : lvs::launch_scan_tnf<__nv_bfloat16, false, false>(qt, kpl, p, fp, xp, iq, grid, smem, st, smem_optin);
}
