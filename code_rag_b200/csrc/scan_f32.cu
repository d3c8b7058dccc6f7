// Instantiations of the fused scan kernel (scan_kernel.cuh) for one storage / metric; see scan_launch.cuh.
#include "scan_launch.cuh"

LVS_SCAN_ENTRY(lvs_launch_scan_f32) {
    return filter ? lvs::launch_scan_tnf<float, false, true>(qt, kpl, p, fp, xp, iq, grid, smem, st, smem_optin)
                  : lvs::launch_scan_tnf<float, false, false>(qt, kpl, p, fp, xp, iq, grid, smem, st, smem_optin);
}
