// Host-side helpers shared by the translation units of liblattice_b200.so (defined in lvs_api.cu).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct lvs_collection;

int lvs_fail(int code, const char* fmt, ...);              // sets the thread-local message behind lvs_last_error(), returns code
bool lvs_lib_ready();                                      // lvs_init() succeeded
int lvs_lib_sm_count();
size_t lvs_lib_smem_optin();
void lvs_lib_bind_thread();                                // make the calling thread current on the library's device
PFN_cuTensorMapEncodeTiled_v12000 lvs_lib_encode_tiled();  // nullptr (with the message set) when the driver entry point is missing
// K4 from vectors that are already on the device (the encoder's output): rows / codes / ties are host arrays as in lvs_upsert
int lvs_upsert_device_vectors(lvs_collection* c, const void* d_vecs, int dtype, int64_t n, const int64_t* rows, const uint32_t* codes,
                              const uint64_t* ties);

#define LVS_CU(expr)                                                                                            \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return lvs_fail(e__ == cudaErrorMemoryAllocation ? LVS_ENOMEM : LVS_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                                       \
    } while (0)
