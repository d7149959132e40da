// ABI version / status strings for the hgs_raster C ABI (include/hgs_raster.h).
#include "hgs_common.cuh"
#include "../../include/hgs_raster.h"

unsigned long long g_hgs_launches = 0;

HGS_API int hgs_abi_version(void) { return HGS_ABI_VERSION; }

HGS_API unsigned long long hgs_debug_launch_count(void) { return __atomic_load_n(&g_hgs_launches, __ATOMIC_RELAXED); }

HGS_API const char* hgs_status_string(int status) {
    switch (status) {
        case 0: return "success";
        case HGS_ERR_INVALID_ARG: return "hgs: invalid argument";
        case HGS_ERR_TOO_LARGE: return "hgs: problem too large for 32-bit intersection indices";
        case HGS_ERR_WORKSPACE: return "hgs: workspace too small";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "hgs: unknown status";
}
