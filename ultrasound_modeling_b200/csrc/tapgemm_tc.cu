// placeholder until the tcgen05 kernels land (replaced below)
#include "tbi_common.cuh"
bool tbi_tapgemm_tc_supported(const tbi_tapgemm*, const char** why) { *why = "not built"; return false; }
bool tbi_tapwgrad_tc_supported(const tbi_tapwgrad*, const char** why) { *why = "not built"; return false; }
int tbi_tapgemm_tc(const tbi_tapgemm*, cudaStream_t) { return tbi_set_error(TBI_ERR_UNSUPPORTED, "tc"); }
int tbi_tapwgrad_tc(const tbi_tapwgrad*, cudaStream_t) { return tbi_set_error(TBI_ERR_UNSUPPORTED, "tc"); }
int64_t tbi_tapwgrad_tc_workspace(const tbi_tapwgrad*) { return 0; }
