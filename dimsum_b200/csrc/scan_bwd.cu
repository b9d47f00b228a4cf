// Selective-scan backward for sm_100a (placeholder until the kernel lands in this round).
#include "common.cuh"

using namespace dimsum;

extern "C" int dimsum_selective_scan_bwd(const dimsum_scan_bwd_params *p, void *stream) {
    (void)p; (void)stream;
    return fail(DIMSUM_ERR_UNSUPPORTED, "selective_scan_bwd: not implemented yet");
}
