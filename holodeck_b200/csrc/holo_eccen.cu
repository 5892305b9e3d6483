// holo_eccen.cu -- K5 eccentric harmonic sum (placeholder until the kernel lands; see include/holo_b200.h)
#include <cuda_runtime.h>

#include "holo_api.cuh"

extern "C" {

int64_t holo_eccen_workspace_bytes(int M, int Q, int Z, int F, int nharms, int nreals) {
    (void)M; (void)Q; (void)Z; (void)F; (void)nharms; (void)nreals;
    return 256;
}

int holo_sam_calc_gwb_single_eccen(const double*, const double*, const double*, const double*,
                                   const double*, const double*, const double*, const double*, int,
                                   int, int, int, int, int, int, int64_t, uint64_t, double*, void*,
                                   int64_t, void*) {
    holo::set_error("holo_sam_calc_gwb_single_eccen: not implemented yet");
    return HOLO_ERR_ARG;
}

}  // extern "C"
