// api.cu — version / build identification of libmisob200.so.
#include "common.cuh"

extern "C" int mb_abi_version(void) { return MB_ABI_VERSION; }

extern "C" const char* mb_build_info(void) {
    return "libmisob200 sm_100a cuda " __DATE__
           ";kernels=k_nms_mask,k_nms_sweep,k_rank_in_segment,k_roi_align_staged,k_roi_align_direct,"
           "k_rpn_hist,k_rpn_select,k_rpn_decode,k_det_candidates,k_crop_plan,k_crop_gather";
}
