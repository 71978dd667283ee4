// placeholder — replaced by the real kernels (see include/misob200.h)
#include "common.cuh"
extern "C" size_t mb_det_workspace_bytes(const mb_det_params*) { return 0; }
extern "C" int mb_det_postprocess(const mb_det_params*, const float*, const float*, const float*, const int32_t*, int32_t, float*, float*, float*, int64_t*, int32_t*, void*, size_t, mb_stream_t) { return MB_ERR_UNSUPPORTED; }
