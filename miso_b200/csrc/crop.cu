// placeholder — replaced by the real kernels (see include/misob200.h)
#include "common.cuh"
extern "C" int mb_crop_plan(const mb_crop_params*, const float*, const float*, const int32_t*, int32_t*, float*, int32_t*, int64_t*, int64_t*, mb_stream_t) { return MB_ERR_UNSUPPORTED; }
extern "C" int mb_crop_gather(const mb_crop_params*, const int32_t*, const int32_t*, const int64_t*, const int64_t*, uint8_t*, int64_t, mb_stream_t) { return MB_ERR_UNSUPPORTED; }
