// placeholder — replaced by the real kernels (see include/misob200.h)
#include "common.cuh"
extern "C" size_t mb_rpn_workspace_bytes(const mb_rpn_params*) { return 0; }
extern "C" int mb_rpn_proposals(const mb_rpn_params*, float*, float*, int32_t*, int64_t*, void*, size_t, mb_stream_t) { return MB_ERR_UNSUPPORTED; }
