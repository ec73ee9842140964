// placeholder - replaced by the real VJP kernels
#include "mho_common.cuh"
#include "mho_internal.h"
extern "C" int mho_cheb_backward(mho_ctx_t*, const mho_batch_t*, const mho_layer_t*, int32_t, const float*, const float*,
                                 const void*, const float*, float*, float*, float*, mho_stream_t) {
    mho_set_error("mho_cheb_backward: not implemented yet");
    return MHO_ERR_INVALID;
}
