// placeholder - replaced by the real optimizer kernel
#include "mho_common.cuh"
#include "mho_internal.h"
extern "C" int mho_adam_replay(mho_ctx_t*, const mho_layer_t*, int32_t, const mho_adam_t*, float*, float*, float*,
                               const float*, int32_t, int64_t, mho_stream_t) {
    mho_set_error("mho_adam_replay: not implemented yet");
    return MHO_ERR_INVALID;
}
