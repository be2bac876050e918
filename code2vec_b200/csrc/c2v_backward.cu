// c2v_backward.cu -- placeholder until the fused backward lands (next commit).
#include "c2v_common.cuh"
namespace c2v {
size_t encode_backward_workspace_bytes(const c2v_dims *, int, int) { return 1024; }
int launch_encode_backward(const c2v_dims *, const c2v_params *, const EncodeArgs &, int, const float *,
                           const float *, const float *, const float *, const c2v_grads *, void *, size_t,
                           cudaStream_t)
{
    set_error("encode backward is not built yet");
    return C2V_EUNSUPPORTED;
}
}  // namespace c2v
