#include "conv_umma.cuh"
namespace rife {
int umma_pack_weights(const float*, int, int, int, std::vector<__half>*, UmmaLayer*) { RIFE_FAIL(RIFE_ERR_DTYPE, "fp16 tcgen05 path not built yet"); }
int launch_conv_umma(const UmmaLayer&, const ConvIO&, cudaStream_t) { RIFE_FAIL(RIFE_ERR_DTYPE, "fp16 tcgen05 path not built yet"); }
int launch_conv_umma_f32io(const UmmaLayer&, const ConvIO&, cudaStream_t) { RIFE_FAIL(RIFE_ERR_DTYPE, "fp16 tcgen05 path not built yet"); }
}
