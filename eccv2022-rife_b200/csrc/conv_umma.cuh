// tcgen05 / TMEM / TMA implicit-GEMM convolution (RIFE_PRECISION_FP16 path) -- interface.
#pragma once
#include <vector>

#include "common.cuh"

namespace rife {

// Device-side view of one conv layer packed for the tcgen05 kernel.
struct UmmaLayer {
  const __half* w = nullptr;    // [tap][kchunk][npad rows][KC] fp16, K-major rows (B operand), see conv_umma.cu
  const float* bias = nullptr;  // [npad]
  const float* beta = nullptr;  // [npad] or null
  int kind = CONV_3x3;
  int taps = 9;
  int cin = 0;   // padded input channels of the activation tensor
  int kc = 0;    // channels per K chunk (16, 32 or 64): TMA box inner extent / swizzle span
  int n = 0;     // valid GEMM columns
  int npad = 0;  // GEMM columns (multiple of 16)
};

// host repack: fp32 [tap][cin][npad] -> fp16 [tap][cin/kc][npad][kc]; fills the geometry fields of *L
int umma_pack_weights(const float* w_f32, int taps, int cin, int npad, std::vector<__half>* packed, UmmaLayer* L);
// fp16 NHWC activations in / out (pipeline path)
int launch_conv_umma(const UmmaLayer& L, const ConvIO& io, cudaStream_t st);
// fp32 NHWC in / out wrapper used by the op-level test entry points (converts to fp16 and back)
int launch_conv_umma_f32io(const UmmaLayer& L, const ConvIO& io, cudaStream_t st);

}  // namespace rife
