// tcgen05 / TMEM / TMA implicit-GEMM convolution (RIFE_PRECISION_FP16 path) -- interface.
#pragma once
#include <vector>

#include "common.cuh"

namespace rife {

// Host-side packed conv layer in the implicit-GEMM layout  w[tap][cin][npad]  (fp32), shared by both paths.
struct HostLayer {
  int kind = CONV_3x3, cin = 0, n = 0, npad = 0, taps = 9;
  int convt_co = 0;  // > 0: ConvTranspose2d(4,2,1) layer with this many output channels per parity (n = 4*convt_co)
  std::vector<float> w, bias, beta;
};

// Device-side view of one conv layer packed for the tcgen05 kernel.
struct UmmaLayer {
  const __half* w = nullptr;    // [tap][kchunk][npad][kc] fp16: one K-major B-operand tile per (tap, K chunk)
  const float* bias = nullptr;  // [npad]
  const float* beta = nullptr;  // [npad] or null
  int kind = CONV_3x3;
  int taps = 9;
  int cin = 0;     // padded input channels of the activation tensor (multiple of 16)
  int kc = 0;      // channels per K chunk (16, 32 or 64) = TMA box inner extent = swizzle span / 2
  int n = 0;       // valid GEMM columns
  int npad = 0;    // GEMM columns issued (multiple of 16, <= 256)
  int ps_cols = 0; // 1: columns permuted to the PixelShuffle-friendly order  ((par*4 + dy*2+dx)*16 + k), npad = 256
  unsigned long long kmask = ~0ull;  // K blocks (tap * nchunks + chunk) with non-zero weights; zero blocks are skipped
};

// fp32 [tap][cin][npad] -> fp16 [tap][cin/kc][npad'][kc] (+ optional PixelShuffle column permutation);
// fills the geometry of *L and the (possibly permuted) bias / beta vectors.
// bf16: the packed weights carry bf16 bit patterns (the layer then runs with ConvIO::in_bf16)
int umma_pack_weights(const HostLayer& h, bool ps_layout, std::vector<__half>* packed, std::vector<float>* bias,
                      std::vector<float>* beta, UmmaLayer* L, bool bf16 = false);
// fp16 NHWC activations in / out (pipeline path)
int launch_conv_umma(const UmmaLayer& L, const ConvIO& io, cudaStream_t st);
// fp32 NHWC in / out wrapper used by the op-level test entry points (casts to fp16 and back, synchronous)
int launch_conv_umma_f32io(const UmmaLayer& L, const ConvIO& io, cudaStream_t st);

}  // namespace rife
