// The whole Head (cnn0 -> cnn1 -> cnn2 -> cnn3 + SRC assembly; train_log/IFNet_HDv3.py:24-43) as ONE kernel -- interface.
// See conv_head.cu.
#pragma once
#include <vector>

#include "conv_umma.cuh"

namespace rife {

struct HeadFused {
  const __half* w = nullptr;    // device: the four layers' row-stacked ("Toeplitz") weight strips, see conv_head.cu
  const float* bias = nullptr;  // device: [4][16]
  int bf16 = 0;                 // RIFE_PRECISION_BF16: cnn1..cnn3 read bf16 activations (cnn0 reads the fp16 s2d frame)
  int w_bytes = 0;
};

// h[0]: cnn0 as CONV_S2D (2x2 taps on the s2d frame, 16 -> 16); h[1], h[2]: 3x3 16 -> 16; h[3]: the ConvTranspose as 3x3
// with N = 4 parities x 4 channels (api.cu: pack_conv3x3_s2d / pack_conv3x3 / pack_convT)
bool head_fused_fits(const HostLayer* h);
int head_fused_pack(const HostLayer* h, bool bf16, std::vector<__half>* packed, std::vector<float>* bias);
// head_in: the s2d frames, planar-8 [F][2][H2][W2][8] fp16 (prep_planar_kernel); src: [F][2 H2][2 W2][8] fp16
int launch_head_fused(const HeadFused& L, const void* head_in, void* src, int F, int H2, int W2, cudaStream_t st);

}  // namespace rife
