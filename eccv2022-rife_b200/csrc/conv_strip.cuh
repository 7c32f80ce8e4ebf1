// Fused pair of kx-stacked 3x3 tcgen05 convolutions on strip tiles (C = 16 / 32) -- interface.  See conv_strip.cu.
#pragma once
#include <vector>

#include "conv_umma.cuh"

namespace rife {

// Two consecutive 3x3 c -> c layers (c = 16 or 32) run by ONE launch; the intermediate activation stays in shared memory.
struct StripPair {
  const __half* w = nullptr;            // device: [layer][ky][c/8][3c][8] (fp16 or bf16 bit patterns)
  const float* bias[2] = {nullptr, nullptr};
  const float* beta[2] = {nullptr, nullptr};
  int res[2] = {0, 0};                  // 1: ResConv tail lrelu((acc + bias) * beta + x), 0: lrelu(acc + bias)
  int c = 0;
};

bool strip_fits(const HostLayer& h);
// fp32 [tap][cin][npad] -> [ky][cin/8][kx * c + co][8] 16-bit (the three horizontal taps of a kernel row side by side in N)
int strip_pack_weights(const HostLayer& h, std::vector<__half>* packed, bool bf16 = false);
// planar-8 [B][c/8][H][W][8] in (io.in) and out (io.out); io.in_bf16: both are bf16 (and so are the weights)
int launch_conv_strip2(const StripPair& L, const ConvIO& io, cudaStream_t st);

}  // namespace rife
