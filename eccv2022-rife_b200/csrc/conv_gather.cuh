// Block-input gather fused into conv0.0 (scale-1 block, 28 -> 16 channels, stride 2) -- interface.  See conv_gather.cu.
#pragma once
#include "conv_halo.cuh"

namespace rife {

// true when `conv00` (halo-packed) has the shape the fused kernel is compiled for: 2x2 taps over the 4 x 32-channel
// space-to-depth block input, 16 output channels
bool gather_conv_fits(const HaloLayer& conv00);
// true when the block's inputs have the form the fused kernel stages by TMA: the fp32 flow buffer holds the sum over the levels
// [0, level - 1) and the newest level (level - 1) is a 16-channel tmp tensor at exactly half the resolution (H, W)
bool gather_conv_levels_fit(int level, const TmpLevel* levels, int fbuf_levels, int H, int W);
// Builds the block input of block `level` (block scale 1, fp16 fast path) tile by tile in shared memory and runs conv0.0 on
// it: out = the conv0.0 activation, stored space-to-depth planar-8 exactly as EPI_LRELU_S2D stores it ([B][8][H/4][W/4][8]).
// fbuf_in: full-resolution flow summed over the levels [0, fbuf_levels = level - 1); fbuf_out (!= fbuf_in): receives the sum
// over [0, level) for every pixel.  out_bf16: store the activation as bf16.
int launch_gather_conv00(const PixelCtx& c, int level, const TmpLevel* levels, const float* fbuf_in, float* fbuf_out,
                         int fbuf_levels, const HaloLayer& conv00, void* out, int out_bf16, cudaStream_t st);

}  // namespace rife
