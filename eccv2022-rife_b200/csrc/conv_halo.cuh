// Persistent halo-tile tcgen05 convolution (thin layers, planar-8 activations) -- interface.  See conv_halo.cu.
#pragma once
#include <vector>

#include "conv_umma.cuh"

namespace rife {

struct HaloLayer {
  const __half* w = nullptr;    // smem image: [entry][cin/8][ne][8] fp16 (no-swizzle K-major B tiles)
  const float* bias = nullptr;  // [ncols]
  const float* beta = nullptr;  // [ncols] or null
  int kind = CONV_3x3;
  int cin = 0;        // multiple of 16, <= 128
  int nentries = 0;   // MMA schedule: taps (9 / 4) or ConvTranspose (parity, tap) pairs (16)
  int ne = 0;         // GEMM N of each MMA
  int ncols = 0;      // accumulator columns per tile (ne, or 4 * 64 for the parity schedule)
  int n_valid = 0;
  int w_bytes = 0;
  int ps_cols = 0;    // 1: PixelShuffle layout (4 x 64 columns), 2: narrow PixelShuffle layout (4 x 32 columns)
  int stream = 0;     // 1: weights streamed through a smem ring (do not fit next to the activation stages)
  signed char e_dy[16] = {}, e_dx[16] = {};
  unsigned char e_col16[16] = {}, e_first[16] = {};
};

bool halo_fits(const HostLayer& h, int ps_layout);  // ps_layout: 0 plain, 1 ConvT+PixelShuffle (16-ch tmp), 2 narrow (8-ch tmp: flow + mask)
int halo_mode(const HostLayer& h, int ps_layout);  // 0 no, 1 resident weights, 2 streamed weights
// bf16: the packed weights carry bf16 bit patterns (the layer then runs with ConvIO::in_bf16)
int halo_pack_weights(const HostLayer& h, int ps_layout, std::vector<__half>* packed, std::vector<float>* bias,
                      std::vector<float>* beta, HaloLayer* L, bool bf16 = false);
// planar-8 fp16 activations [B][C/8][H][W][8] in; planar-8 (LRELU / LINEAR / RES / LRELU_S2D) or pixel-major
// (PS / HEADT / CONVT) fp16 out
int launch_conv_halo(const HaloLayer& L, const ConvIO& io, cudaStream_t st);
// n (2..8) ResConv layers of one shape in ONE cooperative launch: layer l reads io.in (l even) / io.out (l odd) and writes
// the other tensor; a grid-wide barrier replaces the kernel boundary.  `sync`: one device uint32, zero-initialised once
// (two counters; every launch leaves them zero again, so launches can be replayed from a CUDA graph); `sync_host` is unused.
// Returns 1 (nothing launched) when the shape has no compile-time instantiation or its grid is not co-resident.
int launch_conv_halo_chain(const HaloLayer* const* layers, int n, const ConvIO& io, unsigned int* sync,
                           unsigned int* sync_host, cudaStream_t st);
int launch_conv_halo_f32io(const HaloLayer& L, const ConvIO& io, cudaStream_t st);

}  // namespace rife
