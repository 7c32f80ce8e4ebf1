// Scene-change / static-frame detector of the reference's video loop (inference_video.py:226-228,242-243):
//   I_small = F.interpolate(I, (32, 32), mode='bilinear', align_corners=False)   on the zero-padded frame
//   ssim    = ssim_matlab(I0_small[:, :3], I1_small[:, :3])                        (model/pytorch_msssim/__init__.py:81-134)
// as ONE kernel, one CTA per frame pair: both 3x32x32 thumbnails are built in shared memory straight from the
// full-resolution frames (4 taps per thumbnail pixel), the five 11^3 Gaussian-window means (mu1, mu2, E[x^2], E[y^2],
// E[xy]; colour treated as the third spatial axis with replicate padding) are evaluated separably, and the SSIM map is
// reduced to its mean.  The reference runs ~25 small launches and a device->host sync per frame for the same number.
#include "common.cuh"

namespace rife {
namespace {

constexpr int TS = 32;     // thumbnail side
constexpr int WIN = 11;    // Gaussian window, sigma 1.5 (create_window_3d)

struct SsimArgs {
  float g[WIN];        // 1-D window (gaussian(11, 1.5)), fp32 like the reference's
  float gc[3][3];      // the same window folded over the replicate-padded 3-channel axis: out[c] = sum_c' gc[c][c'] in[c']
};

template <typename Tio>
__device__ __forceinline__ float thumb_px(const Tio* __restrict__ img, const ImgLayout& lay, int c, int ty, int tx,
                                          int Hp, int Wp) {
  // ATen upsample_bilinear2d with an explicit output size: scale = in / out, src = scale * (dst + 0.5) - 0.5, >= 0
  const float sy = (float)Hp / (float)TS, sx = (float)Wp / (float)TS;
  float fy = sy * ((float)ty + 0.5f) - 0.5f, fx = sx * ((float)tx + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = min((int)fy, Hp - 1), x0 = min((int)fx, Wp - 1);
  const int y1 = y0 + (y0 < Hp - 1 ? 1 : 0), x1 = x0 + (x0 < Wp - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  auto px = [&](int y, int x) -> float {   // the frame is zero-padded to (Hp, Wp) on the right / bottom
    return (y < lay.h && x < lay.w) ? to_f(img[c * lay.sc + y * lay.sy + x * lay.sx]) : 0.f;
  };
  return (1.f - ly) * ((1.f - lx) * px(y0, x0) + lx * px(y0, x1)) + ly * ((1.f - lx) * px(y1, x0) + lx * px(y1, x1));
}

template <typename Tio>
__global__ void __launch_bounds__(1024) scene_ssim_kernel(const Tio* __restrict__ img0, const Tio* __restrict__ img1,
                                                          ImgLayout lay, int Hp, int Wp, SsimArgs a, float* __restrict__ out) {
  __shared__ float tA[3][TS][TS], tB[3][TS][TS], buf[3][TS][TS];
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const Tio* i0 = img0 + b * lay.sb;
  const Tio* i1 = img1 + b * lay.sb;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    tA[c][y][x] = thumb_px(i0, lay, c, y, x, Hp, Wp);
    tB[c][y][x] = thumb_px(i1, lay, c, y, x, Hp, Wp);
  }
  __syncthreads();
  float m[5][3];  // filtered quantity q at this thread's (y, x) for the three channels
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    // pass along x (replicate padding)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const int xx = min(max(x + k - WIN / 2, 0), TS - 1);
        const float va = tA[c][y][xx], vb = tB[c][y][xx];
        const float v = q == 0 ? va : (q == 1 ? vb : (q == 2 ? va * va : (q == 3 ? vb * vb : va * vb)));
        s += a.g[k] * v;
      }
      buf[c][y][x] = s;
    }
    __syncthreads();
    // pass along y, then the channel axis in registers
    float t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) s += a.g[k] * buf[c][min(max(y + k - WIN / 2, 0), TS - 1)][x];
      t[c] = s;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) m[q][c] = a.gc[c][0] * t[0] + a.gc[c][1] * t[1] + a.gc[c][2] * t[2];
    __syncthreads();
  }
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;   // L = 1: frames in [0, 1]
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mu1 = m[0][c], mu2 = m[1][c];
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = m[2][c] - mu1_sq, s2 = m[3][c] - mu2_sq, s12 = m[4][c] - mu12;
    const float v1 = 2.0f * s12 + C2, v2 = s1 + s2 + C2;
    acc += ((2.f * mu12 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (x == 0) red[y] = acc;
  __syncthreads();
  if (y == 0) {
    float s = red[x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (x == 0) out[b] = s / (float)(3 * TS * TS);
  }
}

}  // namespace

int launch_scene_ssim(const void* img0, const void* img1, int io_dtype, int B, int C, int h, int w, int Hp, int Wp,
                      float* out, cudaStream_t st) {
  SsimArgs a;
  float sum = 0.f;
  for (int k = 0; k < WIN; ++k) {
    a.g[k] = (float)exp(-(double)((k - WIN / 2) * (k - WIN / 2)) / (2.0 * 1.5 * 1.5));   // gaussian(), fp32 tensor
    sum += a.g[k];
  }
  for (int k = 0; k < WIN; ++k) a.g[k] = a.g[k] / sum;
  for (int c = 0; c < 3; ++c) {
    for (int j = 0; j < 3; ++j) a.gc[c][j] = 0.f;
    for (int k = 0; k < WIN; ++k) {
      int cc = c + k - WIN / 2;
      cc = cc < 0 ? 0 : (cc > 2 ? 2 : cc);
      a.gc[c][cc] += a.g[k];
    }
  }
  const ImgLayout lay = make_layout(io_dtype, C, h, w);
  if (io_dtype == RIFE_F32)
    scene_ssim_kernel<float><<<B, 1024, 0, st>>>((const float*)img0, (const float*)img1, lay, Hp, Wp, a, out);
  else if (io_dtype == RIFE_F16)
    scene_ssim_kernel<__half><<<B, 1024, 0, st>>>((const __half*)img0, (const __half*)img1, lay, Hp, Wp, a, out);
  else if (io_dtype == RIFE_U8)
    scene_ssim_kernel<uint8_t><<<B, 1024, 0, st>>>((const uint8_t*)img0, (const uint8_t*)img1, lay, Hp, Wp, a, out);
  else
    RIFE_FAIL(RIFE_ERR_DTYPE, "scene_ssim: bad io dtype");
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

}  // namespace rife
