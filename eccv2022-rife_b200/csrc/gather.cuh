// Device helpers of the block-input gather shared by pixel_kernels.cu (stand-alone gather kernels) and conv_gather.cu
// (the gather fused into block 4's conv0.0): bilinear index / weight helpers, the packed-half2 warp taps, the per-level
// flow accumulation.
// Reference semantics: train_log/IFNet_HDv3.py:79-82,86-89,141-152; model/warplayer.py:8-22.
#pragma once
#include "common.cuh"

namespace rife {
namespace {

struct Lerp {
  int i0, i1;
  float l0, l1;
};

// ATen area_pixel_compute_source_index(align_corners=False, scale given) + compute_source_index_and_lambda
// for an up-sampling by integer factor s (rs = 1/s).
__device__ __forceinline__ Lerp make_lerp(int d, int in_size, float rs) {
  float src = rs * ((float)d + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  int i0 = (int)src;
  i0 = min(i0, in_size - 1);
  Lerp L;
  L.i0 = i0;
  L.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  L.l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  L.l0 = 1.f - L.l1;
  return L;
}

struct WarpGeom {
  float halfw, halfh;          // (W-1)/2, (H-1)/2   (model/warplayer.py:18-19)
  float inv_halfw, inv_halfh;  // reciprocals, used by the fp16-storage pipeline (1 ulp of the normalised coordinate)
  float wm1, hm1;              // W-1, H-1
  int src_div;                 // output b reads SRC frames b / src_div (several timesteps share one frame pair)
};

constexpr int MAXLEV = 5;
struct LevelsArg {
  const void* ptr[MAXLEV];
  int h[MAXLEV], w[MAXLEV], s[MAXLEV];
  float rs[MAXLEV], sc[MAXLEV];  // 1/s and s as floats (both exact: s is a power of two)
};

struct Lerp2 {  // two adjacent positions d, d+1 sharing their taps
  int i0, i1;
  float a0, a1;  // lambda1 of position d and d+1
};
__device__ __forceinline__ Lerp2 make_lerp2(int d, int in_size, float rs) {
  const Lerp l0 = make_lerp(d, in_size, rs), l1 = make_lerp(d + 1, in_size, rs);
  Lerp2 r;
  r.i0 = l0.i0; r.i1 = l0.i1; r.a0 = l0.l1; r.a1 = l1.l1;
  return r;
}

// fp16 fast path: the sample position is x + flow_x directly.  model/warplayer.py:18-22 reaches the same number through
// g = linspace(-1,1,W)[x] + flow_x / ((W-1)/2) and ((g+1)/2)*(W-1); the two differ by fp32 rounding only (2e-4 of a pixel
// at 1920 wide) -- far below the fp16 storage rounding of this path.  The fp32 mode and the final merge
// kernel keep the reference's operation order.
__device__ __forceinline__ void warp8_h2(const __half* __restrict__ src, int H, int W, const WarpGeom& g, int x, int y,
                                         float fx, float fy, __half2 (&out)[4]) {
  float ix = (float)x + fx;
  float iy = (float)y + fy;
  ix = fminf(g.wm1, fmaxf(ix, 0.f));
  iy = fminf(g.hm1, fmaxf(iy, 0.f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float ax = ix - x0f, ay = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const int dx = (x0 + 1 < W) ? SRC_C : 0, dy = (y0 + 1 < H) ? W * SRC_C : 0;  // beyond the edge: weight is 0 anyway
  const __half* p = src + (y0 * W + x0) * SRC_C;
  const uint4 u00 = *reinterpret_cast<const uint4*>(p), u01 = *reinterpret_cast<const uint4*>(p + dx);
  const uint4 u10 = *reinterpret_cast<const uint4*>(p + dy), u11 = *reinterpret_cast<const uint4*>(p + dy + dx);
  const __half2 w00 = __float2half2_rn((1.f - ax) * (1.f - ay)), w01 = __float2half2_rn(ax * (1.f - ay));
  const __half2 w10 = __float2half2_rn((1.f - ax) * ay), w11 = __float2half2_rn(ax * ay);
  const __half2 *h00 = reinterpret_cast<const __half2*>(&u00), *h01 = reinterpret_cast<const __half2*>(&u01);
  const __half2 *h10 = reinterpret_cast<const __half2*>(&u10), *h11 = reinterpret_cast<const __half2*>(&u11);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    out[i] = __hfma2(w11, h11[i], __hfma2(w10, h10[i], __hfma2(w01, h01[i], __hmul2(w00, h00[i]))));
}

// NS = 1 variant (block 4: one sample per pixel): mask / feat of the newest level are blended in packed half2 arithmetic
// straight from the loaded words (they are stored as fp16 right afterwards); the flow stays fp32.  Saves the 48
// half -> float conversions and 24 of the 48 FMAs per pixel of the float form.
__device__ __forceinline__ void level_taps_h1(const __half* __restrict__ r0, const __half* __restrict__ r1, int c0, int c1,
                                              bool newest, const float (&wq)[4], float sc, bool first, float (&flow)[4],
                                              __half2 (&mfh)[6]) {
  float acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half* p = ((k >> 1) ? r1 : r0) + ((k & 1) ? c1 : c0);
    float e[4];
    if (newest) {
      const uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + 8);
      const __half2* hl = reinterpret_cast<const __half2*>(&lo);
      const __half2* hh = reinterpret_cast<const __half2*>(&hi);
      const float2 f01 = __half22float2(hl[0]), f23 = __half22float2(hl[1]);
      e[0] = f01.x; e[1] = f01.y; e[2] = f23.x; e[3] = f23.y;
      const __half2 wk = __float2half2_rn(wq[k]);
      const __half2 v[6] = {hl[2], hl[3], hh[0], hh[1], hh[2], hh[3]};   // channels 4..15
#pragma unroll
      for (int i = 0; i < 6; ++i) mfh[i] = k == 0 ? __hmul2(wk, v[i]) : __hfma2(wk, v[i], mfh[i]);
    } else {
      ld4(p, e);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = k == 0 ? wq[0] * e[c] : fmaf(wq[k], e[c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) flow[c] = first ? acc[c] * sc : fmaf(acc[c], sc, flow[c]);
}

inline WarpGeom make_geom(int H, int W) {
  WarpGeom g;
  g.halfw = (float)((W - 1.0) / 2.0);
  g.halfh = (float)((H - 1.0) / 2.0);
  g.inv_halfw = 1.f / g.halfw;
  g.inv_halfh = 1.f / g.halfh;
  g.wm1 = (float)(W - 1);
  g.hm1 = (float)(H - 1);
  g.src_div = 1;
  return g;
}

inline LevelsArg pack_levels(const TmpLevel* levels, int n) {
  LevelsArg L{};
  for (int i = 0; i < n && i < MAXLEV; ++i) {
    L.ptr[i] = levels[i].ptr;
    L.h[i] = levels[i].h;
    L.w[i] = levels[i].w;
    L.s[i] = levels[i].s;
    L.rs[i] = 1.f / (float)levels[i].s;
    L.sc[i] = levels[i].s > 0 ? (float)levels[i].s : 1.f / (float)(-levels[i].s);   // "* scale" of IFNet_HDv3.py:87
  }
  return L;
}

}  // namespace
}  // namespace rife
