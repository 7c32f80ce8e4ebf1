// HBM-bound pixel kernels of the RIFE HDv3 hot path: input packing, the fused
// {flow accumulate + warp + concat + 1/s bilinear resize} block-input gather, and the fused
// {flow accumulate + warp + sigmoid-merge} output kernel.  Reference semantics:
//   train_log/IFNet_HDv3.py:79-82,86-89 (F.interpolate sites), :141-159 (concat / warp / merge),
//   model/warplayer.py:8-22 (warp).
//
// Design note ("lazy flow"): the reference materialises flow/mask/feat (13 channels) at full resolution
// after every block (IFNet_HDv3.py:86-89,152).  Here each block only writes its low-resolution 13-channel
// `tmp` tensor; every consumer re-evaluates   flow(p) = sum_j s_j * bilerp(tmp_j[0:4], p)   (same
// left-to-right summation order as the reference's `flow = flow + fd`) for exactly the pixels it needs.
// For s >= 4 the 1/s resize only touches the centre 2x2 of each s x s cell, so the full-resolution
// tensors are never needed.
#include <stdlib.h>

#include "common.cuh"
#include "gather.cuh"

#ifdef RIFE_PDL_NO_EARLY_PIX
#define RIFE_PIX_TRIGGER()
#else
#define RIFE_PIX_TRIGGER() pdl_trigger()
#endif
// resident blocks per SM the block-input gather is compiled for (register cap): measured best on B200 at 1080p --
// s = 1 (block 4): 9 blocks (56 regs), s = 2 (block 3): 7 blocks (72 regs); 104 / 61 regs without the cap cost 2-4 % of a step
#ifndef RIFE_BI_MINB1
#define RIFE_BI_MINB1 9
#define RIFE_BI_MINB2 7
#endif
#ifndef RIFE_BI_MINB2R
#define RIFE_BI_MINB2R 9
#endif
#ifndef RIFE_PREP_MINB
#define RIFE_PREP_MINB 3
#endif
namespace rife {

namespace {

// fp32 storage => reproduce the reference's arithmetic exactly; fp16 storage => cheaper, rounding-equivalent forms
template <typename T> struct ExactDiv { static constexpr bool value = true; };
template <> struct ExactDiv<__half> { static constexpr bool value = false; };


// bilinear up-sample read of channels [c0, c0+N) of a tmp level at full-resolution pixel (y, x)
template <typename Tt, int N>
__device__ __forceinline__ void sample_tmp(const Tt* __restrict__ base, int h, int w, int s, float rs, int c0, int y,
                                           int x, float (&out)[N]) {
  static_assert(N == 4 || N == 16, "N");
  if (s < 0) {
    // level stored at u = -s times the full resolution: F.interpolate(tmp, scale_factor=1/u) (IFNet_HDv3.py:86 with
    // scale < 1) reads the centre 2x2 of each u x u cell with all lambdas = 0.5
    const int u = -s, o = u / 2 - 1;
    const Tt* p00 = base + ((y * u + o) * w + (x * u + o)) * TMP_C + c0;
    float v[4][N];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const Tt* p = p00 + ((k >> 1) * w + (k & 1)) * TMP_C;
      if constexpr (N == 4) {
        ld4(p, v[k]);
      } else {
        float a[8];
        ld8(p, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[k][i] = a[i];
        ld8(p + 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[k][8 + i] = a[i];
      }
    }
    // ATen: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11), all weights 0.5
#pragma unroll
    for (int c = 0; c < N; ++c)
      out[c] = 0.5f * (0.5f * v[0][c] + 0.5f * v[1][c]) + 0.5f * (0.5f * v[2][c] + 0.5f * v[3][c]);
    return;
  }
  if (s == 1) {
    const Tt* p = base + (y * w + x) * TMP_C + c0;
    if constexpr (N == 4) {
      ld4(p, out);
    } else {
      float a[8], b[8];
      ld8(p, a);
      ld8(p + 8, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) { out[i] = a[i]; out[8 + i] = b[i]; }
    }
    return;
  }
  const Lerp ly = make_lerp(y, h, rs), lx = make_lerp(x, w, rs);
  const Tt* r0 = base + ly.i0 * w * TMP_C + c0;   // per-image element offsets fit in 32 bits
  const Tt* r1 = base + ly.i1 * w * TMP_C + c0;
  const Tt* p00 = r0 + lx.i0 * TMP_C;
  const Tt* p01 = r0 + lx.i1 * TMP_C;
  const Tt* p10 = r1 + lx.i0 * TMP_C;
  const Tt* p11 = r1 + lx.i1 * TMP_C;
  float v00[N], v01[N], v10[N], v11[N];
  if constexpr (N == 4) {
    ld4(p00, v00); ld4(p01, v01); ld4(p10, v10); ld4(p11, v11);
  } else {
    float a[8];
#define RIFE_LD16(P, V)                                   \
  ld8(P, a);                                              \
  _Pragma("unroll") for (int i = 0; i < 8; ++i) V[i] = a[i]; \
  ld8(P + 8, a);                                          \
  _Pragma("unroll") for (int i = 0; i < 8; ++i) V[8 + i] = a[i];
    RIFE_LD16(p00, v00) RIFE_LD16(p01, v01) RIFE_LD16(p10, v10) RIFE_LD16(p11, v11)
#undef RIFE_LD16
  }
  if constexpr (ExactDiv<Tt>::value) {
    // ATen upsample_bilinear2d association (fp32 pipeline: bit-faithful)
#pragma unroll
    for (int c = 0; c < N; ++c)
      out[c] = ly.l0 * (lx.l0 * v00[c] + lx.l1 * v01[c]) + ly.l1 * (lx.l0 * v10[c] + lx.l1 * v11[c]);
  } else {
    // fp16-storage pipeline: four tap weights, 4 instead of 6 operations per channel (differs by rounding only)
    const float w00 = ly.l0 * lx.l0, w01 = ly.l0 * lx.l1, w10 = ly.l1 * lx.l0, w11 = ly.l1 * lx.l1;
#pragma unroll
    for (int c = 0; c < N; ++c) out[c] = fmaf(w11, v11[c], fmaf(w10, v10[c], fmaf(w01, v01[c], w00 * v00[c])));
  }
}



// One bilinear tap set of warp() on the 8-channel SRC layout.  Follows the reference's arithmetic order:
// g = base + flow / ((dim-1)/2); ix = ((g+1)/2)*(dim-1); clamp; 4 taps with zero weight beyond the edge.
template <typename T>
__device__ __forceinline__ void warp8(const T* __restrict__ src, int H, int W, const WarpGeom& g, float gx, float gy,
                                      float fx, float fy, float (&out)[8]) {
  const float nx = gx + (ExactDiv<T>::value ? __fdiv_rn(fx, g.halfw) : __fmul_rn(fx, g.inv_halfw));
  const float ny = gy + (ExactDiv<T>::value ? __fdiv_rn(fy, g.halfh) : __fmul_rn(fy, g.inv_halfh));
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.f), 0.5f), g.wm1);
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.f), 0.5f), g.hm1);
  ix = fminf(g.wm1, fmaxf(ix, 0.f));
  iy = fminf(g.hm1, fmaxf(iy, 0.f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = x0f + 1.f, y1f = y0f + 1.f;
  const float wnw = (x1f - ix) * (y1f - iy), wne = (ix - x0f) * (y1f - iy);
  const float wsw = (x1f - ix) * (iy - y0f), wse = (ix - x0f) * (iy - y0f);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const bool xin = (x0 + 1) < W, yin = (y0 + 1) < H;
  const T* p = src + (y0 * W + x0) * SRC_C;
  float v[8];
  ld8(p, v);
#pragma unroll
  for (int c = 0; c < 8; ++c) out[c] = v[c] * wnw;
  if (xin) {
    ld8(p + SRC_C, v);
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] += v[c] * wne;
  }
  if (yin) {
    ld8(p + W * SRC_C, v);
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] += v[c] * wsw;
    if (xin) {
      ld8(p + W * SRC_C + SRC_C, v);
#pragma unroll
      for (int c = 0; c < 8; ++c) out[c] += v[c] * wse;
    }
  }
}


// flow(p) after `nlev` blocks, plus all 16 channels of the newest level (mask = ch 4, feat = ch 5..12)
// `j0` > 0: `flow` already holds the sum over levels [0, j0) (read from the full-resolution fp32 flow buffer).
template <typename Tt>
__device__ __forceinline__ void eval_state(const LevelsArg& L, int nlev, int j0, int b, int y, int x, float (&flow)[4],
                                           float (&last)[16]) {
#pragma unroll 1
  for (int j = j0; j < nlev - 1; ++j) {
    float f[4];
    const Tt* base = reinterpret_cast<const Tt*>(L.ptr[j]) + (size_t)b * L.h[j] * L.w[j] * TMP_C;
    sample_tmp<Tt, 4>(base, L.h[j], L.w[j], L.s[j], L.rs[j], 0, y, x, f);
    const float sc = L.sc[j];
#pragma unroll
    for (int c = 0; c < 4; ++c) flow[c] = (j == 0) ? f[c] * sc : flow[c] + f[c] * sc;  // flow = flow + fd (IFNet_HDv3.py:152)
  }
  const int j = nlev - 1;
  const Tt* base = reinterpret_cast<const Tt*>(L.ptr[j]) + (size_t)b * L.h[j] * L.w[j] * TMP_C;
  sample_tmp<Tt, 16>(base, L.h[j], L.w[j], L.s[j], L.rs[j], 0, y, x, last);
  const float sc = L.sc[j];
#pragma unroll
  for (int c = 0; c < 4; ++c) flow[c] = (j == 0) ? last[c] * sc : flow[c] + last[c] * sc;
}

// The first kernel of a call also resets the call's tile-scheduling counters (conv_halo.cu): every conv of the previous call
// on this plan has completed by then (stream order; pdl_wait() makes completion transitive).
__device__ __forceinline__ void zero_tile_ctrs(unsigned int* __restrict__ ctrs, int n) {
  if (ctrs && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) ctrs[i] = 0u;
}

// ---------------------------------------------------------------------------------------------------
// prep: NCHW images -> SRC tensors (ch 4..7; ch 0..3 zero until Head writes them) + s2d Head input
template <typename Tio, typename T>
__global__ void __launch_bounds__(256) prep_kernel(const Tio* __restrict__ img0, const Tio* __restrict__ img1,
                                                   T* __restrict__ src_all, T* __restrict__ head_in, int F, int nf0,
                                                   int C, int H, int W, int planar, int write_src, ImgLayout lay,
                                                   unsigned int* __restrict__ ctrs, int nctrs, const void* const* iot) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  zero_tile_ctrs(ctrs, nctrs);
  if (iot) {  // CUDA-graph replay: the frame pointers of this call come from the plan's table (common.cuh: PixelCtx::io_table)
    img0 = reinterpret_cast<const Tio*>(iot[0]);
    img1 = reinterpret_cast<const Tio*>(iot[1]);
  }
  const int W2 = W >> 1, H2 = H >> 1;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nq = (size_t)F * H2 * W2;
  if (q >= nq) return;
  const int x2 = (int)(q % W2);
  const int y2 = (int)((q / W2) % H2);
  const int f = (int)(q / ((size_t)W2 * H2));    // frame index: img0's nf0 frames first, then img1's
  const int b = f < nf0 ? f : f - nf0;
  const Tio* img = f < nf0 ? img0 : img1;
  T* src = src_all + (size_t)(f - b) * H * W * SRC_C;
  float hv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) hv[i] = 0.f;
#pragma unroll
  for (int py = 0; py < 2; ++py) {
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int y = 2 * y2 + py, x = 2 * x2 + px;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      if (y < lay.h && x < lay.w) {  // beyond the caller's frame: zero padding (F.pad of the reference's callers)
        const Tio* p = img + b * lay.sb + y * lay.sy + x * lay.sx;
        for (int c = 0; c < C; ++c) v[4 + c] = to_f(p[c * lay.sc]);
      }
      // fused mode: Head.cnn3's epilogue assembles the whole SRC pixel (feature + image) from the s2d Head input
      if (write_src) st8(src + (((size_t)b * H + y) * W + x) * SRC_C, v);
#pragma unroll
      for (int c = 0; c < 3; ++c) hv[(py * 2 + px) * 3 + c] = v[4 + c];
      hv[12 + py * 2 + px] = v[7];  // alpha rides in the 4 spare channels (Head.cnn0 has zero weights there)
    }
  }
  // pixel-major [2B][H/2][W/2][16] or planar-8 [2B][2][H/2][W/2][8] (conv_halo.cu)
  T* hp = planar ? head_in + (((size_t)f * 2 * H2 + y2) * W2 + x2) * 8
                 : head_in + (((size_t)f * H2 + y2) * W2 + x2) * 16;
  const size_t gstride = planar ? (size_t)H2 * W2 * 8 : 8;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = hv[i];
  st8(hp, a);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = hv[8 + i];
  st8(hp + gstride, a);
}


// fp16 fused-mode input packing: NCHW frames -> planar-8 s2d Head input [2B][2][H/2][W/2][8] only (SRC is assembled
// by Head.cnn3's epilogue).  One thread converts a 2 x 8 pixel strip: 16-byte loads per channel row, 16-byte stores.
template <typename Tio>
__global__ void __launch_bounds__(256, RIFE_PREP_MINB) prep_planar_kernel(const Tio* __restrict__ img0, const Tio* __restrict__ img1,
                                                          __half* __restrict__ head_in, int nf0, int C, int H, int W,
                                                          unsigned int* __restrict__ ctrs, int nctrs, const void* const* iot) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  zero_tile_ctrs(ctrs, nctrs);
  if (iot) {
    img0 = reinterpret_cast<const Tio*>(iot[0]);
    img1 = reinterpret_cast<const Tio*>(iot[1]);
  }
  const int W8 = W >> 3, H2 = H >> 1, W2 = W >> 1;
  const int xs = blockIdx.x * blockDim.x + threadIdx.x;  // strip index along x (8 pixels)
  const int y2 = blockIdx.y, fb = blockIdx.z;            // frame index: img0's nf0 frames first, then img1's
  if (xs >= W8) return;
  const Tio* img = fb < nf0 ? img0 + (size_t)fb * C * H * W : img1 + (size_t)(fb - nf0) * C * H * W;
  float px[2][4][8];  // [row][channel][x]
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < C) {
        const Tio* p = img + ((size_t)c * H + 2 * y2 + r) * W + xs * 8;
        if constexpr (sizeof(Tio) == 2) {
          ld8(reinterpret_cast<const __half*>(p), px[r][c]);
        } else {
          ld8(reinterpret_cast<const float*>(p), px[r][c]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) px[r][c][i] = 0.f;
      }
    }
  __half* o = head_in + (((size_t)fb * 2 * H2 + y2) * W2 + xs * 4) * 8;
  const size_t gstride = (size_t)H2 * W2 * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // 4 s2d pixels of this strip
    float hv[16];
#pragma unroll
    for (int par = 0; par < 4; ++par) {
      const int r = par >> 1, xx = 2 * k + (par & 1);
      hv[par * 3 + 0] = px[r][0][xx]; hv[par * 3 + 1] = px[r][1][xx]; hv[par * 3 + 2] = px[r][2][xx];
      hv[12 + par] = px[r][3][xx];
    }
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = hv[i];
    st8(o + k * 8, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = hv[8 + i];
    st8(o + gstride + k * 8, a);
  }
}

// uint8 HWC RGB frames (B, h, w, 3) as a video decoder delivers them, h <= H rows (zero-padded below, F.pad of the
// reference's callers), w == W: one thread converts a 2 x 8 pixel strip = 2 x 24 bytes read as 8-byte words.
__global__ void __launch_bounds__(256, RIFE_PREP_MINB) prep_planar_u8_kernel(const uint8_t* __restrict__ img0,
                                                                             const uint8_t* __restrict__ img1,
                                                                             __half* __restrict__ head_in, int nf0, int h,
                                                                             int H, int W, unsigned int* __restrict__ ctrs,
                                                                             int nctrs, const void* const* iot) {
  pdl_wait();
  RIFE_PIX_TRIGGER();
  zero_tile_ctrs(ctrs, nctrs);
  if (iot) {
    img0 = reinterpret_cast<const uint8_t*>(iot[0]);
    img1 = reinterpret_cast<const uint8_t*>(iot[1]);
  }
  const int W8 = W >> 3, H2 = H >> 1, W2 = W >> 1;
  const int xs = blockIdx.x * blockDim.x + threadIdx.x;
  const int y2 = blockIdx.y, fb = blockIdx.z;
  if (xs >= W8) return;
  const uint8_t* img = fb < nf0 ? img0 + (size_t)fb * h * W * 3 : img1 + (size_t)(fb - nf0) * h * W * 3;
  uint32_t wd[2][6];  // [row][32-bit word]: 24 bytes = 8 RGB pixels
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int y = 2 * y2 + r;
    if (y < h) {
      const uint2* p = reinterpret_cast<const uint2*>(img + ((size_t)y * W + xs * 8) * 3);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint2 u = __ldg(p + i);
        wd[r][2 * i] = u.x; wd[r][2 * i + 1] = u.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) wd[r][i] = 0u;
    }
  }
  auto byte_at = [&](int r, int k) -> float {  // k = x*3 + c, compile-time after unrolling
    return to_f((uint8_t)((wd[r][k >> 2] >> ((k & 3) * 8)) & 0xffu));
  };
  __half* o = head_in + (((size_t)fb * 2 * H2 + y2) * W2 + xs * 4) * 8;
  const size_t gstride = (size_t)H2 * W2 * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a[8], b2[8];
    // s2d channel = par*3 + c for par = (row, x parity); channels 12..15 (alpha) are zero for RGB frames
    a[0] = byte_at(0, (2 * k) * 3 + 0); a[1] = byte_at(0, (2 * k) * 3 + 1); a[2] = byte_at(0, (2 * k) * 3 + 2);
    a[3] = byte_at(0, (2 * k + 1) * 3 + 0); a[4] = byte_at(0, (2 * k + 1) * 3 + 1); a[5] = byte_at(0, (2 * k + 1) * 3 + 2);
    a[6] = byte_at(1, (2 * k) * 3 + 0); a[7] = byte_at(1, (2 * k) * 3 + 1);
    b2[0] = byte_at(1, (2 * k) * 3 + 2);
    b2[1] = byte_at(1, (2 * k + 1) * 3 + 0); b2[2] = byte_at(1, (2 * k + 1) * 3 + 1); b2[3] = byte_at(1, (2 * k + 1) * 3 + 2);
    b2[4] = b2[5] = b2[6] = b2[7] = 0.f;
    st8(o + k * 8, a);
    st8(o + gstride + k * 8, b2);
  }
}

// ---------------------------------------------------------------------------------------------------
// block input: one thread per low-resolution pixel of the (H/s, W/s) block input.
template <typename T, typename Tt, int NS, bool LEVEL0>
__global__ void __launch_bounds__(128) block_input_kernel(LevelsArg L, int level, int s, const T* __restrict__ src0,
                                                          const T* __restrict__ src1, const float* __restrict__ gridx,
                                                          const float* __restrict__ gridy,
                                                          const float* __restrict__ tstep, T* __restrict__ xin, int B,
                                                          int H, int W, WarpGeom geom, int planar,
                                                          float* __restrict__ fbuf, int fbuf_levels, int fbuf_store) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  const int Hs = H / s, Ws = W / s;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= Ws) return;
  const T* s0 = src0 + (size_t)(b / geom.src_div) * H * W * SRC_C;
  const T* s1 = src1 + (size_t)(b / geom.src_div) * H * W * SRC_C;
  const float t = tstep[b];
  constexpr int ns = NS;                      // samples per axis: F.interpolate(1/s) reads the centre 2x2 (lambda = 0.5)
  const int off = (NS == 1) ? 0 : (s / 2 - 1);
  float acc[28];
  float row[28];
#pragma unroll 1
  for (int sy = 0; sy < ns; ++sy) {
#pragma unroll 1
    for (int sx = 0; sx < ns; ++sx) {
      const int y = Y * s + off + sy, x = X * s + off + sx;
      float v[28];
      if constexpr (LEVEL0) {
        // IFNet_HDv3.py:141: cat(img0[:, :3], img1[:, :3], f0, f1, timestep)
        float a[8], c[8];
        ld8(s0 + (y * W + x) * SRC_C, a);
        ld8(s1 + (y * W + x) * SRC_C, c);
        v[0] = a[4]; v[1] = a[5]; v[2] = a[6];
        v[3] = c[4]; v[4] = c[5]; v[5] = c[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[6 + i] = a[i]; v[10 + i] = c[i]; }
        v[14] = t;
#pragma unroll
        for (int i = 15; i < 28; ++i) v[i] = 0.f;
      } else {
        // IFNet_HDv3.py:145-147: cat(warped_img0, warped_img1, wf0, wf1, timestep, mask, feat) ++ flow (:82)
        float flow[4] = {0.f, 0.f, 0.f, 0.f}, last[16];
        float* fp = fbuf + (((size_t)b * H + y) * W + x) * 4;
        if (fbuf_levels > 0) ld4(fp, flow);
        eval_state<Tt>(L, level, fbuf_levels, b, y, x, flow, last);
        if (fbuf_store) st4(fp, flow);   // full-resolution flow after `level` blocks, reused by the next stage
        const float gx = gridx[x], gy = gridy[y];
        float a[8], c[8];
        warp8<T>(s0, H, W, geom, gx, gy, flow[0], flow[1], a);
        warp8<T>(s1, H, W, geom, gx, gy, flow[2], flow[3], c);
        v[0] = a[4]; v[1] = a[5]; v[2] = a[6];
        v[3] = c[4]; v[4] = c[5]; v[5] = c[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[6 + i] = a[i]; v[10 + i] = c[i]; }
        v[14] = t;
        v[15] = last[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[16 + i] = last[5 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[24 + i] = flow[i];
      }
      if (ns == 1) {
#pragma unroll
        for (int i = 0; i < 28; ++i) acc[i] = v[i];
      } else if (sx == 0) {
#pragma unroll
        for (int i = 0; i < 28; ++i) row[i] = 0.5f * v[i];
      } else {
        // upsample_bilinear2d: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11) with all lambdas = 0.5
#pragma unroll
        for (int i = 0; i < 28; ++i) {
          const float r = row[i] + 0.5f * v[i];
          acc[i] = (sy == 0) ? 0.5f * r : acc[i] + 0.5f * r;
        }
      }
    }
  }
  // flow channels: interpolate(flow, 1/s) * 1. / s   (IFNet_HDv3.py:81)
  const float inv_s = 1.f / (float)s;
#pragma unroll
  for (int i = 24; i < 28; ++i) acc[i] = acc[i] * inv_s;
  // s2d store, channel = (py*2+px)*32 + c: pixel-major [b][Y/2][X/2][128] or planar-8 [b][16][Y/2][X/2][8]
  const int par = (Y & 1) * 2 + (X & 1);
  const size_t hw2 = (size_t)(Hs >> 1) * (Ws >> 1), pix2 = (size_t)(Y >> 1) * (Ws >> 1) + (X >> 1);
  T* dst = planar ? xin + (((size_t)b * 16 + par * 4) * hw2 + pix2) * 8
                  : xin + (((size_t)b * hw2 + pix2) * 4 + par) * XIN_C;
  const size_t gstride = planar ? hw2 * 8 : 8;
  float o[8];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[k * 8 + i];
    st8(dst + k * gstride, o);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = acc[24 + i];
#pragma unroll
  for (int i = 4; i < 8; ++i) o[i] = 0.f;
  st8(dst + 3 * gstride, o);
}


// ---------------------------------------------------------------------------------------------------
// fp16-storage fast path of the block-input gather (level >= 1).  Same semantics as block_input_kernel, with
//   * the four samples of the 1/s resize sharing their bilinear taps of every coarser tmp level (a 2x2 sample
//     block never straddles a cell of a coarser level), and mask/feat averaged through the mean tap weights;
//   * warp taps blended with packed half2 FMAs (values in [0,1], weights to 11 bits: below the fp16 storage
//     rounding of the block input itself); coordinates and flow stay fp32.
#define SPLIT8(e, lo, hi)                                  \
  do {                                                     \
    lo[0] = e[0]; lo[1] = e[1]; lo[2] = e[2]; lo[3] = e[3]; \
    hi[0] = e[4]; hi[1] = e[5]; hi[2] = e[6]; hi[3] = e[7]; \
  } while (0)

// One coarser tmp level, tap-major: every tap is loaded once (16 + 16 bytes for the newest level, whose mask / feat
// channels are wanted too; 8 bytes of flow otherwise) and folded into the NQ samples' flow sums and into mf straight away,
// in the order 00, 01, 10, 11 of the bilinear formula.  wq[q][k]: weight of tap k for sample q; am[k]: mean weight of tap k.
template <int NQ>
__device__ __forceinline__ void level_taps(const __half* __restrict__ r0, const __half* __restrict__ r1, int c0, int c1,
                                           bool newest, const float (&wq)[NQ][4], const float (&am)[4], float sc,
                                           bool first, float (&flow)[NQ][4], float (&mf)[12]) {
  float acc[NQ][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half* p = ((k >> 1) ? r1 : r0) + ((k & 1) ? c1 : c0);
    float e[8];
    if (newest) {
      float u[8];
      ld8(p, e);
      ld8(p + 8, u);
#pragma unroll
      for (int c = 0; c < 4; ++c) mf[c] = k == 0 ? am[0] * e[4 + c] : fmaf(am[k], e[4 + c], mf[c]);
#pragma unroll
      for (int c = 0; c < 8; ++c) mf[4 + c] = k == 0 ? am[0] * u[c] : fmaf(am[k], u[c], mf[4 + c]);
    } else {
      float v[4];
      ld4(p, v);
      e[0] = v[0]; e[1] = v[1]; e[2] = v[2]; e[3] = v[3];
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[q][c] = k == 0 ? wq[q][0] * e[c] : fmaf(wq[q][k], e[c], acc[q][c]);
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) flow[q][c] = first ? acc[q][c] * sc : fmaf(acc[q][c], sc, flow[q][c]);
}


template <int NS>
__global__ void __launch_bounds__(128, NS == 1 ? RIFE_BI_MINB1 : RIFE_BI_MINB2) block_input_h_kernel(LevelsArg L, int level, int s,
                                                            const __half* __restrict__ src0,
                                                            const __half* __restrict__ src1,
                                                            const float* __restrict__ gridx,
                                                            const float* __restrict__ gridy,
                                                            const float* __restrict__ tstep, __half* __restrict__ xin,
                                                            int B, int H, int W, WarpGeom geom, int planar,
                                                            float* __restrict__ fbuf, int fbuf_levels, int fbuf_store) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  constexpr int NQ = NS * NS;
  const int Hs = H / s, Ws = W / s;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= Ws) return;
  const int fr = geom.src_div == 1 ? b : b / geom.src_div;   // uniform branch: no integer division in the common case
  const __half* s0 = src0 + (size_t)fr * H * W * SRC_C;
  const __half* s1 = src1 + (size_t)fr * H * W * SRC_C;
  const int off = (NS == 1) ? 0 : (s / 2 - 1);
  const int x0 = X * s + off, y0 = Y * s + off;
  float flow[NQ][4];
  if (fbuf_levels > 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) ld4(fbuf + (((size_t)b * H + y0 + q / NS) * W + x0 + q % NS) * 4, flow[q]);
  }
  __half2 mfh[6];  // channels 4..15 of the newest level (mask, feat x8, pad), averaged over the samples, packed half2
#pragma unroll 1
  for (int j = fbuf_levels; j < level; ++j) {
    const __half* base = reinterpret_cast<const __half*>(L.ptr[j]) + (size_t)b * L.h[j] * L.w[j] * TMP_C;
    const int h = L.h[j], w = L.w[j], sj = L.s[j];
    const float sc = L.sc[j];
    const bool newest = j == level - 1;
    if (sj == 1) {  // only possible as the newest level of a scale-1 eval; not reached from block inputs
      continue;
    }
    if constexpr (NS == 1) {
      const Lerp2 ly = make_lerp2(y0, h, L.rs[j]), lx = make_lerp2(x0, w, L.rs[j]);
      const float ay = ly.a0, ax = lx.a0;
      const float wq[4] = {(1.f - ay) * (1.f - ax), (1.f - ay) * ax, ay * (1.f - ax), ay * ax};
      level_taps_h1(base + ly.i0 * w * TMP_C, base + ly.i1 * w * TMP_C, lx.i0 * TMP_C, lx.i1 * TMP_C, newest, wq, sc, j == 0,
                    flow[0], mfh);
      } else {
      // four samples: per-level tap arrays kept whole (the tap-major form costs this variant registers: 0.36 vs 0.34 ms)
      const Lerp2 ly = make_lerp2(y0, h, L.rs[j]), lx = make_lerp2(x0, w, L.rs[j]);
      const __half* r0 = base + ly.i0 * w * TMP_C;
      const __half* r1 = base + ly.i1 * w * TMP_C;
      float v00[4], v01[4], v10[4], v11[4];
      ld4(r0 + lx.i0 * TMP_C, v00); ld4(r0 + lx.i1 * TMP_C, v01);
      ld4(r1 + lx.i0 * TMP_C, v10); ld4(r1 + lx.i1 * TMP_C, v11);
      float wsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float ay = (q / NS) ? ly.a1 : ly.a0, ax = (q % NS) ? lx.a1 : lx.a0;
        const float w00 = (1.f - ay) * (1.f - ax), w01 = (1.f - ay) * ax, w10 = ay * (1.f - ax), w11 = ay * ax;
        wsum[0] += w00; wsum[1] += w01; wsum[2] += w10; wsum[3] += w11;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float f = fmaf(w11, v11[c], fmaf(w10, v10[c], fmaf(w01, v01[c], w00 * v00[c])));
          flow[q][c] = (j == 0) ? f * sc : fmaf(f, sc, flow[q][c]);
        }
      }
      if (newest) {
        // mask / feat (channels 4..15): mean-weight bilinear blend in packed half2, straight from the loaded words
        const float inv = 1.f / (float)NQ;
        const __half2 a00 = __float2half2_rn(wsum[0] * inv), a01 = __float2half2_rn(wsum[1] * inv);
        const __half2 a10 = __float2half2_rn(wsum[2] * inv), a11 = __float2half2_rn(wsum[3] * inv);
        const __half* q00 = r0 + lx.i0 * TMP_C + 4;
        const __half* q01 = r0 + lx.i1 * TMP_C + 4;
        const __half* q10 = r1 + lx.i0 * TMP_C + 4;
        const __half* q11 = r1 + lx.i1 * TMP_C + 4;
        const uint2 t00 = *reinterpret_cast<const uint2*>(q00), t01 = *reinterpret_cast<const uint2*>(q01);
        const uint2 t10 = *reinterpret_cast<const uint2*>(q10), t11 = *reinterpret_cast<const uint2*>(q11);
        const uint4 u00 = *reinterpret_cast<const uint4*>(q00 + 4), u01 = *reinterpret_cast<const uint4*>(q01 + 4);
        const uint4 u10 = *reinterpret_cast<const uint4*>(q10 + 4), u11 = *reinterpret_cast<const uint4*>(q11 + 4);
        const __half2 *h00 = reinterpret_cast<const __half2*>(&t00), *h01 = reinterpret_cast<const __half2*>(&t01);
        const __half2 *h10 = reinterpret_cast<const __half2*>(&t10), *h11 = reinterpret_cast<const __half2*>(&t11);
        const __half2 *g00 = reinterpret_cast<const __half2*>(&u00), *g01 = reinterpret_cast<const __half2*>(&u01);
        const __half2 *g10 = reinterpret_cast<const __half2*>(&u10), *g11 = reinterpret_cast<const __half2*>(&u11);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          mfh[i] = __hfma2(a11, h11[i], __hfma2(a10, h10[i], __hfma2(a01, h01[i], __hmul2(a00, h00[i]))));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mfh[2 + i] = __hfma2(a11, g11[i], __hfma2(a10, g10[i], __hfma2(a01, g01[i], __hmul2(a00, g00[i]))));
      }
      }
  }
  // warps + accumulation over the samples
  __half2 A[4], Cc[4];
  const __half2 qw = __float2half2_rn(1.f / (float)NQ);
  float fs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int y = y0 + q / NS, x = x0 + q % NS;
    if (fbuf_store) st4(fbuf + (((size_t)b * H + y) * W + x) * 4, flow[q]);
    __half2 a[4], c[4];
    warp8_h2(s0, H, W, geom, x, y, flow[q][0], flow[q][1], a);   // IFNet_HDv3.py:145,155
    warp8_h2(s1, H, W, geom, x, y, flow[q][2], flow[q][3], c);   // IFNet_HDv3.py:146,156
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (NQ == 1) { A[i] = a[i]; Cc[i] = c[i]; }
      else if (q == 0) { A[i] = __hmul2(qw, a[i]); Cc[i] = __hmul2(qw, c[i]); }
      else { A[i] = __hfma2(qw, a[i], A[i]); Cc[i] = __hfma2(qw, c[i], Cc[i]); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) fs[i] += flow[q][i];
  }
  // interpolate(flow, 1/s) * 1/s   (IFNet_HDv3.py:81)
  const float fsc = 1.f / ((float)NQ * (float)s);
  const float t = tstep[b];
  __half2 o[16];
  o[0] = A[2];                                   // warped img0 r, g
  o[1] = __halves2half2(__low2half(A[3]), __low2half(Cc[2]));   // img0 b, img1 r
  o[2] = __halves2half2(__high2half(Cc[2]), __low2half(Cc[3])); // img1 g, b
  o[3] = A[0]; o[4] = A[1];                      // warped f0
  o[5] = Cc[0]; o[6] = Cc[1];                    // warped f1
  o[7] = __halves2half2(__float2half_rn(t), __low2half(mfh[0]));      // timestep, mask
  o[8] = __halves2half2(__high2half(mfh[0]), __low2half(mfh[1]));     // feat 0..7
  o[9] = __halves2half2(__high2half(mfh[1]), __low2half(mfh[2]));
  o[10] = __halves2half2(__high2half(mfh[2]), __low2half(mfh[3]));
  o[11] = __halves2half2(__high2half(mfh[3]), __low2half(mfh[4]));
  o[12] = __floats2half2_rn(fs[0] * fsc, fs[1] * fsc);
  o[13] = __floats2half2_rn(fs[2] * fsc, fs[3] * fsc);
  o[14] = __float2half2_rn(0.f);
  o[15] = o[14];
  const int par = (Y & 1) * 2 + (X & 1);
  const size_t hw2 = (size_t)(Hs >> 1) * (Ws >> 1), pix2 = (size_t)(Y >> 1) * (Ws >> 1) + (X >> 1);
  __half* dst = planar ? xin + (((size_t)b * 16 + par * 4) * hw2 + pix2) * 8 : xin + (((size_t)b * hw2 + pix2) * 4 + par) * XIN_C;
  const size_t gstride = planar ? hw2 * 8 : 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 u;
    __half2* hh = reinterpret_cast<__half2*>(&u);
    hh[0] = o[4 * k]; hh[1] = o[4 * k + 1]; hh[2] = o[4 * k + 2]; hh[3] = o[4 * k + 3];
    *reinterpret_cast<uint4*>(dst + k * gstride) = u;
  }
}

// s >= 2 variant with the 2x2 sample block split over a lane pair: lane 2k + r evaluates COLUMN r (two samples, two pairs
// of warps) of cell k and the pair adds its halves with 15 half2 shuffles.  For s = 2 consecutive lanes are consecutive
// full-resolution pixels, so every 16-byte tap load of a warp covers one contiguous 512-byte span (block_input_h_kernel<2>
// strides its lanes by two pixels: twice the L1 wavefronts, which is what bounds these kernels -- ncu: lsu data-pipe 72 %).
__device__ __forceinline__ __half2 shfl_xor_h2(__half2 v, int m) {
  const unsigned u = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<unsigned*>(&v), m);
  return *reinterpret_cast<const __half2*>(&u);
}
__global__ void __launch_bounds__(128, RIFE_BI_MINB2R) block_input_h2c_kernel(LevelsArg L, int level, int s,
                                                            const __half* __restrict__ src0,
                                                            const __half* __restrict__ src1,
                                                            const float* __restrict__ gridx,
                                                            const float* __restrict__ gridy,
                                                            const float* __restrict__ tstep, __half* __restrict__ xin,
                                                            int B, int H, int W, WarpGeom geom, int planar,
                                                            float* __restrict__ fbuf, int fbuf_levels, int fbuf_store) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  const int Hs = H / s, Ws = W / s;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = tid & 1;
  const bool valid = (tid >> 1) < Ws;            // a lane pair is valid or not as a whole; nobody leaves before the shuffles
  const int X = valid ? (tid >> 1) : Ws - 1;
  const int Y = blockIdx.y, b = blockIdx.z;
  const int fr = geom.src_div == 1 ? b : b / geom.src_div;   // uniform branch: no integer division in the common case
  const __half* s0 = src0 + (size_t)fr * H * W * SRC_C;
  const __half* s1 = src1 + (size_t)fr * H * W * SRC_C;
  const int off = s / 2 - 1;
  const int x = X * s + off + r, y0 = Y * s + off;
  float flow[2][4];
  float* fp = fbuf + (((size_t)b * H + y0) * W + x) * 4;
  if (fbuf_levels > 0) {
    ld4(fp, flow[0]);
    ld4(fp + (size_t)W * 4, flow[1]);
  }
  float mf[12];  // this column's share of channels 4..15 of the newest level (mask, feat x8, pad), weights sum to 1/2
#pragma unroll 1
  for (int j = fbuf_levels; j < level; ++j) {
    const int h = L.h[j], w = L.w[j];
    if (L.s[j] == 1) continue;
    const __half* base = reinterpret_cast<const __half*>(L.ptr[j]) + (size_t)b * h * w * TMP_C;
    const float sc = L.sc[j];
    const Lerp2 ly = make_lerp2(y0, h, L.rs[j]);
    const Lerp lx = make_lerp(x, w, L.rs[j]);
    float wq[2][4], am[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float ay = q ? ly.a1 : ly.a0, ax = lx.l1;
      wq[q][0] = (1.f - ay) * (1.f - ax); wq[q][1] = (1.f - ay) * ax; wq[q][2] = ay * (1.f - ax); wq[q][3] = ay * ax;
#pragma unroll
      for (int k = 0; k < 4; ++k) am[k] += wq[q][k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) am[k] *= 0.25f;   // this column holds two of the four samples
    level_taps<2>(base + ly.i0 * w * TMP_C, base + ly.i1 * w * TMP_C, lx.i0 * TMP_C, lx.i1 * TMP_C, j == level - 1, wq, am,
                  sc, j == 0, flow, mf);
  }
  __half2 A[4], Cc[4];
  const __half2 qw = __float2half2_rn(0.25f);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (fbuf_store && valid) st4(fp + (size_t)q * W * 4, flow[q]);
    __half2 a[4], c[4];
    warp8_h2(s0, H, W, geom, x, y0 + q, flow[q][0], flow[q][1], a);   // IFNet_HDv3.py:145,155
    warp8_h2(s1, H, W, geom, x, y0 + q, flow[q][2], flow[q][3], c);   // IFNet_HDv3.py:146,156
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (q == 0) { A[i] = __hmul2(qw, a[i]); Cc[i] = __hmul2(qw, c[i]); }
      else { A[i] = __hfma2(qw, a[i], A[i]); Cc[i] = __hfma2(qw, c[i], Cc[i]); }
    }
  }
  // interpolate(flow, 1/s) * 1/s   (IFNet_HDv3.py:81): this column's half of the 2x2 mean
  const float fsc = 0.25f / (float)s;
  const float t = tstep[b];
  __half2 o[16];
  o[7] = __floats2half2_rn(0.5f * t, mf[0]);     // timestep (exact after the pair's add), mask
  o[8] = __floats2half2_rn(mf[1], mf[2]);        // feat 0..7
  o[9] = __floats2half2_rn(mf[3], mf[4]);
  o[10] = __floats2half2_rn(mf[5], mf[6]);
  o[11] = __floats2half2_rn(mf[7], mf[8]);
  o[12] = __floats2half2_rn((flow[0][0] + flow[1][0]) * fsc, (flow[0][1] + flow[1][1]) * fsc);
  o[13] = __floats2half2_rn((flow[0][2] + flow[1][2]) * fsc, (flow[0][3] + flow[1][3]) * fsc);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    A[i] = __hadd2(A[i], shfl_xor_h2(A[i], 1));
    Cc[i] = __hadd2(Cc[i], shfl_xor_h2(Cc[i], 1));
  }
#pragma unroll
  for (int i = 7; i < 14; ++i) o[i] = __hadd2(o[i], shfl_xor_h2(o[i], 1));
  o[0] = A[2];                                   // warped img0 r, g
  o[1] = __halves2half2(__low2half(A[3]), __low2half(Cc[2]));   // img0 b, img1 r
  o[2] = __halves2half2(__high2half(Cc[2]), __low2half(Cc[3])); // img1 g, b
  o[3] = A[0]; o[4] = A[1];                      // warped f0
  o[5] = Cc[0]; o[6] = Cc[1];                    // warped f1
  o[14] = __float2half2_rn(0.f);
  o[15] = o[14];
  if (!valid) return;
  const int par = (Y & 1) * 2 + (X & 1);
  const size_t hw2 = (size_t)(Hs >> 1) * (Ws >> 1), pix2 = (size_t)(Y >> 1) * (Ws >> 1) + (X >> 1);
  __half* dst = planar ? xin + (((size_t)b * 16 + par * 4) * hw2 + pix2) * 8 : xin + (((size_t)b * hw2 + pix2) * 4 + par) * XIN_C;
  const size_t gstride = planar ? hw2 * 8 : 8;
  // lane r of the pair stores the 16-byte groups 2r, 2r + 1
  uint4 u0, u1;
  __half2* h0 = reinterpret_cast<__half2*>(&u0);
  __half2* h1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h0[i] = r ? o[8 + i] : o[i];
    h1[i] = r ? o[12 + i] : o[4 + i];
  }
  *reinterpret_cast<uint4*>(dst + (2 * r) * gstride) = u0;
  *reinterpret_cast<uint4*>(dst + (2 * r + 1) * gstride) = u1;
}

// ---------------------------------------------------------------------------------------------------
// eval: full-resolution flow / mask / warped images / merged after `nlev` blocks (one thread per pixel)
template <typename T, typename Tt, typename Tio>
__global__ void __launch_bounds__(256) eval_kernel(LevelsArg L, int nlev, const T* __restrict__ src0,
                                                   const T* __restrict__ src1, const float* __restrict__ gridx,
                                                   const float* __restrict__ gridy, Tio* __restrict__ oflow,
                                                   Tio* __restrict__ omask, Tio* __restrict__ ow0,
                                                   Tio* __restrict__ ow1, Tio* __restrict__ omerged, int B, int C,
                                                   int H, int W, WarpGeom geom, const float* __restrict__ fbuf,
                                                   int fbuf_levels, ImgLayout lay, const void* const* iot) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  if (iot) omerged = reinterpret_cast<Tio*>(const_cast<void*>(iot[2]));
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= lay.w || y >= lay.h) return;   // outputs are cropped to the caller's frame
  const size_t plane = (size_t)lay.h * lay.w;
  float flow[4] = {0.f, 0.f, 0.f, 0.f}, last[16];
  if (fbuf_levels > 0) ld4(fbuf + (((size_t)b * H + y) * W + x) * 4, flow);
  eval_state<Tt>(L, nlev, fbuf_levels, b, y, x, flow, last);
  const size_t pix = (size_t)y * lay.w + x;  // flow / mask outputs are always NCHW (B,4|1,h,w)
  if (oflow) {
#pragma unroll
    for (int c = 0; c < 4; ++c) oflow[((size_t)b * 4 + c) * plane + pix] = from_f<Tio>(flow[c]);
  }
  if (omask) omask[(size_t)b * plane + pix] = from_f<Tio>(last[4]);
  if (ow0 || ow1 || omerged) {
    const float gx = gridx[x], gy = gridy[y];
    float a[8], c[8];
    warp8<T>(src0 + (size_t)(b / geom.src_div) * H * W * SRC_C, H, W, geom, gx, gy, flow[0], flow[1], a);  // IFNet_HDv3.py:155
    warp8<T>(src1 + (size_t)(b / geom.src_div) * H * W * SRC_C, H, W, geom, gx, gy, flow[2], flow[3], c);  // IFNet_HDv3.py:156
    const float m = 1.f / (1.f + expf(-last[4]));                                         // IFNet_HDv3.py:158
    for (int ch = 0; ch < C; ++ch) {
      const float w0 = a[4 + ch], w1 = c[4 + ch];
      const size_t o = (size_t)(b * lay.sb + ch * lay.sc + y * lay.sy + x * lay.sx);
      if (ow0) ow0[o] = from_f<Tio>(w0);
      if (ow1) ow1[o] = from_f<Tio>(w1);
      if (omerged) omerged[o] = from_f<Tio>(w0 * m + w1 * (1.f - m));                     // IFNet_HDv3.py:159
    }
  }
}


// fp16-storage fast path of the final merge (Model.inference only wants merged[-1]): the full-resolution flow of
// the first nlev-1 blocks is in `fbuf`, the last block's tmp is at full resolution (s = 1), and only the image
// channels (upper 8 bytes of each 16-byte SRC pixel) are blended.  fp32 arithmetic, fp32 coordinates.
__device__ __forceinline__ void warp_img4(const __half* __restrict__ src, int H, int W, const WarpGeom& g, float gx,
                                          float gy, float fx, float fy, float (&out)[4]) {
  const float nx = gx + __fmul_rn(fx, g.inv_halfw);
  const float ny = gy + __fmul_rn(fy, g.inv_halfh);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.f), 0.5f), g.wm1);
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.f), 0.5f), g.hm1);
  ix = fminf(g.wm1, fmaxf(ix, 0.f));
  iy = fminf(g.hm1, fmaxf(iy, 0.f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float ax = ix - x0f, ay = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const int dx = (x0 + 1 < W) ? SRC_C : 0, dy = (y0 + 1 < H) ? W * SRC_C : 0;  // beyond the edge the weight is 0
  const __half* p = src + (y0 * W + x0) * SRC_C + 4;
  float v00[4], v01[4], v10[4], v11[4];
  ld4(p, v00); ld4(p + dx, v01); ld4(p + dy, v10); ld4(p + dy + dx, v11);
  const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = fmaf(w11, v11[c], fmaf(w10, v10[c], fmaf(w01, v01[c], w00 * v00[c])));
}

template <typename Tio, int PITCH>
__global__ void __launch_bounds__(256) merge_h_kernel(const __half* __restrict__ tmp, const __half* __restrict__ src0,
                                                      const __half* __restrict__ src1, const float* __restrict__ gridx,
                                                      const float* __restrict__ gridy, const float* __restrict__ fbuf,
                                                      Tio* __restrict__ omerged, int C, int H, int W, WarpGeom geom,
                                                      ImgLayout lay, const void* const* iot) {
  pdl_wait();  // inputs are earlier kernels' outputs (see common.cuh: PDL contract)
  RIFE_PIX_TRIGGER();
  if (iot) omerged = reinterpret_cast<Tio*>(const_cast<void*>(iot[2]));
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= lay.w || y >= lay.h) return;
  const size_t pix = ((size_t)b * H + y) * W + x;
  float flow[4], t8[8];
  ld4(fbuf + pix * 4, flow);
  ld8(tmp + pix * PITCH, t8);   // channels 0..3 flow of the last block (scale 1), 4 mask logit
#pragma unroll
  for (int c = 0; c < 4; ++c) flow[c] += t8[c];                                    // IFNet_HDv3.py:152
  const float gx = gridx[x], gy = gridy[y];
  const int fr = geom.src_div == 1 ? b : b / geom.src_div;
  float a[4], c[4];
  warp_img4(src0 + (size_t)fr * H * W * SRC_C, H, W, geom, gx, gy, flow[0], flow[1], a);   // IFNet_HDv3.py:155
  warp_img4(src1 + (size_t)fr * H * W * SRC_C, H, W, geom, gx, gy, flow[2], flow[3], c);   // IFNet_HDv3.py:156
  const float m = 1.f / (1.f + __expf(-t8[4]));                                    // IFNet_HDv3.py:158
  float o[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) o[ch] = a[ch] * m + c[ch] * (1.f - m);            // IFNet_HDv3.py:159
  // the two frame layouts of the API boundary, spelled out (generic 64-bit strides cost this issue-bound kernel ~25 %
  // of its instructions): uint8 = HWC (B,h,w,C), float = NCHW (B,C,h,w), both cropped to the caller's (h, w)
  if constexpr (sizeof(Tio) == 1) {
    Tio* q = omerged + (((size_t)b * lay.h + y) * lay.w + x) * C;
    q[0] = from_f<Tio>(o[0]); q[1] = from_f<Tio>(o[1]); q[2] = from_f<Tio>(o[2]);
    if (C == 4) q[3] = from_f<Tio>(o[3]);
  } else {
    const size_t plane = (size_t)lay.h * lay.w;
    Tio* q = omerged + ((size_t)b * C * lay.h + y) * lay.w + x;
    q[0] = from_f<Tio>(o[0]); q[plane] = from_f<Tio>(o[1]); q[2 * plane] = from_f<Tio>(o[2]);
    if (C == 4) q[3 * plane] = from_f<Tio>(o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------
// stand-alone operators (NCHW, op-level API)
template <typename Tio>
__global__ void __launch_bounds__(256) warp_op_kernel(const Tio* __restrict__ in, const Tio* __restrict__ flow,
                                                      Tio* __restrict__ out, const float* __restrict__ gridx,
                                                      const float* __restrict__ gridy, int B, int C, int H, int W,
                                                      WarpGeom g) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)H * W;
  if (idx >= (size_t)B * plane) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int b = (int)(idx / plane);
  const size_t pix = (size_t)y * W + x;
  const float fx = to_f(flow[((size_t)b * 2 + 0) * plane + pix]);
  const float fy = to_f(flow[((size_t)b * 2 + 1) * plane + pix]);
  const float nx = gridx[x] + __fdiv_rn(fx, g.halfw);
  const float ny = gridy[y] + __fdiv_rn(fy, g.halfh);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.f), 0.5f), g.wm1);
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.f), 0.5f), g.hm1);
  ix = fminf(g.wm1, fmaxf(ix, 0.f));
  iy = fminf(g.hm1, fmaxf(iy, 0.f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = x0f + 1.f, y1f = y0f + 1.f;
  const float wnw = (x1f - ix) * (y1f - iy), wne = (ix - x0f) * (y1f - iy);
  const float wsw = (x1f - ix) * (iy - y0f), wse = (ix - x0f) * (iy - y0f);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const bool xin = (x0 + 1) < W, yin = (y0 + 1) < H;
  for (int c = 0; c < C; ++c) {
    const Tio* p = in + ((size_t)b * C + c) * plane + (size_t)y0 * W + x0;
    float r = to_f(p[0]) * wnw;
    if (xin) r += to_f(p[1]) * wne;
    if (yin) {
      r += to_f(p[W]) * wsw;
      if (xin) r += to_f(p[W + 1]) * wse;
    }
    out[((size_t)b * C + c) * plane + pix] = from_f<Tio>(r);
  }
}

// F.interpolate(bilinear, align_corners=False) with scale_factor = num/den (power-of-two ratio), NCHW fp32
__global__ void __launch_bounds__(256) resize_op_kernel(const float* __restrict__ x, float* __restrict__ out, int BC,
                                                        int H, int W, int Ho, int Wo, float ratio, float post_mul) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)BC * Ho * Wo) return;
  const int ox = (int)(idx % Wo);
  const int oy = (int)((idx / Wo) % Ho);
  const int bc = (int)(idx / ((size_t)Wo * Ho));
  const Lerp ly = make_lerp(oy, H, ratio), lx = make_lerp(ox, W, ratio);
  const float* p = x + (size_t)bc * H * W;
  const float v00 = p[(size_t)ly.i0 * W + lx.i0], v01 = p[(size_t)ly.i0 * W + lx.i1];
  const float v10 = p[(size_t)ly.i1 * W + lx.i0], v11 = p[(size_t)ly.i1 * W + lx.i1];
  out[idx] = (ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11)) * post_mul;
}

// torch.linspace(-1, 1, n) scalar formula (fallback when the caller passes no host grid)
__global__ void fill_grid_kernel(float* g, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n == 1) { g[0] = -1.f; return; }
  const float step = 2.f / (float)(n - 1);
  g[i] = (i < n / 2) ? __fadd_rn(-1.f, __fmul_rn(step, (float)i)) : __fsub_rn(1.f, __fmul_rn(step, (float)(n - 1 - i)));
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                           int C, int H, int W, int Cpad, int s2d) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * H * W * Cpad) return;
  const int c = (int)(idx % Cpad);
  const int xx = (int)((idx / Cpad) % W);
  const int yy = (int)((idx / ((size_t)Cpad * W)) % H);
  const int b = (int)(idx / ((size_t)Cpad * W * H));
  const float v = c < C ? x[(((size_t)b * C + c) * H + yy) * W + xx] : 0.f;
  if (!s2d) {
    y[idx] = v;
  } else {
    y[((((size_t)b * (H >> 1) + (yy >> 1)) * (W >> 1) + (xx >> 1)) * 4 + ((yy & 1) * 2 + (xx & 1))) * Cpad + c] = v;
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                           int C, int H, int W, int Cpad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * C * H * W) return;
  const int xx = (int)(idx % W);
  const int yy = (int)((idx / W) % H);
  const int c = (int)((idx / ((size_t)W * H)) % C);
  const int b = (int)(idx / ((size_t)W * H * C));
  y[idx] = x[(((size_t)b * H + yy) * W + xx) * Cpad + c];
}


inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// scale > 1 (block scale s < 1): xin = F.interpolate(x, scale_factor=up) of the full-resolution 32-channel block input
// (IFNet_HDv3.py:79-81: flow channels additionally * 1/scale = up).  `xfull` is s2d pixel-major [B][H/2][W/2][4][32].
template <typename T>
__global__ void __launch_bounds__(128) upsample_xin_kernel(const T* __restrict__ xfull, T* __restrict__ xin, int B, int H,
                                                           int W, int up, int planar) {
  pdl_wait();
  const int Hu = H * up, Wu = W * up;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= Wu) return;
  const float ratio = 1.f / (float)up;
  const Lerp ly = make_lerp(Y, H, ratio), lx = make_lerp(X, W, ratio);
  auto px = [&](int y, int x) -> const T* {
    return xfull + ((((size_t)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * XIN_C;
  };
  const T *p00 = px(ly.i0, lx.i0), *p01 = px(ly.i0, lx.i1), *p10 = px(ly.i1, lx.i0), *p11 = px(ly.i1, lx.i1);
  const int par = (Y & 1) * 2 + (X & 1);
  const size_t hw2 = (size_t)(Hu >> 1) * (Wu >> 1), pix2 = (size_t)(Y >> 1) * (Wu >> 1) + (X >> 1);
  T* dst = planar ? xin + (((size_t)b * 16 + par * 4) * hw2 + pix2) * 8 : xin + (((size_t)b * hw2 + pix2) * 4 + par) * XIN_C;
  const size_t gstride = planar ? hw2 * 8 : 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v00[8], v01[8], v10[8], v11[8], o[8];
    ld8(p00 + 8 * k, v00); ld8(p01 + 8 * k, v01); ld8(p10 + 8 * k, v10); ld8(p11 + 8 * k, v11);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = ly.l0 * (lx.l0 * v00[i] + lx.l1 * v01[i]) + ly.l1 * (lx.l0 * v10[i] + lx.l1 * v11[i]);
      if (k == 3 && i < 4) o[i] = o[i] * (float)up;      // flow: interpolate(flow, 1/s) * 1/s
    }
    st8(dst + k * gstride, o);
  }
}

int launch_upsample_xin(const PixelCtx& c, const void* xfull, void* xin, int up, bool planar, cudaStream_t st) {
  const dim3 grid(nblk((size_t)c.W * up, 128), (unsigned)(c.H * up), (unsigned)c.B);
  if (c.half_act)
    RIFE_CUDA_TRY(launch_pdl(upsample_xin_kernel<__half>, grid, dim3(128), 0, st, (const __half*)xfull, (__half*)xin, c.B, c.H,
                             c.W, up, planar ? 1 : 0));
  else
    RIFE_CUDA_TRY(launch_pdl(upsample_xin_kernel<float>, grid, dim3(128), 0, st, (const float*)xfull, (float*)xin, c.B, c.H, c.W,
                             up, planar ? 1 : 0));
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_prep(const PixelCtx& c, const void* img0, const void* img1, int io_dtype, void* src_all, void* head_in,
                cudaStream_t st) {
  const int fh = c.h ? c.h : c.H, fw = c.w ? c.w : c.W;
  const ImgLayout lay = make_layout(io_dtype, c.C, fh, fw);
  if (c.planar_head && c.half_act && c.W % 8 == 0 && fh == c.H && fw == c.W && io_dtype != RIFE_U8) {
    const dim3 g(nblk((size_t)(c.W / 8), 256), (unsigned)(c.H / 2), (unsigned)c.F);
    if (io_dtype == RIFE_F16)
      RIFE_CUDA_TRY(launch_pdl(prep_planar_kernel<__half>, g, dim3(256), 0, st, (const __half*)img0, (const __half*)img1, (__half*)head_in, c.nf0, c.C, c.H, c.W, c.tile_ctrs, c.n_tile_ctrs, c.io_table));
    else if (io_dtype == RIFE_F32)
      RIFE_CUDA_TRY(launch_pdl(prep_planar_kernel<float>, g, dim3(256), 0, st, (const float*)img0, (const float*)img1, (__half*)head_in, c.nf0, c.C, c.H, c.W, c.tile_ctrs, c.n_tile_ctrs, c.io_table));
    else
      RIFE_FAIL(RIFE_ERR_DTYPE, "launch_prep: bad io dtype");
    RIFE_LAUNCH_CHECK();
    return RIFE_OK;
  }
  if (c.planar_head && c.half_act && io_dtype == RIFE_U8 && c.C == 3 && fw == c.W && c.W % 8 == 0 &&
      ((uintptr_t)img0 % 8) == 0 && ((uintptr_t)img1 % 8) == 0 && ((size_t)fh * fw * 3) % 8 == 0) {
    const dim3 g(nblk((size_t)(c.W / 8), 256), (unsigned)(c.H / 2), (unsigned)c.F);
    RIFE_CUDA_TRY(launch_pdl(prep_planar_u8_kernel, g, dim3(256), 0, st, (const uint8_t*)img0, (const uint8_t*)img1,
                             (__half*)head_in, c.nf0, fh, c.H, c.W, c.tile_ctrs, c.n_tile_ctrs, c.io_table));
    RIFE_LAUNCH_CHECK();
    return RIFE_OK;
  }
  const size_t nq = (size_t)c.F * (c.H / 2) * (c.W / 2);
  dim3 grid(nblk(nq, 256));
#define RIFE_PREP(TIO, T)                                                                                        \
  RIFE_CUDA_TRY(launch_pdl(prep_kernel<TIO, T>, grid, dim3(256), 0, st, (const TIO*)img0, (const TIO*)img1,            \
                           (T*)src_all, (T*)head_in, c.F, c.nf0, c.C, c.H, c.W, c.planar_head ? 1 : 0,                \
                           c.planar_head ? 0 : 1, lay, c.tile_ctrs, c.n_tile_ctrs, c.io_table))
  if (io_dtype == RIFE_F32 && !c.half_act) RIFE_PREP(float, float);
  else if (io_dtype == RIFE_F32 && c.half_act) RIFE_PREP(float, __half);
  else if (io_dtype == RIFE_F16 && !c.half_act) RIFE_PREP(__half, float);
  else if (io_dtype == RIFE_F16 && c.half_act) RIFE_PREP(__half, __half);
  else if (io_dtype == RIFE_U8 && !c.half_act) RIFE_PREP(uint8_t, float);
  else if (io_dtype == RIFE_U8 && c.half_act) RIFE_PREP(uint8_t, __half);
  else RIFE_FAIL(RIFE_ERR_DTYPE, "launch_prep: bad io dtype");
#undef RIFE_PREP
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}


int launch_block_input(const PixelCtx& c, int level, const TmpLevel* levels, int s, void* xin, bool planar,
                       float* fbuf, int fbuf_levels, bool fbuf_store, cudaStream_t st) {
  LevelsArg L = pack_levels(levels, level);
  WarpGeom g = make_geom(c.H, c.W);
  g.src_div = c.src_div;
  const dim3 grid(nblk((size_t)(c.W / s), 128), (unsigned)(c.H / s), (unsigned)c.B);
#define RIFE_BI(T, NS_, L0_)                                                                                       \
  RIFE_CUDA_TRY(launch_pdl(block_input_kernel<T, T, NS_, L0_>, grid, dim3(128), 0, st, L, level, s, (const T*)c.src0, \
                           (const T*)c.src1, c.gridx, c.gridy, c.tstep, (T*)xin, c.B, c.H, c.W, g, planar ? 1 : 0,   \
                           fbuf, fbuf_levels, fbuf_store ? 1 : 0))
  const bool l0 = level == 0, s1 = s == 1;
  static const bool no_fast = getenv("RIFE_NO_FAST_INPUT") != nullptr;
  if (c.half_act && !l0 && !no_fast && !c.any_up) {
#define RIFE_BIH(NS_)                                                                                                 \
  RIFE_CUDA_TRY(launch_pdl(block_input_h_kernel<NS_>, grid, dim3(128), 0, st, L, level, s, (const __half*)c.src0,    \
                           (const __half*)c.src1, c.gridx, c.gridy, c.tstep, (__half*)xin, c.B, c.H, c.W, g,          \
                           planar ? 1 : 0, fbuf, fbuf_levels, fbuf_store ? 1 : 0))
    // s = 2 (every pixel sampled; issue-bound): one thread per 2x2 block (0.36 ms vs 0.39 ms at 8 x 1080p);
    // s >= 4 (sparse samples; L1-wavefront-bound): the lane-pair split (b1 + b2: 0.156 vs 0.171 ms)
    static const int pair_min_s = getenv("RIFE_BI_PAIR_MIN_S") ? atoi(getenv("RIFE_BI_PAIR_MIN_S")) : 4;
    if (s1) {
      RIFE_BIH(1);
    } else if (s < pair_min_s) {
      RIFE_BIH(2);
    } else {
      const dim3 grid2(nblk((size_t)(c.W / s) * 2, 128), (unsigned)(c.H / s), (unsigned)c.B);
      RIFE_CUDA_TRY(launch_pdl(block_input_h2c_kernel, grid2, dim3(128), 0, st, L, level, s, (const __half*)c.src0,
                               (const __half*)c.src1, c.gridx, c.gridy, c.tstep, (__half*)xin, c.B, c.H, c.W, g,
                               planar ? 1 : 0, fbuf, fbuf_levels, fbuf_store ? 1 : 0));
    }
#undef RIFE_BIH
  } else if (c.half_act) {
    if (l0 && s1) RIFE_BI(__half, 1, true); else if (l0) RIFE_BI(__half, 2, true);
    else if (s1) RIFE_BI(__half, 1, false); else RIFE_BI(__half, 2, false);
  } else {
    if (l0 && s1) RIFE_BI(float, 1, true); else if (l0) RIFE_BI(float, 2, true);
    else if (s1) RIFE_BI(float, 1, false); else RIFE_BI(float, 2, false);
  }
#undef RIFE_BI
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_eval(const PixelCtx& c, int nlev, const TmpLevel* levels, const EvalOut& o, int io_dtype, const float* fbuf,
                int fbuf_levels, cudaStream_t st) {
  const size_t n = (size_t)c.B * c.H * c.W;
  LevelsArg L = pack_levels(levels, nlev);
  WarpGeom g = make_geom(c.H, c.W);
  g.src_div = c.src_div;
  const int fh = c.h ? c.h : c.H, fw = c.w ? c.w : c.W;
  const ImgLayout lay = make_layout(io_dtype, c.C, fh, fw);
  if (io_dtype == RIFE_U8 && (o.flow || o.mask)) RIFE_FAIL(RIFE_ERR_DTYPE, "launch_eval: flow / mask outputs need a float io dtype");
  const dim3 egrid(nblk((size_t)fw, 256), (unsigned)fh, (unsigned)c.B);
  static const bool no_fast_merge = getenv("RIFE_NO_FAST_MERGE") != nullptr;
  if (c.half_act && !c.any_up && !no_fast_merge && o.merged && !o.flow && !o.mask && !o.warped0 && !o.warped1 && nlev >= 2 &&
      fbuf_levels == nlev - 1 && levels[nlev - 1].s == 1) {
    const __half* tmp = (const __half*)levels[nlev - 1].ptr;
#define RIFE_MERGE(TIO, PITCH)                                                                                       \
  RIFE_CUDA_TRY(launch_pdl(merge_h_kernel<TIO, PITCH>, egrid, dim3(256), 0, st, tmp, (const __half*)c.src0,           \
                           (const __half*)c.src1, c.gridx, c.gridy, fbuf, (TIO*)o.merged, c.C, c.H, c.W, g, lay, c.io_table))
    if (levels[nlev - 1].pitch == 8) {
      if (io_dtype == RIFE_F32) RIFE_MERGE(float, 8);
      else if (io_dtype == RIFE_F16) RIFE_MERGE(__half, 8);
      else RIFE_MERGE(uint8_t, 8);
    } else {
      if (io_dtype == RIFE_F32) RIFE_MERGE(float, TMP_C);
      else if (io_dtype == RIFE_F16) RIFE_MERGE(__half, TMP_C);
      else RIFE_MERGE(uint8_t, TMP_C);
    }
#undef RIFE_MERGE
    RIFE_LAUNCH_CHECK();
    return RIFE_OK;
  }
  for (int i = 0; i < nlev; ++i)
    if (levels[i].pitch != TMP_C) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "launch_eval: an 8-channel tmp level needs the fast merge path");
#define RIFE_EVAL(T, TIO)                                                                                         \
  RIFE_CUDA_TRY(launch_pdl(eval_kernel<T, T, TIO>, egrid, dim3(256), 0, st, L, nlev, (const T*)c.src0,            \
                           (const T*)c.src1, c.gridx, c.gridy, (TIO*)o.flow, (TIO*)o.mask, (TIO*)o.warped0,       \
                           (TIO*)o.warped1, (TIO*)o.merged, c.B, c.C, c.H, c.W, g, fbuf, fbuf_levels, lay, c.io_table))
  if (!c.half_act && io_dtype == RIFE_F32) RIFE_EVAL(float, float);
  else if (!c.half_act && io_dtype == RIFE_F16) RIFE_EVAL(float, __half);
  else if (c.half_act && io_dtype == RIFE_F32) RIFE_EVAL(__half, float);
  else if (c.half_act && io_dtype == RIFE_F16) RIFE_EVAL(__half, __half);
  else if (!c.half_act && io_dtype == RIFE_U8) RIFE_EVAL(float, uint8_t);
  else if (c.half_act && io_dtype == RIFE_U8) RIFE_EVAL(__half, uint8_t);
  else RIFE_FAIL(RIFE_ERR_DTYPE, "launch_eval: bad io dtype");
#undef RIFE_EVAL
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

// timesteps of one call as kernel parameters: no copy-engine work on the compute stream (a small cudaMemcpyAsync would
// queue behind the frame uploads of the next calls that are in flight on the H2D engine)
struct TsArg { float v[256]; };
__global__ void set_tstep_kernel(TsArg a, float* __restrict__ dst, int n) {
  const int i = threadIdx.x;
  if (i < n) dst[i] = a.v[i];
}
int launch_set_tstep(float* dst, const float* host, int n, cudaStream_t st) {
  for (int o = 0; o < n; o += 256) {
    TsArg a;
    const int m = n - o < 256 ? n - o : 256;
    for (int i = 0; i < m; ++i) a.v[i] = host[o + i];
    set_tstep_kernel<<<1, 256, 0, st>>>(a, dst + o, m);
  }
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

// the frame pointers of one call, written through kernel parameters into the plan's table (read by the first / last
// kernel of a replayed CUDA graph)
__global__ void set_io_table_kernel(const void** tab, const void* a, const void* b, const void* c) {
  if (threadIdx.x == 0) { tab[0] = a; tab[1] = b; tab[2] = c; }
}
int launch_set_io_table(const void** tab, const void* img0, const void* img1, const void* out, cudaStream_t st) {
  set_io_table_kernel<<<1, 32, 0, st>>>(tab, img0, img1, out);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_warp_op(const void* input, const void* flow, void* out, int B, int C, int H, int W, int dtype,
                   const float* gridx, const float* gridy, cudaStream_t st) {
  const size_t n = (size_t)B * H * W;
  WarpGeom g = make_geom(H, W);
  if (dtype == RIFE_F32)
    warp_op_kernel<float><<<nblk(n, 256), 256, 0, st>>>((const float*)input, (const float*)flow, (float*)out, gridx,
                                                         gridy, B, C, H, W, g);
  else if (dtype == RIFE_F16)
    warp_op_kernel<__half><<<nblk(n, 256), 256, 0, st>>>((const __half*)input, (const __half*)flow, (__half*)out,
                                                          gridx, gridy, B, C, H, W, g);
  else
    RIFE_FAIL(RIFE_ERR_DTYPE, "rife_warp: dtype must be RIFE_F32 or RIFE_F16");
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_resize_op(const float* x, float* out, int B, int C, int H, int W, int num, int den, float post_mul,
                     cudaStream_t st) {
  const int Ho = (int)((long long)H * num / den), Wo = (int)((long long)W * num / den);
  const size_t n = (size_t)B * C * Ho * Wo;
  const float ratio = (float)den / (float)num;  // 1 / scale_factor
  resize_op_kernel<<<nblk(n, 256), 256, 0, st>>>(x, out, B * C, H, W, Ho, Wo, ratio, post_mul);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_fill_grid(float* g, int n, cudaStream_t st) {
  fill_grid_kernel<<<nblk(n, 128), 128, 0, st>>>(g, n);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int H, int W, int Cpad, bool s2d,
                            cudaStream_t st) {
  const size_t n = (size_t)B * H * W * Cpad;
  nchw_to_nhwc_kernel<<<nblk(n, 256), 256, 0, st>>>(x, y, B, C, H, W, Cpad, s2d ? 1 : 0);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_nhwc_to_nchw_f32(const float* x, float* y, int B, int C, int H, int W, int Cpad, cudaStream_t st) {
  const size_t n = (size_t)B * C * H * W;
  nhwc_to_nchw_kernel<<<nblk(n, 256), 256, 0, st>>>(x, y, B, C, H, W, Cpad);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

}  // namespace rife
