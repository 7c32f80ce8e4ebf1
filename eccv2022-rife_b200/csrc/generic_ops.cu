// Generic NHWC operators (fp32 or fp16 per tensor) used by the ECCV-paper network path (model/IFNet.py, model/refine.py of the reference):
// every tensor is [B][H][W][pitch] and an operator reads / writes a channel slice [off, off + C), so that torch.cat
// (model/IFNet.py:75-79, model/refine.py:73-80) costs nothing: producers write straight into the consumer's slice.
#include <string>

#include "common.cuh"

namespace rife {
namespace {

inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

struct Lerp {
  int i0, i1;
  float l0, l1;
};
// ATen area_pixel_compute_source_index(align_corners=False, scale_factor given): ratio = 1 / scale_factor
__device__ __forceinline__ Lerp make_lerp(int d, int in_size, float ratio) {
  float src = ratio * ((float)d + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  int i0 = min((int)src, in_size - 1);
  Lerp L;
  L.i0 = i0;
  L.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  L.l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  L.l0 = 1.f - L.l1;
  return L;
}

template <typename Tio, typename T>
__global__ void __launch_bounds__(256) nchw_to_slice_kernel(const Tio* __restrict__ src, T* __restrict__ dst, int B,
                                                            int C, int H, int W, int pitch, int off) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int c = (int)((i / ((size_t)W * H)) % C), b = (int)(i / ((size_t)W * H * C));
  dst[(((size_t)b * H + y) * W + x) * pitch + off + c] = from_f<T>(to_f(src[i]));
}

template <typename Tio, typename T>
__global__ void __launch_bounds__(256) slice_to_nchw_kernel(const T* __restrict__ src, int pitch, int off,
                                                            Tio* __restrict__ dst, int B, int C, int H, int W, int sig) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int c = (int)((i / ((size_t)W * H)) % C), b = (int)(i / ((size_t)W * H * C));
  float v = to_f(src[(((size_t)b * H + y) * W + x) * pitch + off + c]);
  if (sig) v = 1.f / (1.f + expf(-v));
  dst[i] = from_f<Tio>(v);
}

// ---- 8-channel chunks: one thread moves channels [c0, c0 + 8) of one pixel; 16-byte (fp16) / 2 x 16-byte (fp32)
// vector accesses when the slice is 8-aligned (`vec`), scalar otherwise (raw 3 / 6 / 13 / 17-channel slices).
template <typename T>
__device__ __forceinline__ void load_chunk(const T* __restrict__ p, int nv, bool vec, float (&v)[8]) {
  if (vec) {
    ld8(p, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < nv ? to_f(p[i]) : 0.f;
  }
}
template <typename T>
__device__ __forceinline__ void store_chunk(T* __restrict__ p, int nv, bool vec, const float (&v)[8]) {
  if (vec) {
    st8(p, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) p[i] = from_f<T>(v[i]);
  }
}

// F.interpolate(bilinear, align_corners=False) of a channel slice; dst = (acc ? dst : 0) + value * mul
template <typename T, typename Td>
__global__ void __launch_bounds__(256) resize_nhwc_kernel(const T* __restrict__ src, int sp, int so,
                                                          Td* __restrict__ dst, int dp, int dof, int B, int C, int Hs,
                                                          int Ws, int Hd, int Wd, float ratio, float mul, int acc,
                                                          int vsrc, int vdst) {
  const int nch = (C + 7) >> 3;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * Hd * Wd * nch) return;
  const int ch = (int)(i % nch);
  const size_t pix = i / nch;
  const int x = (int)(pix % Wd), y = (int)((pix / Wd) % Hd), b = (int)(pix / ((size_t)Wd * Hd));
  const int c0 = ch * 8, nv = min(8, C - c0);
  float v[8];
  if (Hs == Hd && Ws == Wd) {
    load_chunk(src + (((size_t)b * Hs + y) * Ws + x) * sp + so + c0, nv, vsrc, v);
  } else {
    const Lerp ly = make_lerp(y, Hs, ratio), lx = make_lerp(x, Ws, ratio);
    const T* p = src + (size_t)b * Hs * Ws * sp + so + c0;
    float v00[8], v01[8], v10[8], v11[8];
    load_chunk(p + ((size_t)ly.i0 * Ws + lx.i0) * sp, nv, vsrc, v00);
    load_chunk(p + ((size_t)ly.i0 * Ws + lx.i1) * sp, nv, vsrc, v01);
    load_chunk(p + ((size_t)ly.i1 * Ws + lx.i0) * sp, nv, vsrc, v10);
    load_chunk(p + ((size_t)ly.i1 * Ws + lx.i1) * sp, nv, vsrc, v11);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ly.l0 * (lx.l0 * v00[k] + lx.l1 * v01[k]) + ly.l1 * (lx.l0 * v10[k] + lx.l1 * v11[k]);
  }
  Td* d = dst + (((size_t)b * Hd + y) * Wd + x) * dp + dof + c0;
  if (acc) {
    float o[8];
    load_chunk(d, nv, vdst, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = o[k] + v[k] * mul;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = v[k] * mul;
  }
  store_chunk(d, nv, vdst, v);
}

// warp() of model/warplayer.py:8-22 on a channel slice; flow = 2 channels (x, y) of another NHWC tensor
template <typename T, typename Tf, typename Td>
__global__ void __launch_bounds__(256) warp_nhwc_kernel(const T* __restrict__ src, int sp, int so,
                                                        const Tf* __restrict__ flow, int fp, int fo,
                                                        Td* __restrict__ dst, int dp, int dof, int B, int C, int H,
                                                        int W, const float* __restrict__ gridx,
                                                        const float* __restrict__ gridy, float halfw, float halfh,
                                                        int vsrc, int vdst) {
  const int nch = (C + 7) >> 3;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W * nch) return;
  const int ch = (int)(i % nch);
  const size_t pix = i / nch;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
  const int c0 = ch * 8, nv = min(8, C - c0);
  const Tf* f = flow + pix * fp + fo;
  const float nx = gridx[x] + __fdiv_rn(to_f(f[0]), halfw), ny = gridy[y] + __fdiv_rn(to_f(f[1]), halfh);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.f), 0.5f), (float)(H - 1));
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wnw = (x0f + 1.f - ix) * (y0f + 1.f - iy), wne = (ix - x0f) * (y0f + 1.f - iy);
  const float wsw = (x0f + 1.f - ix) * (iy - y0f), wse = (ix - x0f) * (iy - y0f);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const bool xin = x0 + 1 < W, yin = y0 + 1 < H;
  const T* p = src + (((size_t)b * H + y0) * W + x0) * sp + so + c0;
  float r[8], t[8];
  load_chunk(p, nv, vsrc, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = t[k] * wnw;
  if (xin) {
    load_chunk(p + sp, nv, vsrc, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] += t[k] * wne;
  }
  if (yin) {
    load_chunk(p + (size_t)W * sp, nv, vsrc, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] += t[k] * wsw;
    if (xin) {
      load_chunk(p + (size_t)W * sp + sp, nv, vsrc, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] += t[k] * wse;
    }
  }
  store_chunk(dst + pix * dp + dof + c0, nv, vdst, r);
}

// space-to-depth of a channel slice: dst [B][H/2][W/2][4*cpad], channel = (py*2+px)*cpad + c (pad channels zero);
// cpad is a multiple of 8, so every destination chunk is one aligned vector
template <typename T, typename Td>
__global__ void __launch_bounds__(256) s2d_nhwc_kernel(const T* __restrict__ src, int sp, int so, int C,
                                                       Td* __restrict__ dst, int cpad, int B, int H, int W, int vsrc) {
  const int nch = cpad >> 3;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W * nch) return;
  const int ch = (int)(i % nch);
  const size_t pix = i / nch;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
  const int c0 = ch * 8, nv = max(0, min(8, C - c0));
  float v[8];
  if (nv == 8 && vsrc) {
    ld8(src + pix * sp + so + c0, v);
  } else {
    const T* p = src + pix * sp + so + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = k < nv ? to_f(p[k]) : 0.f;
  }
  st8(dst + ((((size_t)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * cpad + c0, v);
}

// constant plane(s): dst[..., off : off + C] = value  (the timestep plane of model/IFNet_m.py:64)
template <typename Td>
__global__ void __launch_bounds__(256) fill_nhwc_kernel(Td* __restrict__ dst, int dp, int dof, size_t npix, int C, float value) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  dst[(i / C) * dp + dof + (int)(i % C)] = from_f<Td>(value);
}

// merged = w0*sigmoid(mask) + w1*(1-sigmoid(mask))  (model/IFNet.py:98-99); with `refine` (pre-sigmoid Unet output):
// clamp(merged + sigmoid(refine)*2 - 1, 0, 1)  (model/refine.py:82, model/IFNet.py:106-107).  Output NCHW.
template <typename Tio, typename T, typename Tr>
__global__ void __launch_bounds__(256) eccv_merge_kernel(const T* __restrict__ w0, int p0, int o0,
                                                         const T* __restrict__ w1, int p1, int o1,
                                                         const T* __restrict__ mask, int pm, int om,
                                                         const Tr* __restrict__ refine, int pr, int orf,
                                                         Tio* __restrict__ out, int B, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)H * W;
  if (i >= (size_t)B * plane) return;
  const size_t b = i / plane, pix = i % plane;
  const float m = 1.f / (1.f + expf(-to_f(mask[i * pm + om])));
  for (int c = 0; c < 3; ++c) {
    float v = to_f(w0[i * p0 + o0 + c]) * m + to_f(w1[i * p1 + o1 + c]) * (1.f - m);
    if (refine) {
      const float s = 1.f / (1.f + expf(-to_f(refine[i * pr + orf + c])));
      v = fminf(fmaxf(v + (s * 2.f - 1.f), 0.f), 1.f);
    }
    out[(b * 3 + c) * plane + pix] = from_f<Tio>(v);
  }
}

}  // namespace

// a slice can be moved in 8-channel vectors when base, pitch, offset and width are all 8-element aligned
inline int vec_ok(const void* base, size_t esz, int pitch, int off, int C) {
  return (pitch % 8 == 0 && off % 8 == 0 && C % 8 == 0 && ((uintptr_t)base % (8 * esz)) == 0) ? 1 : 0;
}

// dtype code -> element type tag (generic lambdas below pick the kernel instantiation)
template <typename F>
int with_type(int dt, const char* what, F&& f) {
  if (dt == RIFE_F32) return f(float{});
  if (dt == RIFE_F16) return f(__half{});
  RIFE_FAIL(RIFE_ERR_DTYPE, std::string(what) + ": tensor dtype must be RIFE_F32 or RIFE_F16");
}

int launch_nchw_to_nhwc_slice(const void* src, int io_dtype, void* dst, int ddt, int B, int C, int H, int W, int pitch,
                              int off, cudaStream_t st) {
  const size_t n = (size_t)B * C * H * W;
  return with_type(io_dtype, "nchw_to_nhwc_slice", [&](auto a) -> int {
    return with_type(ddt, "nchw_to_nhwc_slice", [&](auto d) -> int {
      using Ta = decltype(a);
      using Td = decltype(d);
      nchw_to_slice_kernel<Ta, Td><<<nblk(n, 256), 256, 0, st>>>((const Ta*)src, (Td*)dst, B, C, H, W, pitch, off);
      RIFE_LAUNCH_CHECK();
      return (int)RIFE_OK;
    });
  });
}

int launch_nhwc_slice_to_nchw(const void* src, int sdt, int pitch, int off, void* dst, int io_dtype, int B, int C, int H,
                              int W, int sigmoid, cudaStream_t st) {
  const size_t n = (size_t)B * C * H * W;
  return with_type(io_dtype, "nhwc_slice_to_nchw", [&](auto a) -> int {
    return with_type(sdt, "nhwc_slice_to_nchw", [&](auto s) -> int {
      using Ta = decltype(a);
      using Ts = decltype(s);
      slice_to_nchw_kernel<Ta, Ts><<<nblk(n, 256), 256, 0, st>>>((const Ts*)src, pitch, off, (Ta*)dst, B, C, H, W, sigmoid);
      RIFE_LAUNCH_CHECK();
      return (int)RIFE_OK;
    });
  });
}

int launch_resize_nhwc(const void* src, int sdt, int spitch, int soff, void* dst, int ddt, int dpitch, int doff, int B,
                       int C, int Hs, int Ws, int Hd, int Wd, float mul, int accumulate, cudaStream_t st) {
  const size_t n = (size_t)B * Hd * Wd * ((C + 7) / 8);
  const float ratio = (float)Hs / (float)Hd;  // power-of-two factors: exact
  return with_type(sdt, "nhwc_resize", [&](auto s) -> int {
    return with_type(ddt, "nhwc_resize", [&](auto d) -> int {
      using Ts = decltype(s);
      using Td = decltype(d);
      const int vs = vec_ok(src, sizeof(Ts), spitch, soff, C), vd = vec_ok(dst, sizeof(Td), dpitch, doff, C);
      resize_nhwc_kernel<Ts, Td><<<nblk(n, 256), 256, 0, st>>>((const Ts*)src, spitch, soff, (Td*)dst, dpitch, doff, B, C, Hs,
                                                               Ws, Hd, Wd, ratio, mul, accumulate, vs, vd);
      RIFE_LAUNCH_CHECK();
      return (int)RIFE_OK;
    });
  });
}

int launch_warp_nhwc(const void* src, int sdt, int spitch, int soff, const void* flow, int fdt, int fpitch, int foff,
                     void* dst, int ddt, int dpitch, int doff, int B, int C, int H, int W, const float* gridx,
                     const float* gridy, cudaStream_t st) {
  const size_t n = (size_t)B * H * W * ((C + 7) / 8);
  const float hw = (float)((W - 1.0) / 2.0), hh = (float)((H - 1.0) / 2.0);
  return with_type(sdt, "nhwc_warp", [&](auto s) -> int {
    return with_type(fdt, "nhwc_warp", [&](auto f) -> int {
      return with_type(ddt, "nhwc_warp", [&](auto d) -> int {
        using Ts = decltype(s);
        using Tf = decltype(f);
        using Td = decltype(d);
        const int vs = vec_ok(src, sizeof(Ts), spitch, soff, C), vd = vec_ok(dst, sizeof(Td), dpitch, doff, C);
        warp_nhwc_kernel<Ts, Tf, Td><<<nblk(n, 256), 256, 0, st>>>((const Ts*)src, spitch, soff, (const Tf*)flow, fpitch, foff,
                                                                   (Td*)dst, dpitch, doff, B, C, H, W, gridx, gridy, hw, hh,
                                                                   vs, vd);
        RIFE_LAUNCH_CHECK();
        return (int)RIFE_OK;
      });
    });
  });
}

int launch_s2d_nhwc(const void* src, int sdt, int spitch, int soff, int C, void* dst, int ddt, int cpad, int B, int H,
                    int W, cudaStream_t st) {
  if (cpad % 8 || ((uintptr_t)dst & 31)) RIFE_FAIL(RIFE_ERR_SHAPE, "nhwc_s2d: cpad must be a multiple of 8 and dst 32-byte aligned");
  const size_t n = (size_t)B * H * W * (cpad / 8);
  return with_type(sdt, "nhwc_s2d", [&](auto s) -> int {
    return with_type(ddt, "nhwc_s2d", [&](auto d) -> int {
      using Ts = decltype(s);
      using Td = decltype(d);
      s2d_nhwc_kernel<Ts, Td><<<nblk(n, 256), 256, 0, st>>>((const Ts*)src, spitch, soff, C, (Td*)dst, cpad, B, H, W,
                                                            vec_ok(src, sizeof(Ts), spitch, soff, 8));
      RIFE_LAUNCH_CHECK();
      return (int)RIFE_OK;
    });
  });
}

int launch_fill_nhwc(void* dst, int ddt, int dpitch, int doff, int B, int C, int H, int W, float value, cudaStream_t st) {
  const size_t npix = (size_t)B * H * W;
  return with_type(ddt, "nhwc_fill", [&](auto d) -> int {
    using Td = decltype(d);
    fill_nhwc_kernel<Td><<<nblk(npix * C, 256), 256, 0, st>>>((Td*)dst, dpitch, doff, npix, C, value);
    RIFE_LAUNCH_CHECK();
    return (int)RIFE_OK;
  });
}

int launch_eccv_merge(const void* w0, int p0, int o0, const void* w1, int p1, int o1, const void* mask, int pm, int om,
                      int wdt, const void* refine, int rdt, int pr, int orf, void* out, int io_dtype, int B, int H, int W,
                      cudaStream_t st) {
  const size_t n = (size_t)B * H * W;
  return with_type(io_dtype, "eccv_merge", [&](auto a) -> int {
    return with_type(wdt, "eccv_merge", [&](auto w) -> int {
      return with_type(refine ? rdt : (int)RIFE_F32, "eccv_merge", [&](auto r) -> int {
        using Ta = decltype(a);
        using Tw = decltype(w);
        using Tr = decltype(r);
        eccv_merge_kernel<Ta, Tw, Tr><<<nblk(n, 256), 256, 0, st>>>((const Tw*)w0, p0, o0, (const Tw*)w1, p1, o1,
                                                                    (const Tw*)mask, pm, om, (const Tr*)refine, pr, orf,
                                                                    (Ta*)out, B, H, W);
        RIFE_LAUNCH_CHECK();
        return (int)RIFE_OK;
      });
    });
  });
}

}  // namespace rife
