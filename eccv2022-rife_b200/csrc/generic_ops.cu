// Generic fp32 NHWC operators used by the ECCV-paper network path (model/IFNet.py, model/refine.py of the reference):
// every tensor is [B][H][W][pitch] and an operator reads / writes a channel slice [off, off + C), so that torch.cat
// (model/IFNet.py:75-79, model/refine.py:73-80) costs nothing: producers write straight into the consumer's slice.
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

template <typename Tio>
__global__ void __launch_bounds__(256) nchw_to_slice_kernel(const Tio* __restrict__ src, float* __restrict__ dst, int B,
                                                            int C, int H, int W, int pitch, int off) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int c = (int)((i / ((size_t)W * H)) % C), b = (int)(i / ((size_t)W * H * C));
  dst[(((size_t)b * H + y) * W + x) * pitch + off + c] = to_f(src[i]);
}

template <typename Tio>
__global__ void __launch_bounds__(256) slice_to_nchw_kernel(const float* __restrict__ src, int pitch, int off,
                                                            Tio* __restrict__ dst, int B, int C, int H, int W, int sig) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int c = (int)((i / ((size_t)W * H)) % C), b = (int)(i / ((size_t)W * H * C));
  float v = src[(((size_t)b * H + y) * W + x) * pitch + off + c];
  if (sig) v = 1.f / (1.f + expf(-v));
  dst[i] = from_f<Tio>(v);
}

// F.interpolate(bilinear, align_corners=False) of a channel slice; dst = (acc ? dst : 0) + value * mul
__global__ void __launch_bounds__(256) resize_nhwc_kernel(const float* __restrict__ src, int sp, int so,
                                                          float* __restrict__ dst, int dp, int dof, int B, int C, int Hs,
                                                          int Ws, int Hd, int Wd, float ratio, float mul, int acc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * Hd * Wd * C) return;
  const int c = (int)(i % C), x = (int)((i / C) % Wd), y = (int)((i / ((size_t)C * Wd)) % Hd);
  const int b = (int)(i / ((size_t)C * Wd * Hd));
  float v;
  if (Hs == Hd && Ws == Wd) {
    v = src[(((size_t)b * Hs + y) * Ws + x) * sp + so + c];
  } else {
    const Lerp ly = make_lerp(y, Hs, ratio), lx = make_lerp(x, Ws, ratio);
    const float* p = src + (size_t)b * Hs * Ws * sp + so + c;
    const float v00 = p[((size_t)ly.i0 * Ws + lx.i0) * sp], v01 = p[((size_t)ly.i0 * Ws + lx.i1) * sp];
    const float v10 = p[((size_t)ly.i1 * Ws + lx.i0) * sp], v11 = p[((size_t)ly.i1 * Ws + lx.i1) * sp];
    v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
  }
  float* d = dst + (((size_t)b * Hd + y) * Wd + x) * dp + dof + c;
  v = v * mul;
  *d = acc ? (*d + v) : v;
}

// warp() of model/warplayer.py:8-22 on a channel slice; flow = 2 channels (x, y) of another NHWC tensor
__global__ void __launch_bounds__(256) warp_nhwc_kernel(const float* __restrict__ src, int sp, int so,
                                                        const float* __restrict__ flow, int fp, int fo,
                                                        float* __restrict__ dst, int dp, int dof, int B, int C, int H,
                                                        int W, const float* __restrict__ gridx,
                                                        const float* __restrict__ gridy, float halfw, float halfh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((size_t)W * H));
  const float* f = flow + i * fp + fo;
  const float nx = gridx[x] + __fdiv_rn(f[0], halfw), ny = gridy[y] + __fdiv_rn(f[1], halfh);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.f), 0.5f), (float)(H - 1));
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wnw = (x0f + 1.f - ix) * (y0f + 1.f - iy), wne = (ix - x0f) * (y0f + 1.f - iy);
  const float wsw = (x0f + 1.f - ix) * (iy - y0f), wse = (ix - x0f) * (iy - y0f);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const bool xin = x0 + 1 < W, yin = y0 + 1 < H;
  const float* p = src + (((size_t)b * H + y0) * W + x0) * sp + so;
  float* d = dst + i * dp + dof;
  for (int c = 0; c < C; ++c) {
    float r = p[c] * wnw;
    if (xin) r += p[sp + c] * wne;
    if (yin) {
      r += p[(size_t)W * sp + c] * wsw;
      if (xin) r += p[(size_t)W * sp + sp + c] * wse;
    }
    d[c] = r;
  }
}

// space-to-depth of a channel slice: dst [B][H/2][W/2][4*cpad], channel = (py*2+px)*cpad + c (pad channels zero)
__global__ void __launch_bounds__(256) s2d_nhwc_kernel(const float* __restrict__ src, int sp, int so, int C,
                                                       float* __restrict__ dst, int cpad, int B, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W * cpad) return;
  const int c = (int)(i % cpad), x = (int)((i / cpad) % W), y = (int)((i / ((size_t)cpad * W)) % H);
  const int b = (int)(i / ((size_t)cpad * W * H));
  const float v = c < C ? src[(((size_t)b * H + y) * W + x) * sp + so + c] : 0.f;
  dst[((((size_t)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * cpad + c] = v;
}

// merged = w0*sigmoid(mask) + w1*(1-sigmoid(mask))  (model/IFNet.py:98-99); with `refine` (pre-sigmoid Unet output):
// clamp(merged + sigmoid(refine)*2 - 1, 0, 1)  (model/refine.py:82, model/IFNet.py:106-107).  Output NCHW.
template <typename Tio>
__global__ void __launch_bounds__(256) eccv_merge_kernel(const float* __restrict__ w0, int p0, int o0,
                                                         const float* __restrict__ w1, int p1, int o1,
                                                         const float* __restrict__ mask, int pm, int om,
                                                         const float* __restrict__ refine, int pr, int orf,
                                                         Tio* __restrict__ out, int B, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)H * W;
  if (i >= (size_t)B * plane) return;
  const size_t b = i / plane, pix = i % plane;
  const float m = 1.f / (1.f + expf(-mask[i * pm + om]));
  for (int c = 0; c < 3; ++c) {
    float v = w0[i * p0 + o0 + c] * m + w1[i * p1 + o1 + c] * (1.f - m);
    if (refine) {
      const float s = 1.f / (1.f + expf(-refine[i * pr + orf + c]));
      v = fminf(fmaxf(v + (s * 2.f - 1.f), 0.f), 1.f);
    }
    out[(b * 3 + c) * plane + pix] = from_f<Tio>(v);
  }
}

}  // namespace

int launch_nchw_to_nhwc_slice(const void* src, int io_dtype, float* dst, int B, int C, int H, int W, int pitch, int off,
                              cudaStream_t st) {
  const size_t n = (size_t)B * C * H * W;
  if (io_dtype == RIFE_F32) nchw_to_slice_kernel<float><<<nblk(n, 256), 256, 0, st>>>((const float*)src, dst, B, C, H, W, pitch, off);
  else if (io_dtype == RIFE_F16) nchw_to_slice_kernel<__half><<<nblk(n, 256), 256, 0, st>>>((const __half*)src, dst, B, C, H, W, pitch, off);
  else RIFE_FAIL(RIFE_ERR_DTYPE, "nchw_to_nhwc_slice: bad dtype");
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_nhwc_slice_to_nchw(const float* src, int pitch, int off, void* dst, int io_dtype, int B, int C, int H, int W,
                              int sigmoid, cudaStream_t st) {
  const size_t n = (size_t)B * C * H * W;
  if (io_dtype == RIFE_F32) slice_to_nchw_kernel<float><<<nblk(n, 256), 256, 0, st>>>(src, pitch, off, (float*)dst, B, C, H, W, sigmoid);
  else if (io_dtype == RIFE_F16) slice_to_nchw_kernel<__half><<<nblk(n, 256), 256, 0, st>>>(src, pitch, off, (__half*)dst, B, C, H, W, sigmoid);
  else RIFE_FAIL(RIFE_ERR_DTYPE, "nhwc_slice_to_nchw: bad dtype");
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_resize_nhwc(const float* src, int spitch, int soff, float* dst, int dpitch, int doff, int B, int C, int Hs,
                       int Ws, int Hd, int Wd, float mul, int accumulate, cudaStream_t st) {
  const size_t n = (size_t)B * Hd * Wd * C;
  const float ratio = (float)Hs / (float)Hd;  // power-of-two factors: exact
  resize_nhwc_kernel<<<nblk(n, 256), 256, 0, st>>>(src, spitch, soff, dst, dpitch, doff, B, C, Hs, Ws, Hd, Wd, ratio, mul,
                                                   accumulate);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_warp_nhwc(const float* src, int spitch, int soff, const float* flow, int fpitch, int foff, float* dst,
                     int dpitch, int doff, int B, int C, int H, int W, const float* gridx, const float* gridy,
                     cudaStream_t st) {
  const size_t n = (size_t)B * H * W;
  warp_nhwc_kernel<<<nblk(n, 256), 256, 0, st>>>(src, spitch, soff, flow, fpitch, foff, dst, dpitch, doff, B, C, H, W, gridx,
                                                 gridy, (float)((W - 1.0) / 2.0), (float)((H - 1.0) / 2.0));
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_s2d_nhwc(const float* src, int spitch, int soff, int C, float* dst, int cpad, int B, int H, int W,
                    cudaStream_t st) {
  const size_t n = (size_t)B * H * W * cpad;
  s2d_nhwc_kernel<<<nblk(n, 256), 256, 0, st>>>(src, spitch, soff, C, dst, cpad, B, H, W);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_eccv_merge(const float* w0, int p0, int o0, const float* w1, int p1, int o1, const float* mask, int pm, int om,
                      const float* refine, int pr, int orf, void* out, int io_dtype, int B, int H, int W, cudaStream_t st) {
  const size_t n = (size_t)B * H * W;
  if (io_dtype == RIFE_F32)
    eccv_merge_kernel<float><<<nblk(n, 256), 256, 0, st>>>(w0, p0, o0, w1, p1, o1, mask, pm, om, refine, pr, orf, (float*)out, B, H, W);
  else if (io_dtype == RIFE_F16)
    eccv_merge_kernel<__half><<<nblk(n, 256), 256, 0, st>>>(w0, p0, o0, w1, p1, o1, mask, pm, om, refine, pr, orf, (__half*)out, B, H, W);
  else RIFE_FAIL(RIFE_ERR_DTYPE, "eccv_merge: bad dtype");
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

}  // namespace rife
