// Shared declarations for the RIFE B200 kernels (device helpers, layer descriptors, launch API).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/rife_b200.h"

namespace rife {

// ------------------------------------------------------------------ error plumbing (api.cu)
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;
#define RIFE_CUDA_TRY(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      cudaGetLastError(); /* clear the sticky last-error so later launches are not blamed */   \
      rife::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                        \
      return RIFE_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)
#define RIFE_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != RIFE_OK) return _s; \
  } while (0)
#define RIFE_FAIL(code, msg)  \
  do {                        \
    rife::set_error(msg);     \
    return (code);            \
  } while (0)
// after a kernel launch: count it and pick up launch-configuration errors
#define RIFE_LAUNCH_CHECK()                              \
  do {                                                   \
    rife::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    RIFE_CUDA_TRY(cudaGetLastError());                   \
  } while (0)

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Pipeline kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs may become
// resident and run their prologue (barrier init, TMEM allocation, weight loads, index math) while the previous kernel
// of the stream drains.  Contract: EVERY kernel launched through launch_pdl() executes pdl_wait() before it touches
// memory written (or still being read) by earlier kernels -- the wait returns once the preceding grid has completed
// and flushed, which also makes completion order transitive along the stream.  RIFE_NO_PDL=1 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------ layouts
// Full-resolution "source" tensor per frame: [B][H][W][8]: ch 0..3 = Head feature f, 4..6 = image
// channels 0..2, 7 = image channel 3 (alpha; 0 when C == 3).  One 16 B (fp16) / 32 B (fp32) vector
// per bilinear tap of warp().
constexpr int SRC_C = 8;
// Low-resolution block output ("tmp" of IFBlock.forward before the final upsample):
// [B][H/s][W/s][16]: ch 0..3 flow (un-scaled), 4 mask, 5..12 feat, 13..15 zero.
constexpr int TMP_C = 16;
// Block input after the 1/s resize, channel-padded 28 -> 32 and stored space-to-depth (s2d) so that
// the stride-2 conv0.0 becomes a stride-1 2x2-tap conv: [B][H/2s][W/2s][4 * 32], ch = (py*2+px)*32 + c.
constexpr int XIN_C = 32;

// ------------------------------------------------------------------ vector load/store helpers
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  float4 a = reinterpret_cast<const float4*>(p)[0];
  float4 b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
// 16-bit activation storage that is fp16 or bf16 by a (warp-uniform) run-time flag: RIFE_PRECISION_BF16 keeps the conv
// activations inside the Head / IFBlocks in bf16 (same range as fp32), everything else of the fp16 pipeline in fp16
__device__ __forceinline__ uint32_t pack2x(float a, float b, bool bf) {
  if (bf) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2x(uint32_t u, bool bf) {
  if (bf) return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ uint4 pack8x(const float (&v)[8], bool bf) {
  uint4 u;
  u.x = pack2x(v[0], v[1], bf); u.y = pack2x(v[2], v[3], bf); u.z = pack2x(v[4], v[5], bf); u.w = pack2x(v[6], v[7], bf);
  return u;
}
__device__ __forceinline__ void unpack8x(const uint4& u, bool bf, float (&v)[8]) {
  float2 f;
  f = unpack2x(u.x, bf); v[0] = f.x; v[1] = f.y;
  f = unpack2x(u.y, bf); v[2] = f.x; v[3] = f.y;
  f = unpack2x(u.z, bf); v[4] = f.x; v[5] = f.y;
  f = unpack2x(u.w, bf); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void st8x(__half* p, const float (&v)[8], bool bf) { *reinterpret_cast<uint4*>(p) = pack8x(v, bf); }
// host: fp32 -> the 16-bit pattern of the layer's operand type (round to nearest even), carried in a __half slot
inline __half host_cvt16(float x, bool bf) {
  if (!bf) return __float2half_rn(x);
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);   // RNE (NaN / Inf pass through)
  __half_raw r;
  r.x = (unsigned short)(u >> 16);
  return __half(r);
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const __half* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__half* p, const float (&v)[4]) {
  uint2 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
  h[0] = __floats2half2_rn(v[0], v[1]);
  h[1] = __floats2half2_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
// uint8 frames: the callers' `frame / 255.` (inference_img.py:59-60, inference_video.py:191-193)
__device__ __forceinline__ float to_f(uint8_t x) { return __fdiv_rn((float)x, 255.f); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
// `(x * 255).byte()` of the callers (inference_img.py:111, inference_video.py:267): truncation; clamped to [0, 255]
template <> __device__ __forceinline__ uint8_t from_f<uint8_t>(float x) {
  return (uint8_t)fminf(fmaxf(__fmul_rn(x, 255.f), 0.f), 255.f);
}

__device__ __forceinline__ float lrelu02(float x) { return x >= 0.f ? x : x * 0.2f; }

// ------------------------------------------------------------------ conv layer descriptor
enum ConvKind : int {
  CONV_3x3 = 0,   // taps (dy,dx) in {-1,0,1}^2, tap index = (dy+1)*3 + (dx+1)
  CONV_S2D = 1,   // stride-2 3x3 conv expressed on the space-to-depth input: taps in {-1,0}^2, index (dy+1)*2+(dx+1)
};
enum ConvEpi : int {
  EPI_LRELU = 0,      // y = lrelu(acc + bias)                         -> NHWC [.., N]
  EPI_LRELU_S2D = 1,  // same, stored space-to-depth for a following stride-2 conv
  EPI_RES = 2,        // y = lrelu((acc + bias) * beta + res)           (ResConv, IFNet_HDv3.py:53-54)
  EPI_PS = 3,         // ConvTranspose(4,2,1)+PixelShuffle(2) scatter into the TMP_C-channel tmp tensor
  EPI_HEADT = 4,      // Head.cnn3 ConvTranspose scatter into channels 0..3 of the SRC tensor
  EPI_LINEAR = 5,     // y = acc + bias (no activation)                 -> NHWC [.., N]
  EPI_CONVT = 6,      // generic ConvTranspose(4,2,1) parity scatter to NHWC [B][2H][2W][N/4] (op-level entry)
  // ECCV-paper network (model/IFNet.py, model/refine.py): per-channel PReLU tails (slopes in `beta`)
  EPI_PRELU = 7,        // y = prelu(acc + bias)
  EPI_PRELU_ADD = 8,    // y = prelu(acc + bias) + res        (convblock(x) + x, model/IFNet.py:46)
  EPI_CONVT_PRELU = 9,  // ConvTranspose parity scatter + prelu   (deconv(), model/refine.py:19-23)
  EPI_PS8 = 10,         // EPI_PS of the network's last block: flow (4) + mask (1) only, 8-channel tmp pixel (halo kernel)
};

// Device-side view of one packed conv layer (fp32 FMA path).
struct ConvLayerF32 {
  const float* w = nullptr;     // [taps][Cin][Npad]
  const float* bias = nullptr;  // [Npad] (ConvT: bias replicated per parity)
  const float* beta = nullptr;  // [Npad] or null
  int kind = CONV_3x3;
  int cin = 0;    // padded input channels as stored in the activation tensor
  int n = 0;      // valid output columns of the implicit GEMM
  int npad = 0;   // multiple of 4
};

struct ConvIO {
  const void* in = nullptr;   // [B][H][W][cin]
  const void* res = nullptr;  // [B][H][W][n] (EPI_RES)
  void* out = nullptr;
  int B = 0, H = 0, W = 0;    // spatial size of the conv *input* (= GEMM rows grid)
  int epi = EPI_LRELU;
  int out_c = 0;              // channel count of the destination pixel (NHWC pitch)
  int co_per_par = 0;         // EPI_PS / EPI_CONVT / EPI_HEADT: output channels per parity
  int in_pitch = 0;           // channels per pixel of the tensor `in` points into (0 = cin: dense); fp32 path only
  int res_pitch = 0;          // same for `res` (0 = out_c)
  int out_f32 = 0;            // NHWC tcgen05 kernel, EPI_LINEAR / EPI_CONVT only: the destination tensor is fp32
  int planar_out = 0;         // NHWC tcgen05 kernel, EPI_LRELU only: store planar-8 [B][C/8][H][W][8] (feeds conv_halo.cu)
  unsigned int* tile_ctr = nullptr;  // halo kernel: device counter (zero when the kernel starts) for dynamic tile scheduling
  int in_bf16 = 0;            // tcgen05 kernels: `in` (and `res` of EPI_RES) and the packed weights are bf16 instead of fp16
  int out_bf16 = 0;           // tcgen05 kernels: planar-8 / NHWC activation outputs are stored as bf16 (tmp / SRC stay fp16)
};

int launch_conv_f32(const ConvLayerF32& L, const ConvIO& io, cudaStream_t st);

// ------------------------------------------------------------------ pixel kernels (pixel_kernels.cu)
struct TmpLevel {
  const void* ptr;  // [B][h][w][TMP_C]
  int h, w;
  int s;            // up-sampling factor to full resolution (power of two >= 1); negative: the level is stored at
                    // -s times the full resolution (scale > 1) and is DOWN-sampled by -s (centre 2x2 mean)
  int pitch = TMP_C;  // channels per pixel: TMP_C, or 8 for the last block's flow + mask only level (EPI_PS8)
};

// Frame tensors at the API boundary.  RIFE_F32 / RIFE_F16: NCHW (B,C,h,w); RIFE_U8: HWC (B,h,w,C) as video decoders
// deliver them.  (h, w) <= (H, W): the pipeline zero-pads on the right / bottom to (H, W) and crops its outputs, which is
// what the reference's callers do around Model.inference (inference_img.py:62-67,111).
struct ImgLayout {
  long long sb, sc, sy, sx;  // element strides of batch, channel, row, column
  int h, w;
};
inline ImgLayout make_layout(int io_dtype, int C, int h, int w) {
  ImgLayout L;
  L.h = h; L.w = w;
  if (io_dtype == RIFE_U8) { L.sx = C; L.sc = 1; L.sy = (long long)w * C; L.sb = (long long)h * w * C; }
  else { L.sx = 1; L.sy = w; L.sc = (long long)h * w; L.sb = (long long)C * h * w; }
  return L;
}

struct PixelCtx {
  int B, C, H, W;      // B = outputs (virtual pairs) = frame pairs x timesteps per pair
  int F = 0, nf0 = 0;  // SRC frames in total / of them taken from img0 (the rest from img1)
  int src_div = 1;     // output b warps SRC frames b / src_div of src0 and src1
  bool any_up = false; // scale > 1: some block works above full resolution -> generic (non-fast) pixel kernels only
  int h = 0, w = 0;    // frame size seen by the caller (0 = H, W)
  const float* gridx;  // device, W entries (model/warplayer.py:11)
  const float* gridy;  // device, H entries (model/warplayer.py:13)
  const float* tstep;  // device, B entries
  const void* src0;    // [B][H][W][SRC_C]
  const void* src1;
  bool half_act;       // activation / src / tmp storage type: false = float, true = __half
  bool planar_head = false;  // Head input stored planar-8 (conv_halo.cu) instead of pixel-major
  unsigned int* tile_ctrs = nullptr;  // the call's tile-scheduling counters, zeroed by the input-packing kernel
  int n_tile_ctrs = 0;
  // CUDA-graph replay (api.cu): device table {img0, img1, merged} that overrides the frame pointers baked into the
  // captured kernel nodes; null = use the launch arguments
  const void* const* io_table = nullptr;
};

// NCHW image (io dtype) -> SRC tensor channels 4..7 (+ zeros in 0..3) and the s2d Head input [2B][H/2][W/2][16]
int launch_prep(const PixelCtx& c, const void* img0, const void* img1, int io_dtype, void* src_all, void* head_in,
                cudaStream_t st);
// builds the s2d block input of block `level` from the lazily-accumulated tmp levels [0, level)
// `fbuf` ([B][H][W][4] fp32): full-resolution flow summed over levels [0, fbuf_levels); read when fbuf_levels > 0,
// (re)written with the sum over [0, level) when `fbuf_store` (only valid when every pixel is sampled: s <= 2)
int launch_block_input(const PixelCtx& c, int level, const TmpLevel* levels, int s, void* xin, bool planar,
                       float* fbuf, int fbuf_levels, bool fbuf_store, cudaStream_t st);
// scale > 1: F.interpolate(x, scale_factor=up) of the full-resolution block input `xfull` (s2d pixel-major, as
// launch_block_input(s = 1, planar = false) wrote it) into the s2d block input `xin` at (up*H, up*W); flow channels * up
int launch_upsample_xin(const PixelCtx& c, const void* xfull, void* xin, int up, bool planar, cudaStream_t st);
// evaluates flow / mask / warped images / merged at full resolution after `nlev` blocks
struct EvalOut {
  void* flow = nullptr;     // (B,4,H,W)
  void* mask = nullptr;     // (B,1,H,W)
  void* warped0 = nullptr;  // (B,C,H,W)
  void* warped1 = nullptr;
  void* merged = nullptr;   // (B,C,H,W): w0*sigmoid(mask) + w1*(1-sigmoid(mask))
};
int launch_eval(const PixelCtx& c, int nlev, const TmpLevel* levels, const EvalOut& o, int io_dtype, const float* fbuf,
                int fbuf_levels, cudaStream_t st);

int launch_warp_op(const void* input, const void* flow, void* out, int B, int C, int H, int W, int dtype,
                   const float* gridx, const float* gridy, cudaStream_t st);
int launch_resize_op(const float* x, float* out, int B, int C, int H, int W, int num, int den, float post_mul,
                     cudaStream_t st);
int launch_fill_grid(float* g, int n, cudaStream_t st);
// dst[0..n) = host[0..n) through kernel parameters (no copy engine involved)
int launch_set_tstep(float* dst, const float* host, int n, cudaStream_t st);
int launch_set_io_table(const void** tab, const void* img0, const void* img1, const void* out, cudaStream_t st);
// layout converters for the op-level entry points
int launch_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int H, int W, int Cpad, bool s2d, cudaStream_t st);
int launch_nhwc_to_nchw_f32(const float* x, float* y, int B, int C, int H, int W, int Cpad, cudaStream_t st);

// ------------------------------------------------------------------ generic NHWC ops (generic_ops.cu)
// Tensors are NHWC [B][H][W][pitch], each RIFE_F32 or RIFE_F16 (its own *dt argument); an op touches the channel
// slice [off, off + C) (offsets and pitches in elements).
int launch_nchw_to_nhwc_slice(const void* src, int io_dtype, void* dst, int ddt, int B, int C, int H, int W, int pitch,
                              int off, cudaStream_t st);
int launch_nhwc_slice_to_nchw(const void* src, int sdt, int pitch, int off, void* dst, int io_dtype, int B, int C, int H,
                              int W, int sigmoid, cudaStream_t st);
int launch_resize_nhwc(const void* src, int sdt, int spitch, int soff, void* dst, int ddt, int dpitch, int doff, int B,
                       int C, int Hs, int Ws, int Hd, int Wd, float mul, int accumulate, cudaStream_t st);
int launch_warp_nhwc(const void* src, int sdt, int spitch, int soff, const void* flow, int fdt, int fpitch, int foff,
                     void* dst, int ddt, int dpitch, int doff, int B, int C, int H, int W, const float* gridx,
                     const float* gridy, cudaStream_t st);
int launch_s2d_nhwc(const void* src, int sdt, int spitch, int soff, int C, void* dst, int ddt, int cpad, int B, int H,
                    int W, cudaStream_t st);
// scene_ssim.cu: ssim_matlab of the 32x32 thumbnails of B frame pairs -> out[B] (device)
int launch_scene_ssim(const void* img0, const void* img1, int io_dtype, int B, int C, int h, int w, int Hp, int Wp,
                      float* out, cudaStream_t st);
int launch_fill_nhwc(void* dst, int ddt, int dpitch, int doff, int B, int C, int H, int W, float value, cudaStream_t st);
int launch_eccv_merge(const void* w0, int p0, int o0, const void* w1, int p1, int o1, const void* mask, int pm, int om,
                      int wdt, const void* refine, int rdt, int pr, int orf, void* out, int io_dtype, int B, int H, int W,
                      cudaStream_t st);

}  // namespace rife
