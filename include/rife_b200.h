/*
 * rife_b200.h -- C ABI of the B200-native RIFE HDv3 (v4.25) IFNet inference hot path.
 *
 * One shared library (librife_b200.so), plain pointers and sizes, no torch / C++ types.
 * Every entry point is asynchronous on the given CUDA stream (no device sync inside)
 * unless stated otherwise, returns 0 on success or a negative rife_status, and leaves a
 * human-readable message retrievable with rife_last_error() (thread-local).
 * There is NO CPU fallback: host pointers where device pointers are expected are an error.
 *
 * The reference (Dk0071942/ECCV2022-RIFE) is pure Python on PyTorch and has no native
 * interface; each function below names the reference call site (file:line, relative to
 * the reference repo) whose work it replaces.  A maintainer of the reference binds these
 * with ctypes.
 *
 * Tensor conventions at the boundary: NCHW contiguous, float32 or float16 (rife_dtype),
 * device pointers (tensor.data_ptr()), stream = torch.cuda.current_stream().cuda_stream.
 */
#ifndef RIFE_B200_H
#define RIFE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIFE_B200_ABI_VERSION 1

typedef enum rife_status {
  RIFE_OK = 0,
  RIFE_ERR_INVALID_ARG = -1,   /* null pointer, bad enum, bad handle                     */
  RIFE_ERR_SHAPE = -2,         /* H or W not a multiple of 64/scale, unsupported C/scale */
  RIFE_ERR_DTYPE = -3,         /* unsupported dtype / precision combination              */
  RIFE_ERR_CUDA = -4,          /* a CUDA runtime / driver call failed                    */
  RIFE_ERR_WEIGHTS = -5,       /* missing / mis-shaped tensor in the state dict          */
  RIFE_ERR_WORKSPACE = -6,     /* workspace too small or misaligned                      */
  RIFE_ERR_NOT_DEVICE_PTR = -7 /* a pointer that must be device memory is not            */
} rife_status;

/* RIFE_F32 / RIFE_F16: NCHW float frames in [0,1].  RIFE_U8: HWC uint8 frames (B,h,w,C) as cv2 / video decoders deliver
 * them; the library applies the callers' `/255.` on input and `(x*255).byte()` on output (inference_video.py:191,267). */
typedef enum rife_dtype { RIFE_F32 = 0, RIFE_F16 = 1, RIFE_U8 = 2 } rife_dtype;

/* Arithmetic mode of the whole pipeline.
 *  RIFE_PRECISION_FP32: fp32 storage, fp32 FMA convs        (<= 1e-4 max-abs vs reference fp32)
 *  RIFE_PRECISION_FP16: fp16 NHWC activations, tcgen05 tensor-core convs with fp32
 *                       accumulation in TMEM, fp32 coordinates (<= 2e-2 max-abs, >= 45 dB PSNR)
 *  RIFE_PRECISION_BF16: as FP16, but the conv activations inside the Head and the IFBlocks (the only tensors without
 *                       an a-priori bound: 8 ResConv residual accumulations per block) are bf16 operands of bf16 x bf16
 *                       tcgen05 MMAs -- fp32 range, no overflow whatever the weights; frames, features, flow / mask
 *                       levels and warp coordinates keep their fp16 / fp32 storage.  Same stated bar as FP16.  (The
 *                       reference has only the --fp16 switch, inference_video.py:85-86, and that mode overflows to NaN on
 *                       the 256x448 ECCV fixture, tests/golden/eccv_ref_fp16.json.) */
typedef enum rife_precision { RIFE_PRECISION_FP32 = 0, RIFE_PRECISION_FP16 = 1, RIFE_PRECISION_BF16 = 2 } rife_precision;

typedef struct rife_weights rife_weights; /* immutable packed weights on one device          */
typedef struct rife_plan rife_plan;       /* shape-specialised launch plan bound to a workspace */

/* ---- library ------------------------------------------------------------------------- */
int rife_abi_version(void);
const char* rife_last_error(void);
/* Number of kernels this library has launched on this process since load (all threads). */
uint64_t rife_launch_count(void);

/* ---- weights: replaces Model.load_model's load_state_dict (train_log/RIFE_HDv3.py:36-50) */
/* `names[i]` are state-dict keys WITHOUT the "module." prefix (e.g. "block0.conv0.0.0.weight",
 * "encode.cnn3.bias"); `data[i]` are HOST float32 pointers in the reference's own layouts
 * (Conv2d OIHW, ConvTranspose2d IOHW, beta (1,c,1,1)); `numel[i]` element counts.  All 158
 * tensors of SURVEY Appendix A must be present.  Repacks into kernel layouts on the current
 * device (synchronous). */
int rife_weights_create(rife_weights** out, int n, const char* const* names,
                        const float* const* data, const int64_t* numel, int precision);
int rife_weights_destroy(rife_weights* w);

/* ---- plan ------------------------------------------------------------------------------ */
/* Bytes of device workspace a plan for this problem needs (0 on error, see rife_last_error). */
size_t rife_plan_workspace_bytes(int B, int C, int H, int W, float scale, int precision);
/* B frames-pairs per call, C image channels (3, or 4 for RGBA), H and W multiples of 64/scale,
 * scale in {0.25, 0.5, 1, 2, 4} (scale_list = [16,8,4,2,1]/scale, train_log/RIFE_HDv3.py:58; inference_video.py:92).
 * `workspace` is caller-owned device memory (256-byte aligned) that must outlive the plan.
 * `gridx_host` (W floats) / `gridy_host` (H floats) are the warp base grids
 * torch.linspace(-1,1,W) / (-1,1,H) of model/warplayer.py:11-14 computed by the caller on the
 * host (so that they are bit-identical to the reference's); NULL = computed by the library.
 * Synchronous (uploads small tables, encodes TMA descriptors). */
int rife_plan_create(rife_plan** out, const rife_weights* w, int B, int C, int H, int W,
                     float scale, void* workspace, size_t workspace_bytes,
                     const float* gridx_host, const float* gridy_host);
/* Caller-side batching with Head-feature reuse (train_log/IFNet_HDv3.py:128-129 recomputes
 * Head(frame) on every call): `pairs` frame pairs x `timesteps` timesteps per pair = pairs*timesteps outputs per
 * call, pair-major (out[p*timesteps + t], timestep_host likewise) -- the arbitrary-timestep 4X / 8X case of
 * benchmark/HD_multi_4X.py:47 computes each pair's features once.  sequence != 0: the pairs are the `pairs`
 * consecutive pairs of ONE run of pairs+1 frames passed in img0 (img1 = NULL), as inference_video.py:215-260 walks a
 * video: Head runs once per frame instead of twice.  rife_plan_create == (B, 1, 0). */
size_t rife_plan_workspace_bytes_ex(int pairs, int timesteps, int sequence, int C, int H, int W, float scale,
                                    int precision);
int rife_plan_create_ex(rife_plan** out, const rife_weights* w, int pairs, int timesteps, int sequence, int C,
                        int H, int W, float scale, void* workspace, size_t workspace_bytes,
                        const float* gridx_host, const float* gridy_host);
/* IFNet.forward with an arbitrary scale_list (train_log/IFNet_HDv3.py:119,139-148 accepts any five numbers): entry i is
 * the down-scaling factor of block i, here a power of two in [1/4, 64]; H and W must be multiples of 4 * entry for every
 * block.  Lists other than the halving pyramid [16,8,4,2,1]/scale run on the generic (level-by-level) pixel kernels. */
size_t rife_plan_workspace_bytes_list(int pairs, int timesteps, int sequence, int C, int H, int W,
                                      const float* scale_list5, int precision);
int rife_plan_create_list(rife_plan** out, const rife_weights* w, int pairs, int timesteps, int sequence, int C,
                          int H, int W, const float* scale_list5, void* workspace, size_t workspace_bytes,
                          const float* gridx_host, const float* gridy_host);
int rife_plan_destroy(rife_plan* p);
/* Frame size seen by the caller, (h, w) <= the plan's padded (H, W) (default: H, W).  Inputs are then (B,C,h,w) /
 * (B,h,w,C) tensors, zero-padded on the right / bottom inside the first kernel and every output is cropped back to
 * (h, w) -- the F.pad / [:h, :w] the reference's callers wrap around Model.inference (inference_img.py:62-67,111,
 * inference_video.py:199-202,267). */
int rife_plan_set_frame_size(rife_plan* p, int h, int w);
/* Batch scheduling of one call: the blocks working at <= 1/4 resolution (latency-bound, < 7 % of the pixels) always run
 * over the whole batch; the 1/2- and full-resolution stages (HBM / L2-bound) run in passes of `pairs` frame pairs through
 * the same work buffers.  0 = library default (env RIFE_CHUNK_PAIRS, else the whole batch).  Results do not depend on it. */
int rife_plan_set_chunk(rife_plan* p, int pairs);
/* Per-kernel timing: when enabled, every launch of the pipeline is bracketed by CUDA events on the launch
 * stream.  After the stream has been synchronised, rife_plan_get_profile returns (up to `cap`) the stage names
 * (pointers valid until the next run) and durations in ms of the LAST run; return value = number of stages. */
int rife_plan_set_profiling(rife_plan* p, int enable);
int rife_plan_get_profile(rife_plan* p, const char** names, float* ms, int cap);

/* ---- Model.inference (train_log/RIFE_HDv3.py:56-60) ------------------------------------ */
/* img0, img1, out: device, NCHW (B,C,H,W) of `io_dtype`.  timestep_host: B floats (a Python
 * float timestep is broadcast by the caller; a (B,1,1,1) tensor is passed element-wise,
 * train_log/IFNet_HDv3.py:124-127).  out = merged[-1]. */
int rife_inference(rife_plan* p, const void* img0, const void* img1, const float* timestep_host,
                   void* out, int io_dtype, void* stream);

/* ---- IFNet.forward (train_log/IFNet_HDv3.py:119-169), full return structure ------------- */
/* Any pointer may be NULL (not materialised).  flow[i]: (B,4,H,W); mask: (B,1,H,W) pre-sigmoid
 * logits of the last block; warped0/1[i]: (B,C,H,W) = merged[i] tuples; merged: (B,C,H,W). */
typedef struct rife_ifnet_outputs {
  void* flow[5];
  void* mask;
  void* warped0[5];
  void* warped1[5];
  void* merged;
} rife_ifnet_outputs;
int rife_ifnet_forward(rife_plan* p, const void* img0, const void* img1,
                       const float* timestep_host, const rife_ifnet_outputs* outs, int io_dtype,
                       void* stream);

/* ---- end-to-end with HOST buffers (the caller side of inference_img.py:57-60,105-111) ---- */
/* img0_host/img1_host/out_host: host NCHW buffers of io_dtype (pinned memory recommended).
 * Enqueues H2D copies into plan-owned staging, the inference, and the D2H copy of the result
 * on `stream`; returns without synchronising. */
int rife_inference_host(rife_plan* p, const void* img0_host, const void* img1_host,
                        const float* timestep_host, void* out_host, int io_dtype, void* stream);
/* Same, as a three-stage software pipeline for a frame server (inference_video.py:139-143,253-270 keeps a reader
 * thread, the model and a writer thread busy at once): H2D copies on stream_h2d, kernels on stream_compute, the D2H
 * copy on stream_d2h, ordered by plan-owned events (including re-use of this plan's staging buffers by its next
 * call).  Rotate >= 2 plans over ONE compute stream: the kernels of consecutive calls then run back to back and never
 * interleave (which would evict each other's L2-resident activations) while their PCIe copies overlap the compute. */
int rife_inference_host_pipelined(rife_plan* p, const void* img0_host, const void* img1_host,
                                  const float* timestep_host, void* out_host, int io_dtype,
                                  void* stream_h2d, void* stream_compute, void* stream_d2h);

/* ---- single operators (op-level parity tests; drop-in for the reference's helpers) ------ */
/* warp(tenInput, tenFlow), model/warplayer.py:8-22.  input (B,C,H,W), flow (B,2,H,W) in pixels,
 * out (B,C,H,W); all device, NCHW, `dtype`.  gridx/gridy: DEVICE float tables (W / H entries)
 * or NULL (library formula). */
int rife_warp(const void* input, const void* flow, void* out, int B, int C, int H, int W,
              int dtype, const float* gridx_dev, const float* gridy_dev, void* stream);
/* F.interpolate(x, scale_factor=num/den, mode="bilinear", align_corners=False) for power-of-two
 * factors (train_log/IFNet_HDv3.py:79,81,86); out = result * post_mul.  NCHW float32, device. */
int rife_resize_bilinear(const void* x, void* out, int B, int C, int H, int W, int num, int den,
                         float post_mul, void* stream);
/* conv() / ResConv of train_log/IFNet_HDv3.py:9-14,45-54 on NCHW float32 device tensors:
 * y = leaky_relu((conv3x3(x, w, stride, pad=1) + bias) * beta + residual, 0.2)
 * (beta / residual may be NULL => plain conv + bias + LeakyReLU; act=0 skips the LeakyReLU).
 * w: device OIHW float32.  `precision` selects the fp32 FMA kernel or the fp16 tcgen05 kernel.
 * Synchronous (repacks on the fly; test entry point, not the product path). */
int rife_conv3x3(const float* x, const float* w, const float* bias, const float* beta,
                 const float* residual, float* y, int B, int Cin, int H, int W, int Cout,
                 int stride, int act, int precision, void* stream);
/* ConvTranspose2d(Cin,Cout,4,2,1) (+ optional PixelShuffle(2)) of train_log/IFNet_HDv3.py:30,73-76.
 * w: device IOHW float32; y: (B,Cout,2H,2W) or, with pixel_shuffle=1, (B,Cout/4,4H,4W). */
int rife_conv_transpose4x4s2(const float* x, const float* w, const float* bias, float* y, int B,
                             int Cin, int H, int W, int Cout, int pixel_shuffle, int precision,
                             void* stream);

/* ---- scene-change / static-frame detector of the video loop (inference_video.py:226-228,242-243) ----
 * ssim_out[b] = ssim_matlab(F.interpolate(pad(img0[b]), (32,32), bilinear)[:3], same of img1[b])
 * (model/pytorch_msssim/__init__.py:81-134) in one launch.  Frames: (B,C,h,w) NCHW float32/float16 in [0,1] or
 * (B,h,w,C) uint8; they are taken as zero-padded on the right / bottom to multiples of `pad_multiple` first, as
 * inference_video.py:199-203 pads them (max(32, 32/scale)); pad_multiple = 1 means no padding.  ssim_out: device. */
int rife_scene_ssim(const void* img0, const void* img1, int io_dtype, int B, int C, int h, int w,
                    int pad_multiple, float* ssim_out, void* stream);

/* ---- generic NHWC layer / operator API: the ECCV-paper network (model/IFNet.py, model/refine.py, -------------
 *      model/RIFE.py:56-65; BASELINE config 5).  NHWC tensors [B][H][W][pitch]; an operator touches the
 *      channel slice [off, off + C) so that torch.cat (model/IFNet.py:75-79, model/refine.py:73-80) is free. */
typedef struct rife_layer rife_layer;
/* kind 0: nn.Conv2d(cin,cout,3,1,1); kind 1: nn.Conv2d(cin,cout,3,2,1) reading the space-to-depth tensor produced by
 * rife_nhwc_s2d (cpad = cin_store); kind 2: nn.ConvTranspose2d(cin,cout,4,2,1).  w_host in the reference's layout
 * (OIHW / IOHW), bias_host [cout] or NULL, slope_host [cout] = nn.PReLU(cout).weight or NULL.  cin_store = channels
 * of the input slice incl. zero padding: a multiple of 8 with RIFE_PRECISION_FP32 (fp32 tensors, FMA kernel); with
 * RIFE_PRECISION_FP16 (fp16 tensors, tcgen05 kernel) a multiple of 16 for kinds 0 / 2 and of 8 for kind 1, and the
 * layer then writes round_up(cout, 16) channels (kind 2: per parity round_up(cout, 8)), the padding as zeros. */
int rife_layer_create(rife_layer** out, int kind, const float* w_host, const float* bias_host,
                      const float* slope_host, int cin, int cout, int cin_store, int precision);
int rife_layer_destroy(rife_layer* l);
/* act: 0 none, 1 LeakyReLU(0.2), 2 PReLU, 3 PReLU then "+ res" (convblock(x) + x, model/IFNet.py:46).
 * H, W: spatial size of the input grid (for kind 1: of the s2d tensor, i.e. of the output).  in / res are of the
 * layer's precision; out_dtype likewise, except that an fp16 layer with act 0 (the flow / mask / refine heads) may
 * write RIFE_F32 so that its result is not rounded before it is scaled and accumulated. */
int rife_layer_run(const rife_layer* l, const void* in, int in_pitch, void* out, int out_dtype, int out_pitch,
                   int B, int H, int W, const void* res, int res_pitch, int act, void* stream);
/* Every NHWC tensor carries its own element type (RIFE_F32 or RIFE_F16; the *_dtype argument after its pointer):
 * with RIFE_PRECISION_FP16 layers the conv activations are fp16 while the full-resolution flow / mask / warped-frame
 * tensor stays fp32.  Pitches and offsets are in elements. */
int rife_nhwc_from_nchw(const void* src, int io_dtype, void* dst, int dst_dtype, int B, int C, int H, int W,
                        int pitch, int off, void* stream);
int rife_nhwc_to_nchw(const void* src, int src_dtype, int pitch, int off, void* dst, int io_dtype, int B, int C,
                      int H, int W, int sigmoid, void* stream);
/* F.interpolate(bilinear, align_corners=False) by a power-of-two factor; dst = (accumulate ? dst : 0) + value * mul */
int rife_nhwc_resize(const void* src, int src_dtype, int spitch, int soff, void* dst, int dst_dtype, int dpitch,
                     int doff, int B, int C, int Hs, int Ws, int Hd, int Wd, float mul, int accumulate,
                     void* stream);
/* warp() of model/warplayer.py:8-22; flow = channels [foff, foff+2) of an NHWC tensor at the same resolution */
int rife_nhwc_warp(const void* src, int src_dtype, int spitch, int soff, const void* flow, int flow_dtype,
                   int fpitch, int foff, void* dst, int dst_dtype, int dpitch, int doff, int B, int C, int H, int W,
                   const float* gridx_dev, const float* gridy_dev, void* stream);
int rife_nhwc_s2d(const void* src, int src_dtype, int spitch, int soff, int C, void* dst, int dst_dtype, int cpad,
                  int B, int H, int W, void* stream);
/* dst[..., off : off + C] = value: the constant timestep plane of model/IFNet_m.py:64 */
int rife_nhwc_fill(void* dst, int dst_dtype, int pitch, int off, int B, int C, int H, int W, float value,
                   void* stream);
/* merged = w0*sigmoid(mask) + w1*(1-sigmoid(mask)) (model/IFNet.py:98-99); with refine (pre-sigmoid Unet output):
 * clamp(merged + sigmoid(refine)*2 - 1, 0, 1) (model/IFNet.py:105-107).  w0 / w1 / mask share w_dtype.
 * out: NCHW (B,3,H,W) of io_dtype. */
int rife_eccv_merge(const void* w0, int p0, int o0, const void* w1, int p1, int o1, const void* mask, int pm,
                    int om, int w_dtype, const void* refine, int refine_dtype, int pr, int orf, void* out,
                    int io_dtype, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIFE_B200_H */
