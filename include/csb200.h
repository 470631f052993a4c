/*
 * csb200.h -- C ABI of libcsb200.so, the B200 (sm_100a) replacement for the reference's
 * kernel-JIT operator boundary.
 *
 * What it replaces (paths relative to the reference repo):
 *   utils/cupy_utils.py:7-13   launch_kernel(name, src)(grid=, block=, args=[int32 n, tensor.data_ptr(), ...])
 *                              -- a cupy.RawKernel launch on raw device pointers of contiguous torch tensors.
 *   The five call sites of that boundary are listed next to each entry point below.
 *
 * Conventions
 *   - every entry point returns an int status (CSB_OK == 0) and enqueues all of its work on `stream`
 *     (a cudaStream_t passed as void*; the Python side passes torch.cuda.current_stream().cuda_stream).
 *     No entry point synchronises, allocates or frees device memory, or reads results back to the host:
 *     the caller owns every buffer including scratch, so a whole frame is CUDA-graph capturable.
 *   - all tensors are contiguous fp32 unless stated, layouts as in the reference
 *     (points [B,3,N] channel-planar, images [B,C,H,W]).
 *   - errors: non-zero status, message via csb_last_error() (thread-local).  Kernels never assert/trap
 *     (the reference kernels use device-side assert, anime_3dkenburns/models/utils.py:73-74).
 */
#ifndef CSB200_H
#define CSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSB_API __attribute__((visibility("default")))

#define CSB_OK 0
#define CSB_ERR_INVALID 1   /* bad shape / null pointer / misaligned buffer */
#define CSB_ERR_CUDA 2      /* a CUDA runtime call failed; see csb_last_error() */
#define CSB_ERR_ARCH 3      /* device is not sm_100 */

CSB_API int csb_version(void);
CSB_API const char* csb_last_error(void);
/* Number of kernel launches this library has enqueued since load (process-wide, monotonic). */
CSB_API uint64_t csb_launch_count(void);
/* Per-kernel device timing for bench.py's roofline: between begin and end every launch is followed by a CUDA event on
 * its stream; end() synchronises and writes {"kernel": {"ms": total, "count": launches}, ...} into json. */
CSB_API int csb_profile_begin(void* stream);
CSB_API int csb_profile_end(char* json, size_t cap);

/* ---------------------------------------------------------------------------------------------
 * render_pointcloud -- anime_3dkenburns/models/utils.py:56-315 (kernel_pointrender_updateZee :63-149,
 * kernel_pointrender_updateDegrid :152-212, kernel_pointrender_updateOutput :215-313, host tail :315),
 * with process_shift's tensor part (anime_3dkenburns/common.py:76-81) optionally folded in.
 *
 *   points  [B,3,N]   data [B,C,N]                      (inputs, read-only)
 *   shift   3 host floats (sx,sy,sz) or NULL: when given, every point is first transformed exactly as
 *           common.py:78-81 does:  x = x*(z/(z+1e-7)) + sx, y likewise, z = z + sz
 *   shift_dev  the same 3 floats in DEVICE memory (takes precedence), e.g. written by csb_shift_from_scalars,
 *           so a pipeline never has to read the depth range back to the host
 *   zkey    [B,H,W] int32 scratch       zee [B,H,W] fp32 scratch (post-degrid z-buffer; also an output for tests)
 *   acc     [B,H,W,CP] fp32 scratch, CP = csb_render_acc_channels(C) (channel-interleaved accumulator)
 *   render  [B,C,H,W]  existing [B,1,H,W]               (outputs; either may be NULL to skip)
 * ------------------------------------------------------------------------------------------- */
CSB_API int csb_render_acc_channels(int C);
CSB_API int csb_pointcloud_render(const float* points, const float* data, int B, int N, int C, int H, int W,
                          double focal, double baseline, const float* shift, const float* shift_dev,
                          int32_t* zkey, float* zee, float* acc, float* render, float* existing, void* stream);
/* The three stages on their own (tests pin each against the reference kernel of the same name). */
CSB_API int csb_pointcloud_zpass(const float* points, int B, int N, int H, int W, double focal, double baseline,
                         const float* shift, int32_t* zkey, void* stream);
CSB_API int csb_pointcloud_degrid(const int32_t* zkey, int B, int H, int W, float* zee, void* stream);

/* Autozoom score -- anime_3dkenburns/common.py:116-131: for each of S candidate shifts, render the raw point
 * cloud and count pixels with tenExisting > 0.  Only coverage is needed, so no colour is splatted.
 *   shifts [S,3] host floats;  zkey [S,H,W] int32 scratch;  zee [S,H,W] fp32 scratch; cover [S,H,W] uint8 scratch;
 *   counts [S] int32 device output (stays on device; the caller reads it once). */
CSB_API int csb_autozoom_coverage(const float* points, int N, int H, int W, double focal, double baseline,
                          const float* shifts, int S, int32_t* zkey, float* zee, uint8_t* cover, int32_t* counts,
                          void* stream);

/* Context render of the Inpaint net -- anime_3dkenburns/models/pointcloud_inpainting.py:135-146: render_pointcloud of a C-channel payload given as
 * fp16 channel-interleaved [N, ld] (straight from the conv engine), then tenExisting = (existing > 0) * median-5(existing > 0), render *= tenExisting,
 * written as the NHWC fp16 input of netInput: out16 [H,W,ldo] = {render[0..C), tenExisting, 0...}; existing [H,W] fp32 (0/1).
 * scratch: zkey/zee/acc as csb_pointcloud_render (CP = csb_render_acc_channels(C)), flags [H*W] uint8.  shift: 3 host floats. */
CSB_API int csb_inpaint_context_render(const float* points, const void* data16, int ld, int N, int C, int H, int W, double focal, double baseline,
                               const float* shift, int32_t* zkey, float* zee, float* acc, uint8_t* flags, void* out16, int ldo, float* existing,
                               void* stream);

/* fill_disocclusion -- anime_3dkenburns/common.py:145-247 (kernel_discfill_updateOutput :149-245).
 *   input [B,C,H,W], depth [B,1,H,W] -> output [B,C,H,W] (output may not alias input). */
CSB_API int csb_disocclusion_fill(const float* input, const float* depth, int B, int C, int H, int W, float* output, void* stream);

/* depth_adjustment_animesseg -- anime_3dkenburns/kenburns_effect.py:39-91 (non-median branch), all K instance masks applied in order
 * on the device without host syncs.  disparity [H,W] fp32 in place; masks [K,H,W] uint8 0/1 (torch.bool); state: 16 bytes of device scratch.
 * (The reference's `.sum() == 0` skip and row tests are evaluated as "any plane > 0", identical for non-negative disparities.) */
CSB_API int csb_depth_adjust_instances(float* disparity, const uint8_t* masks, int K, int H, int W, void* state, void* stream);

/* The `use_medium=True` variant of depth_adjustment_animesseg (kenburns_effect.py:80): per instance, in order, every masked pixel with a positive
 * disparity is set to the (lower) median of those disparities -- exact, by radix select on the device.  state: 1040 bytes of scratch. */
CSB_API int csb_depth_adjust_median(float* disparity, const uint8_t* masks, int K, int H, int W, void* state, void* stream);
/* The same for a batch without host synchronisation: disparity [N,H,W] in place, masks [N,Kmax,H,W], num [N] device int32 (instances per image,
 * e.g. straight from csb_rtmdet_select), state: csb_depth_adjust_state_words(N, Kmax, H, W) int32 of device scratch.  N <= number of SMs.
 * When every disparity under a mask is > 0 (what the depth estimators produce) the recurrence is evaluated order-free in four parallel passes
 * (K-bit cover word per pixel, predecessor relation, K scalar max steps, write-back: csrc/kb_adjust.cu); otherwise -- decided on the device -- one
 * cooperative launch walks the instances in order.  Both are bit-identical to the reference's torch formulation. */
CSB_API long long csb_depth_adjust_state_words(int N, int Kmax, int H, int W);
CSB_API int csb_depth_adjust_batch(float* disparity, const uint8_t* masks, const int* num, int N, int Kmax, int H, int W, int32_t* state, void* stream);

/* process_shift scalar part -- anime_3dkenburns/common.py:60-72, evaluated on the device in double precision from the
 * `scalars` written by csb_disparity_to_cloud (closest point = argmin of the depth crop): fltShiftU/V as in the reference,
 * fltDepthFrom = depthmin, fltDepthTo = depthmin * depth_ratio.  shift_dev: 3 device floats. */
CSB_API int csb_shift_from_scalars(const float* scalars, int W, int H, double focal, double shiftU, double shiftV, double depth_ratio,
                           float* shift_dev, void* stream);

/* process_shift tensor part -- anime_3dkenburns/common.py:76-81.  shift = 3 host floats. */
CSB_API int csb_points_shift(const float* points, int B, int N, const float* shift, float* out, void* stream);

/* depth_to_points -- anime_3dkenburns/models/utils.py:43-50.  depth [B,1,H,W] -> points [B,3,H,W]. */
CSB_API int csb_depth_to_points(const float* depth, int B, int H, int W, double focal, float* points, void* stream);

/* spatial_filter -- anime_3dkenburns/models/utils.py:9-40.  kind: 0 'laplacian', 3 'median-3', 5 'median-5'. */
CSB_API int csb_spatial_filter(const float* input, int B, int C, int H, int W, int kind, float* output, void* stream);

/* generate_kenburns_config math -- anime_3dkenburns/kenburns_effect.py:928-937, fused:
 *   raw disparity [H,W] -> disparity, depth, valid [H,W]; points, unaltered [3,H,W];
 *   scalars (device, 8 floats): dispmin, dispmax, depthmin, depthmax, argmin x,y, argmax x,y of depth[128:-128,128:-128]
 *   (cv2.minMaxLoc tie rule: first occurrence in row-major order).  scratch: 64 uint64 (device).
 *   Optional (both or neither): image_hwc [H,W,3] u8 BGR -> data4 [4,H*W] = {B,G,R}*float32(1/255) planar ++ depth, the
 *   render payload the frame loop builds at kenburns_effect.py:921,1036. */
CSB_API int csb_disparity_to_cloud(const float* raw, int H, int W, double focal, double baseline, float* disparity, float* depth,
                           float* valid, float* points, float* unaltered, float* scalars, uint64_t* scratch,
                           const uint8_t* image_hwc, float* data4, void* stream);

/* Frame tail -- anime_3dkenburns/kenburns_effect.py:1040 ((x*255).clip(0,255).astype(uint8), CHW -> HWC)
 * and :1069-1070 (cv2.getRectSubPix centre crop + cv2.resize INTER_LINEAR back to W x H, both on uint8 with
 * OpenCV's fixed-point arithmetic).  render [>=3,H,W] fp32 (first three planes used).
 *   csb_frame_pack_u8:      render -> frame [H,W,3] u8
 *   csb_frame_crop_resize:  frame [H,W,3] u8 -> out [H,W,3] u8 (crop pw x ph around (cx,cy), resize to W x H) */
CSB_API int csb_frame_pack_u8(const float* render, int H, int W, uint8_t* frame, void* stream);
CSB_API int csb_frame_crop_resize(const uint8_t* frame, int H, int W, int pw, int ph, double cx, double cy, uint8_t* out, void* stream);

/* cv2.resize(src, (Wo,Ho), interpolation=INTER_LINEAR) on uint8 HWC images, bit-exact (used for scaledown_maxsize, utils/io_utils.py:254-274). */
CSB_API int csb_resize_u8c3(const uint8_t* src, int H, int W, uint8_t* dst, int Ho, int Wo, void* stream);

/* Input / output glue of the per-image Ken-Burns networks (csrc/kb_netio.cu), replacing the eager per-tensor ops around the reference's Inpaint and
 * Refine nets (anime_3dkenburns/models/pointcloud_inpainting.py:117-131, 190-200; disparity_refinement.py:99-100, 128-135):
 *   csb_tensor_stats     out[3] = {mean, population std (torch.std(unbiased=False)), max} of x[n] (double accumulation); scratch: 3 doubles
 *   csb_pack_norm16      out [HW][16] fp16 = [(a - mean_a) / (std_a + eps) (ca planes of [ca][HW] fp32) | the same for b (cb planes) | zeros];
 *                        stats_a / stats_b = csb_tensor_stats outputs (null: copy without normalisation)
 *   csb_inpaint_payload  payload [HW][72] fp16 = [x16[:, 0:4] | ctx [HW][64] | zeros(4)]: the interleaved payload of csb_inpaint_context_render
 *   csb_net_output       out [C][HW] fp32 = post((a [+ b]) [HW][C] * (std + eps) + mean), post: 0 none, 1 clip to [0, 1], 2 threshold at 0
 *   csb_inpaint_points   valid-masked cloud of the raw frame (:117-120): points [3][HW] from disp [H][W]; stats = csb_tensor_stats(disp) (uses the max) */
CSB_API int csb_tensor_stats(const float* x, long long n, double* scratch, float* out, void* stream);
CSB_API int csb_pack_norm16(const float* a, int ca, const float* stats_a, const float* b, int cb, const float* stats_b, long long HW, float eps, void* out16,
                            void* stream);
CSB_API int csb_inpaint_payload(const void* x16, const void* ctx64, long long HW, void* payload72, void* stream);
CSB_API int csb_net_output(const float* a, const float* b, int C, long long HW, const float* stats, float eps, int post, float* out_nchw, void* stream);
CSB_API int csb_inpaint_points(const float* disp, int H, int W, double focal, double baseline, const float* stats, float* points, void* stream);

/* LeReS post-processing -- depth_modules/leres/__init__.py:117-140 (min/max normalise to 16 bit, cv2.convertScaleAbs to 8 bit,
 * bitwise_not) + kenburns_effect.py:572-577 (cv2.resize back to the frame size, astype(float32)), bit-exact against numpy + OpenCV: INTER_AREA for
 * the upscaling / same-size branch (H >= h and W >= w), INTER_LANCZOS4 (8-tap fixed point, both axes) when h > H as the reference selects it.
 * logits [N,h,w] fp32 -> out [N,H,W] fp32 (8-bit values); minmax: 2*N uint32, q8: N*h*w bytes of scratch. */
CSB_API int csb_leres_depth_tail(const float* logits, int N, int h, int w, int H, int W, unsigned* minmax, uint8_t* q8, float* out, void* stream);

/* `depth[depth == 0] = depth[depth > 0].min()` (anime_3dkenburns/kenburns_effect.py:577, :815), per image, in place: x [N, per] fp32,
 * scratch: N uint32.  Two launches, no host read. */
CSB_API int csb_zero_to_min_positive(float* x, int N, long long per, unsigned* scratch, void* stream);

/* AnimeInstances.resize (animeinsseg/anime_instances.py:268-280): masks [K,H0,W0] bool bytes -> F.interpolate(mode='area') > thr (0.3) ->
 * out [K,H,W] bool bytes, bit-identical to the torch ops; boxes_in/boxes_out (optional, [K,4] int32 xywh): the reference's scaling
 * (columns 0,2 by H/H0, columns 1,3 by W/W0 -- its x/y swap kept) and torch.round. */
CSB_API int csb_masks_area_resize(const uint8_t* masks, int K, int H0, int W0, uint8_t* out, int H, int W, float thr, const int* boxes_in, int* boxes_out,
                                  void* stream);

/* AnimeInstances.compose_masks (anime_instances.py:282-298): logical OR of K masks of P = H*W bool bytes -> out [P]. */
CSB_API int csb_masks_compose(const uint8_t* masks, int K, long long P, uint8_t* out, void* stream);

/* Point-cloud growth after an inpaint pass (kenburns_effect.py:462-512): dst[c] = concat(old[c][0..n_old), src[c][existing == 0]) for up to 8 fp32
 * planes, element order = the boolean-mask gather's (ascending pixel index): an order-preserving stream compaction on the device.
 *   csb_cloud_append_scratch_ints(P)  ints of scratch;   csb_cloud_append_count: phase 1, the hole count lands in *total_out (device int);
 *   csb_cloud_append: phase 2 (dst planes hold n_old + total floats). */
CSB_API long long csb_cloud_append_scratch_ints(long long P);
CSB_API int csb_cloud_append_count(const float* existing, long long P, int* scratch, int* total_out, void* stream);
CSB_API int csb_cloud_append(const float* existing, long long P, const int* scratch, const float* const* src, const float* const* old_planes, float* const* dst,
                             int planes, long long n_old, void* stream);

/* ZoeDepth / MiDaS DPT-BEiT-L encoder pieces (the encoder is torch.hub `intel-isl/MiDaS` `DPT_BEiT_L_384`, loaded at
 * depth_modules/zoedepth/models/base_models/midas.py:341 and NOT vendored in the reference: restated from timm's BEiT + MiDaS v3.1's DPT).
 * Every Linear / Conv runs on csb_conv2d_nhwc; these are the remaining ops (csrc/zoe_attn.cu, csrc/zoe_io.cu).
 *   csb_attention_bias   out[b,q,h*64+:] = softmax_k(Q.K/sqrt(64)*... + bias[h,q,k]) V; qkv [B,T,3*heads*64] fp16 (q|k|v), bias [heads,Tp,Tp] fp16 with
 *                        Tp a multiple of 64 and the columns k >= T set to a large negative number; out [B,T,heads*64] fp16.
 *   csb_tokens_assemble  tokens [B,1+P,C] = cat(cls [C], patches [B,P,C])            (BeitEmbeddings)
 *   csb_readout_concat   out [B,P,2C] = [tokens[b,1+p] | tokens[b,0]]                (DPT 'project' readout before its Linear + GELU)
 *   csb_pixel_shuffle_nhwc  x [B,h,w,k*k*C] -> y [B,h*k,w*k,C]: the tail of ConvTranspose2d(kernel=stride=k) computed as a 1x1 conv
 *   csb_zoe_prep         img [H,W,3] u8 -> patch rows [1|2][Hn/16][Wn/16][768] fp16 (channel (r*16+s)*3+c): /255, reflect pad (pad_h, pad_w), bilinear
 *                        align_corners=True resize to Hn x Wn, (x-0.5)/0.5; with flip_aug a second, horizontally flipped entry
 *                        (depth_model.py:57-112, midas.py:164-186)
 *   csb_zoe_finish       net depth [1|2][Hn][Wn] fp32 -> out [H][W]: bicubic (A=-0.75, align_corners=False) to the padded size, crop, un-flip, mean
 *   csb_zoe_disparity    depth -> disparity (kenburns_effect.py:812-818); scratch: one uint32 */
CSB_API int csb_attention_bias(const void* qkv, int B, int T, int heads, int head_dim, const void* bias, int Tp, float scale, void* out, void* stream);
/*   csb_attention_bias_tc  the same contract on tcgen05 (csrc/zoe_attn_tc.cu: S = Q K^T and O = P V as tcgen05.mma with TMEM accumulators, TMA-staged
 *                        operands, one query row per thread); bias padded to a multiple of 128; vt_scratch: csb_attention_tc_scratch_bytes(B, T, heads)
 *                        bytes of device memory owned by the caller (V transposed to [B][heads][64][keys]). */
CSB_API long long csb_attention_tc_scratch_bytes(int B, int T, int heads);
CSB_API int csb_attention_bias_tc(const void* qkv, int B, int T, int heads, int head_dim, const void* bias, int Tp, float scale, void* vt_scratch, void* out,
                          void* stream);
CSB_API int csb_tokens_assemble(const void* patches, const void* cls, int B, int P, int C, void* tokens, void* stream);
CSB_API int csb_readout_concat(const void* tokens, int B, int P, int C, void* out, void* stream);
CSB_API int csb_pixel_shuffle_nhwc(const void* x, int B, int h, int w, int k, int C, void* y, void* stream);
CSB_API int csb_zoe_prep(const uint8_t* img, int H, int W, int pad_h, int pad_w, int Hn, int Wn, int flip_aug, void* patches, void* stream);
CSB_API int csb_zoe_finish(const float* depth_net, int flip_aug, int Hn, int Wn, int H, int W, int pad_h, int pad_w, float* out, void* stream);
CSB_API int csb_zoe_disparity(const float* depth, long long n, double focal, double baseline, float* disparity, unsigned* scratch, void* stream);

/* Bokeh depth-of-field of the frame loop -- kenburns_effect.py:1042-1067, utils/effects.py:12-84,143-182,
 * depth_modules/zoedepth/utils/misc.py:97-150 -- on the device; the reference does all but the three gathers in numpy on the host.
 *   ws: csb_bokeh_workspace_bytes(H, W, K) bytes of device memory, shared by the three calls of a frame.
 *   csb_depth_colorize_u8:  depth [H,W] fp32 -> colorize(depth, cmap='gray_r')[..., 0] [H,W] u8: exact np.percentile(2) / (85) (linear
 *                           method) -> normalise in fp32 -> matplotlib Colormap index rule -> lut256 (the gray_r byte table, host-built);
 *                           -99 (invalid_val) -> 128.
 *   csb_focal_plane_range:  depth8 + masks [K,H,W] u8 (0/1) -> start_end[2] (device doubles): focalplane_end = max_k np.median(depth8[mask_k]),
 *                           focalplane_start = 255 if |255-end| > |end| else 0; K == 0 -> (0, 255)   (kenburns_effect.py:1045-1059)
 *   csb_bokeh_blur:         frame [H,W,3] u8 + depth8 -> out [H,W,3] u8 = bokeh_blur(frame, depth8, nsamples, lightness, depth_factor,
 *                           use_cuda=True, focal_plane).  highlight_lut[256] = np.power(arange(256, f32)/255, lightness) (device, host-built);
 *                           inv_lightness = float32(1/lightness); focal plane = focal_int*range[1] + (1-focal_int)*range[0] when focal_range
 *                           (device) is given (kenburns_effect.py:1065-1066), else `focal_plane`.  kernel_bokeh's planar-buffer /
 *                           interleaved-index mix-up (effects.py:36-38 on a np2flatten_tensor buffer) is reproduced. */
CSB_API size_t csb_bokeh_workspace_bytes(int H, int W, int K);
CSB_API int csb_depth_colorize_u8(const float* depth, int H, int W, const uint8_t* lut256, uint8_t* out8, void* ws, void* stream);
CSB_API int csb_focal_plane_range(const uint8_t* depth8, const uint8_t* masks, int K, int H, int W, double* start_end, void* ws, void* stream);
CSB_API int csb_bokeh_blur(const uint8_t* frame, const uint8_t* depth8, int H, int W, int nsamples, const float* highlight_lut, float inv_lightness,
                           const double* focal_range, double focal_int, double focal_plane, int depth_factor, uint8_t* out, void* ws,
                           void* stream);

/* Fused Ken-Burns frame: shift + render (C=4: BGR + depth) + normalise + disocclusion fill + u8 pack, then
 * crop+resize -- the body of the reference's frame loop, kenburns_effect.py:1028-1040,1069-1070, as 6 launches.
 *   points [1,3,N], data [1,4,N]; scratch as csb_pointcloud_render (C=4); packed [H,W,3] u8 scratch;
 *   out [H,W,3] u8, or NULL to stop after `packed` (the bokeh stage sits between the two);  depth_out [H,W] fp32 or NULL (filled
 *   depth plane, needed only by the bokeh stage). */
CSB_API int csb_kenburns_frame(const float* points, const float* data, int N, int H, int W, double focal, double baseline,
                       const float* shift, const float* shift_dev, int pw, int ph, double cx, double cy,
                       int32_t* zkey, float* zee, float* acc, uint8_t* packed, uint8_t* out, float* depth_out, void* stream);

/* The whole frame loop of kenburns_effect.py:1015-1072 in one call: F frames with per-frame camera shifts `shifts` [F,3] (host floats) -> out
 * [F,H,W,3] u8 (device).  host_out (pinned, optional) + copy_stream: each finished frame is copied to the host on copy_stream, overlapping the
 * next frame's render; `stream` waits for the last copy. */
CSB_API int csb_kenburns_frames(const float* points, const float* data, int N, int H, int W, double focal, double baseline, const float* shifts, int F,
                                int pw, int ph, double cx, double cy, int32_t* zkey, float* zee, float* acc, uint8_t* packed, uint8_t* out,
                                uint8_t* host_out, void* copy_stream, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense contractions on the tcgen05 tensor cores (csrc/tc_conv.cu).  In the reference every conv / linear layer of
 * the detector, ISNet, LeReS and Inpaint nets is a torch.nn.Conv2d / Linear dispatched to cuDNN/cuBLAS fp32 NCHW
 * (e.g. animeinsseg/models/animeseg_refine/isnet.py:95-108, depth_modules/leres/leres/network_auxi.py:100-124,
 * anime_3dkenburns/models/pointcloud_inpainting.py:85-90 and the mmdet modules restated in SURVEY.md Appendix A).
 *
 * csb_conv2d_nhwc:  y = act(conv(x, w) + bias [+ residual]) with fp32 accumulation.
 *   x        NHWC fp16/bf16 [N,Hin,Win,in_ld], channels in_coff .. in_coff+Cin are convolved (Cin % 16 == 0)
 *   w        [Cout][R][S][Cin] fp16/bf16 (K-major rows; BatchNorm / layer-scale already folded in)
 *   bias     [Cout] fp32 or NULL;  act_param: per-channel PReLU slope [Cout] or NULL
 *   residual NHWC like y (res_ld/res_coff), added before (res_mode 1) or after (res_mode 2) the activation
 *   y        NHWC fp16/bf16 [N,Hout,Wout,out_ld] written at channel offset out_coff (concat fusion), or
 *   y_f32    the same in fp32 (exactly one of y / y_f32 is used when both are given: y_f32 wins)
 * ------------------------------------------------------------------------------------------- */
#define CSB_ACT_NONE 0
#define CSB_ACT_RELU 1
#define CSB_ACT_SILU 2
#define CSB_ACT_GELU 3        /* exact erf GELU */
#define CSB_ACT_PRELU 4
#define CSB_ACT_SIGMOID 5
#define CSB_ACT_SOFTPLUS 6
#define CSB_ACT_HARDSIGMOID 7

typedef struct csb_conv_desc {
    int N, Hin, Win, Cin;
    int in_ld, in_coff;
    int Cout, R, S, stride, pad, dil;
    int out_ld, out_coff;
    int act;
    int res_mode, res_ld, res_coff;
    int dtype;                /* 0 fp16, 1 bf16 */
    int groups;               /* 0/1 dense; > 1: grouped conv (Cin == Cout, Cin % 64 == 0, 64 % (Cin/groups) == 0) with w given as
                                 [Cout][R][S][64] = the 64x64 block-diagonal slice of each 64 output channels (ResNeXt 32x8d 3x3 convs,
                                 depth_modules/leres/leres/Resnext_torch.py:70-118) */
} csb_conv_desc;

CSB_API int csb_conv2d_nhwc(const csb_conv_desc* desc, const void* x, const void* w, const float* bias, const float* act_param,
                    const void* residual, void* y, float* y_f32, void* stream);

/* Tuning / A-B switch of the conv engine's CTA-pair path (tcgen05 cta_group::2: two SMs share one weight tile): 0 never, 1 (default) for the
 * wide-N layers, 2 whenever the shape allows it.  Same values as the CSB_CTA_PAIR environment variable; returns the previous mode.  Results do not
 * depend on the mode beyond fp32 accumulation order (identical: the k-order is unchanged). */
CSB_API int csb_conv_set_pair_mode(int mode);

/* A-B switch of the conv engine's GELU epilogue (exact-erf GELU, mmpretrain ConvNeXtBlock act_cfg=dict(type='GELU')): 1 (default) the fitted
 * atanh(erf(x / sqrt2)) argument through one tanh.approx MUFU op per element, 0 the round-1 mix of a degree-19 polynomial and the sigmoid form
 * (|error| <= 2.6e-5).  Both are below the fp16 rounding of the output; the tanh form issues fewer FMA-pipe instructions (profiles/r2_gelu_form_ab.md).
 * CSB_GELU_FORM=poly / tanh in the environment selects the initial form; returns the previous form. */
CSB_API int csb_conv_set_gelu_form(int form);

/* Halo-tile path of the conv engine (csrc/tc_halo.cu) for stride-1 RxS convolutions: ONE activation halo box per 64-channel chunk feeds all R*S
 * taps through shifted tcgen05 shared-memory descriptors (the per-tap path moves R*S x the activation bytes from L2 to shared memory).
 *   desc->groups <= 1: w = [Cout][R][S][Cin] as for csb_conv2d_nhwc (which routes its eligible shapes with Cout <= 128 here by itself);
 *   desc->groups  > 1: Cin == Cout, w COMPACT = [Cout][R][S][gk], gk = max(16, Cin / groups): per output channel the gk input channels of the
 *                      gk-aligned block that holds its group (zeros outside the group when Cin / groups < 16).  groups == Cin is a depthwise conv
 *                      (ConvNeXt 7x7, mmpretrain ConvNeXtBlock; CSPNeXt 5x5), ResNeXt-101 32x8d gives gk = 16 / 16 / 32 / 64
 *                      (depth_modules/leres/leres/Resnext_torch.py:70-118).
 * act: none / relu / silu / prelu.  stats (optional, grouped only): [N*H*W][Cout/64][2] fp32 (sum, sum of squares) of the rounded outputs per
 * pixel and 64-channel slice -- the input csb_conv2d_ln_nhwc expects after a depthwise conv.
 * csb_conv_halo_supported: 1 if the shape can take this path (stride 1, R, S >= 2, (S-1) * dil <= 8, Cin % 64 == 0, dense Cout <= 256).
 * csb_conv_set_halo_mode: 0 off, 1 on (default; CSB_CONV_HALO), 2 on with a non-zero descriptor base offset (diagnostic: the B200 applies the
 * swizzle to absolute shared-memory address bits, mode 2 gives wrong results and exists to document that); returns the previous mode. */
CSB_API int csb_conv_halo_supported(const csb_conv_desc* desc);
CSB_API int csb_conv_set_halo_mode(int mode);
CSB_API int csb_conv2d_halo_nhwc(const csb_conv_desc* desc, const void* x, const void* w, const float* bias, const float* act_param, const void* residual,
                         void* y, float* y_f32, float* stats, void* stream);

/* LayerNorm folded into the 1x1 conv that consumes it (the ConvNeXt block: depthwise 7x7 -> LayerNorm -> Linear C->4C -> GELU, mmpretrain
 * ConvNeXtBlock).  csb_dwconv_stats_nhwc is csb_dwconv_nhwc (K = 7, no activation) that also writes, per pixel and
 * per 64-channel chunk, (sum, sum of squares) of its fp16 outputs: stats [N*H*W][C/64][2] fp32.  csb_conv2d_ln_nhwc then computes
 *     y = act( rstd * (x W'^T - mean * colsum) + bias' ),   W' = W diag(gamma) (packed fp16),  colsum[co] = sum_ci W'[co][ci],
 *     bias' = bias + W beta,   mean / rstd = LayerNorm statistics of the row from `stats` (biased variance, eps inside the sqrt)
 * on the UN-normalised x (desc: dense 1x1, stride 1, Cin % 64 == 0, act none or gelu): the LayerNorm pass over the activations is gone. */
CSB_API int csb_dwconv_stats_nhwc(const void* x, int ldx, int xoff, const float* w, const float* bias, int N, int H, int W, int C, int K, void* y, int ldy,
                          int yoff, float* stats, void* stream);
CSB_API int csb_conv2d_ln_nhwc(const csb_conv_desc* desc, const void* x, const void* w, const float* bias, const float* colsum, const float* stats,
                       float eps, const void* residual, void* y, void* stream);

/* The ConvNeXt block's MLP in ONE launch (csrc/tc_conv.cu: k_mlp_tc): y = residual + W2 GELU(LayerNorm(x) W1^T + b1) + b2 with the LayerNorm
 * folded as in csb_conv2d_ln_nhwc (w1 = W1 diag(gamma) packed [hidden][C], b1 = bias + W1 beta, colsum, stats from csb_dwconv_stats_nhwc) and the
 * layer scale folded into w2 [C][hidden] / b2.  The 4C-wide hidden activations stay in tensor memory / shared memory.  C = 128 or 256 with
 * hidden = 4 C (the two wide-image stages of ConvNeXt-B, mmpretrain ConvNeXtBlock); x, residual, y: NHWC fp16 (dtype 0) / bf16 (1) with `pixels`
 * rows, channel strides *_ld and offsets *_coff (multiples of 8); y may alias residual.  Bit-identical to csb_conv2d_ln_nhwc + csb_conv2d_nhwc.
 * csb_convnext_mlp_supported: 1 if the shape qualifies and the path is enabled (CSB_FUSE_MLP, default 1; csb_convnext_mlp_set_mode returns the
 * previous mode). */
CSB_API int csb_convnext_mlp_supported(int C, int hidden);
CSB_API int csb_convnext_mlp_set_mode(int mode);
CSB_API int csb_convnext_mlp_nhwc(const void* x, int x_ld, int x_coff, long long pixels, int C, int hidden, const void* w1, const float* b1, const float* colsum,
                          const float* stats, float eps, const void* w2, const float* b2, const void* residual, int res_ld, int res_coff, void* y, int y_ld,
                          int y_coff, int dtype, void* stream);

/* HBM-bound NHWC fp16 layers between the tensor-core convs (csrc/nn_elem.cu).  Channel-slice addressing: ld = channels of the
 * buffer, off = first channel, so concats (CSPNeXtPAFPN, MaskFeatModule, CSPLayer -- SURVEY.md Appendix A.3-A.6) need no copy.
 *   csb_dwconv_nhwc     depthwise KxK (K = 3/5/7, stride 1, zero pad K/2) + bias, then LayerNorm over C (ln_gamma/ln_beta != NULL:
 *                       the ConvNeXt block head, Appendix A.4) and/or an activation (CSPNeXtBlock depthwise 5x5 + folded BN + SiLU).
 *                       w is [K][K][C] fp32.
 *   csb_layernorm_nhwc  LayerNorm2d over the channels of each pixel (ConvNeXt stem / downsample / output norms).
 *   csb_resample_nhwc   mode 0 nearest, 1 bilinear align_corners=False, 2 bilinear align_corners=True, into a channel slice.
 *   csb_image_prep_nhwc uint8 HWC -> fp16 NHWC with CP channels (zero padded), (x - mean)/std, optional R/B swap
 *                       (mmdet DetDataPreprocessor; animeinsseg/__init__.py:63-76). */
CSB_API int csb_dwconv_nhwc(const void* x, int ldx, int xoff, const float* w, const float* bias, const float* ln_gamma, const float* ln_beta,
                    float eps, int act, int N, int H, int W, int C, int K, void* y, int ldy, int yoff, void* stream);
CSB_API int csb_layernorm_nhwc(const void* x, int ldx, int xoff, const float* gamma, const float* beta, float eps, long long npix, int C,
                       void* y, int ldy, int yoff, void* stream);
CSB_API int csb_resample_nhwc(const void* x, int ldx, int xoff, int N, int Hi, int Wi, int C, int Ho, int Wo, int mode, void* y, int ldy, int yoff,
                      void* stream);
CSB_API int csb_image_prep_nhwc(const uint8_t* img, long long npix, const float* mean3, const float* std3, int swap_rb, int CP, void* y, void* stream);
/*   csb_image_prep_s2d_nhwc  the same normalisation written space-to-depth for a P x P stride-P patchify stem (ConvNeXt 4x4 s4): y [N,H/P,W/P,CP],
 *                       channel (r*P+s)*3+c = pixel (P*oy+r, P*ox+s), zero padded to CP; the stem conv then is a 1x1 GEMM over CP channels. */
CSB_API int csb_image_prep_s2d_nhwc(const uint8_t* img, int N, int H, int W, int P, const float* mean3, const float* std3, int swap_rb, int CP, void* y,
                            void* stream);
/*   csb_maxpool_nhwc    MaxPool2d(kernel 3, stride 2, padding 1) (ResNeXt stem, Resnext_torch.py:160).
 *   csb_add_nhwc        y = a + b on channel slices (FFM skip add, network_auxi.py:207).
 *   csb_resample_f32    single-channel fp32 bilinear resize, align_corners selectable (AO output upsample, network_auxi.py:251). */
CSB_API int csb_maxpool_nhwc(const void* x, int N, int H, int W, int C, void* y, void* stream);
/*   csb_maxpool2d_nhwc  general MaxPool2d(K, stride, pad, ceil_mode) on channel slices (ISNet's 2x2 ceil-mode pools, isnet.py:130). */
CSB_API int csb_maxpool2d_nhwc(const void* x, int ldx, int xoff, int N, int H, int W, int C, int K, int stride, int pad, int ceil_mode, void* y, int ldy,
                       int yoff, void* stream);
/*   csb_gap_nhwc / csb_scale_channels_nhwc  mmdet ChannelAttention of the CSPNeXt CSPLayer: y[n,c] = mean_hw x (fp32
 *                       accumulation in acc [N,C], fp16 result), and x[n,h,w,c] *= s[n,c] in place; the fc + Hardsigmoid between them is a conv launch. */
CSB_API int csb_gap_nhwc(const void* x, int ldx, int xoff, int N, int H, int W, int C, float* acc, void* y, void* stream);
CSB_API int csb_scale_channels_nhwc(void* x, int ldx, int xoff, int N, int H, int W, int C, const void* scale, void* stream);
CSB_API int csb_add_nhwc(const void* a, int lda, int aoff, const void* b, int ldb, int boff, long long npix, int C, void* y, int ldy, int yoff, void* stream);
/*   csb_prelu_nhwc      y = PReLU(x), per-channel slopes (pre-activations of the Inpaint GridNet, pointcloud_inpainting.py:10-13). */
CSB_API int csb_prelu_nhwc(const void* x, int ldx, int xoff, const float* slope, long long npix, int C, void* y, int ldy, int yoff, void* stream);
CSB_API int csb_resample_f32(const float* x, int N, int Hi, int Wi, int Ho, int Wo, int align_corners, float* y, void* stream);

/* RTMDet-Ins post-processing (csrc/det_post.cu) -- SURVEY.md §8a rows A5-A8; mmdet predict_by_feat / mmcv batched_nms restated in
 * SURVEY.md Appendix A.7, mask head animeinsseg/models/rtmdet_inshead_custom.py:253-303, mask tail animeinsseg/__init__.py:361-370.
 * Single class.  Per level l (L <= 4): cls[l] [N,h,w,1], reg[l] [N,h,w,4] (ltrb distances in pixels), ker[l] [N,h,w,169], all fp32 NHWC.
 *   csb_rtmdet_select: score = sigmoid(cls) > score_thr, top nms_pre per level (stable descending), distance2bbox + clamp to the image,
 *     min_bbox_size filter (< 0 disables), greedy NMS (IoU > iou_thr suppressed), first max_per_img kept.
 *     scratch: cand [N, L*nms_pre, 10] fp32, cand_count [N*L] int32.
 *     out: boxes [N,max_per_img,4] xyxy, scores [N,max_per_img], priors [N,max_per_img,4] (x,y,stride,stride),
 *          kernels [N,max_per_img,169], num [N] int32 (device; rows >= num[n] are zero).
 *   csb_rtmdet_masks: logits [N,max_per_img,h,w] (scratch/output) = dynamic-conv mask head on mask_feat [N,h,w,8] fp32; then
 *     F.interpolate(scale_factor=stride0) -> F.interpolate(size=(resized_h,resized_w)) -> crop [:out_h,:out_w] -> sigmoid > mask_thr
 *     -> masks [N,max_per_img,out_h,out_w] uint8 0/1 (torch.bool compatible). */
CSB_API int csb_rtmdet_select(const float* const* cls, const float* const* reg, const float* const* ker, const int* hs, const int* ws, const int* strides,
                      int L, int N, float score_thr, int nms_pre, float iou_thr, int max_per_img, float min_bbox_size, int img_h, int img_w,
                      float* cand, int* cand_count, float* boxes, float* scores, float* priors, float* kernels, int* num, void* stream);
CSB_API int csb_rtmdet_masks(const float* mask_feat, const float* kernels, const float* priors, const int* num, int N, int max_per_img, int h, int w,
                     int stride0, int out_h, int out_w, int resized_h, int resized_w, float mask_thr, float* logits, uint8_t* masks, void* stream);

/* ISNet mask refinement glue (csrc/det_refine.cu) -- animeinsseg/__init__.py:37-55 (prepare_refine_batch), :638-665 (_postprocess_refine).
 *   csb_refine_prep: img [h,w,3] u8 (the image already scaled down to <= S), masks [K,H,W] u8 0/1 -> x16 [K,S,S,16] fp16 NHWC =
 *                    {B,G,R}/255, mask resized like cv2.resize(float, INTER_LINEAR), zero padded bottom/right (resize_pad, utils/io_utils.py:277-292).
 *   csb_refine_post: d1 [K,S,S] fp32 logits -> sigmoid -> crop [:h,:w] -> bilinear(align_corners=True) to (H,W) -> > mask_thr -> masks_out [K,H,W] u8. */
CSB_API int csb_refine_prep(const uint8_t* img_hw3, int h, int w, const uint8_t* masks, int K, int H, int W, int S, void* x16, void* stream);
CSB_API int csb_refine_post(const float* d1, int K, int S, int h, int w, int H, int W, float mask_thr, uint8_t* masks_out, void* stream);

/* ZoeDepth metric head, elementwise stages (csrc/zoe_head.cu) -- depth_modules/zoedepth/models/zoedepth/zoedepth_v1.py:150-192,
 * layers/attractor.py:139-208, layers/dist_layers.py:29-121.  All fp32 NHWC unless stated.
 *   csb_zoe_attractor      A [N,h,w,na] (softplus outputs), b_prev [N,hp,wp,nbins] -> bilinear(align_corners=True) -> b_new [N,h,w,nbins] =
 *                          c + mean_a dx/(1 + alpha dx^2), dx = A_a - c  (alpha = 300: the function default the reference actually uses).
 *   csb_zoe_cond_input     out176 [N,H,W,176] fp16 = [feat32 fp16 (0..31) | interp(emb128 fp16 [N,he,we,128]) (32..159) | interp(rel [N,hr,wr]) (160) | 0-pad]:
 *                          the reference's cat([feat, rel, emb]) with the rel channel moved behind the embedding (16 B aligned groups); permute the input
 *                          channels of conditional_log_binomial.mlp.0 accordingly.
 *   csb_zoe_logbinom_depth pt4 [N,H,W,4] (softplus outputs: p0,p1,t0,t1), b_centers [N,hb,wb,nbins] -> depth [N,H,W]. */
CSB_API int csb_zoe_attractor(const float* A, int na, const float* b_prev, int hp, int wp, int N, int h, int w, int nbins, float alpha, float* b_new, void* stream);
CSB_API int csb_zoe_cond_input(const void* feat32, const float* rel, int hr, int wr, const void* emb128, int he, int we, int N, int H, int W, void* out176,
                       void* stream);
CSB_API int csb_zoe_logbinom_depth(const float* pt4, const float* b_centers, int hb, int wb, int N, int H, int W, int nbins, float p_eps, float min_temp,
                           float max_temp, float* depth, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSB200_H */
