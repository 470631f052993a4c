// RTMDet-Ins post-processing for sm_100a: candidate selection, box decode, NMS, dynamic mask head and the
// full-resolution mask tail -- everything after the head convs of `AnimeInsSeg._det_forward`.
//
// Reference path (third-party mmdet 3.3.0 / mmcv 2.1.0, restated in SURVEY.md Appendix A.7-A.8, plus the in-repo parts):
//   predict_by_feat: per level sigmoid, score_thr filter, sort + top nms_pre, distance2bbox, clamp        (call sites animeinsseg/__init__.py:308-315)
//   batched_nms (mmcv._ext) greedy IoU > thr, keep max_per_img                                             (animeinsseg/__init__.py:265-294)
//   _mask_predict_by_feat_single: rel-coord + 8 prototype maps -> 3 dynamic 1x1 convs per instance         (animeinsseg/models/rtmdet_inshead_custom.py:253-303)
//   mask tail: x8 bilinear, resize, crop, sigmoid, > thr                                                   (animeinsseg/__init__.py:361-370)
// The reference runs these as dozens of small ATen kernels with host syncs (nonzero, sort, .item()); here it is 4 launches per batch and the
// instance count stays on the device.  Single-class model (utils/constants.py:3-5), so the class-offset trick of batched_nms is the identity.
//
//   k_select_level   one CTA per (image, level): sigmoid + threshold + compaction into shared memory, bitonic sort of 64-bit keys
//                    (score bits | inverted location index = the stable descending order of the reference), decode of the first nms_pre.
//   k_nms            one CTA per image: merge the levels in P3,P4,P5 order, min-size filter, stable sort by score, greedy NMS with early exit
//                    at max_per_img, gather of boxes / scores / priors / 169 dynamic-conv parameters.
//   k_mask_logits    fused A7: relative coordinates computed on the fly + the 10->8->8->1 MLP per pixel; the [K*10,h,w] tensor the
//                    reference materialises (65 MB at K=100) never exists.
//   k_mask_tail      A8: two chained bilinear resamplings evaluated per output pixel from the stride-8 logits, sigmoid, threshold, bool store.
//                    Algorithmic bytes: K*(h*w*4) read + K*H*W written (HBM-bound).
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kMaxLevels = 4;
constexpr int kGen = 169;

struct Levels {
    const float* cls[kMaxLevels];
    const float* reg[kMaxLevels];
    const float* ker[kMaxLevels];
    int h[kMaxLevels], w[kMaxLevels], stride[kMaxLevels];
    int L;
};

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// candidate record written by k_select_level: box[4], score, prior x, prior y, stride, location, level
constexpr int kRec = 10;

__global__ void __launch_bounds__(1024) k_select_level(Levels lv, int N, float score_thr, int nms_pre, int img_h, int img_w, int cap, float* __restrict__ cand,
                                                       int* __restrict__ cand_count) {
    extern __shared__ unsigned long long keys[];          // `cap` keys
    __shared__ int s_count;
    const int lvl = blockIdx.x % lv.L, img = blockIdx.x / lv.L;
    const int h = lv.h[lvl], w = lv.w[lvl], nloc = h * w, stride = lv.stride[lvl];
    const float* cls = lv.cls[lvl] + (size_t) img * nloc;
    const float* reg = lv.reg[lvl] + (size_t) img * nloc * 4;
    // Levels with more locations than the shared-memory array holds (det_size > 1024: 160 x 160 at stride 8) go through in chunks: the best nms_pre
    // keys so far stay at the front, the next chunk's candidates are appended and the array is sorted again.  The key is a total order (score bits,
    // inverted location), so the result does not depend on the chunking; a level that fits is one chunk, as before.
    const int chunk = nloc <= cap ? nloc : cap - nms_pre;
    int take = 0;
    for (int c0 = 0; c0 < nloc; c0 += chunk) {
        __syncthreads();
        if (threadIdx.x == 0) s_count = take;
        __syncthreads();
        const int c1 = c0 + chunk < nloc ? c0 + chunk : nloc;
        for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            const float s = 1.0f / (1.0f + expf(-cls[i]));                  // scores = cls.sigmoid()
            if (s > score_thr) keys[atomicAdd(&s_count, 1)] = ((unsigned long long) __float_as_uint(s) << 32) | (unsigned) (0xffffffffu - (unsigned) i);
        }
        __syncthreads();
        const int count = s_count;
        int p2 = 1;
        while (p2 < count) p2 <<= 1;
        for (int i = count + threadIdx.x; i < p2; i += blockDim.x) keys[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(keys, p2);
        take = count < nms_pre ? count : nms_pre;
    }
    float* out = cand + ((size_t) img * lv.L + lvl) * nms_pre * kRec;
    for (int i = threadIdx.x; i < take; i += blockDim.x) {
        const unsigned long long k = keys[i];
        const int loc = (int) (0xffffffffu - (unsigned) (k & 0xffffffffull));
        const float score = __uint_as_float((unsigned) (k >> 32));
        const float px = (float) ((loc % w) * stride), py = (float) ((loc / w) * stride);       // MlvlPointGenerator(offset=0)
        const float4 d = *reinterpret_cast<const float4*>(reg + (size_t) loc * 4);
        float x1 = __fsub_rn(px, d.x), y1 = __fsub_rn(py, d.y), x2 = __fadd_rn(px, d.z), y2 = __fadd_rn(py, d.w);    // distance2bbox
        x1 = fminf(fmaxf(x1, 0.f), (float) img_w); x2 = fminf(fmaxf(x2, 0.f), (float) img_w);
        y1 = fminf(fmaxf(y1, 0.f), (float) img_h); y2 = fminf(fmaxf(y2, 0.f), (float) img_h);
        float* r = out + (size_t) i * kRec;
        r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2; r[4] = score; r[5] = px; r[6] = py; r[7] = (float) stride;
        r[8] = __int_as_float(loc); r[9] = __int_as_float(lvl);
    }
    if (threadIdx.x == 0) cand_count[img * lv.L + lvl] = take;
}

__global__ void __launch_bounds__(1024) k_nms(Levels lv, int nms_pre, float iou_thr, int max_per_img, float min_bbox_size, const float* __restrict__ cand,
                                              const int* __restrict__ cand_count, float* __restrict__ boxes, float* __restrict__ scores,
                                              float* __restrict__ priors, float* __restrict__ kernels, int* __restrict__ num_out) {
    extern __shared__ unsigned long long keys[];          // [cap] sort keys, then reused
    const int img = blockIdx.x;
    const int cap_total = lv.L * nms_pre;
    int p2 = 1;
    while (p2 < cap_total) p2 <<= 1;
    float* sbox = reinterpret_cast<float*>(keys + p2);     // [cap_total][4] sorted boxes
    int* sidx = reinterpret_cast<int*>(sbox + (size_t) cap_total * 4);   // [cap_total] record index (lvl * nms_pre + i)
    unsigned char* alive = reinterpret_cast<unsigned char*>(sidx + cap_total);
    __shared__ int s_n, s_cur, s_kept;
    __shared__ int s_keep[1024];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // concatenated order P3, P4, P5; key = (score bits, inverted concat position) -> stable descending sort
    for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        unsigned long long k = 0ull;
        if (i < cap_total) {
            const int lvl = i / nms_pre, j = i % nms_pre;
            if (j < cand_count[img * lv.L + lvl]) {
                const float* r = cand + (((size_t) img * lv.L + lvl) * nms_pre + j) * kRec;
                const float bw = __fsub_rn(r[2], r[0]), bh = __fsub_rn(r[3], r[1]);
                if (min_bbox_size < 0.f || (bw > min_bbox_size && bh > min_bbox_size))
                    k = ((unsigned long long) __float_as_uint(r[4]) << 32) | (unsigned) (0xffffffffu - (unsigned) i);
            }
        }
        keys[i] = k;
    }
    __syncthreads();
    bitonic_sort_desc(keys, p2);
    for (int i = threadIdx.x; i < cap_total; i += blockDim.x) {
        const unsigned long long k = keys[i];
        if (k != 0ull) {
            const int pos = (int) (0xffffffffu - (unsigned) (k & 0xffffffffull));
            const float* r = cand + ((size_t) img * lv.L * nms_pre + pos) * kRec;
            sbox[i * 4 + 0] = r[0]; sbox[i * 4 + 1] = r[1]; sbox[i * 4 + 2] = r[2]; sbox[i * 4 + 3] = r[3];
            sidx[i] = pos;
            alive[i] = 1;
            atomicMax(&s_n, i + 1);
        } else {
            alive[i] = 0;
        }
    }
    __syncthreads();
    const int n = s_n;
    if (threadIdx.x == 0) { s_cur = 0; s_kept = 0; }
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) {
            int c = s_cur;
            while (c < n && !alive[c]) ++c;
            s_cur = c;
        }
        __syncthreads();
        const int i = s_cur;
        if (i >= n || s_kept >= max_per_img) break;
        const float ax1 = sbox[i * 4], ay1 = sbox[i * 4 + 1], ax2 = sbox[i * 4 + 2], ay2 = sbox[i * 4 + 3];
        const float aarea = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
        for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
            if (!alive[j]) continue;
            const float bx1 = sbox[j * 4], by1 = sbox[j * 4 + 1], bx2 = sbox[j * 4 + 2], by2 = sbox[j * 4 + 3];
            const float iw = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f), ih = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
            const float inter = __fmul_rn(iw, ih);
            const float barea = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
            const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
            if (iou > iou_thr) alive[j] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_keep[s_kept] = i; s_kept = s_kept + 1; s_cur = i + 1; }
        __syncthreads();
    }
    __syncthreads();
    const int kept = s_kept;
    if (threadIdx.x == 0) num_out[img] = kept;
    for (int t = threadIdx.x; t < max_per_img * kGen; t += blockDim.x) {
        const int k = t / kGen, g = t % kGen;
        float v = 0.f;
        if (k < kept) {
            const float* r = cand + ((size_t) img * lv.L * nms_pre + sidx[s_keep[k]]) * kRec;
            const int loc = __float_as_int(r[8]), lvl = __float_as_int(r[9]);
            v = lv.ker[lvl][((size_t) img * lv.h[lvl] * lv.w[lvl] + loc) * kGen + g];
        }
        kernels[((size_t) img * max_per_img + k) * kGen + g] = v;
    }
    for (int k = threadIdx.x; k < max_per_img; k += blockDim.x) {
        float b[4] = {0, 0, 0, 0}, s = 0.f, p[4] = {0, 0, 0, 0};
        if (k < kept) {
            const float* r = cand + ((size_t) img * lv.L * nms_pre + sidx[s_keep[k]]) * kRec;
            b[0] = r[0]; b[1] = r[1]; b[2] = r[2]; b[3] = r[3]; s = r[4]; p[0] = r[5]; p[1] = r[6]; p[2] = r[7]; p[3] = r[7];
        }
        for (int c = 0; c < 4; ++c) { boxes[((size_t) img * max_per_img + k) * 4 + c] = b[c]; priors[((size_t) img * max_per_img + k) * 4 + c] = p[c]; }
        scores[(size_t) img * max_per_img + k] = s;
    }
}

// A7 fused: one CTA per (instance, 256-pixel strip).  mask_feat NHWC [N,h,w,8] fp32.
__global__ void __launch_bounds__(256) k_mask_logits(const float* __restrict__ mask_feat, const float* __restrict__ kernels, const float* __restrict__ priors,
                                                     const int* __restrict__ num, int max_per_img, int h, int w, int stride0, float* __restrict__ logits) {
    const int inst = blockIdx.y, img = blockIdx.z;
    if (inst >= num[img]) return;
    __shared__ float prm[kGen];
    const float* K = kernels + ((size_t) img * max_per_img + inst) * kGen;
    for (int i = threadIdx.x; i < kGen; i += blockDim.x) prm[i] = K[i];
    __syncthreads();
    const float* P = priors + ((size_t) img * max_per_img + inst) * 4;
    const float px = P[0], py = P[1], den = __fmul_rn(P[2], 8.0f);
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= h * w) return;
    const int x = pix % w, y = pix / w;
    float in[10];
    in[0] = __fdiv_rn(__fsub_rn(px, (float) (x * stride0)), den);           // relative_coord :273-275
    in[1] = __fdiv_rn(__fsub_rn(py, (float) (y * stride0)), den);
    const float4 f0 = *reinterpret_cast<const float4*>(mask_feat + ((size_t) img * h * w + pix) * 8);
    const float4 f1 = *reinterpret_cast<const float4*>(mask_feat + ((size_t) img * h * w + pix) * 8 + 4);
    in[2] = f0.x; in[3] = f0.y; in[4] = f0.z; in[5] = f0.w; in[6] = f1.x; in[7] = f1.y; in[8] = f1.z; in[9] = f1.w;
    // parse_dynamic_params: [w0 (8x10), w1 (8x8), w2 (1x8), b0 (8), b1 (8), b2 (1)]
    const float *w0 = prm, *w1 = prm + 80, *w2 = prm + 144, *b0 = prm + 152, *b1 = prm + 160, *b2 = prm + 168;
    float h0[8], h1[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float a = b0[o];
#pragma unroll
        for (int i = 0; i < 10; ++i) a = fmaf(w0[o * 10 + i], in[i], a);
        h0[o] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float a = b1[o];
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(w1[o * 8 + i], h0[i], a);
        h1[o] = fmaxf(a, 0.f);
    }
    float a = b2[0];
#pragma unroll
    for (int i = 0; i < 8; ++i) a = fmaf(w2[i], h1[i], a);
    logits[((size_t) img * max_per_img + inst) * h * w + pix] = a;
}

// torch upsample_bilinear2d source index, align_corners=False
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float r = fmaxf(scale * ((float) dst + 0.5f) - 0.5f, 0.f);
    i0 = (int) r;
    i0 = i0 < in_size - 1 ? i0 : in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = r - (float) i0;
    l0 = 1.0f - l1;
}

// value of the x8-upsampled logit map (stage 1, scale_factor=8 -> source scale 1/8) at integer position (y,x) of the 8h x 8w grid
__device__ __forceinline__ float up8(const float* __restrict__ lg, int h, int w, int up, int y, int x) {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    const float sc = 1.0f / (float) up;
    src_index(sc, y, h, y0, y1, ly0, ly1);
    src_index(sc, x, w, x0, x1, lx0, lx1);
    return ly0 * (lx0 * __ldg(lg + y0 * w + x0) + lx1 * __ldg(lg + y0 * w + x1)) + ly1 * (lx0 * __ldg(lg + y1 * w + x0) + lx1 * __ldg(lg + y1 * w + x1));
}

__device__ __forceinline__ unsigned char mask_decide(float v, float thr, float thr_logit) {
    // sigmoid(v) > thr.  Away from the decision boundary the sign of v - logit(thr) decides; within 1e-3 of it the reference expression is evaluated.
    const float d = v - thr_logit;
    if (fabsf(d) > 1e-3f) return d > 0.f ? 1 : 0;
    return (1.0f / (1.0f + expf(-v))) > thr ? 1 : 0;
}

// 4 consecutive output pixels per thread -> one 32-bit store; the vertical interpolation set-up is shared by the 4 pixels.
__global__ void __launch_bounds__(256) k_mask_tail(const float* __restrict__ logits, const int* __restrict__ num, int max_per_img, int h, int w, int up,
                                                   int H2, int W2, float sy2, float sx2, int identity2, int H, int W, float thr, float thr_logit,
                                                   unsigned char* __restrict__ masks) {
    const int inst = blockIdx.y, img = blockIdx.z;
    if (inst >= num[img]) return;
    const float* lg = logits + ((size_t) img * max_per_img + inst) * h * w;
    unsigned char* out = masks + ((size_t) img * max_per_img + inst) * H * W;
    const int HU = h * up, WU = w * up;
    const int Wq = (W + 3) / 4;
    const bool vec = (W % 4 == 0) && (((size_t) H * W) % 4 == 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * Wq; i += gridDim.x * blockDim.x) {
        const int xq = i % Wq, y = i / Wq;
        unsigned char r[4] = {0, 0, 0, 0};
        if (identity2) {
            int y0, y1;
            float ly0, ly1;
            src_index(1.0f / (float) up, y, h, y0, y1, ly0, ly1);
            const float* r0 = lg + y0 * w;
            const float* r1 = lg + y1 * w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = xq * 4 + j;
                if (x >= W) break;
                int x0, x1;
                float lx0, lx1;
                src_index(1.0f / (float) up, x, w, x0, x1, lx0, lx1);
                const float v = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
                r[j] = mask_decide(v, thr, thr_logit);
            }
        } else {       // second F.interpolate(size=[H2,W2], align_corners=False) on the x8 map, then crop [:H,:W]
            int y0, y1;
            float ly0, ly1;
            src_index(sy2, y, HU, y0, y1, ly0, ly1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = xq * 4 + j;
                if (x >= W) break;
                int x0, x1;
                float lx0, lx1;
                src_index(sx2, x, WU, x0, x1, lx0, lx1);
                const float v = ly0 * (lx0 * up8(lg, h, w, up, y0, x0) + lx1 * up8(lg, h, w, up, y0, x1)) + ly1 * (lx0 * up8(lg, h, w, up, y1, x0) + lx1 * up8(lg, h, w, up, y1, x1));
                r[j] = mask_decide(v, thr, thr_logit);
            }
        }
        if (vec) *reinterpret_cast<uchar4*>(out + (size_t) y * W + xq * 4) = make_uchar4(r[0], r[1], r[2], r[3]);
        else
            for (int j = 0; j < 4 && xq * 4 + j < W; ++j) out[(size_t) y * W + xq * 4 + j] = r[j];
    }
}

// Fast path of the mask tail for the common case (no second resize, x8): output pixels 8c+4 .. 8c+11 of a row all interpolate between logit
// columns c and c+1 with weights (2j+1)/16, so one thread blends the two row-interpolated columns once and walks the 8 pixels with one FMA each
// (the generic kernel spends ~25 instructions per pixel on coordinates).  Decisions within 1e-3 of the threshold are re-evaluated with the
// reference's exact expression; the 4-pixel borders use the generic coordinate rule.
__global__ void __launch_bounds__(256) k_mask_tail_up8(const float* __restrict__ logits, const int* __restrict__ num, int max_per_img, int h, int w, int H, int W,
                                                       float thr, float thr_logit, unsigned char* __restrict__ masks) {
    const int inst = blockIdx.y, img = blockIdx.z;
    if (inst >= num[img]) return;
    const float* lg = logits + ((size_t) img * max_per_img + inst) * h * w;
    unsigned char* out = masks + ((size_t) img * max_per_img + inst) * H * W;
    const int cells = w + 1;                                   // c = -1 .. w-1
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * cells; i += gridDim.x * blockDim.x) {
        const int c = i % cells - 1, y = i / cells;
        int y0, y1;
        float ly0, ly1;
        src_index(0.125f, y, h, y0, y1, ly0, ly1);
        const float* r0 = lg + y0 * w;
        const float* r1 = lg + y1 * w;
        auto exact = [&](int x) {
            int x0, x1;
            float lx0, lx1;
            src_index(0.125f, x, w, x0, x1, lx0, lx1);
            return ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
        };
        const int xb = 8 * c + 4;
        if (c < 0 || c >= w - 1) {                              // left / right border: 4 pixels, generic rule
            for (int j = 0; j < 8; ++j) {
                const int x = xb + j;
                if (x < 0 || x >= W) continue;
                out[(size_t) y * W + x] = mask_decide(exact(x), thr, thr_logit);
            }
            continue;
        }
        const float a = ly0 * __ldg(r0 + c) + ly1 * __ldg(r1 + c), b = ly0 * __ldg(r0 + c + 1) + ly1 * __ldg(r1 + c + 1);
        const float d = b - a;
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = fmaf(d, (float) (2 * j + 1) * 0.0625f, a);
            unsigned char bit;
            const float dd = v - thr_logit;
            if (fabsf(dd) > 1e-3f) bit = dd > 0.f ? 1 : 0;
            else bit = mask_decide(exact(xb + j), thr, thr_logit);
            if (j < 4) lo |= (unsigned) bit << (8 * j);
            else hi |= (unsigned) bit << (8 * (j - 4));
        }
        unsigned* o = reinterpret_cast<unsigned*>(out + (size_t) y * W + xb);      // xb = 8c + 4: 4-byte aligned when W % 4 == 0
        o[0] = lo;
        o[1] = hi;
    }
}

// 16 output pixels per thread, one aligned 16 B store.  Pixels 16t .. 16t+15 of a row interpolate between the logit columns 2t-1 .. 2t+2
// (clamped at the borders, which is exactly what the align_corners=False source-index clamp does): [second half of cell 2t-1 | cell 2t | first half of
// cell 2t+1].  The interpolant is linear inside a cell, so if its six segment end points all lie on one side of the threshold (by more than the 1e-3
// margin) the whole 16-pixel run is constant -- the common case away from mask boundaries costs ~2 instructions per pixel; otherwise the pixels
// are walked individually, with the reference's exact expression inside the margin.
__global__ void __launch_bounds__(256) k_mask_tail_up8x16(const float* __restrict__ logits, const int* __restrict__ num, int max_per_img, int h, int w, int H, int W,
                                                          float thr, float thr_logit, unsigned char* __restrict__ masks) {
    const int inst = blockIdx.y, img = blockIdx.z;
    if (inst >= num[img]) return;
    const float* lg = logits + ((size_t) img * max_per_img + inst) * h * w;
    unsigned char* out = masks + ((size_t) img * max_per_img + inst) * H * W;
    const int runs = W / 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * runs; i += gridDim.x * blockDim.x) {
        const int t = i % runs, y = i / runs;
        int y0, y1;
        float ly0, ly1;
        src_index(0.125f, y, h, y0, y1, ly0, ly1);
        const float* r0 = lg + y0 * w;
        const float* r1 = lg + y1 * w;
        float A[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = min(max(2 * t - 1 + k, 0), w - 1);
            A[k] = ly0 * __ldg(r0 + c) + ly1 * __ldg(r1 + c);
        }
        const float d0 = A[1] - A[0], d1 = A[2] - A[1], d2 = A[3] - A[2];
        // segment end points: cell 2t-1 at j = 4, 7; cell 2t at j = 0, 7; cell 2t+1 at j = 0, 3   (weight (2j+1)/16)
        const float e0 = fmaf(d0, 0.5625f, A[0]), e1 = fmaf(d0, 0.9375f, A[0]), e2 = fmaf(d1, 0.0625f, A[1]), e3 = fmaf(d1, 0.9375f, A[1]),
                    e4 = fmaf(d2, 0.0625f, A[2]), e5 = fmaf(d2, 0.4375f, A[2]);
        const float lo = fminf(fminf(fminf(e0, e1), fminf(e2, e3)), fminf(e4, e5)), hi = fmaxf(fmaxf(fmaxf(e0, e1), fmaxf(e2, e3)), fmaxf(e4, e5));
        uint4 r;
        if (lo - thr_logit > 1e-3f) r = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
        else if (thr_logit - hi > 1e-3f) r = make_uint4(0u, 0u, 0u, 0u);
        else {
            auto exact = [&](int x) {
                int x0, x1;
                float lx0, lx1;
                src_index(0.125f, x, w, x0, x1, lx0, lx1);
                return ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
            };
            unsigned wds[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int seg = p < 4 ? 0 : (p < 12 ? 1 : 2), j = p < 4 ? p + 4 : (p < 12 ? p - 4 : p - 12);
                const float a = seg == 0 ? A[0] : (seg == 1 ? A[1] : A[2]), dd = seg == 0 ? d0 : (seg == 1 ? d1 : d2);
                const float v = fmaf(dd, (float) (2 * j + 1) * 0.0625f, a);
                const float m = v - thr_logit;
                unsigned bit;
                if (fabsf(m) > 1e-3f) bit = m > 0.f ? 1u : 0u;
                else bit = mask_decide(exact(16 * t + p), thr, thr_logit);
                wds[p >> 2] |= bit << (8 * (p & 3));
            }
            r = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        }
        *reinterpret_cast<uint4*>(out + (size_t) y * W + 16 * t) = r;
    }
}

}  // namespace

extern "C" int csb_rtmdet_select(const float* const* cls, const float* const* reg, const float* const* ker, const int* hs, const int* ws, const int* strides,
                                 int L, int N, float score_thr, int nms_pre, float iou_thr, int max_per_img, float min_bbox_size, int img_h, int img_w,
                                 float* cand, int* cand_count, float* boxes, float* scores, float* priors, float* kernels, int* num, void* stream) {
    CSB_REQUIRE(cls && reg && ker && hs && ws && strides && cand && cand_count && boxes && scores && priors && kernels && num, "null pointer");
    CSB_REQUIRE(L >= 1 && L <= kMaxLevels && N > 0 && nms_pre > 0 && nms_pre <= 4096 && max_per_img > 0 && max_per_img <= 1024, "bad sizes");
    Levels lv{};
    lv.L = L;
    int maxloc = 0;
    for (int l = 0; l < L; ++l) {
        lv.cls[l] = cls[l]; lv.reg[l] = reg[l]; lv.ker[l] = ker[l]; lv.h[l] = hs[l]; lv.w[l] = ws[l]; lv.stride[l] = strides[l];
        CSB_REQUIRE(cls[l] && reg[l] && ker[l] && hs[l] > 0 && ws[l] > 0, "bad level");
        maxloc = hs[l] * ws[l] > maxloc ? hs[l] * ws[l] : maxloc;
    }
    int p2 = 1;
    while (p2 < maxloc) p2 <<= 1;
    if (p2 > 16384) p2 = 16384;                            // larger levels are sorted in chunks of (16384 - nms_pre) locations (k_select_level)
    CSB_REQUIRE(nms_pre > 0 && nms_pre <= 4096, "nms_pre must be in [1, 4096]");
    cudaStream_t st = (cudaStream_t) stream;
    static unsigned char attr_done[64] = {};
    if (csb::first_use_on_device(attr_done)) {          // per device, not per process
        cudaFuncSetAttribute(k_select_level, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    k_select_level<<<N * L, 1024, (size_t) p2 * 8, st>>>(lv, N, score_thr, nms_pre, img_h, img_w, p2, cand, cand_count);
    CSB_TRY(csb::launched("k_select_level", st));
    const int cap = L * nms_pre;
    int q2 = 1;
    while (q2 < cap) q2 <<= 1;
    const size_t smem = (size_t) q2 * 8 + (size_t) cap * (16 + 4 + 1) + 16;
    CSB_REQUIRE(smem <= 200 * 1024, "nms_pre * levels too large for the shared-memory NMS");
    k_nms<<<N, 1024, smem, st>>>(lv, nms_pre, iou_thr, max_per_img, min_bbox_size, cand, cand_count, boxes, scores, priors, kernels, num);
    return csb::launched("k_nms", st);
}

extern "C" int csb_rtmdet_masks(const float* mask_feat, const float* kernels, const float* priors, const int* num, int N, int max_per_img, int h, int w,
                                int stride0, int out_h, int out_w, int resized_h, int resized_w, float mask_thr, float* logits, uint8_t* masks,
                                void* stream) {
    CSB_REQUIRE(mask_feat && kernels && priors && num && logits && masks, "null pointer");
    CSB_REQUIRE(N > 0 && max_per_img > 0 && h > 0 && w > 0 && stride0 > 0 && out_h > 0 && out_w > 0, "bad shape");
    CSB_REQUIRE(out_h <= resized_h && out_w <= resized_w, "crop larger than the resized map");
    cudaStream_t st = (cudaStream_t) stream;
    k_mask_logits<<<dim3((h * w + 255) / 256, max_per_img, N), 256, 0, st>>>(mask_feat, kernels, priors, num, max_per_img, h, w, stride0, logits);
    CSB_TRY(csb::launched("k_mask_logits", st));
    const int HU = h * stride0, WU = w * stride0;
    const int identity2 = resized_h == HU && resized_w == WU;
    // F.interpolate(size=...) without scale_factor: source scale = in/out
    const float sy2 = (float) HU / (float) resized_h, sx2 = (float) WU / (float) resized_w;
    if (identity2 && stride0 == 8 && out_h == h * 8 && out_w == w * 8) {
        const float tl = (mask_thr > 0.f && mask_thr < 1.f) ? logf(mask_thr / (1.0f - mask_thr)) : (mask_thr <= 0.f ? -INFINITY : INFINITY);
        int g8 = (out_h * (w + 1) + 255) / 256;
        g8 = g8 > 128 ? 128 : g8;
        if (out_w % 16 == 0 && (((uintptr_t) masks | (size_t) out_h * out_w) & 15) == 0) {
            int g16 = (out_h * (out_w / 16) + 255) / 256;
            g16 = g16 > 64 ? 64 : g16;
            k_mask_tail_up8x16<<<dim3(g16, max_per_img, N), 256, 0, st>>>(logits, num, max_per_img, h, w, out_h, out_w, mask_thr, tl, masks);
            return csb::launched("k_mask_tail", st);
        }
        k_mask_tail_up8<<<dim3(g8, max_per_img, N), 256, 0, st>>>(logits, num, max_per_img, h, w, out_h, out_w, mask_thr, tl, masks);
        return csb::launched("k_mask_tail", st);
    }
    int gx = (out_h * ((out_w + 3) / 4) + 255) / 256;
    gx = gx > 256 ? 256 : gx;
    const float thr_logit = (mask_thr > 0.f && mask_thr < 1.f) ? logf(mask_thr / (1.0f - mask_thr)) : (mask_thr <= 0.f ? -INFINITY : INFINITY);
    k_mask_tail<<<dim3(gx, max_per_img, N), 256, 0, st>>>(logits, num, max_per_img, h, w, stride0, resized_h, resized_w, sy2, sx2, identity2, out_h, out_w,
                                                          mask_thr, thr_logit, masks);
    return csb::launched("k_mask_tail", st);
}
