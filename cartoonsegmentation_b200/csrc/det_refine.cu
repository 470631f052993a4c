// ISNet mask-refinement glue for sm_100a: the tensors around the ISNetDIS forward of
// `AnimeInsSeg._postprocess_refine` (animeinsseg/__init__.py:638-665) and `prepare_refine_batch` (:37-55) / `resize_pad` (utils/io_utils.py:277-292).
// The reference copies every mask to the host, resizes with OpenCV on the CPU, stacks with numpy, uploads, and downloads every prediction again
// (2K host<->device copies per image); here both ends are one kernel each and nothing leaves the device.
//
//   k_refine_prep  x[k] = [B, G, R, mask_k, 0 x 12] as NHWC fp16 at S x S: image = uint8 already scaled to (h,w) (bit-exact cv2 path: csb_resize_u8c3),
//                  * float32(1/255); mask_k = cv2.resize(float mask, INTER_LINEAR) evaluated on the fly from the bool mask (float bilinear, OpenCV's
//                  coordinate rule); bottom/right zero padding to S (resize_pad pads with 0 only there).  32 B written per pixel per instance.
//   k_refine_post  preds = sigmoid(d1)[..., :h, :w] -> bilinear(align_corners=True) to (H,W) -> > mask_thr -> bool.  K*H*W bytes written.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ void cv_linear_coord(int d, double scale, int src_size, int& i0, int& i1, float& a0, float& a1, bool clamp_weights) {
    float f = (float) ((d + 0.5) * scale - 0.5);
    int s = (int) floorf(f);
    f -= s;
    if (clamp_weights) {               // x direction (resize.cpp: fx = 0 when the tap leaves the image)
        if (s < 0) { f = 0; s = 0; }
        if (s >= src_size - 1) { f = 0; s = src_size - 1; }
        i0 = s;
        i1 = s + 1 < src_size ? s + 1 : s;
    } else {                           // y direction: rows are clamped, weights kept
        i0 = s < 0 ? 0 : (s > src_size - 1 ? src_size - 1 : s);
        i1 = s + 1 < 0 ? 0 : (s + 1 > src_size - 1 ? src_size - 1 : s + 1);
    }
    a0 = 1.f - f;
    a1 = f;
}

__global__ void __launch_bounds__(256) k_refine_prep(const uint8_t* __restrict__ img, int h, int w, const uint8_t* __restrict__ masks, int K, int H, int W,
                                                     int S, __half* __restrict__ x) {
    const long long total = (long long) K * S * S;
    const double sx = 1.0 / ((double) w / W), sy = 1.0 / ((double) h / H);
    const bool same = (h == H && w == W);
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int px = (int) (i % S), py = (int) ((i / S) % S), k = (int) (i / ((long long) S * S));
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (px < w && py < h) {
            const uint8_t* p = img + ((size_t) py * w + px) * 3;
            const float s255 = (float) (1.0 / 255.0);            // img.astype(float32) / 255.  (:41) -- division by 255.f, restated as in numpy
            v[0] = (float) p[0] / 255.0f; v[1] = (float) p[1] / 255.0f; v[2] = (float) p[2] / 255.0f;
            (void) s255;
            const uint8_t* m = masks + (size_t) k * H * W;
            if (same) {
                v[3] = m[(size_t) py * W + px] ? 1.0f : 0.0f;
            } else {
                int x0, x1, y0, y1;
                float a0, a1, b0, b1;
                cv_linear_coord(px, sx, W, x0, x1, a0, a1, true);
                cv_linear_coord(py, sy, H, y0, y1, b0, b1, false);
                const float r0 = (m[(size_t) y0 * W + x0] ? a0 : 0.f) + (m[(size_t) y0 * W + x1] ? a1 : 0.f);
                const float r1 = (m[(size_t) y1 * W + x0] ? a0 : 0.f) + (m[(size_t) y1 * W + x1] ? a1 : 0.f);
                v[3] = r0 * b0 + r1 * b1;
            }
        }
        __half2* o = reinterpret_cast<__half2*>(x + (size_t) i * 16);
        o[0] = __floats2half2_rn(v[0], v[1]);
        o[1] = __floats2half2_rn(v[2], v[3]);
        const __half2 z = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int j = 2; j < 8; ++j) o[j] = z;
    }
}

__global__ void __launch_bounds__(256) k_refine_post(const float* __restrict__ d1, int K, int S, int h, int w, int H, int W, float thr, uint8_t* __restrict__ out) {
    const long long total = (long long) K * H * W;
    const float sh = H > 1 ? (float) (h - 1) / (float) (H - 1) : 0.f, sw = W > 1 ? (float) (w - 1) / (float) (W - 1) : 0.f;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int ox = (int) (i % W), oy = (int) ((i / W) % H), k = (int) (i / ((long long) W * H));
        const float* D = d1 + (size_t) k * S * S;
        const float fy = oy * sh, fx = ox * sw;
        const int y0 = min((int) fy, h - 1), x0 = min((int) fx, w - 1), y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
        const float ly = fy - y0, lx = fx - x0;
        auto sg = [&](int yy, int xx) { return 1.0f / (1.0f + expf(-D[(size_t) yy * S + xx])); };
        const float v = (1.f - ly) * ((1.f - lx) * sg(y0, x0) + lx * sg(y0, x1)) + ly * ((1.f - lx) * sg(y1, x0) + lx * sg(y1, x1));
        out[i] = v > thr ? 1 : 0;
    }
}

}  // namespace

extern "C" int csb_refine_prep(const uint8_t* img_hw3, int h, int w, const uint8_t* masks, int K, int H, int W, int S, void* x16, void* stream) {
    CSB_REQUIRE(img_hw3 && masks && x16 && K > 0 && h > 0 && w > 0 && h <= S && w <= S && H > 0 && W > 0, "bad arguments");
    k_refine_prep<<<csb::wave_grid((long long) K * S * S, 256, 8), 256, 0, (cudaStream_t) stream>>>(img_hw3, h, w, masks, K, H, W, S, (__half*) x16);
    return csb::launched("k_refine_prep", (cudaStream_t) stream);
}

extern "C" int csb_refine_post(const float* d1, int K, int S, int h, int w, int H, int W, float mask_thr, uint8_t* masks_out, void* stream) {
    CSB_REQUIRE(d1 && masks_out && K > 0 && h > 0 && w > 0 && h <= S && w <= S && H > 0 && W > 0, "bad arguments");
    k_refine_post<<<csb::wave_grid((long long) K * H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(d1, K, S, h, w, H, W, mask_thr, masks_out);
    return csb::launched("k_refine_post", (cudaStream_t) stream);
}
