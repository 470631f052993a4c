// LeReS post-processing on the device (reference depth_modules/leres/__init__.py:117-140 `apply_leres` tail and
// anime_3dkenburns/kenburns_effect.py:572-577): the reference downloads the network output and runs numpy + OpenCV per image on the host
//     out   = 65535 * (depth - min) / (max - min)            (float32; zeros if max - min <= eps)
//     d8    = bitwise_not(convertScaleAbs(out.astype(uint16), alpha = 255 / 65535))
//     depth = cv2.resize(d8, (W, H), INTER_AREA if upscaling else INTER_LANCZOS4).astype(float32)
// which stalls the GPU pipeline behind the host.  Here: per-image min/max (order-preserving keys), quantisation with the same float32 operation
// order and OpenCV's rounding (cvRound = round-half-even of src * (float) alpha), and OpenCV's INTER_AREA-upscale = INTER_LINEAR fixed-point
// kernel with area-mode source positions (imgproc/src/resize.cpp), all bit-exact against cv2 (oracle: orc_resize_area_up_u8c1, pinned to cv2
// in the CPU suite).  Downscaling (INTER_LANCZOS4, inputs smaller than the estimator size) stays on the host path.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned okey(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// mm [N][2]: min key (init 0xffffffff), max key (init 0)
__global__ void __launch_bounds__(256) k_lt_minmax(const float* __restrict__ x, long long per, unsigned* __restrict__ mm) {
    const int img = blockIdx.y;
    const float* X = x + (long long) img * per;
    unsigned lo = 0xffffffffu, hi = 0u;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x) {
        const unsigned k = okey(X[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 2 * img, lo);
        atomicMax(mm + 2 * img + 1, hi);
    }
}

__global__ void __launch_bounds__(256) k_lt_quant(const float* __restrict__ x, long long per, const unsigned* __restrict__ mm, uint8_t* __restrict__ q) {
    const int img = blockIdx.y;
    const float mn = okey_inv(mm[2 * img]), mx = okey_inv(mm[2 * img + 1]);
    const float range = __fsub_rn(mx, mn);
    const bool ok = (double) range > 2.220446049250313e-16;                 // np.finfo("float").eps
    const float alpha = (float) (255.0 / 65535.0);
    const float* X = x + (long long) img * per;
    uint8_t* Q = q + (long long) img * per;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x) {
        float out = 0.f;
        if (ok) out = __fdiv_rn(__fmul_rn(65535.0f, __fsub_rn(X[i], mn)), range);
        const unsigned u16 = (unsigned) out & 0xffffu;                        // astype("uint16"): truncation (values are in [0, 65535])
        int v = __float2int_rn(__fmul_rn((float) u16, alpha));                // convertScaleAbs: saturate_cast<uchar>(|src * a|), cvRound
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        Q[i] = (uint8_t) (255 - v);                                           // bitwise_not
    }
}

__device__ __forceinline__ void area_coef(int d, double scale, double inv, int n, int& s, short& a0, short& a1) {
    s = (int) floor(d * scale);
    float f = (float) ((d + 1) - (s + 1) * inv);
    f = f <= 0 ? 0.f : f - floorf(f);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n - 1) { f = 0.f; s = n - 1; }
    a0 = (short) __float2int_rn((1.f - f) * 2048.f);
    a1 = (short) __float2int_rn(f * 2048.f);
}

__global__ void __launch_bounds__(256) k_lt_resize(const uint8_t* __restrict__ q, int h, int w, int H, int W, float* __restrict__ out) {
    const int img = blockIdx.y;
    const uint8_t* S = q + (long long) img * h * w;
    float* O = out + (long long) img * H * W;
    const double inv_x = (double) W / w, inv_y = (double) H / h, scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < (long long) H * W; i += (long long) gridDim.x * blockDim.x) {
        const int dx = (int) (i % W), dy = (int) (i / W);
        int sx, sy;
        short a0, a1, b0, b1;
        area_coef(dx, scale_x, inv_x, w, sx, a0, a1);
        sy = (int) floor(dy * scale_y);
        float fy = (float) ((dy + 1) - (sy + 1) * inv_y);
        fy = fy <= 0 ? 0.f : fy - floorf(fy);
        b0 = (short) __float2int_rn((1.f - fy) * 2048.f);
        b1 = (short) __float2int_rn(fy * 2048.f);
        const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
        const int x1 = sx + 1 < w ? sx + 1 : sx;
        const int s0 = S[(long long) y0 * w + sx] * a0 + S[(long long) y0 * w + x1] * a1;
        const int s1 = S[(long long) y1 * w + sx] * a0 + S[(long long) y1 * w + x1] * a1;
        int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        O[i] = (float) v;
    }
}

__global__ void __launch_bounds__(256) k_lt_tofloat(const uint8_t* __restrict__ q, long long n, float* __restrict__ out) {
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) out[i] = (float) q[i];
}

// depth[depth == 0] = depth[depth > 0].min()  (kenburns_effect.py:577), per image: positive floats order like their bit patterns
__global__ void __launch_bounds__(256) k_min_positive(const float* __restrict__ x, long long per, unsigned* __restrict__ slot) {
    const int img = blockIdx.y;
    const float* X = x + (long long) img * per;
    unsigned lo = 0x7f800000u;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x) {
        const float v = X[i];
        if (v > 0.f) lo = min(lo, __float_as_uint(v));
    }
    for (int o = 16; o; o >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if ((threadIdx.x & 31) == 0) atomicMin(slot + img, lo);
}
__global__ void __launch_bounds__(256) k_zero_fill(float* __restrict__ x, long long per, const unsigned* __restrict__ slot) {
    const int img = blockIdx.y;
    float* X = x + (long long) img * per;
    const float m = __uint_as_float(slot[img]);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x)
        if (X[i] == 0.f) X[i] = m;
}

}  // namespace

extern "C" int csb_zero_to_min_positive(float* x, int N, long long per, unsigned* scratch, void* stream) {
    CSB_REQUIRE(x && scratch && N > 0 && per > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(scratch, 0x7f, sizeof(unsigned) * N, st), "memset"));      // 0x7f7f7f7f = 3.39e38 (+inf stand-in: no positive pixel)
    csb::memset_done(st);
    int gx = (4 * csb::num_sms() + N - 1) / N;
    gx = gx < 1 ? 1 : gx;
    k_min_positive<<<dim3(gx, N), 256, 0, st>>>(x, per, scratch);
    CSB_TRY(csb::launched("k_min_positive", st));
    k_zero_fill<<<dim3(gx, N), 256, 0, st>>>(x, per, scratch);
    return csb::launched("k_zero_fill", st);
}

extern "C" int csb_leres_depth_tail(const float* logits, int N, int h, int w, int H, int W, unsigned* minmax, uint8_t* q8, float* out, void* stream) {
    CSB_REQUIRE(logits && minmax && q8 && out && N > 0 && h > 0 && w > 0, "bad arguments");
    CSB_REQUIRE(H >= h && W >= w, "only the INTER_AREA (upscaling / same size) branch of kenburns_effect.py:575-576 runs on the device");
    cudaStream_t st = (cudaStream_t) stream;
    const long long per = (long long) h * w;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(minmax, 0, sizeof(unsigned) * 2 * N, st), "memset"));
    csb::memset_done(st);
    // min slots start at 0xffffffff
    CSB_TRY(csb::cuda_ok(cudaMemset2DAsync(minmax, 2 * sizeof(unsigned), 0xff, sizeof(unsigned), N, st), "memset2d"));
    csb::memset_done(st);
    int gx = (2 * csb::num_sms() + N - 1) / N;
    gx = gx < 1 ? 1 : gx;
    k_lt_minmax<<<dim3(gx, N), 256, 0, st>>>(logits, per, minmax);
    CSB_TRY(csb::launched("k_lt_minmax", st));
    k_lt_quant<<<dim3(gx, N), 256, 0, st>>>(logits, per, minmax, q8);
    CSB_TRY(csb::launched("k_lt_quant", st));
    if (H == h && W == w) {
        k_lt_tofloat<<<csb::wave_grid(per * N, 256, 4), 256, 0, st>>>(q8, per * N, out);
        return csb::launched("k_lt_tofloat", st);
    }
    int gy = (4 * csb::num_sms() + N - 1) / N;
    k_lt_resize<<<dim3(gy < 1 ? 1 : gy, N), 256, 0, st>>>(q8, h, w, H, W, out);
    return csb::launched("k_lt_resize", st);
}
