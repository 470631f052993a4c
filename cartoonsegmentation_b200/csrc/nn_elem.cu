// HBM-bound NHWC fp16 layers of the detector / depth nets for sm_100a: everything between the tensor-core convs that is
// NOT a dense contraction ("Depthwise 5x5/7x7, LayerNorm, GELU, BN-less elementwise, upsample/concat count
// toward HBM, not tensor pipe").  In the reference these are separate cuDNN / ATen kernels in fp32 NCHW
// (mmpretrain ConvNeXt block: depthwise 7x7 -> LayerNorm -> ...; CSPNeXtBlock depthwise 5x5 + BN + SiLU; CSPNeXtPAFPN nearest
// upsample + concat; MaskFeatModule bilinear upsample + concat -- SURVEY.md Appendix A.3-A.6).
//
// Layout: activations NHWC fp16, channel-slice addressing (ld = channels of the buffer, coff = first channel) so a kernel
// can read from / write into a slice of a wider tensor: concats never materialise a copy.
// All kernels: one warp per pixel, lanes own interleaved 8-channel (16 B) vectors -> every global access is a 16 B vector and a
// warp touches contiguous 512 B runs of one pixel's channels.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

struct alignas(16) H8 {
    __half2 v[4];
};

__device__ __forceinline__ void unpack8(const H8& h, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __half22float2(h.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ H8 pack8(const float (&f)[8]) {
    H8 h;
#pragma unroll
    for (int i = 0; i < 4; ++i) h.v[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return h;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float act_f(float x, int act) {
    switch (act) {
        case CSB_ACT_RELU: return fmaxf(x, 0.0f);
        case CSB_ACT_SILU: return x / (1.0f + __expf(-x));
        case CSB_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
        case CSB_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-x));
        default: return x;
    }
}

constexpr int kMaxVec = 8;   // up to 8 x 8 channels per lane = C <= 2048

// Depthwise KxK (stride 1, zero pad K/2) + bias, then optionally LayerNorm over C (eps, affine) or an activation.
//   x [N,H,W,ldx] (+xoff), w [K][K][C] fp32, y [N,H,W,ldy] (+yoff).  Algorithmic bytes: 2*2*C per pixel (read + write) + weights.
template <int K>
__global__ void __launch_bounds__(256) k_dwconv(const __half* __restrict__ x, int ldx, int xoff, const float* __restrict__ w,
                                                const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                float eps, int act, int N, int H, int W, int C, __half* __restrict__ y, int ldy, int yoff) {
    const int lane = threadIdx.x & 31;
    const long long npix = (long long) N * H * W;
    const int nvec = C / 256 + ((C % 256) ? 1 : 0);     // 8-channel vectors per lane (lane handles channels 8*(lane + 32*j) ..)
    for (long long pix = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += (long long) gridDim.x * (blockDim.x >> 5)) {
        const int px = (int) (pix % W), py = (int) ((pix / W) % H);
        const long long img = pix / ((long long) W * H);
        float acc[kMaxVec][8];
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int c0 = 8 * (lane + 32 * j);
            if (j < nvec && c0 < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[j][e] = bias ? __ldg(bias + c0 + e) : 0.0f;
            }
        }
        for (int r = 0; r < K; ++r) {
            const int iy = py + r - K / 2;
            if (iy < 0 || iy >= H) continue;
            for (int s = 0; s < K; ++s) {
                const int ix = px + s - K / 2;
                if (ix < 0 || ix >= W) continue;
                const __half* xp = x + ((img * H + iy) * W + ix) * ldx + xoff;
                const float* wp = w + (size_t) (r * K + s) * C;
#pragma unroll
                for (int j = 0; j < kMaxVec; ++j) {
                    const int c0 = 8 * (lane + 32 * j);
                    if (j < nvec && c0 < C) {
                        float f[8];
                        unpack8(*reinterpret_cast<const H8*>(xp + c0), f);
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + c0)), w1 = __ldg(reinterpret_cast<const float4*>(wp + c0 + 4));
                        acc[j][0] = fmaf(f[0], w0.x, acc[j][0]); acc[j][1] = fmaf(f[1], w0.y, acc[j][1]);
                        acc[j][2] = fmaf(f[2], w0.z, acc[j][2]); acc[j][3] = fmaf(f[3], w0.w, acc[j][3]);
                        acc[j][4] = fmaf(f[4], w1.x, acc[j][4]); acc[j][5] = fmaf(f[5], w1.y, acc[j][5]);
                        acc[j][6] = fmaf(f[6], w1.z, acc[j][6]); acc[j][7] = fmaf(f[7], w1.w, acc[j][7]);
                    }
                }
            }
        }
        float mean = 0.0f, rstd = 1.0f;
        if (ln_g) {
            float s1 = 0.0f;
#pragma unroll
            for (int j = 0; j < kMaxVec; ++j)
                if (j < nvec && 8 * (lane + 32 * j) < C)
#pragma unroll
                    for (int e = 0; e < 8; ++e) s1 += acc[j][e];
            mean = warp_sum(s1) / (float) C;
            float s2 = 0.0f;
#pragma unroll
            for (int j = 0; j < kMaxVec; ++j)
                if (j < nvec && 8 * (lane + 32 * j) < C)
#pragma unroll
                    for (int e = 0; e < 8; ++e) { float d = acc[j][e] - mean; s2 += d * d; }
            rstd = rsqrtf(warp_sum(s2) / (float) C + eps);
        }
        __half* yp = y + pix * ldy + yoff;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int c0 = 8 * (lane + 32 * j);
            if (j < nvec && c0 < C) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float v = acc[j][e];
                    if (ln_g) v = (v - mean) * rstd * __ldg(ln_g + c0 + e) + __ldg(ln_b + c0 + e);
                    o[e] = act_f(v, act);
                }
                *reinterpret_cast<H8*>(yp + c0) = pack8(o);
            }
        }
    }
}

// LayerNorm over the channels of each pixel (LayerNorm2d / channels-last LN).  4*C bytes per pixel.
__global__ void __launch_bounds__(256) k_layernorm(const __half* __restrict__ x, int ldx, int xoff, const float* __restrict__ g,
                                                   const float* __restrict__ b, float eps, long long npix, int C, __half* __restrict__ y, int ldy, int yoff) {
    const int lane = threadIdx.x & 31;
    const int nvec = C / 256 + ((C % 256) ? 1 : 0);
    for (long long pix = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix; pix += (long long) gridDim.x * (blockDim.x >> 5)) {
        const __half* xp = x + pix * ldx + xoff;
        float v[kMaxVec][8];
        float s1 = 0.0f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int c0 = 8 * (lane + 32 * j);
            if (j < nvec && c0 < C) {
                unpack8(*reinterpret_cast<const H8*>(xp + c0), v[j]);
#pragma unroll
                for (int e = 0; e < 8; ++e) s1 += v[j][e];
            }
        }
        const float mean = warp_sum(s1) / (float) C;
        float s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j)
            if (j < nvec && 8 * (lane + 32 * j) < C)
#pragma unroll
                for (int e = 0; e < 8; ++e) { float d = v[j][e] - mean; s2 += d * d; }
        const float rstd = rsqrtf(warp_sum(s2) / (float) C + eps);
        __half* yp = y + pix * ldy + yoff;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
            const int c0 = 8 * (lane + 32 * j);
            if (j < nvec && c0 < C) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd * __ldg(g + c0 + e) + __ldg(b + c0 + e);
                *reinterpret_cast<H8*>(yp + c0) = pack8(o);
            }
        }
    }
}

// Resample NHWC into a channel slice: mode 0 nearest (F.interpolate 'nearest': src = floor(dst * in/out)), 1 bilinear
// align_corners=False, 2 bilinear align_corners=True.  One thread per (pixel, 8-channel vector).
__global__ void __launch_bounds__(256) k_resample(const __half* __restrict__ x, int ldx, int xoff, int N, int Hi, int Wi, int C, int Ho, int Wo, int mode,
                                                  __half* __restrict__ y, int ldy, int yoff) {
    const int cv = C / 8;
    const long long total = (long long) N * Ho * Wo * cv;
    const float sh = mode == 2 ? (Ho > 1 ? (float) (Hi - 1) / (float) (Ho - 1) : 0.0f) : (float) Hi / (float) Ho;
    const float sw = mode == 2 ? (Wo > 1 ? (float) (Wi - 1) / (float) (Wo - 1) : 0.0f) : (float) Wi / (float) Wo;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int c0 = (int) (i % cv) * 8;
        const long long p = i / cv;
        const int ox = (int) (p % Wo), oy = (int) ((p / Wo) % Ho);
        const long long n = p / ((long long) Wo * Ho);
        const __half* xb = x + n * Hi * Wi * ldx + xoff + c0;
        H8 out;
        if (mode == 0) {
            const int iy = min((int) floorf(oy * sh), Hi - 1), ix = min((int) floorf(ox * sw), Wi - 1);
            out = *reinterpret_cast<const H8*>(xb + ((long long) iy * Wi + ix) * ldx);
        } else {
            float fy = mode == 2 ? oy * sh : fmaxf((oy + 0.5f) * sh - 0.5f, 0.0f);
            float fx = mode == 2 ? ox * sw : fmaxf((ox + 0.5f) * sw - 0.5f, 0.0f);
            const int y0 = min((int) fy, Hi - 1), x0 = min((int) fx, Wi - 1);
            const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
            const float ly = fy - y0, lx = fx - x0;
            float a[8], b[8], c[8], d[8], o[8];
            unpack8(*reinterpret_cast<const H8*>(xb + ((long long) y0 * Wi + x0) * ldx), a);
            unpack8(*reinterpret_cast<const H8*>(xb + ((long long) y0 * Wi + x1) * ldx), b);
            unpack8(*reinterpret_cast<const H8*>(xb + ((long long) y1 * Wi + x0) * ldx), c);
            unpack8(*reinterpret_cast<const H8*>(xb + ((long long) y1 * Wi + x1) * ldx), d);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                o[e] = (1.0f - ly) * ((1.0f - lx) * a[e] + lx * b[e]) + ly * ((1.0f - lx) * c[e] + lx * d[e]);
            out = pack8(o);
        }
        *reinterpret_cast<H8*>(y + p * ldy + yoff + c0) = out;
    }
}

// Detector input: uint8 HWC (BGR) -> fp16 NHWC with CP channels (zero padded), (x - mean) / std, optional channel swap.
// Replaces mmdet DetDataPreprocessor.  3 B read + 2*CP B written per pixel.
__global__ void __launch_bounds__(256) k_image_prep(const uint8_t* __restrict__ img, long long npix, float m0, float m1, float m2, float s0, float s1,
                                                    float s2, int swap_rb, int CP, __half* __restrict__ y) {
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < npix; i += (long long) gridDim.x * blockDim.x) {
        float c0 = (float) img[i * 3 + 0], c1 = (float) img[i * 3 + 1], c2 = (float) img[i * 3 + 2];
        if (swap_rb) { float t = c0; c0 = c2; c2 = t; }
        float o[8] = {(c0 - m0) / s0, (c1 - m1) / s1, (c2 - m2) / s2, 0.f, 0.f, 0.f, 0.f, 0.f};
        __half* yp = y + i * CP;
        *reinterpret_cast<H8*>(yp) = pack8(o);
        const H8 z = pack8({0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f});
        for (int c = 8; c < CP; c += 8) *reinterpret_cast<H8*>(yp + c) = z;
    }
}

}  // namespace

extern "C" int csb_dwconv_nhwc(const void* x, int ldx, int xoff, const float* w, const float* bias, const float* ln_gamma, const float* ln_beta,
                               float eps, int act, int N, int H, int W, int C, int K, void* y, int ldy, int yoff, void* stream) {
    CSB_REQUIRE(x && w && y, "null pointer");
    CSB_REQUIRE((K == 3 || K == 5 || K == 7) && C % 8 == 0 && C <= 2048, "K must be 3, 5 or 7; C a multiple of 8 up to 2048");
    CSB_REQUIRE(ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0, "channel strides/offsets must be multiples of 8");
    CSB_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "LayerNorm needs both gamma and beta");
    cudaStream_t st = (cudaStream_t) stream;
    const long long npix = (long long) N * H * W;
    const int grid = csb::wave_grid(npix * 32, 256, 8);
    const __half* xh = (const __half*) x;
    __half* yh = (__half*) y;
    if (K == 3) k_dwconv<3><<<grid, 256, 0, st>>>(xh, ldx, xoff, w, bias, ln_gamma, ln_beta, eps, act, N, H, W, C, yh, ldy, yoff);
    else if (K == 5) k_dwconv<5><<<grid, 256, 0, st>>>(xh, ldx, xoff, w, bias, ln_gamma, ln_beta, eps, act, N, H, W, C, yh, ldy, yoff);
    else k_dwconv<7><<<grid, 256, 0, st>>>(xh, ldx, xoff, w, bias, ln_gamma, ln_beta, eps, act, N, H, W, C, yh, ldy, yoff);
    return csb::launched("k_dwconv", st);
}

extern "C" int csb_layernorm_nhwc(const void* x, int ldx, int xoff, const float* gamma, const float* beta, float eps, long long npix, int C, void* y,
                                  int ldy, int yoff, void* stream) {
    CSB_REQUIRE(x && gamma && beta && y, "null pointer");
    CSB_REQUIRE(C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0, "C, strides and offsets must be multiples of 8");
    k_layernorm<<<csb::wave_grid(npix * 32, 256, 8), 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, gamma, beta, eps, npix, C, (__half*) y,
                                                                                     ldy, yoff);
    return csb::launched("k_layernorm", (cudaStream_t) stream);
}

extern "C" int csb_resample_nhwc(const void* x, int ldx, int xoff, int N, int Hi, int Wi, int C, int Ho, int Wo, int mode, void* y, int ldy, int yoff,
                                 void* stream) {
    CSB_REQUIRE(x && y, "null pointer");
    CSB_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0 && mode >= 0 && mode <= 2, "bad channel layout or mode");
    k_resample<<<csb::wave_grid((long long) N * Ho * Wo * (C / 8), 256, 8), 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, N, Hi, Wi, C, Ho, Wo,
                                                                                                             mode, (__half*) y, ldy, yoff);
    return csb::launched("k_resample", (cudaStream_t) stream);
}

extern "C" int csb_image_prep_nhwc(const uint8_t* img, long long npix, const float* mean3, const float* std3, int swap_rb, int CP, void* y, void* stream) {
    CSB_REQUIRE(img && mean3 && std3 && y, "null pointer");
    CSB_REQUIRE(CP % 8 == 0 && CP >= 8, "CP must be a multiple of 8");
    k_image_prep<<<csb::wave_grid(npix, 256, 8), 256, 0, (cudaStream_t) stream>>>(img, npix, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], swap_rb,
                                                                                 CP, (__half*) y);
    return csb::launched("k_image_prep", (cudaStream_t) stream);
}
