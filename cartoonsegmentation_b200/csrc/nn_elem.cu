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
#include <cstdlib>

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
// One warp owns T consecutive output pixels of a row and ALL channels (lane l -> channel vectors 8*(l + 32 j)), so the LayerNorm
// reduction is a warp shuffle.  For every filter row the warp loads the T+K-1 input vectors once into registers (converted to fp32
// once) and slides the K taps over them: (T+K-1)/T loads per output row-tap group instead of K -- the kernel is L1-bandwidth bound, this
// is the reuse that matters (measured: the untiled version spent 66% of the detector forward here).
template <int K, int T, int NV>
__global__ void __launch_bounds__(256) k_dwconv(const __half* __restrict__ x, int ldx, int xoff, const float* __restrict__ w,
                                                const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                float eps, int act, int N, int H, int W, int C, __half* __restrict__ y, int ldy, int yoff) {
    constexpr int R = K / 2;
    const int lane = threadIdx.x & 31;
    const int groups_w = (W + T - 1) / T;
    const long long ngroups = (long long) N * H * groups_w;
    for (long long gidx = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gidx < ngroups; gidx += (long long) gridDim.x * (blockDim.x >> 5)) {
        const int gx = (int) (gidx % groups_w), py = (int) ((gidx / groups_w) % H);
        const long long img = gidx / ((long long) groups_w * H);
        const int px0 = gx * T;
        float acc[T][NV][8];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c0 = 8 * (lane + 32 * j);
            float b8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) b8[e] = (bias && c0 < C) ? __ldg(bias + c0 + e) : 0.0f;
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[t][j][e] = b8[e];
        }
#pragma unroll 1
        for (int r = 0; r < K; ++r) {
            const int iy = py + r - R;
            if (iy < 0 || iy >= H) continue;
            const __half* xrow = x + ((img * H + iy) * W) * ldx + xoff;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c0 = 8 * (lane + 32 * j);
                if (c0 >= C) continue;
                float xin[T + K - 1][8];
#pragma unroll
                for (int s = 0; s < T + K - 1; ++s) {
                    const int ix = px0 + s - R;
                    if (ix >= 0 && ix < W) unpack8(*reinterpret_cast<const H8*>(xrow + (size_t) ix * ldx + c0), xin[s]);
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) xin[s][e] = 0.0f;
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float* wp = w + (size_t) (r * K + k) * C + c0;
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int t = 0; t < T; ++t)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[t][j][e] = fmaf(xin[t + k][e], wv[e], acc[t][j][e]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int px = px0 + t;
            if (px >= W) break;
            float mean = 0.0f, rstd = 1.0f;
            if (ln_g) {
                float s1 = 0.0f;
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (8 * (lane + 32 * j) < C)
#pragma unroll
                        for (int e = 0; e < 8; ++e) s1 += acc[t][j][e];
                mean = warp_sum(s1) / (float) C;
                float s2 = 0.0f;
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (8 * (lane + 32 * j) < C)
#pragma unroll
                        for (int e = 0; e < 8; ++e) { float d = acc[t][j][e] - mean; s2 += d * d; }
                rstd = rsqrtf(warp_sum(s2) / (float) C + eps);
            }
            __half* yp = y + (((img * H + py) * W) + px) * ldy + yoff;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c0 = 8 * (lane + 32 * j);
                if (c0 < C) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float v = acc[t][j][e];
                        if (ln_g) v = (v - mean) * rstd * __ldg(ln_g + c0 + e) + __ldg(ln_b + c0 + e);
                        o[e] = act_f(v, act);
                    }
                    *reinterpret_cast<H8*>(yp + c0) = pack8(o);
                }
            }
        }
    }
}

// LayerNorm over the channels of each pixel (LayerNorm2d / channels-last LN).  4*C bytes per pixel.
// One warp normalises PIX pixels per iteration: all PIX*NV 16-byte loads are issued before the first reduction (memory-level parallelism --
// the one-pixel-at-a-time version ran at ~1/7 of HBM speed), lanes own interleaved 8-channel vectors.
template <int NV, int PIX>
__global__ void __launch_bounds__(256) k_layernorm(const __half* __restrict__ x, int ldx, int xoff, const float* __restrict__ g,
                                                   const float* __restrict__ b, float eps, long long npix, int C, __half* __restrict__ y, int ldy, int yoff) {
    const int lane = threadIdx.x & 31;
    const long long ngroups = (npix + PIX - 1) / PIX;
    for (long long grp = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); grp < ngroups; grp += (long long) gridDim.x * (blockDim.x >> 5)) {
        H8 raw[PIX][NV];
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            const long long pix = grp * PIX + p;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c0 = 8 * (lane + 32 * j);
                if (pix < npix && c0 < C) raw[p][j] = *reinterpret_cast<const H8*>(x + pix * ldx + xoff + c0);
            }
        }
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            const long long pix = grp * PIX + p;
            if (pix >= npix) break;
            float v[NV][8];
            float s1 = 0.0f;
#pragma unroll
            for (int j = 0; j < NV; ++j)
                if (8 * (lane + 32 * j) < C) {
                    unpack8(raw[p][j], v[j]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) s1 += v[j][e];
                }
            const float mean = warp_sum(s1) / (float) C;
            float s2 = 0.0f;
#pragma unroll
            for (int j = 0; j < NV; ++j)
                if (8 * (lane + 32 * j) < C)
#pragma unroll
                    for (int e = 0; e < 8; ++e) { float d = v[j][e] - mean; s2 += d * d; }
            const float rstd = rsqrtf(warp_sum(s2) / (float) C + eps);
            __half* yp = y + pix * ldy + yoff;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c0 = 8 * (lane + 32 * j);
                if (c0 < C) {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c0)), g1 = __ldg(reinterpret_cast<const float4*>(g + c0 + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
                    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd * gg[e] + bb[e];
                    *reinterpret_cast<H8*>(yp + c0) = pack8(o);
                }
            }
        }
    }
}

// C <= 128 (the ConvNeXt stem / first downsample norms, the two largest LayerNorm launches of the detector): 16 lanes cover a pixel, so a warp
// normalises two pixels side by side instead of idling half of its lanes; PIX pixel pairs per iteration, reductions over 16 lanes (4 shuffles).
template <int PIX>
__global__ void __launch_bounds__(256) k_layernorm_h16(const __half* __restrict__ x, int ldx, int xoff, const float* __restrict__ g, const float* __restrict__ b, float eps,
                                                       long long npix, int C, __half* __restrict__ y, int ldy, int yoff) {
    const int lane = threadIdx.x & 31, sub = lane >> 4, c0 = 8 * (lane & 15);
    const bool on = c0 < C;
    const long long ngroups = (npix + 2 * PIX - 1) / (2 * PIX);
    float gg[8], bb[8];
    if (on) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c0)), g1 = __ldg(reinterpret_cast<const float4*>(g + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    }
    auto sum16 = [](float v) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    for (long long grp = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); grp < ngroups; grp += (long long) gridDim.x * (blockDim.x >> 5)) {
        H8 raw[PIX];
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            const long long pix = (grp * PIX + p) * 2 + sub;
            if (pix < npix && on) raw[p] = *reinterpret_cast<const H8*>(x + pix * ldx + xoff + c0);
        }
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            const long long pix = (grp * PIX + p) * 2 + sub;
            const bool live = pix < npix && on;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float s1 = 0.0f;
            if (live) {
                unpack8(raw[p], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) s1 += v[e];                      // same per-lane order as k_layernorm
            }
            const float mean = sum16(s1) / (float) C;
            float s2 = 0.0f;
            if (live)
#pragma unroll
                for (int e = 0; e < 8; ++e) { float d = v[e] - mean; s2 += d * d; }
            const float rstd = rsqrtf(sum16(s2) / (float) C + eps);
            if (live) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (v[e] - mean) * rstd * gg[e] + bb[e];
                *reinterpret_cast<H8*>(y + pix * ldy + yoff + c0) = pack8(o);
            }
        }
    }
}

static int launch_layernorm(const __half* x, int ldx, int xoff, const float* g, const float* b, float eps, long long npix, int C, __half* y, int ldy, int yoff,
                            cudaStream_t st) {
    static const int h16 = [] { const char* e = getenv("CSB_LN_H16"); return e ? atoi(e) : 1; }();
    if (h16 && C <= 128 && C % 8 == 0) {
        k_layernorm_h16<8><<<csb::wave_grid((npix + 15) / 16 * 32, 256, 8), 256, 0, st>>>(x, ldx, xoff, g, b, eps, npix, C, y, ldy, yoff);
        return csb::launched("k_layernorm", st);
    }
    const int nv = (C + 255) / 256;
#define CSB_LN(NV_, PIX_) k_layernorm<NV_, PIX_><<<csb::wave_grid((npix + PIX_ - 1) / PIX_ * 32, 256, 8), 256, 0, st>>>(x, ldx, xoff, g, b, eps, npix, C, y, ldy, yoff)
    if (nv == 1) CSB_LN(1, 8);
    else if (nv == 2) CSB_LN(2, 4);
    else if (nv <= 4) CSB_LN(4, 2);
    else CSB_LN(8, 1);
#undef CSB_LN
    return csb::launched("k_layernorm", st);
}

// Resample NHWC into a channel slice: mode 0 nearest (F.interpolate 'nearest': src = floor(dst * in/out)), 1 bilinear
// align_corners=False, 2 bilinear align_corners=True.  One thread per (pixel, 8-channel vector).
__global__ void __launch_bounds__(256) k_resample(const __half* __restrict__ x, int ldx, int xoff, int N, int Hi, int Wi, int C, int Ho, int Wo, int mode,
                                                  __half* __restrict__ y, int ldy, int yoff) {
    const int cv = C / 8;
    const long long total = (long long) N * Ho * Wo * cv;
    const float sh = mode == 2 ? (Ho > 1 ? (float) (Hi - 1) / (float) (Ho - 1) : 0.0f) : (float) Hi / (float) Ho;
    const float sw = mode == 2 ? (Wo > 1 ? (float) (Wi - 1) / (float) (Wo - 1) : 0.0f) : (float) Wi / (float) Wo;
    const unsigned ucv = (unsigned) cv, uWo = (unsigned) Wo, uHo = (unsigned) Ho;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        // index split in 32-bit arithmetic (the host guarantees N*Ho*Wo*cv < 2^32): 64-bit div/mod cost ~200 instructions per thread here
        const unsigned ui = (unsigned) i;
        const int c0 = (int) (ui % ucv) * 8;
        const unsigned up = ui / ucv;
        const long long p = up;
        const int ox = (int) (up % uWo), oy = (int) ((up / uWo) % uHo);
        const long long n = up / (uWo * uHo);
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

// Row-blocked form of k_resample: grid (x blocks, row groups of R, image).  A thread owns one (output column, 8-channel vector) and walks R output
// rows: the column part (x0, x1, lx, one 32-bit division) is computed once, the row part (y0, y1, ly) is block-uniform.  The generic kernel spends
// three 32-bit div/mods and 64-bit address products per 16 output bytes and ran at ~2 TB/s (ncu: 1.10 ms for the 2.3 GB 160^2 -> 320^2 x 256 launch of
// the LeReS decoder).  Same float expressions per output -> bit-identical results.
template <int MODE, int R>
__global__ void __launch_bounds__(256) k_resample_rows(const __half* __restrict__ x, int ldx, int xoff, int Hi, int Wi, int cv, int Ho, int Wo, float sh, float sw,
                                                       __half* __restrict__ y, int ldy, int yoff) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned) Wo * (unsigned) cv) return;
    const int ox = (int) (t / (unsigned) cv), c0 = (int) (t - (unsigned) ox * (unsigned) cv) * 8;
    const long long n = blockIdx.z;
    const __half* xb = x + n * Hi * Wi * ldx + xoff + c0;
    __half* yb = y + (n * Ho * Wo + ox) * ldy + yoff + c0;
    int x0, x1 = 0;
    float lx = 0.f;
    if (MODE == 0) {
        x0 = min((int) floorf(ox * sw), Wi - 1);
    } else {
        const float fx = MODE == 2 ? ox * sw : fmaxf((ox + 0.5f) * sw - 0.5f, 0.0f);
        x0 = min((int) fx, Wi - 1);
        x1 = min(x0 + 1, Wi - 1);
        lx = fx - x0;
    }
    const int oy0 = blockIdx.y * R;
    if (MODE == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int oy = oy0 + r;
            if (oy >= Ho) break;
            const int iy = min((int) floorf(oy * sh), Hi - 1);
            *reinterpret_cast<H8*>(yb + (long long) oy * Wo * ldy) = *reinterpret_cast<const H8*>(xb + ((long long) iy * Wi + x0) * ldx);
        }
        return;
    }
    // bilinear: G rows at a time, all 4 G corner loads issued before the first interpolation (bytes in flight, not instructions, bound this kernel)
    constexpr int G = 4;
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += G) {
        H8 raw[G][4];
        float lyv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int oy = min(oy0 + r0 + g, Ho - 1);                  // rows past the end repeat the last row's loads (never stored)
            const float fy = MODE == 2 ? oy * sh : fmaxf((oy + 0.5f) * sh - 0.5f, 0.0f);
            const int y0 = min((int) fy, Hi - 1), y1 = min(y0 + 1, Hi - 1);
            lyv[g] = fy - y0;
            raw[g][0] = *reinterpret_cast<const H8*>(xb + ((long long) y0 * Wi + x0) * ldx);
            raw[g][1] = *reinterpret_cast<const H8*>(xb + ((long long) y0 * Wi + x1) * ldx);
            raw[g][2] = *reinterpret_cast<const H8*>(xb + ((long long) y1 * Wi + x0) * ldx);
            raw[g][3] = *reinterpret_cast<const H8*>(xb + ((long long) y1 * Wi + x1) * ldx);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int oy = oy0 + r0 + g;
            if (oy >= Ho) break;
            const float ly = lyv[g];
            float a[8], b[8], c[8], d[8], o[8];
            unpack8(raw[g][0], a); unpack8(raw[g][1], b); unpack8(raw[g][2], c); unpack8(raw[g][3], d);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                o[e] = (1.0f - ly) * ((1.0f - lx) * a[e] + lx * b[e]) + ly * ((1.0f - lx) * c[e] + lx * d[e]);
            *reinterpret_cast<H8*>(yb + (long long) oy * Wo * ldy) = pack8(o);
        }
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

// Space-to-depth variant for a P x P stride-P stem conv (ConvNeXt's 4x4 s4 patchify): y [N, H/P, W/P, CP] with channel (r*P + s)*3 + c holding the
// normalised pixel (P*oy + r, P*ox + s), zero padded to CP -- the stem becomes a dense 1x1 GEMM with K = CP instead of P*P taps of 16 padded
// channels, and this kernel writes 2*CP/(P*P) B per input pixel instead of 32.  One thread per (output cell, row r of the patch).
__global__ void __launch_bounds__(256) k_image_prep_s2d(const uint8_t* __restrict__ img, int N, int H, int W, int P, float m0, float m1, float m2, float s0,
                                                        float s1, float s2, int swap_rb, int CP, __half* __restrict__ y) {
    const int Ho = H / P, Wo = W / P;
    const long long total = (long long) N * Ho * Wo * P;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int r = (int) (i % P);
        const long long cell = i / P;
        const int ox = (int) (cell % Wo), oy = (int) ((cell / Wo) % Ho);
        const long long n = cell / ((long long) Wo * Ho);
        const uint8_t* src = img + ((n * H + (long long) oy * P + r) * W + (long long) ox * P) * 3;
        __half* dst = y + cell * CP + r * P * 3;
        for (int e = 0; e < P * 3; ++e) {
            const int c = e % 3;
            const int cs = swap_rb ? 2 - c : c;
            dst[e] = __float2half_rn(((float) src[e - c + cs] - mean[c]) / sd[c]);
        }
        if (r == P - 1) {                                          // zero padding: 8-byte stores when the tail starts 8-byte aligned (P = 2: halves 12 .. CP)
            const int e0 = P * P * 3;
            if ((e0 & 3) == 0 && (CP & 3) == 0) {
                uint2* z = reinterpret_cast<uint2*>(y + cell * CP + e0);
                for (int e = 0; e < (CP - e0) / 4; ++e) z[e] = make_uint2(0u, 0u);
            } else {
                for (int e = e0; e < CP; ++e) y[cell * CP + e] = __float2half_rn(0.f);
            }
        }
    }
}

// The detector's case (P = 4, CP = 64, W % 4 == 0): one thread per output cell.  Each patch row is 12 bytes at a 4-byte aligned address (three 32-bit
// loads), the 48 + 16 halves of the cell leave as eight 16-byte stores; index arithmetic in 32 bits.  (The generic kernel above issues 2-byte stores
// and three 64-bit div/mods per thread: 0.73 ms per 32 frames against ~0.1 ms of HBM time.)
__global__ void __launch_bounds__(256) k_image_prep_s2d4(const uint8_t* __restrict__ img, unsigned ncell, int Ho, int Wo, int H, int W, float m0, float m1,
                                                         float m2, float s0, float s1, float s2, int swap_rb, __half* __restrict__ y) {
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    for (unsigned cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += gridDim.x * blockDim.x) {
        const unsigned ox = cell % (unsigned) Wo, t = cell / (unsigned) Wo, oy = t % (unsigned) Ho, n = t / (unsigned) Ho;
        const uint8_t* src = img + (((size_t) n * H + (size_t) oy * 4) * W + (size_t) ox * 4) * 3;
        float o[64];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(src + (size_t) r * W * 3);
            const uint32_t w0 = __ldg(row), w1 = __ldg(row + 1), w2 = __ldg(row + 2);
            uint8_t px[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) { px[k] = (uint8_t) (w0 >> (8 * k)); px[4 + k] = (uint8_t) (w1 >> (8 * k)); px[8 + k] = (uint8_t) (w2 >> (8 * k)); }
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const int c = e % 3, cs = swap_rb ? 2 - c : c;
                o[r * 12 + e] = ((float) px[e - c + cs] - mean[c]) / sd[c];
            }
        }
#pragma unroll
        for (int e = 48; e < 64; ++e) o[e] = 0.f;
        __half* dst = y + (size_t) cell * 64;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float v[8] = {o[8 * g], o[8 * g + 1], o[8 * g + 2], o[8 * g + 3], o[8 * g + 4], o[8 * g + 5], o[8 * g + 6], o[8 * g + 7]};
            *reinterpret_cast<H8*>(dst + 8 * g) = pack8(v);
        }
    }
}

// The LeReS case (P = 2, 2 x 2 space-to-depth of the 7x7 stride-2 stem): one thread per output cell; each patch row is 6 bytes at a 2-byte aligned
// address (three 16-bit loads), the 12 + (CP - 12) halves of the cell leave as CP / 8 16-byte stores.  NV = CP / 8 (2 or 8).
template <int NV>
__global__ void __launch_bounds__(256) k_image_prep_s2d2(const uint8_t* __restrict__ img, unsigned ncell, int Ho, int Wo, int H, int W, float m0, float m1,
                                                         float m2, float s0, float s1, float s2, int swap_rb, __half* __restrict__ y) {
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    for (unsigned cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += gridDim.x * blockDim.x) {
        const unsigned ox = cell % (unsigned) Wo, t = cell / (unsigned) Wo, oy = t % (unsigned) Ho, n = t / (unsigned) Ho;
        const uint8_t* src = img + (((size_t) n * H + (size_t) oy * 2) * W + (size_t) ox * 2) * 3;
        float o[16];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint16_t* row = reinterpret_cast<const uint16_t*>(src + (size_t) r * W * 3);
            const uint32_t w0 = __ldg(row), w1 = __ldg(row + 1), w2 = __ldg(row + 2);
            const uint8_t px[6] = {(uint8_t) w0, (uint8_t) (w0 >> 8), (uint8_t) w1, (uint8_t) (w1 >> 8), (uint8_t) w2, (uint8_t) (w2 >> 8)};
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                const int c = e % 3, cs = swap_rb ? 2 - c : c;
                o[r * 6 + e] = ((float) px[e - c + cs] - mean[c]) / sd[c];
            }
        }
#pragma unroll
        for (int e = 12; e < 16; ++e) o[e] = 0.f;
        __half* dst = y + (size_t) cell * (NV * 8);
        {
            const float v0[8] = {o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]}, v1[8] = {o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]};
            *reinterpret_cast<H8*>(dst) = pack8(v0);
            *reinterpret_cast<H8*>(dst + 8) = pack8(v1);
        }
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int g = 2; g < NV; ++g) *reinterpret_cast<H8*>(dst + 8 * g) = pack8(z);
    }
}

// MaxPool2d(K, stride, pad[, ceil_mode]) NHWC fp16 on channel slices: one thread per (output pixel, 8-channel vector).
__global__ void __launch_bounds__(256) k_maxpool(const __half* __restrict__ x, int ldx, int xoff, int N, int H, int W, int C, int K, int stride, int pad,
                                                 int Ho, int Wo, __half* __restrict__ y, int ldy, int yoff) {
    const int cv = C / 8;
    const long long total = (long long) N * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int c0 = (int) (i % cv) * 8;
        const long long p = i / cv;
        const int ox = (int) (p % Wo), oy = (int) ((p / Wo) % Ho);
        const long long n = p / ((long long) Wo * Ho);
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
        for (int r = 0; r < K; ++r) {
            const int iy = oy * stride - pad + r;
            if (iy < 0 || iy >= H) continue;
            for (int s2 = 0; s2 < K; ++s2) {
                const int ix = ox * stride - pad + s2;
                if (ix < 0 || ix >= W) continue;
                float f[8];
                unpack8(*reinterpret_cast<const H8*>(x + ((n * H + iy) * W + ix) * ldx + xoff + c0), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], f[e]);
            }
        }
        *reinterpret_cast<H8*>(y + p * ldy + yoff + c0) = pack8(m);
    }
}

__global__ void __launch_bounds__(256) k_add(const __half* __restrict__ a, int lda, int aoff, const __half* __restrict__ b, int ldb, int boff, long long npix,
                                             int C, __half* __restrict__ y, int ldy, int yoff) {
    const int cv = C / 8;
    const long long total = npix * cv;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int c0 = (int) (i % cv) * 8;
        const long long p = i / cv;
        float fa[8], fb[8];
        unpack8(*reinterpret_cast<const H8*>(a + p * lda + aoff + c0), fa);
        unpack8(*reinterpret_cast<const H8*>(b + p * ldb + boff + c0), fb);
#pragma unroll
        for (int e = 0; e < 8; ++e) fa[e] += fb[e];
        *reinterpret_cast<H8*>(y + p * ldy + yoff + c0) = pack8(fa);
    }
}

// y = PReLU(x) with per-channel slopes (the pre-activation of the GridNet blocks, pointcloud_inpainting.py:10-13).  4*C B/px.
__global__ void __launch_bounds__(256) k_prelu(const __half* __restrict__ x, int ldx, int xoff, const float* __restrict__ slope, long long npix, int C,
                                               __half* __restrict__ y, int ldy, int yoff) {
    const int cv = C / 8;
    const long long total = npix * cv;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int c0 = (int) (i % cv) * 8;
        const long long p = i / cv;
        float f[8];
        unpack8(*reinterpret_cast<const H8*>(x + p * ldx + xoff + c0), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * __ldg(slope + c0 + e);
        *reinterpret_cast<H8*>(y + p * ldy + yoff + c0) = pack8(f);
    }
}

__global__ void __launch_bounds__(256) k_resample_f32(const float* __restrict__ x, int N, int Hi, int Wi, int Ho, int Wo, int ac, float* __restrict__ y) {
    const long long total = (long long) N * Ho * Wo;
    const float sh = ac ? (Ho > 1 ? (float) (Hi - 1) / (float) (Ho - 1) : 0.f) : (float) Hi / (float) Ho;
    const float sw = ac ? (Wo > 1 ? (float) (Wi - 1) / (float) (Wo - 1) : 0.f) : (float) Wi / (float) Wo;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int ox = (int) (i % Wo), oy = (int) ((i / Wo) % Ho);
        const float* xb = x + (i / ((long long) Wo * Ho)) * Hi * Wi;
        const float fy = ac ? oy * sh : fmaxf((oy + 0.5f) * sh - 0.5f, 0.f), fx = ac ? ox * sw : fmaxf((ox + 0.5f) * sw - 0.5f, 0.f);
        const int y0 = min((int) fy, Hi - 1), x0 = min((int) fx, Wi - 1), y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
        const float ly = fy - y0, lx = fx - x0;
        y[i] = (1.f - ly) * ((1.f - lx) * xb[y0 * Wi + x0] + lx * xb[y0 * Wi + x1]) + ly * ((1.f - lx) * xb[y1 * Wi + x0] + lx * xb[y1 * Wi + x1]);
    }
}

}  // namespace

static int pool_out(int H, int K, int stride, int pad, int ceil_mode) {
    int o = ceil_mode ? (H + 2 * pad - K + stride - 1) / stride + 1 : (H + 2 * pad - K) / stride + 1;
    if (ceil_mode && (o - 1) * stride >= H + pad) --o;      // PyTorch: the last window must start inside the (left-padded) input
    return o;
}

extern "C" int csb_image_prep_s2d_nhwc(const uint8_t* img, int N, int H, int W, int P, const float* mean3, const float* std3, int swap_rb, int CP, void* y,
                                       void* stream) {
    CSB_REQUIRE(img && mean3 && std3 && y, "null pointer");
    CSB_REQUIRE(N > 0 && P > 0 && H % P == 0 && W % P == 0 && CP >= P * P * 3 && CP % 8 == 0, "H and W must be multiples of P, CP >= 3 P^2");
    const long long ncell = (long long) N * (H / P) * (W / P);
    if (P == 4 && CP == 64 && ncell < (1ll << 31) && ((uintptr_t) img & 3) == 0 && ((uintptr_t) y & 15) == 0)
        k_image_prep_s2d4<<<csb::wave_grid(ncell, 256, 8), 256, 0, (cudaStream_t) stream>>>(img, (unsigned) ncell, H / P, W / P, H, W, mean3[0], mean3[1], mean3[2],
                                                                                              std3[0], std3[1], std3[2], swap_rb, (__half*) y);
    else if (P == 2 && (CP == 16 || CP == 64) && W % 2 == 0 && ncell < (1ll << 31) && ((uintptr_t) img & 1) == 0 && ((uintptr_t) y & 15) == 0) {
        // (every patch row starts at an even byte offset: (row * W + 2 ox) * 3 with W even)
        if (CP == 16) k_image_prep_s2d2<2><<<csb::wave_grid(ncell, 256, 8), 256, 0, (cudaStream_t) stream>>>(img, (unsigned) ncell, H / P, W / P, H, W, mean3[0], mean3[1],
                                                                                                      mean3[2], std3[0], std3[1], std3[2], swap_rb, (__half*) y);
        else k_image_prep_s2d2<8><<<csb::wave_grid(ncell, 256, 8), 256, 0, (cudaStream_t) stream>>>(img, (unsigned) ncell, H / P, W / P, H, W, mean3[0], mean3[1],
                                                                                                     mean3[2], std3[0], std3[1], std3[2], swap_rb, (__half*) y);
    } else
        k_image_prep_s2d<<<csb::wave_grid((long long) N * (H / P) * (W / P) * P, 256, 8), 256, 0, (cudaStream_t) stream>>>(
            img, N, H, W, P, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], swap_rb, CP, (__half*) y);
    return csb::launched("k_image_prep", (cudaStream_t) stream);
}

extern "C" int csb_maxpool2d_nhwc(const void* x, int ldx, int xoff, int N, int H, int W, int C, int K, int stride, int pad, int ceil_mode, void* y, int ldy,
                                  int yoff, void* stream) {
    CSB_REQUIRE(x && y && C % 8 == 0 && N > 0 && H > 0 && W > 0 && K > 0 && stride > 0 && (ldx | xoff | ldy | yoff) % 8 == 0, "bad arguments");
    const int Ho = pool_out(H, K, stride, pad, ceil_mode), Wo = pool_out(W, K, stride, pad, ceil_mode);
    k_maxpool<<<csb::wave_grid((long long) N * Ho * Wo * (C / 8), 256, 8), 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, N, H, W, C, K, stride, pad,
                                                                                                            Ho, Wo, (__half*) y, ldy, yoff);
    return csb::launched("k_maxpool", (cudaStream_t) stream);
}

extern "C" int csb_maxpool_nhwc(const void* x, int N, int H, int W, int C, void* y, void* stream) {
    return csb_maxpool2d_nhwc(x, C, 0, N, H, W, C, 3, 2, 1, 0, y, C, 0, stream);
}

// ---- mmdet ChannelAttention (CSPNeXt CSPLayer): x * hardsigmoid(fc(global_avg_pool(x))) -- GAP and the per-(image, channel) scaling; the fc is a conv launch
namespace {
// partial sums: grid (chunks, N); thread t owns channel octets t, t+blockDim, ...; fp32 atomics into acc[N][C]
__global__ void k_gap_partial(const __half* __restrict__ x, int ldx, int xoff, int HW, int C, float* __restrict__ acc) {
    const int n = blockIdx.y, c8n = C / 8;
    const long long per = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long) blockIdx.x * per, p1 = min(p0 + per, (long long) HW);
    for (int c8 = threadIdx.x; c8 < c8n; c8 += blockDim.x) {
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const __half* base = x + ((long long) n * HW) * ldx + xoff + c8 * 8;
        for (long long p = p0; p < p1; ++p) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + p * ldx);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                s[2 * j] += f.x;
                s[2 * j + 1] += f.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&acc[(long long) n * C + c8 * 8 + j], s[j]);
    }
}
__global__ void k_gap_finish(const float* __restrict__ acc, long long n, float inv, __half* __restrict__ out) {
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) out[i] = __float2half_rn(acc[i] * inv);
}
__global__ void k_scale_channels(__half* __restrict__ x, int ldx, int xoff, long long HW, int C, long long total8, const __half* __restrict__ s) {
    const int c8n = C / 8;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long) gridDim.x * blockDim.x) {
        const int c8 = (int) (i % c8n);
        const long long pix = i / c8n;
        const long long n = pix / HW;
        uint4* px = reinterpret_cast<uint4*>(x + pix * ldx + xoff + c8 * 8);
        uint4 v = *px;
        const uint4 w = *reinterpret_cast<const uint4*>(s + n * C + c8 * 8);
        __half2* a = reinterpret_cast<__half2*>(&v);
        const __half2* b = reinterpret_cast<const __half2*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = __half22float2(a[j]), fb = __half22float2(b[j]);
            a[j] = __floats2half2_rn(fa.x * fb.x, fa.y * fb.y);
        }
        *px = v;
    }
}
}  // namespace

extern "C" int csb_gap_nhwc(const void* x, int ldx, int xoff, int N, int H, int W, int C, float* acc, void* y, void* stream) {
    CSB_REQUIRE(x && acc && y && N > 0 && H > 0 && W > 0 && C % 8 == 0 && (ldx | xoff) % 8 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(acc, 0, sizeof(float) * (size_t) N * C, st), "memset"));
    csb::memset_done(st);
    int chunks = (4 * csb::num_sms() + N - 1) / N;
    chunks = chunks > H * W ? H * W : (chunks < 1 ? 1 : chunks);
    k_gap_partial<<<dim3(chunks, N), 128, 0, st>>>((const __half*) x, ldx, xoff, H * W, C, acc);
    CSB_TRY(csb::launched("k_gap_partial", st));
    k_gap_finish<<<csb::wave_grid((long long) N * C, 256, 1), 256, 0, st>>>(acc, (long long) N * C, 1.0f / (float) (H * W), (__half*) y);
    return csb::launched("k_gap_finish", st);
}

extern "C" int csb_scale_channels_nhwc(void* x, int ldx, int xoff, int N, int H, int W, int C, const void* scale, void* stream) {
    CSB_REQUIRE(x && scale && N > 0 && H > 0 && W > 0 && C % 8 == 0 && (ldx | xoff) % 8 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    const long long total8 = (long long) N * H * W * (C / 8);
    k_scale_channels<<<csb::wave_grid(total8, 256, 8), 256, 0, st>>>((__half*) x, ldx, xoff, (long long) H * W, C, total8, (const __half*) scale);
    return csb::launched("k_scale_channels", st);
}

extern "C" int csb_add_nhwc(const void* a, int lda, int aoff, const void* b, int ldb, int boff, long long npix, int C, void* y, int ldy, int yoff, void* stream) {
    CSB_REQUIRE(a && b && y && C % 8 == 0 && (lda | aoff | ldb | boff | ldy | yoff) % 8 == 0, "bad arguments");
    k_add<<<csb::wave_grid(npix * (C / 8), 256, 8), 256, 0, (cudaStream_t) stream>>>((const __half*) a, lda, aoff, (const __half*) b, ldb, boff, npix, C, (__half*) y, ldy, yoff);
    return csb::launched("k_add", (cudaStream_t) stream);
}

extern "C" int csb_prelu_nhwc(const void* x, int ldx, int xoff, const float* slope, long long npix, int C, void* y, int ldy, int yoff, void* stream) {
    CSB_REQUIRE(x && y && slope && C % 8 == 0 && (ldx | xoff | ldy | yoff) % 8 == 0, "bad arguments");
    k_prelu<<<csb::wave_grid(npix * (C / 8), 256, 8), 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, slope, npix, C, (__half*) y, ldy, yoff);
    return csb::launched("k_prelu", (cudaStream_t) stream);
}

extern "C" int csb_resample_f32(const float* x, int N, int Hi, int Wi, int Ho, int Wo, int align_corners, float* y, void* stream) {
    CSB_REQUIRE(x && y && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bad arguments");
    k_resample_f32<<<csb::wave_grid((long long) N * Ho * Wo, 256, 8), 256, 0, (cudaStream_t) stream>>>(x, N, Hi, Wi, Ho, Wo, align_corners, y);
    return csb::launched("k_resample_f32", (cudaStream_t) stream);
}

// Spatially tiled depthwise KxK: each thread owns 2 channels (one half2) and a 2 x 8 block of output pixels; a warp owns 64 consecutive channels of
// one tile (128 B coalesced rows).  A CTA works on ONE 64-channel chunk (blockIdx.y): its K*K x 64 filter taps sit in shared memory (conflict-free
// LDS.64) and are reused by every tile the CTA walks.  The kernel is L1/shared-bandwidth bound unless operands are reused from registers, so the
// loop runs over filter ROWS with both output rows' input rows resident (fp32, converted once): each tap is read once per tile and feeds 32 FMAs, the
// two input rows slide down by one per iteration (one new row is fetched -- prefetched as half2 during the previous iteration's FMAs -- and the other
// is kept), i.e. per 16 x 2 outputs: (2+K-1) x (8+K-1) half2 loads + K*K LDS.64 for 32*K*K FMAs.  The row loop stays rolled (compact code).
// No LayerNorm here (a pixel's channels are spread over several CTAs): the ConvNeXt block runs this kernel followed by k_layernorm in place.
// sm_100 packed fp32 FMA: d.xy = a.xy * b.xy + c.xy in one issue slot.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

constexpr int kDwThreads = 128;
// STATS: also emit, per pixel and per 64-channel chunk, (sum, sum of squares) of the fp16-rounded outputs -> stats[pixel][C/64][2] fp32.  The
// following C -> 4C GEMM applies the LayerNorm of the ConvNeXt block in its epilogue from these (csb_conv2d_ln_nhwc), so the separate
// LayerNorm pass over the activations (read + write of 2 C B/px) disappears.
// LDC > 0: the channel stride of x and y is the compile-time constant LDC (dense tensors, ldx == ldy == C): the (2 + K - 1) x (8 + K - 1) halo loads and
// the 16 stores of a tile then address [base + immediate] instead of spending 3-4 integer instructions each on `ix * ldx` -- SASS of the generic build:
// 224 FFMA2 in a 703-instruction filter-row loop, i.e. the kernel was issue-bound by address arithmetic at 33 % of the FFMA peak.
template <int K, int ACT, bool STATS = false, int LDC = 0>       // ACT: the fused activation, compiled in (CSB_ACT_NONE / CSB_ACT_SILU), or -1 = decided at run time
__global__ void __launch_bounds__(kDwThreads, 3) k_dwconv_tile(const __half* __restrict__ x, int ldx_, int xoff, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int act, int N, int H, int W, int C_, __half* __restrict__ y, int ldy_,
                                                        int yoff, float* __restrict__ stats = nullptr) {
    const int ldx = LDC > 0 ? LDC : ldx_, ldy = LDC > 0 ? LDC : ldy_, C = LDC > 0 ? LDC : C_;
    constexpr int R = K / 2, TSY = 2, TSX = 8, INX = TSX + K - 1;
    __shared__ float2 wsm[K * K * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    const int c0 = chunk * 64 + lane * 2;
    for (int i = threadIdx.x; i < K * K * 32; i += blockDim.x) wsm[i] = __ldg(reinterpret_cast<const float2*>(w + (size_t) (i / 32) * C + chunk * 64 + (i % 32) * 2));
    __syncthreads();
    const int tiles_x = (W + TSX - 1) / TSX, tiles_y = (H + TSY - 1) / TSY;
    const long long ntiles = (long long) N * tiles_y * tiles_x;
    const float2 b2 = bias ? __ldg(reinterpret_cast<const float2*>(bias + c0)) : make_float2(0.f, 0.f);
    for (long long task = (long long) blockIdx.x * (kDwThreads / 32) + warp; task < ntiles; task += (long long) gridDim.x * (kDwThreads / 32)) {
        const int tx = (int) (task % tiles_x), ty = (int) ((task / tiles_x) % tiles_y);
        const long long img = task / ((long long) tiles_x * tiles_y);
        const int ox0 = tx * TSX, oy0 = ty * TSY;
        const __half* xb = x + (img * H * W) * ldx + xoff + c0;
        const bool interior = ox0 - R >= 0 && ox0 + TSX + R <= W && oy0 - R >= 0 && oy0 + TSY + R <= H;       // warp-uniform
        unsigned colmask = 0;
#pragma unroll
        for (int ix = 0; ix < INX; ++ix) colmask |= ((ox0 + ix - R) >= 0 && (ox0 + ix - R) < W) ? (1u << ix) : 0u;
        auto load_row = [&](int iy, __half2 (&row)[INX]) {              // input row iy (0 .. TSY+K-2) of the tile's halo window
            const int gy = oy0 + iy - R;
            const __half* xrow = xb + ((long long) gy * W + (ox0 - R)) * ldx;
            if (interior) {
#pragma unroll
                for (int ix = 0; ix < INX; ++ix) row[ix] = *reinterpret_cast<const __half2*>(xrow + (size_t) ix * ldx);
            } else {
                const unsigned m = (gy >= 0 && gy < H) ? colmask : 0u;
#pragma unroll
                for (int ix = 0; ix < INX; ++ix)
                    row[ix] = ((m >> ix) & 1u) ? *reinterpret_cast<const __half2*>(xrow + (long long) ix * ldx) : __floats2half2_rn(0.f, 0.f);
            }
        };
        float2 acc0[TSX], acc1[TSX], xa[INX], xc[INX];                  // xa / xc: input rows feeding output rows 0 / 1 at the current filter row
        __half2 nxt[INX];
#pragma unroll
        for (int j = 0; j < TSX; ++j) acc0[j] = acc1[j] = b2;
        load_row(0, nxt);
#pragma unroll
        for (int ix = 0; ix < INX; ++ix) xa[ix] = __half22float2(nxt[ix]);
        load_row(1, nxt);
#pragma unroll
        for (int ix = 0; ix < INX; ++ix) xc[ix] = __half22float2(nxt[ix]);
        load_row(2, nxt);
        // filter row r: output row 0 reads input row r (held in `top`), output row 1 reads input row r+1 (`bot`)
        auto taps = [&](int r, const float2 (&top)[INX], const float2 (&bot)[INX]) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float2 wv = wsm[(r * K + k) * 32 + lane];
#pragma unroll
                for (int t = 0; t < TSX; ++t) {                       // the channel pair is one packed FFMA2
                    acc0[t] = ffma2(top[t + k], wv, acc0[t]);
                    acc1[t] = ffma2(bot[t + k], wv, acc1[t]);
                }
            }
        };
        // the two-row window slides down one row per filter row; unrolled by two so the buffers swap roles instead of being copied
#pragma unroll 1
        for (int r = 0; r + 1 < K; r += 2) {
            taps(r, xa, xc);
#pragma unroll
            for (int ix = 0; ix < INX; ++ix) xa[ix] = __half22float2(nxt[ix]);          // input row r+2
            load_row(r + 3, nxt);                                                       // r+3 <= K always holds here
            taps(r + 1, xc, xa);
#pragma unroll
            for (int ix = 0; ix < INX; ++ix) xc[ix] = __half22float2(nxt[ix]);          // input row r+3
            if (r + 4 <= K) load_row(r + 4, nxt);
        }
        taps(K - 1, xa, xc);                                                            // K is odd: rows K-1 / K sit in xa / xc
        if (ACT != CSB_ACT_NONE) {
            const int a = ACT < 0 ? act : ACT;
#pragma unroll
            for (int j = 0; j < TSX; ++j) {
                acc0[j].x = act_f(acc0[j].x, a), acc0[j].y = act_f(acc0[j].y, a);
                acc1[j].x = act_f(acc1[j].x, a), acc1[j].y = act_f(acc1[j].y, a);
            }
        }
        __half* yb = y + ((img * H + oy0) * W + ox0) * ldy + yoff + c0;
        const bool row1 = oy0 + 1 < H;
        if constexpr (STATS) {
            float v[32];                                                // [0,16): per-pixel sum over this lane's 2 channels, [16,32): sum of squares
#pragma unroll
            for (int j = 0; j < TSX; ++j) {
                const __half2 h0 = __floats2half2_rn(acc0[j].x, acc0[j].y), h1 = __floats2half2_rn(acc1[j].x, acc1[j].y);
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                v[j] = f0.x + f0.y; v[16 + j] = fmaf(f0.x, f0.x, f0.y * f0.y);
                v[8 + j] = f1.x + f1.y; v[24 + j] = fmaf(f1.x, f1.x, f1.y * f1.y);
                if (ox0 + j < W) {
                    *reinterpret_cast<__half2*>(yb + (size_t) j * ldy) = h0;
                    if (row1) *reinterpret_cast<__half2*>(yb + ((size_t) W + j) * ldy) = h1;
                }
            }
            // transposed butterfly: 32 values x 32 lanes -> lane l ends with the warp total of value l (31 shuffles instead of 160)
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int m = 16 >> s;                                  // lane bit examined = number of values kept
                const bool up = (lane & m) != 0;
#pragma unroll
                for (int i = 0; i < m; ++i) {
                    const float send = up ? v[i] : v[i + m], keep = up ? v[i + m] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
            }
            const int pl = lane & 15, py = oy0 + (pl >> 3), px = ox0 + (pl & 7);
            if (py < H && px < W) stats[((((long long) img * H + py) * W + px) * (C >> 6) + chunk) * 2 + (lane >> 4)] = v[0];
        } else {
#pragma unroll
            for (int j = 0; j < TSX; ++j) {
                if (ox0 + j >= W) break;
                *reinterpret_cast<__half2*>(yb + (size_t) j * ldy) = __floats2half2_rn(acc0[j].x, acc0[j].y);
                if (row1) *reinterpret_cast<__half2*>(yb + ((size_t) W + j) * ldy) = __floats2half2_rn(acc1[j].x, acc1[j].y);
            }
        }
    }
}

template <int K>
static int launch_dwconv(const __half* xh, int ldx, int xoff, const float* w, const float* bias, const float* g, const float* b, float eps, int act, int N, int H,
                         int W, int C, __half* yh, int ldy, int yoff, cudaStream_t st) {
    const int nv = (C + 255) / 256;
#define CSB_DW(T_, NV_)                                                                                                                        \
    do {                                                                                                                                       \
        const long long groups = (long long) N * H * ((W + T_ - 1) / T_);                                                                      \
        k_dwconv<K, T_, NV_><<<csb::wave_grid(groups * 32, 256, 8), 256, 0, st>>>(xh, ldx, xoff, w, bias, g, b, eps, act, N, H, W, C, yh, ldy, yoff); \
    } while (0)
    if (nv == 1) CSB_DW(4, 1);
    else if (nv == 2) CSB_DW(4, 2);
    else if (nv <= 4) CSB_DW(2, 4);
    else CSB_DW(1, 8);
#undef CSB_DW
    return csb::launched("k_dwconv", st);
}

// csrc/dw_halo.cu: the TMA halo-tile kernel; -1000 = shape outside its domain (take the register-tiled path below)
int csb_dwconv_halo_try(const void* x, int ldx, int xoff, const float* w, const float* bias, int act, int N, int H, int W, int C, int K, void* y, int ldy, int yoff,
                        float* stats, cudaStream_t st);

extern "C" int csb_dwconv_nhwc(const void* x, int ldx, int xoff, const float* w, const float* bias, const float* ln_gamma, const float* ln_beta,
                               float eps, int act, int N, int H, int W, int C, int K, void* y, int ldy, int yoff, void* stream) {
    CSB_REQUIRE(x && w && y, "null pointer");
    CSB_REQUIRE((K == 3 || K == 5 || K == 7) && C % 8 == 0 && C <= 2048, "K must be 3, 5 or 7; C a multiple of 8 up to 2048");
    CSB_REQUIRE(ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0, "channel strides/offsets must be multiples of 8");
    CSB_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "LayerNorm needs both gamma and beta");
    cudaStream_t st = (cudaStream_t) stream;
    const __half* xh = (const __half*) x;
    __half* yh = (__half*) y;
    if ((K == 5 || K == 7) && C % 64 == 0 && (ldx | xoff | ldy | yoff) % 2 == 0) {
        // tiled path: depthwise conv (+ bias, + activation when there is no LayerNorm), then LayerNorm in place
        {
            const int hs = csb_dwconv_halo_try(x, ldx, xoff, w, bias, ln_gamma ? CSB_ACT_NONE : act, N, H, W, C, K, y, ldy, yoff, nullptr, st);
            if (hs != -1000) {
                CSB_TRY(hs);
                if (!ln_gamma) return CSB_OK;
                CSB_REQUIRE(act == CSB_ACT_NONE, "LayerNorm followed by an activation is not used on this path");
                return launch_layernorm(yh, ldy, yoff, ln_gamma, ln_beta, eps, (long long) N * H * W, C, yh, ldy, yoff, st);
            }
        }
        static const int ctas = [] { const char* e = getenv("CSB_DW_CTAS"); return e ? atoi(e) : 3; }();        // CTAs per SM (tuning knob)
        const long long ntiles = (long long) N * ((H + 1) / 2) * ((W + 7) / 8);
        const int chunks = C / 64;
        int gx = ctas * csb::num_sms() / chunks;                           // at most ONE wave of resident CTAs (floor), each walking many tiles of its chunk
        const long long need = (ntiles + kDwThreads / 32 - 1) / (kDwThreads / 32);
        gx = gx > need ? (int) need : gx;
        gx = gx < 1 ? 1 : gx;
        const dim3 grid(gx, chunks);
        const int a = ln_gamma ? CSB_ACT_NONE : act;
#define CSB_DW_LAUNCH(KK, AA) k_dwconv_tile<KK, AA><<<grid, kDwThreads, 0, st>>>(xh, ldx, xoff, w, bias, a, N, H, W, C, yh, ldy, yoff)
        static const bool const_ld = [] { const char* e = getenv("CSB_DW_CONST_LD"); return !e || atoi(e) != 0; }();
        if (K == 5 && a == CSB_ACT_SILU && const_ld && C == 128 && ldx == C && ldy == C)       // CSPNeXtPAFPN blocks: dense 128-channel tensors
            k_dwconv_tile<5, CSB_ACT_SILU, false, 128><<<grid, kDwThreads, 0, st>>>(xh, ldx, xoff, w, bias, a, N, H, W, C, yh, ldy, yoff);
        else if (K == 5) { if (a == CSB_ACT_NONE) CSB_DW_LAUNCH(5, CSB_ACT_NONE); else if (a == CSB_ACT_SILU) CSB_DW_LAUNCH(5, CSB_ACT_SILU); else CSB_DW_LAUNCH(5, -1); }
        else { if (a == CSB_ACT_NONE) CSB_DW_LAUNCH(7, CSB_ACT_NONE); else if (a == CSB_ACT_SILU) CSB_DW_LAUNCH(7, CSB_ACT_SILU); else CSB_DW_LAUNCH(7, -1); }
#undef CSB_DW_LAUNCH
        CSB_TRY(csb::launched("k_dwconv_tile", st));
        if (!ln_gamma) return CSB_OK;
        CSB_REQUIRE(act == CSB_ACT_NONE, "LayerNorm followed by an activation is not used on this path");
        return launch_layernorm(yh, ldy, yoff, ln_gamma, ln_beta, eps, (long long) N * H * W, C, yh, ldy, yoff, st);
    }
    if (K == 3) return launch_dwconv<3>(xh, ldx, xoff, w, bias, ln_gamma, ln_beta, eps, act, N, H, W, C, yh, ldy, yoff, st);
    if (K == 5) return launch_dwconv<5>(xh, ldx, xoff, w, bias, ln_gamma, ln_beta, eps, act, N, H, W, C, yh, ldy, yoff, st);
    return launch_dwconv<7>(xh, ldx, xoff, w, bias, ln_gamma, ln_beta, eps, act, N, H, W, C, yh, ldy, yoff, st);
}

extern "C" int csb_dwconv_stats_nhwc(const void* x, int ldx, int xoff, const float* w, const float* bias, int N, int H, int W, int C, int K, void* y, int ldy,
                                     int yoff, float* stats, void* stream) {
    CSB_REQUIRE(x && w && y && stats, "null pointer");
    CSB_REQUIRE(K == 7 && C % 64 == 0 && C <= 2048, "K must be 7 and C a multiple of 64 (the ConvNeXt block)");
    CSB_REQUIRE(ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0, "channel strides/offsets must be multiples of 8");
    cudaStream_t st = (cudaStream_t) stream;
    {
        const int hs = csb_dwconv_halo_try(x, ldx, xoff, w, bias, CSB_ACT_NONE, N, H, W, C, K, y, ldy, yoff, stats, st);
        if (hs != -1000) return hs;
    }
    static const int ctas = [] { const char* e = getenv("CSB_DW_CTAS"); return e ? atoi(e) : 3; }();
    const long long ntiles = (long long) N * ((H + 1) / 2) * ((W + 7) / 8);
    const int chunks = C / 64;
    int gx = ctas * csb::num_sms() / chunks;
    const long long need = (ntiles + kDwThreads / 32 - 1) / (kDwThreads / 32);
    gx = gx > need ? (int) need : gx;
    gx = gx < 1 ? 1 : gx;
#define CSB_DWS(LDC_) k_dwconv_tile<7, CSB_ACT_NONE, true, LDC_><<<dim3(gx, chunks), kDwThreads, 0, st>>>((const __half*) x, ldx, xoff, w, bias, CSB_ACT_NONE, N, H, W, C, \
                                                                                                      (__half*) y, ldy, yoff, stats)
    static const bool const_ld = [] { const char* e = getenv("CSB_DW_CONST_LD"); return !e || atoi(e) != 0; }();
    const bool dense = const_ld && ldx == C && ldy == C;            // the ConvNeXt stages: compile-time channel stride
    if (dense && C == 128) CSB_DWS(128);
    else if (dense && C == 256) CSB_DWS(256);
    else if (dense && C == 512) CSB_DWS(512);
    else if (dense && C == 1024) CSB_DWS(1024);
    else CSB_DWS(0);
#undef CSB_DWS
    return csb::launched("k_dwconv_tile", st);
}

extern "C" int csb_layernorm_nhwc(const void* x, int ldx, int xoff, const float* gamma, const float* beta, float eps, long long npix, int C, void* y,
                                  int ldy, int yoff, void* stream) {
    CSB_REQUIRE(x && gamma && beta && y, "null pointer");
    CSB_REQUIRE(C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0, "C, strides and offsets must be multiples of 8");
    return launch_layernorm((const __half*) x, ldx, xoff, gamma, beta, eps, npix, C, (__half*) y, ldy, yoff, (cudaStream_t) stream);
}

extern "C" int csb_resample_nhwc(const void* x, int ldx, int xoff, int N, int Hi, int Wi, int C, int Ho, int Wo, int mode, void* y, int ldy, int yoff,
                                 void* stream) {
    CSB_REQUIRE(x && y, "null pointer");
    CSB_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && xoff % 8 == 0 && ldy % 8 == 0 && yoff % 8 == 0 && mode >= 0 && mode <= 2, "bad channel layout or mode");
    CSB_REQUIRE((long long) N * Ho * Wo * (C / 8) < (1ll << 32), "output too large for the 32-bit index split (split the batch)");
    static const int rows_mode = [] { const char* e = getenv("CSB_RESAMPLE_ROWS"); return e ? atoi(e) : 1; }();
    const int cv = C / 8;
    if (rows_mode && N <= 65535 && (long long) Wo * cv < (1ll << 31) && Ho >= 8) {
        constexpr int R = 8;
        const float sh = mode == 2 ? (Ho > 1 ? (float) (Hi - 1) / (float) (Ho - 1) : 0.0f) : (float) Hi / (float) Ho;       // as in k_resample
        const float sw = mode == 2 ? (Wo > 1 ? (float) (Wi - 1) / (float) (Wo - 1) : 0.0f) : (float) Wi / (float) Wo;
        const dim3 grid((unsigned) (((long long) Wo * cv + 255) / 256), (unsigned) ((Ho + R - 1) / R), (unsigned) N);
        if (mode == 0) k_resample_rows<0, R><<<grid, 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, Hi, Wi, cv, Ho, Wo, sh, sw, (__half*) y, ldy, yoff);
        else if (mode == 1) k_resample_rows<1, R><<<grid, 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, Hi, Wi, cv, Ho, Wo, sh, sw, (__half*) y, ldy, yoff);
        else k_resample_rows<2, R><<<grid, 256, 0, (cudaStream_t) stream>>>((const __half*) x, ldx, xoff, Hi, Wi, cv, Ho, Wo, sh, sw, (__half*) y, ldy, yoff);
        return csb::launched("k_resample", (cudaStream_t) stream);
    }
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
