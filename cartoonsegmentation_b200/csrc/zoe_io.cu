// ZoeDepth inference wrapper on the device:
//   csb_zoe_prep       `_infer_with_pad_aug` reflect padding (zoedepth/models/depth_model.py:81-87) + `PrepForMidas` (base_models/midas.py:164-186:
//                      bilinear align_corners=True resize to the 32-multiple net size, Normalize(0.5, 0.5)) + the horizontally flipped twin of
//                      `infer_with_flip_aug` (depth_model.py:110-111), written directly as 16 x 16 patch rows for the BEiT patch embedding
//   csb_zoe_finish     bicubic (A = -0.75, align_corners=False) resize of both metric-depth maps back to the padded size (depth_model.py:89-90), crop,
//                      un-flip, average (depth_model.py:112)
//   csb_zoe_disparity  `_depth_est_zoe` tail (anime_3dkenburns/kenburns_effect.py:812-818): zeros -> smallest positive depth, focal*baseline/(d+1e-5),
//                      nan_to_num(0, 0, 0)
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ int reflect_idx(int p, int pad, int n) {          // F.pad(mode='reflect'): border not repeated
    int o = p - pad;
    if (o < 0) o = -o;
    if (o >= n) o = 2 * (n - 1) - o;
    return o;
}

// out [nb][Hn/16][Wn/16][768] fp16, channel = (r*16 + s)*3 + c; batch entry 1 (if nb == 2) is the flipped image.
// img [H,W,3] u8 (the reference feeds BGR/255 without reordering -- kenburns_effect.py:813 quirk, kept).
__global__ void k_zoe_prep(const uint8_t* __restrict__ img, int H, int W, int ph, int pw, int Hn, int Wn, int nb, __half* __restrict__ out) {
    const int Hp = H + 2 * ph, Wp = W + 2 * pw;
    const float sh = Hn > 1 ? (float) (Hp - 1) / (float) (Hn - 1) : 0.f, sw = Wn > 1 ? (float) (Wp - 1) / (float) (Wn - 1) : 0.f;
    const long long n = (long long) nb * Hn * Wn;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % Wn), y = (int) ((i / Wn) % Hn), f = (int) (i / ((long long) Wn * Hn));
        const float fy = y * sh, fx = x * sw;                                // align_corners=True source index
        const int y0 = (int) fy, x0 = (int) fx;
        const int y1 = y0 + (y0 < Hp - 1 ? 1 : 0), x1 = x0 + (x0 < Wp - 1 ? 1 : 0);
        const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
        // the flipped image is flip(padded): column px of it is column Wp-1-px of the padded original
        const int cx0 = f ? Wp - 1 - x0 : x0, cx1 = f ? Wp - 1 - x1 : x1;
        const int oy0 = reflect_idx(y0, ph, H), oy1 = reflect_idx(y1, ph, H), ox0 = reflect_idx(cx0, pw, W), ox1 = reflect_idx(cx1, pw, W);
        const uint8_t* p00 = img + ((long long) oy0 * W + ox0) * 3;
        const uint8_t* p01 = img + ((long long) oy0 * W + ox1) * 3;
        const uint8_t* p10 = img + ((long long) oy1 * W + ox0) * 3;
        const uint8_t* p11 = img + ((long long) oy1 * W + ox1) * 3;
        __half* o = out + ((((long long) f * (Hn / 16) + y / 16) * (Wn / 16) + x / 16) * 256 + (y % 16) * 16 + (x % 16)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = __fdiv_rn((float) p00[c], 255.f), b = __fdiv_rn((float) p01[c], 255.f), cc = __fdiv_rn((float) p10[c], 255.f),
                        d = __fdiv_rn((float) p11[c], 255.f);
            const float v = hy * (hx * a + lx * b) + ly * (hx * cc + lx * d);   // upsample_bilinear2d
            o[c] = __float2half_rn(__fdiv_rn(v - 0.5f, 0.5f));
        }
    }
}

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }            // |x| <= 1
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }      // 1 < |x| < 2

__device__ __forceinline__ float bicubic_at(const float* __restrict__ src, int Hi, int Wi, float sy, float sx) {
    const float A = -0.75f;
    const float fy = floorf(sy), fx = floorf(sx);
    const int iy = (int) fy, ix = (int) fx;
    const float ty = sy - fy, tx = sx - fx;
    float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
    float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int yy = min(max(iy - 1 + j, 0), Hi - 1);
        float row = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) row += src[(long long) yy * Wi + min(max(ix - 1 + k, 0), Wi - 1)] * wx[k];
        acc += row * wy[j];
    }
    return acc;
}

// d [nb][Hn][Wn] fp32 (metric depth of the net; entry 1 = flipped input) -> out [H][W]
__global__ void k_zoe_finish(const float* __restrict__ d, int nb, int Hn, int Wn, int H, int W, int ph, int pw, float* __restrict__ out) {
    const int Hp = H + 2 * ph, Wp = W + 2 * pw;
    const float sh = (float) Hn / (float) Hp, sw = (float) Wn / (float) Wp;
    const bool same = Hn == Hp && Wn == Wp;                                   // depth_model.py:89: interpolate only if the shapes differ
    const long long n = (long long) H * W;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % W), y = (int) (i / W);
        const int py = y + ph, px = x + pw;
        float v;
        if (same) v = d[(long long) py * Wn + px];
        else v = bicubic_at(d, Hn, Wn, sh * (py + 0.5f) - 0.5f, sw * (px + 0.5f) - 0.5f);
        if (nb == 2) {
            const int qx = Wp - 1 - px;                                       // torch.flip(out_flip, dims=[3])
            float u;
            if (same) u = d[(long long) Hn * Wn + (long long) py * Wn + qx];
            else u = bicubic_at(d + (long long) Hn * Wn, Hn, Wn, sh * (py + 0.5f) - 0.5f, sw * (qx + 0.5f) - 0.5f);
            v = (v + u) / 2.f;
        }
        out[i] = v;
    }
}

__global__ void k_min_positive(const float* __restrict__ d, long long n, unsigned* __restrict__ slot) {
    unsigned m = 0x7f800000u;                                                 // +inf; positive floats order like their bit patterns
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float v = d[i];
        if (v > 0.f) m = min(m, __float_as_uint(v));
    }
    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(slot, m);
}

__global__ void k_zoe_disparity(const float* __restrict__ d, long long n, const unsigned* __restrict__ slot, float fb, float* __restrict__ out) {
    const float mn = __uint_as_float(*slot);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        float v = d[i];
        if (v == 0.f) v = mn;                                                 // depth[depth == 0] = depth[depth > 0].min()
        float r = __fdiv_rn(fb, v + 0.00001f);
        if (r != r || isinf(r)) r = 0.f;                                      // nan_to_num_(0, 0, 0)
        out[i] = r;
    }
}

}  // namespace

extern "C" int csb_zoe_prep(const uint8_t* img, int H, int W, int pad_h, int pad_w, int Hn, int Wn, int flip_aug, void* patches, void* stream) {
    CSB_REQUIRE(img && patches, "null pointer");
    CSB_REQUIRE(H > 1 && W > 1 && pad_h >= 0 && pad_w >= 0 && pad_h < H && pad_w < W, "reflect padding must be smaller than the image");
    CSB_REQUIRE(Hn > 0 && Wn > 0 && Hn % 16 == 0 && Wn % 16 == 0, "net size must be a multiple of the 16 x 16 patch");
    cudaStream_t st = (cudaStream_t) stream;
    const int nb = flip_aug ? 2 : 1;
    k_zoe_prep<<<csb::wave_grid((long long) nb * Hn * Wn, 256, 4), 256, 0, st>>>(img, H, W, pad_h, pad_w, Hn, Wn, nb, (__half*) patches);
    return csb::launched("k_zoe_prep", st);
}

extern "C" int csb_zoe_finish(const float* depth_net, int flip_aug, int Hn, int Wn, int H, int W, int pad_h, int pad_w, float* out, void* stream) {
    CSB_REQUIRE(depth_net && out && Hn > 0 && Wn > 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    k_zoe_finish<<<csb::wave_grid((long long) H * W, 256, 4), 256, 0, st>>>(depth_net, flip_aug ? 2 : 1, Hn, Wn, H, W, pad_h, pad_w, out);
    return csb::launched("k_zoe_finish", st);
}

extern "C" int csb_zoe_disparity(const float* depth, long long n, double focal, double baseline, float* disparity, unsigned* scratch, void* stream) {
    CSB_REQUIRE(depth && disparity && scratch && n > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(scratch, 0x7f, sizeof(unsigned), st), "memset"));      // 0x7f7f7f7f: a huge finite float, > any depth
    csb::memset_done(st);
    k_min_positive<<<csb::wave_grid(n, 256, 4), 256, 0, st>>>(depth, n, scratch);
    CSB_TRY(csb::launched("k_min_positive", st));
    k_zoe_disparity<<<csb::wave_grid(n, 256, 4), 256, 0, st>>>(depth, n, scratch, (float) (focal * baseline), disparity);
    return csb::launched("k_zoe_disparity", st);
}
