// ZoeDepth metric-head kernels for sm_100a: the elementwise stages between the 1x1 convs of
// depth_modules/zoedepth/models/zoedepth/zoedepth_v1.py:150-192.
//
//   k_attractor        AttractorLayerUnnormed.forward tail (layers/attractor.py:176-208): b_prev -> bilinear(align_corners=True) to (h,w);
//                      delta_c[bin] = mean_a inv_attractor(A_a - c_bin) with inv_attractor(dx) = dx / (1 + alpha * dx^gamma) evaluated with the
//                      function DEFAULTS alpha=300, gamma=2 (the layer calls dist() without its configured alpha/gamma, attractor.py:45,195 --
//                      reference quirk); b_new = c + delta_c.  The reference broadcasts a [B,na,64,h,w] tensor; here one
//                      thread owns one (pixel, bin).
//   k_zoe_cond_input   cat([outconv_activation(32), interpolate(rel_depth), interpolate(b_embedding 128)]) (zoedepth_v1.py:176-184, both
//                      align_corners=True) as one NHWC fp16 tensor padded to 176 channels, rel_depth moved to channel 160 -- the input of
//                      ConditionalLogBinomial.mlp (whose first layer's input channels are permuted to match).
//   k_logbinom_depth   ConditionalLogBinomial tail + LogBinomial + expectation (layers/dist_layers.py:29-33,58-69,110-121; zoedepth_v1.py:186-192):
//                      p, t linear-norm, Stirling log-binomial over 64 classes, softmax(y / t), depth = sum_k p_k * interpolate(b_centers)_k.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ void ac_coord(int d, int in, int out, int& i0, int& i1, float& l) {
    const float s = out > 1 ? (float) (in - 1) / (float) (out - 1) : 0.f;
    const float f = d * s;
    i0 = min((int) f, in - 1);
    i1 = min(i0 + 1, in - 1);
    l = f - (float) i0;
}

__global__ void __launch_bounds__(256) k_attractor(const float* __restrict__ A, int na, const float* __restrict__ bprev, int hp, int wp, int N, int h, int w,
                                                   int nbins, float alpha, float* __restrict__ bnew) {
    const long long total = (long long) N * h * w * nbins;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int bin = (int) (i % nbins);
        const long long pix = i / nbins;
        const int x = (int) (pix % w), y = (int) ((pix / w) % h);
        const long long n = pix / ((long long) w * h);
        int y0, y1, x0, x1;
        float ly, lx;
        ac_coord(y, hp, h, y0, y1, ly);
        ac_coord(x, wp, w, x0, x1, lx);
        const float* B = bprev + n * hp * wp * nbins + bin;
        const float c = (1.f - ly) * ((1.f - lx) * B[((size_t) y0 * wp + x0) * nbins] + lx * B[((size_t) y0 * wp + x1) * nbins]) +
                        ly * ((1.f - lx) * B[((size_t) y1 * wp + x0) * nbins] + lx * B[((size_t) y1 * wp + x1) * nbins]);
        const float* Ap = A + pix * na;
        float acc = 0.f;
        for (int a = 0; a < na; ++a) {
            const float dx = Ap[a] - c;
            acc += dx / (1.0f + alpha * dx * dx);          // inv_attractor, gamma = 2
        }
        bnew[i] = c + acc / (float) na;                    // kind = 'mean'
    }
}

// One thread writes 8 consecutive output channels (one 16 B store).  Channel order of the OUTPUT: [feat 0..31 | embedding 32..159 | rel_depth 160 | 0-pad]
// -- the reference concatenates [feat, rel, embedding]; moving the single rel channel behind the embedding keeps every 8-channel group of the
// embedding 16 B aligned (4 vector loads per thread instead of 32 two-byte loads: 7.4 -> ~2 ms per 32 net inputs).  The host permutes the input
// channels of conditional_log_binomial.mlp.0 accordingly (depth_modules/zoedepth.py).
__global__ void __launch_bounds__(256) k_zoe_cond_input(const __half* __restrict__ feat, const float* __restrict__ rel, int hr, int wr,
                                                        const __half* __restrict__ emb, int he, int we, int N, int H, int W, __half* __restrict__ out) {
    const long long total = (long long) N * H * W * 22;           // < 2^32 (checked by the host): index split in 32-bit arithmetic
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const unsigned ui = (unsigned) i;
        const int g = (int) (ui % 22u);
        const unsigned upix = ui / 22u;
        const long long pix = upix;
        uint4* o4 = reinterpret_cast<uint4*>(out + pix * 176) + g;
        if (g < 4) {                                                          // outconv_activation, copied
            *o4 = reinterpret_cast<const uint4*>(feat + pix * 32)[g];
            continue;
        }
        const int x = (int) (upix % (unsigned) W), y = (int) ((upix / (unsigned) W) % (unsigned) H);
        const long long n = upix / ((unsigned) W * (unsigned) H);
        int y0, y1, x0, x1;
        float ly, lx;
        ac_coord(y, he, H, y0, y1, ly);
        ac_coord(x, we, W, x0, x1, lx);
        const __half* E = emb + n * he * we * 128;
        const __half* e00 = E + ((size_t) y0 * we + x0) * 128;
        const __half* e01 = E + ((size_t) y0 * we + x1) * 128;
        const __half* e10 = E + ((size_t) y1 * we + x0) * 128;
        const __half* e11 = E + ((size_t) y1 * we + x1) * 128;
        __align__(16) __half v[8];
        if (g < 20) {                                                         // channels 32..159: interpolate(b_embedding), 16 B aligned corner loads
            const int ce = (g - 4) * 8;
            __align__(16) __half a[8], b[8], c[8], d[8];
            *reinterpret_cast<uint4*>(a) = *reinterpret_cast<const uint4*>(e00 + ce);
            *reinterpret_cast<uint4*>(b) = *reinterpret_cast<const uint4*>(e01 + ce);
            *reinterpret_cast<uint4*>(c) = *reinterpret_cast<const uint4*>(e10 + ce);
            *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(e11 + ce);
#pragma unroll
            for (int j = 0; j < 8; ++j)   // same association as the per-pixel form: (1-ly)*((1-lx)*a + lx*b) + ly*((1-lx)*c + lx*d)
                v[j] = __float2half_rn((1.f - ly) * ((1.f - lx) * __half2float(a[j]) + lx * __half2float(b[j])) +
                                       ly * ((1.f - lx) * __half2float(c[j]) + lx * __half2float(d[j])));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float2half_rn(0.f);          // channels 161..175: zero padding
            if (g == 20) {                                                    // channel 160: interpolate(rel_depth), align_corners=True
                int ry0, ry1, rx0, rx1;
                float rly, rlx;
                ac_coord(y, hr, H, ry0, ry1, rly);
                ac_coord(x, wr, W, rx0, rx1, rlx);
                const float* R = rel + n * hr * wr;
                v[0] = __float2half_rn((1.f - rly) * ((1.f - rlx) * R[(size_t) ry0 * wr + rx0] + rlx * R[(size_t) ry0 * wr + rx1]) +
                                       rly * ((1.f - rlx) * R[(size_t) ry1 * wr + rx0] + rlx * R[(size_t) ry1 * wr + rx1]));
            }
        }
        *o4 = *reinterpret_cast<const uint4*>(v);
    }
}

__device__ __forceinline__ float log_binom(float n, float k) {     // dist_layers.py:29-33, eps = 1e-7
    n += 1e-7f;
    k += 1e-7f;
    return n * logf(n) - k * logf(k) - (n - k) * logf(n - k + 1e-7f);
}

// K = 64 (the shipped config): log_binom(K-1, k) does not depend on the pixel -- one evaluation per CTA into shared memory (the same fp32
// expression, so results are bit-identical to the generic kernel below, which evaluated 6 K logf per pixel) -- and y_k is computed once and kept in
// registers for the max and the exp pass.
__global__ void __launch_bounds__(256) k_logbinom_depth64(const float* __restrict__ pt, const float* __restrict__ bc, int hb, int wb, int N, int H, int W,
                                                          float p_eps, float min_temp, float max_temp, float* __restrict__ depth) {
    constexpr int K = 64;
    __shared__ float s_lb[K];
    if (threadIdx.x < K) s_lb[threadIdx.x] = log_binom((float) (K - 1), (float) threadIdx.x);
    __syncthreads();
    const long long total = (long long) N * H * W;
    for (long long pix = blockIdx.x * (long long) blockDim.x + threadIdx.x; pix < total; pix += (long long) gridDim.x * blockDim.x) {
        const unsigned upix = (unsigned) pix;                                 // N*H*W < 2^32 (checked by the host)
        const int x = (int) (upix % (unsigned) W), y = (int) ((upix / (unsigned) W) % (unsigned) H);
        const long long n = upix / ((unsigned) W * (unsigned) H);
        const float4 q = *reinterpret_cast<const float4*>(pt + pix * 4);
        const float p0 = q.x + p_eps, p1 = q.y + p_eps, t0 = q.z + p_eps, t1 = q.w + p_eps;
        float p = p0 / (p0 + p1);
        const float t = (max_temp - min_temp) * (t0 / (t0 + t1)) + min_temp;
        const float omp = fminf(fmaxf(1.f - p, 1e-4f), 1.f);
        p = fminf(fmaxf(p, 1e-4f), 1.f);
        const float lp = logf(p), lq = logf(omp);
        int y0, y1, x0, x1;
        float ly, lx;
        ac_coord(y, hb, H, y0, y1, ly);
        ac_coord(x, wb, W, x0, x1, lx);
        const float* B = bc + n * hb * wb * K;
        const float4* b00 = reinterpret_cast<const float4*>(B + ((size_t) y0 * wb + x0) * K);
        const float4* b01 = reinterpret_cast<const float4*>(B + ((size_t) y0 * wb + x1) * K);
        const float4* b10 = reinterpret_cast<const float4*>(B + ((size_t) y1 * wb + x0) * K);
        const float4* b11 = reinterpret_cast<const float4*>(B + ((size_t) y1 * wb + x1) * K);
        float yk[K];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            yk[k] = (s_lb[k] + (float) k * lp + (float) (K - 1 - k) * lq) / t;
            mx = fmaxf(mx, yk[k]);
        }
        float se = 0.f, sc = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < K / 4; ++k4) {
            const float4 c00 = __ldg(b00 + k4), c01 = __ldg(b01 + k4), c10 = __ldg(b10 + k4), c11 = __ldg(b11 + k4);
            const float cc[4] = {(1.f - ly) * ((1.f - lx) * c00.x + lx * c01.x) + ly * ((1.f - lx) * c10.x + lx * c11.x),
                                 (1.f - ly) * ((1.f - lx) * c00.y + lx * c01.y) + ly * ((1.f - lx) * c10.y + lx * c11.y),
                                 (1.f - ly) * ((1.f - lx) * c00.z + lx * c01.z) + ly * ((1.f - lx) * c10.z + lx * c11.z),
                                 (1.f - ly) * ((1.f - lx) * c00.w + lx * c01.w) + ly * ((1.f - lx) * c10.w + lx * c11.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = expf(yk[4 * k4 + j] - mx);
                se += e;
                sc += e * cc[j];
            }
        }
        depth[pix] = sc / se;
    }
}

__global__ void __launch_bounds__(256) k_logbinom_depth(const float* __restrict__ pt, const float* __restrict__ bc, int hb, int wb, int N, int H, int W, int K,
                                                        float p_eps, float min_temp, float max_temp, float* __restrict__ depth) {
    const long long total = (long long) N * H * W;
    for (long long pix = blockIdx.x * (long long) blockDim.x + threadIdx.x; pix < total; pix += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (pix % W), y = (int) ((pix / W) % H);
        const long long n = pix / ((long long) W * H);
        const float4 q = *reinterpret_cast<const float4*>(pt + pix * 4);
        const float p0 = q.x + p_eps, p1 = q.y + p_eps, t0 = q.z + p_eps, t1 = q.w + p_eps;
        float p = p0 / (p0 + p1);
        const float t = (max_temp - min_temp) * (t0 / (t0 + t1)) + min_temp;
        const float omp = fminf(fmaxf(1.f - p, 1e-4f), 1.f);
        p = fminf(fmaxf(p, 1e-4f), 1.f);
        const float lp = logf(p), lq = logf(omp);
        int y0, y1, x0, x1;
        float ly, lx;
        ac_coord(y, hb, H, y0, y1, ly);
        ac_coord(x, wb, W, x0, x1, lx);
        const float* B = bc + n * hb * wb * K;
        float mx = -INFINITY;
        for (int k = 0; k < K; ++k) {
            const float yk = (log_binom((float) (K - 1), (float) k) + (float) k * lp + (float) (K - 1 - k) * lq) / t;
            mx = fmaxf(mx, yk);
        }
        float se = 0.f, sc = 0.f;
        for (int k = 0; k < K; ++k) {
            const float yk = (log_binom((float) (K - 1), (float) k) + (float) k * lp + (float) (K - 1 - k) * lq) / t;
            const float e = expf(yk - mx);
            const float c = (1.f - ly) * ((1.f - lx) * B[((size_t) y0 * wb + x0) * K + k] + lx * B[((size_t) y0 * wb + x1) * K + k]) +
                            ly * ((1.f - lx) * B[((size_t) y1 * wb + x0) * K + k] + lx * B[((size_t) y1 * wb + x1) * K + k]);
            se += e;
            sc += e * c;
        }
        depth[pix] = sc / se;
    }
}

}  // namespace

extern "C" int csb_zoe_attractor(const float* A, int na, const float* b_prev, int hp, int wp, int N, int h, int w, int nbins, float alpha, float* b_new,
                                 void* stream) {
    CSB_REQUIRE(A && b_prev && b_new && na > 0 && N > 0 && h > 0 && w > 0 && hp > 0 && wp > 0 && nbins > 0, "bad arguments");
    k_attractor<<<csb::wave_grid((long long) N * h * w * nbins, 256, 8), 256, 0, (cudaStream_t) stream>>>(A, na, b_prev, hp, wp, N, h, w, nbins, alpha, b_new);
    return csb::launched("k_attractor", (cudaStream_t) stream);
}

extern "C" int csb_zoe_cond_input(const void* feat32, const float* rel, int hr, int wr, const void* emb128, int he, int we, int N, int H, int W, void* out176,
                                  void* stream) {
    CSB_REQUIRE(feat32 && rel && emb128 && out176 && N > 0 && H > 0 && W > 0, "bad arguments");
    CSB_REQUIRE((long long) N * H * W * 22 < (1ll << 32), "batch too large for the 32-bit index split (split it)");
    k_zoe_cond_input<<<csb::wave_grid((long long) N * H * W * 22, 256, 8), 256, 0, (cudaStream_t) stream>>>((const __half*) feat32, rel, hr, wr, (const __half*) emb128, he, we,
                                                                                                       N, H, W, (__half*) out176);
    return csb::launched("k_zoe_cond_input", (cudaStream_t) stream);
}

extern "C" int csb_zoe_logbinom_depth(const float* pt4, const float* b_centers, int hb, int wb, int N, int H, int W, int nbins, float p_eps, float min_temp,
                                      float max_temp, float* depth, void* stream) {
    CSB_REQUIRE(pt4 && b_centers && depth && N > 0 && H > 0 && W > 0 && nbins > 1, "bad arguments");
    if (nbins == 64 && (long long) N * H * W < (1ll << 32) && ((uintptr_t) b_centers & 15) == 0)
        k_logbinom_depth64<<<csb::wave_grid((long long) N * H * W, 256, 4), 256, 0, (cudaStream_t) stream>>>(pt4, b_centers, hb, wb, N, H, W, p_eps, min_temp, max_temp,
                                                                                                             depth);
    else
        k_logbinom_depth<<<csb::wave_grid((long long) N * H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(pt4, b_centers, hb, wb, N, H, W, nbins, p_eps, min_temp,
                                                                                                           max_temp, depth);
    return csb::launched("k_logbinom_depth", (cudaStream_t) stream);
}
