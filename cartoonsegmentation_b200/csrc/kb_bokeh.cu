// C8 -- bokeh depth-of-field of the Ken-Burns frame loop, entirely on the device (the reference round-trips every frame through numpy:
// kenburns_effect.py:1042-1067 + utils/effects.py:143-182 + zoedepth/utils/misc.py:97-150).
//
//   csb_depth_colorize_u8   colorize(depth, cmap='gray_r')[..., 0]: np.percentile(2 / 85) by a 4-pass radix select on order-preserving keys
//                           (exact order statistics, numpy's `_lerp`), normalise, matplotlib Colormap.__call__(bytes=True) index rule, gray_r LUT
//   csb_focal_plane_range   frame 0: per-instance np.median of the 8-bit depth under each mask (256-bin histograms), the start/end rule
//   csb_bokeh_blur          bokeh_blur(frame, depth8, 32, lightness, focal_plane, use_cuda=True, depth_factor): depth -> blur radius (a 256-entry
//                           table, because the depth is 8-bit), img^lightness (table, built on the host by numpy exactly as the reference does),
//                           the three directional gathers of `kernel_bokeh` (effects.py:16-72, including its planar-buffer / interleaved-index
//                           quirk), average, ^(1/lightness), *255, truncate.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t fkey(float f) {                  // order-preserving; NaN sorts last like numpy
    uint32_t u = __float_as_uint(f);
    if (f != f) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr float kInvalid = -99.0f;                                   // colorize(invalid_val=-99)

// Workspace layout (uint32 words): [0..3] prefix of the 4 queries, [4..7] remaining rank, [8] n_valid, [9] pad, [10..11] vmin (double),
// [12..13] vmax (double), [16 .. 16+1024) radix histograms [4][256], [1040 .. 1296) histogram of the 8-bit depth.
struct SelState {
    uint32_t prefix[4];
    uint32_t rank[4];
    uint32_t nvalid, pad;
    double vmin, vmax;
    uint32_t pad2[2];
    uint32_t hist[4][256];
    uint32_t hist8[256];
    double gamma[2];
};

__global__ void k_sel_hist(const float* __restrict__ v, long long n, SelState* st, int shift) {
    __shared__ uint32_t h[4][256];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    uint32_t pre[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) pre[q] = st->prefix[q];
    const bool first = shift == 24;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float f = v[i];
        if (f == kInvalid) continue;
        const uint32_t k = fkey(f);
        const uint32_t b = (k >> shift) & 255u;
        if (first) {
            atomicAdd(&h[0][b], 1u);                                 // all four queries share the first-pass histogram
        } else {
            const uint32_t hi = k >> (shift + 8);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (hi == (pre[q] >> (shift + 8))) atomicAdd(&h[q][b], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&st->hist[0][0] + i, c);
    }
}

// One block of 128 threads: warp q resolves query q's next byte.  On the first pass the ranks are derived from the valid count:
// numpy 'linear' percentile: virtual index (n-1)*q, previous = floor, next = previous+1 (clipped), gamma = virtual - previous.
__global__ void k_sel_pick(SelState* st, int shift) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool first = shift == 24;
    const uint32_t* h = first ? st->hist[0] : st->hist[q];
    uint32_t c[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = h[lane * 8 + j]; s += c[j]; }
    uint32_t incl = s;                                               // inclusive scan over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t rank;
    if (first) {
        const double qq = (q < 2) ? 0.02 : 0.85;                     // np.true_divide(q, 100)
        const double nn = (double) total;
        const double virt = nn * qq + (1.0 + qq * (1.0 - 1.0 - 1.0)) - 1.0;      // numpy _compute_virtual_index(n, q, alpha=1, beta=1), same order
        double prev = floor(virt);
        if (prev < 0.0) prev = 0.0;
        uint32_t r = (uint32_t) prev + (uint32_t) (q & 1);
        if (r > total - 1) r = total - 1;
        rank = total ? r : 0;
        if (lane == 0 && (q & 1) == 0) st->gamma[q >> 1] = virt - prev;
        if (q == 0 && lane == 0) st->nvalid = total;
    } else {
        rank = st->rank[q];
    }
    const uint32_t excl = incl - s;
    if (total && rank >= excl && rank < incl) {                      // exactly one lane owns the bin
        uint32_t run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (rank >= run && rank < run + c[j]) {
                st->prefix[q] = (first ? 0u : st->prefix[q]) | ((uint32_t) (lane * 8 + j) << shift);
                st->rank[q] = rank - run;
            }
            run += c[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&st->hist[0][0])[i] = 0;
    if (shift == 0) {
        __syncthreads();
        if (threadIdx.x < 2) {                                       // numpy _lerp on float32 operands with a float64 weight
            const float a = fkey_inv(st->prefix[2 * threadIdx.x]), b = fkey_inv(st->prefix[2 * threadIdx.x + 1]);
            const double t = st->gamma[threadIdx.x];
            const float d = b - a;
            const double r = (t >= 0.5) ? (double) b - (double) d * (1.0 - t) : (double) a + (double) d * t;
            if (threadIdx.x == 0) st->vmin = r; else st->vmax = r;
        }
    }
}

// normalise + Colormap.__call__(bytes=True) of 'gray_r', channel 0
__global__ void k_colorize(const float* __restrict__ v, long long n, SelState* st, const uint8_t* __restrict__ lut, uint8_t* __restrict__ out) {
    __shared__ uint8_t l[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) l[i] = lut[i];
    __syncthreads();
    const double vmin = st->vmin, vmax = st->vmax;
    const bool flat = !(vmin != vmax);
    const float vmin32 = (float) vmin, den32 = (float) (vmax - vmin);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float f = v[i];
        uint8_t o;
        if (f == kInvalid) {
            o = 128;                                                 // background_color
        } else {
            const float x = flat ? f * 0.0f : __fdiv_rn(f - vmin32, den32);
            const float xa = x * 256.0f;
            if (xa != xa) o = 0;                                     // bad -> (0,0,0,0)
            else {
                int idx = xa >= 256.0f ? 255 : (xa < 0.0f ? 0 : (int) xa);
                o = l[idx];
            }
        }
        out[i] = o;
    }
}

// which of the 256 depth values occur in the frame (drives the min / max of the blur-radius table)
__global__ void k_hist8(const uint8_t* __restrict__ d8, long long n, uint32_t* __restrict__ hist8) {
    __shared__ uint32_t h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) atomicAdd(&h[d8[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&hist8[i], h[i]);
}

// per-instance 256-bin histogram of depth8 under the mask: grid (chunks, K)
__global__ void k_mask_hist(const uint8_t* __restrict__ d8, const uint8_t* __restrict__ masks, long long n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint8_t* m = masks + (size_t) blockIdx.y * n;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
        if (m[i]) atomicAdd(&h[d8[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[blockIdx.y * 256 + i], h[i]);
}

// kenburns_effect.py:1045-1059: focalplane_end = max over instances of np.median(depth8[mask]) (start value -1; empty masks give nan and never
// win), or 255 without instances; focalplane_start = 255 if |255-end| > |end| else 0.  One thread per instance, then thread 0 reduces.
__global__ void k_focal_range(const uint32_t* __restrict__ hist, int K, double* __restrict__ out) {
    extern __shared__ double med[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const uint32_t* h = hist + k * 256;
        unsigned long long n = 0;
        for (int i = 0; i < 256; ++i) n += h[i];
        double m = nan("");
        if (n) {
            const unsigned long long r0 = (n - 1) / 2, r1 = n / 2;   // np.median: mean of the two middle order statistics
            unsigned long long run = 0;
            int a = -1, b = -1;
            for (int i = 0; i < 256; ++i) {
                run += h[i];
                if (a < 0 && run > r0) a = i;
                if (b < 0 && run > r1) { b = i; break; }
            }
            m = ((double) a + (double) b) / 2.0;
        }
        med[k] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double start = 0.0, end = 255.0;
        if (K > 0) {
            end = -1.0;
            for (int k = 0; k < K; ++k)
                if (med[k] > end) end = med[k];
            start = fabs(255.0 - end) > fabs(0.0 - end) ? 255.0 : 0.0;
        }
        out[0] = start;
        out[1] = end;
    }
}

// effects.py:146-154,163-164: depth8 -> float32 -> max - |d - focal| -> ^depth_factor -> -min -> /max -> 1 - . -> * 0.0005, as a table over
// the 256 possible depth values; min / max run over the values PRESENT in the frame (hist8).
__global__ void k_blur_table(const SelState* __restrict__ st, const double* __restrict__ range, double focal_int, double focal_plane_host,
                             int depth_factor, float* __restrict__ dlut) {
    __shared__ float t[256];
    __shared__ float red[2];
    const int b = threadIdx.x;
    const bool present = st->hist8[b] != 0;
    const double fp = range ? focal_int * range[1] + (1.0 - focal_int) * range[0] : focal_plane_host;   // kenburns_effect.py:1066
    if (b == 0) {
        float dmax = 0.f;
        for (int i = 0; i < 256; ++i)
            if (st->hist8[i]) dmax = (float) i;
        red[0] = dmax;
    }
    __syncthreads();
    float x = red[0] - fabsf((float) b - (float) fp);
    if (depth_factor == 2) x = x * x;                                // np.power(float32, 2) is np.square
    else if (depth_factor != 1) x = (float) pow((double) x, (double) depth_factor);      // correctly rounded float32 pow
    t[b] = x;
    __syncthreads();
    if (b == 0) {
        float mn = INFINITY;
        for (int i = 0; i < 256; ++i)
            if (st->hist8[i]) mn = fminf(mn, t[i]);
        float mx = -INFINITY;
        for (int i = 0; i < 256; ++i)
            if (st->hist8[i]) mx = fmaxf(mx, t[i] - mn);
        red[0] = mn;
        red[1] = mx;
    }
    __syncthreads();
    x = __fdiv_rn(t[b] - red[0], red[1]);
    x = 1.0f - x;
    dlut[b] = present ? x * 0.0005f : 0.0f;
}

// np2flatten_tensor (effects.py:87-98): HWC -> [1,3,HW] planar, through the img^lightness table; depth8 -> blur-radius plane
__global__ void k_bokeh_planes(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ d8, long long n, const float* __restrict__ hlut,
                               const float* __restrict__ dlut, float* __restrict__ img, float* __restrict__ depth) {
    __shared__ float hl[256], dl[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { hl[i] = hlut[i]; dl[i] = dlut[i]; }
    __syncthreads();
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        img[i] = hl[frame[3 * i]];
        img[n + i] = hl[frame[3 * i + 1]];
        img[2 * n + i] = hl[frame[3 * i + 2]];
        depth[i] = dl[d8[i]];
    }
}

// kernel_bokeh (effects.py:16-72).  `img` is the channel-PLANAR buffer above but is indexed as if interleaved ((y*w+x)*3+c), exactly as the
// reference does; the product accumulates in the same order with the same contraction (color = fma(img, w, color)).
__global__ void k_bokeh_pass(long long n, int h, int w, int nsamples, float dx, float dy, const float* __restrict__ img,
                             const float* __restrict__ depth, float* __restrict__ blurred) {
    const int im_size = min(h, w), off = nsamples / 2;
    for (long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < n * 3; idx += (long long) gridDim.x * blockDim.x) {
        const long long smp = idx / 3;
        const int c = (int) (idx % 3);
        const int y = (int) ((smp / w) % h), x = (int) (smp % w);
        const long long fxy = (long long) y * w + x;
        const float d = depth[fxy];
        const float ddx = dx * d, ddy = dy * d;
        float weight = 0.f, color = 0.f;
        for (int s = 0; s < nsamples; ++s) {
            const int sp = (s - off) * im_size;
            const int x_ = x + (int) roundf(ddx * (float) sp);
            const int y_ = y + (int) roundf(ddy * (float) sp);
            if ((x_ >= w) | (y_ >= h) | (x_ < 0) | (y_ < 0)) continue;
            const long long f2 = (long long) y_ * w + x_;
            const float w_ = depth[f2];
            weight += w_;
            color = fmaf(img[f2 * 3 + c], w_, color);
        }
        blurred[idx] = (weight != 0.f) ? __fdiv_rn(color, weight) : img[fxy * 3 + c];
    }
}

// (diag + rhom) / 2 -> ftensor2img (planar -> HWC) -> ^(1/lightness) -> * 255 -> astype(uint8)   (effects.py:172-182)
__global__ void k_bokeh_finish(const float* __restrict__ a, const float* __restrict__ b, long long n, float inv_light, uint8_t* __restrict__ out) {
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = (a[c * n + i] + b[c * n + i]) / 2.0f;
            // np.power(float32, float32(1/lightness)): the platform's powf (glibc: correctly rounded; SVML: <= 1 ulp).  CUDA's powf is up to
            // 4 ulp off, which flips the truncation below on every unblurred pixel (v^13^(1/13)*255 sits ON an integer): use the correctly
            // rounded value, via double.
            const float p = (float) pow((double) v, (double) inv_light) * 255.0f;
            out[3 * i + c] = (uint8_t) (int) p;
        }
    }
}

}  // namespace

// Workspace: SelState (8 KiB) | dlut[256] | img, tmp1, tmp2 [3][n] | depth [n] | instance histograms [K][256]
extern "C" size_t csb_bokeh_workspace_bytes(int H, int W, int K) {
    const size_t n = (size_t) H * W;
    return 8192 + 1024 + sizeof(float) * (3 * 3 * n + n) + (size_t) (K > 0 ? K : 1) * 1024 + 256;
}

namespace {
struct BokehWs {
    SelState* st;
    float* dlut;
    float *img, *t1, *t2, *depth;
    uint32_t* ihist;
};
BokehWs carve(void* ws, int H, int W) {
    const size_t n = (size_t) H * W;
    char* p = (char*) ws;
    BokehWs w;
    w.st = (SelState*) p; p += 8192;
    w.dlut = (float*) p; p += 1024;
    w.img = (float*) p; p += sizeof(float) * 3 * n;
    w.t1 = (float*) p; p += sizeof(float) * 3 * n;
    w.t2 = (float*) p; p += sizeof(float) * 3 * n;
    w.depth = (float*) p; p += sizeof(float) * n;
    w.ihist = (uint32_t*) p;
    return w;
}
}  // namespace

extern "C" int csb_depth_colorize_u8(const float* depth, int H, int W, const uint8_t* lut256, uint8_t* out8, void* ws, void* stream) {
    CSB_REQUIRE(depth && lut256 && out8 && ws, "null pointer");
    CSB_REQUIRE(H > 0 && W > 0, "bad shape");
    static_assert(sizeof(SelState) <= 8192, "workspace header too small");
    cudaStream_t st = (cudaStream_t) stream;
    const long long n = (long long) H * W;
    BokehWs w = carve(ws, H, W);
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(w.st, 0, sizeof(SelState), st), "memset"));
    csb::memset_done(st);
    const int grid = csb::wave_grid(n, 256, 4);
    for (int shift = 24; shift >= 0; shift -= 8) {
        k_sel_hist<<<grid, 256, 0, st>>>(depth, n, w.st, shift);
        CSB_TRY(csb::launched("k_sel_hist", st));
        k_sel_pick<<<1, 128, 0, st>>>(w.st, shift);
        CSB_TRY(csb::launched("k_sel_pick", st));
    }
    k_colorize<<<grid, 256, 0, st>>>(depth, n, w.st, lut256, out8);
    return csb::launched("k_colorize", st);
}

extern "C" int csb_focal_plane_range(const uint8_t* depth8, const uint8_t* masks, int K, int H, int W, double* start_end, void* ws, void* stream) {
    CSB_REQUIRE(depth8 && start_end && ws && (masks || K == 0), "null pointer");
    CSB_REQUIRE(H > 0 && W > 0 && K >= 0 && K <= 4096, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    const long long n = (long long) H * W;
    BokehWs w = carve(ws, H, W);
    if (K > 0) {
        CSB_TRY(csb::cuda_ok(cudaMemsetAsync(w.ihist, 0, (size_t) K * 1024, st), "memset"));
        csb::memset_done(st);
        int chunks = (2 * csb::num_sms() + K - 1) / K;
        chunks = chunks < 1 ? 1 : chunks;
        k_mask_hist<<<dim3(chunks, K), 256, 0, st>>>(depth8, masks, n, w.ihist);
        CSB_TRY(csb::launched("k_mask_hist", st));
    }
    k_focal_range<<<1, 128, sizeof(double) * (K > 0 ? K : 1), st>>>(w.ihist, K, start_end);
    return csb::launched("k_focal_range", st);
}

extern "C" int csb_bokeh_blur(const uint8_t* frame, const uint8_t* depth8, int H, int W, int nsamples, const float* highlight_lut, float inv_lightness,
                              const double* focal_range, double focal_int, double focal_plane, int depth_factor, uint8_t* out, void* ws,
                              void* stream) {
    CSB_REQUIRE(frame && depth8 && highlight_lut && out && ws, "null pointer");
    CSB_REQUIRE(H > 0 && W > 0 && nsamples > 0 && nsamples <= 1024, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    const long long n = (long long) H * W;
    BokehWs w = carve(ws, H, W);
    const int grid = csb::wave_grid(n, 256, 4);
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(w.st->hist8, 0, sizeof(uint32_t) * 256, st), "memset"));
    csb::memset_done(st);
    k_hist8<<<grid, 256, 0, st>>>(depth8, n, w.st->hist8);
    CSB_TRY(csb::launched("k_hist8", st));
    k_blur_table<<<1, 256, 0, st>>>(w.st, focal_range, focal_int, focal_plane, depth_factor, w.dlut);
    CSB_TRY(csb::launched("k_blur_table", st));
    k_bokeh_planes<<<grid, 256, 0, st>>>(frame, depth8, n, highlight_lut, w.dlut, w.img, w.depth);
    CSB_TRY(csb::launched("k_bokeh_planes", st));
    const double PI = 3.141592653589793;         // math.pi
    const int g3 = csb::wave_grid(3 * n, 256, 8);
    k_bokeh_pass<<<g3, 256, 0, st>>>(n, H, W, nsamples, 0.0f, 1.0f, w.img, w.depth, w.t1);                                        // vertical
    CSB_TRY(csb::launched("k_bokeh_pass", st));
    k_bokeh_pass<<<g3, 256, 0, st>>>(n, H, W, nsamples, (float) cos(-PI / 6), (float) sin(-PI / 6), w.t1, w.depth, w.t2);          // diagonal
    CSB_TRY(csb::launched("k_bokeh_pass", st));
    k_bokeh_pass<<<g3, 256, 0, st>>>(n, H, W, nsamples, (float) cos(-PI * 5 / 6), (float) sin(-PI * 5 / 6), w.t2, w.depth, w.img);  // rhomboid
    CSB_TRY(csb::launched("k_bokeh_pass", st));
    k_bokeh_finish<<<grid, 256, 0, st>>>(w.t2, w.img, n, inv_lightness, out);
    return csb::launched("k_bokeh_finish", st);
}
