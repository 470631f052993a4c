// Ken-Burns frame tail for sm_100a (anime_3dkenburns/kenburns_effect.py:1028-1040, 1069-1070):
//   k_pack_u8          (render*255).clip(0,255).astype(uint8), CHW -> HWC                                  (:1040)
//   k_crop_resize      cv2.getRectSubPix (16-bit fixed point) + cv2.resize INTER_LINEAR (11-bit fixed point) (:1069-1070)
//   k_norm_pack_mark + k_fill_holes
//                      normalise (models/utils.py:315) + mask depth (:1039) + fill_disocclusion (common.py:149-245) + pack,
//                      reading the interleaved render accumulator -- the float render/existing/filled tensors of the
//                      reference (4 full-resolution fp32 tensors per frame) never touch HBM.
// The reference does the pack on the host after a D2H of the float frame and the crop/resize in OpenCV on the CPU; the
// integer arithmetic below reproduces OpenCV's uint8 paths bit for bit (pinned against cv2 in tests/test_oracle_cpu.py).
#include "kb_fill.cuh"

int csb_render_accumulate(const float* points, const float* data, int B, int N, int C, int H, int W, double focal, double baseline,
                          const float* shift, const float* shift_dev, int32_t* zkey, float* zee, float* acc, cudaStream_t st);

namespace {

__device__ __forceinline__ uint8_t pack1(float v) {
    v = __fmul_rn(v, 255.0f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);      // NaN -> 0 (fmaxf returns the non-NaN operand)
    return (uint8_t) (int) v;               // truncation, numpy astype
}

// 15 B/px
__global__ void __launch_bounds__(256) k_pack_u8(const float* __restrict__ render, long long HW, uint8_t* __restrict__ frame) {
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < HW; i += (long long) gridDim.x * blockDim.x) {
        frame[i * 3 + 0] = pack1(render[i]);
        frame[i * 3 + 1] = pack1(render[HW + i]);
        frame[i * 3 + 2] = pack1(render[2 * HW + i]);
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct CropParams {
    int ipx, ipy, a11, a12, a21, a22, pw, ph;
    double sx, sy;
};

// getRectSubPix_Cn_<uchar,uchar,int,scale_fixpt,cast_8u> (OpenCV samplers.cpp), one crop pixel, 3 channels.
template <bool COPY>
__device__ __forceinline__ void crop_px(const uint8_t* __restrict__ src, int H, int W, const CropParams& p, int x, int y, int (&o)[3]) {
    int y0 = clampi(p.ipy + y, 0, H - 1), y1 = clampi(p.ipy + y + 1, 0, H - 1);
    int x0 = clampi(p.ipx + x, 0, W - 1), x1 = clampi(p.ipx + x + 1, 0, W - 1);
    if constexpr (COPY) {       // integer crop origin (a11 = 65536, the other weights 0): (v * 65536 + 32768) >> 16 == v, one tap instead of four
        const uint8_t* r00 = src + ((size_t) y0 * W + x0) * 3;
        o[0] = r00[0]; o[1] = r00[1]; o[2] = r00[2];
        return;
    }
    const uint8_t *r00 = src + ((size_t) y0 * W + x0) * 3, *r01 = src + ((size_t) y0 * W + x1) * 3;
    const uint8_t *r10 = src + ((size_t) y1 * W + x0) * 3, *r11 = src + ((size_t) y1 * W + x1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int t = r00[c] * p.a11 + r01[c] * p.a12 + r10[c] * p.a21 + r11[c] * p.a22;
        o[c] = (t + (1 << 15)) >> 16;
    }
}

// resize.cpp: HResizeLinear<uchar,int,short,2048> + VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>.  6 B/px.
template <bool COPY>
__global__ void __launch_bounds__(256) k_crop_resize(const uint8_t* __restrict__ src, int H, int W, CropParams p, int Ho, int Wo, uint8_t* __restrict__ dst) {
    const int total = Ho * Wo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int dx = i % Wo, dy = i / Wo;
        float fx = (float) ((dx + 0.5) * p.sx - 0.5);
        int ix = (int) floorf(fx);
        fx -= ix;
        if (ix < 0) { fx = 0; ix = 0; }
        if (ix >= p.pw - 1) { fx = 0; ix = p.pw - 1; }
        int ax0 = (int) (short) __float2int_rn((1.f - fx) * 2048.f), ax1 = (int) (short) __float2int_rn(fx * 2048.f);
        float fy = (float) ((dy + 0.5) * p.sy - 0.5);
        int iy = (int) floorf(fy);
        fy -= iy;
        int y0 = clampi(iy, 0, p.ph - 1), y1 = clampi(iy + 1, 0, p.ph - 1);
        int b0 = (int) (short) __float2int_rn((1.f - fy) * 2048.f), b1 = (int) (short) __float2int_rn(fy * 2048.f);
        int x1 = ix + 1 < p.pw ? ix + 1 : ix;
        int c00[3], c01[3], c10[3], c11[3];
        crop_px<COPY>(src, H, W, p, ix, y0, c00);
        crop_px<COPY>(src, H, W, p, x1, y0, c01);
        crop_px<COPY>(src, H, W, p, ix, y1, c10);
        crop_px<COPY>(src, H, W, p, x1, y1, c11);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int s0 = c00[c] * ax0 + c01[c] * ax1, s1 = c10[c] * ax0 + c11[c] * ax1;
            int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            dst[(size_t) i * 3 + c] = (uint8_t) clampi(v, 0, 255);
        }
    }
}

// cv2.resize(INTER_LINEAR) with an exact 2x decimation takes OpenCV's INTER_AREA fast path: (a + b + c + d + 2) >> 2 per channel.
__global__ void __launch_bounds__(256) k_resize_half_u8c3(const uint8_t* __restrict__ src, int W, int Ho, int Wo, uint8_t* __restrict__ dst) {
    const int total = Ho * Wo * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % 3, x = (i / 3) % Wo, y = i / (3 * Wo);
        const uint8_t* p = src + ((size_t) (2 * y) * W + 2 * x) * 3 + c;
        dst[i] = (uint8_t) ((p[0] + p[3] + p[(size_t) W * 3] + p[(size_t) W * 3 + 3] + 2) >> 2);
    }
}

// Fused normalise + depth mask + disocclusion fill + u8 pack for C = 4 (BGR + depth), CP = 8, in two launches.
// acc pixel = {b*w, g*w, r*w, depth*w, w, 0, 0, 0} (32 B, one sector).
//
// A pixel is a hole of the composed reference path iff  render[3] * (existing > 0) <= 0  (kenburns_effect.py:1039, common.py:163), i.e.
// iff !(w > 0 && acc3 / (w + 1e-7) > 0).  For finite data that is exactly !(w > 0 && acc3 > 0): the quotient of a normal positive
// acc3 (the float reductions flush subnormals) by w + 1e-7 <= ~5 cannot underflow to zero.
//
//   k_norm_pack_mark  every pixel: valid -> normalise + pack (+ depth); hole -> appended to a device-side hole list.  Also writes a byte
//                     validity mask (1 B/px) so that the ray march below touches 32 pixels per sector instead of 1.  21 B/px.
//   k_fill_holes      16 lanes per hole pixel, one ray direction per lane (the reference walks the 16 directions serially in ONE thread,
//                     common.py:181-235): all lanes busy, no divergence against valid pixels; the serial "first strictly shortest
//                     direction wins" rule becomes a shuffle arg-min on (distance, direction index).
__device__ __forceinline__ bool acc_valid(const float* __restrict__ acc, long long pix) {
    const float* A = acc + (size_t) pix * 8;
    return __ldg(A + 4) > 0.0f && __ldg(A + 3) > 0.0f;
}

__global__ void __launch_bounds__(256) k_norm_pack_mark(const float* __restrict__ acc, long long HW, uint8_t* __restrict__ frame, float* __restrict__ depth_out,
                                                        uint8_t* __restrict__ mask, int* __restrict__ holes, int* __restrict__ nholes) {
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < HW; i += (long long) gridDim.x * blockDim.x) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(acc + (size_t) i * 8));
        const float w = __ldg(acc + (size_t) i * 8 + 4);
        const bool ok = w > 0.0f && q.w > 0.0f;
        mask[i] = ok ? 1 : 0;
        if (!ok) holes[atomicAdd(nholes, 1)] = (int) i;        // aggregated per warp by the compiler (REDUX + one atomic)
        // holes that find no fill source keep their own normalised value (common.py:146 clone), so write it for every pixel
        const float d = __fadd_rn(w, 0.0000001f);
        frame[i * 3 + 0] = pack1(__fdiv_rn(q.x, d));
        frame[i * 3 + 1] = pack1(__fdiv_rn(q.y, d));
        frame[i * 3 + 2] = pack1(__fdiv_rn(q.z, d));
        if (depth_out) depth_out[i] = __fdiv_rn(q.w, d);
    }
}

__global__ void __launch_bounds__(256) k_fill_holes(const float* __restrict__ acc, const uint8_t* __restrict__ mask, const int* __restrict__ holes,
                                                    const int* __restrict__ nholes, int H, int W, uint8_t* __restrict__ frame, float* __restrict__ depth_out) {
    const int n = *nholes;
    const int d = threadIdx.x & 15;                                 // ray direction of this lane
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);          // the 16 lanes cooperating on one hole
    float dx, dy;
    {
        const float tx[16] = {-1, 0, 1, 1, -1, 1, 2, 2, -2, -1, 1, 2, 3, 3, 3, 3}, ty[16] = {1, 1, 1, 0, 2, 2, 1, -1, 3, 3, 3, 3, 2, 1, -1, -2};
        const float nrm = sqrtf((tx[d] * tx[d]) + (ty[d] * ty[d]));  // common.py:174-179
        dx = __fdiv_rn(tx[d], nrm);
        dy = __fdiv_rn(ty[d], nrm);
    }
    auto depthv = [&](int yy, int xx) {                               // masked depth, kenburns_effect.py:1039
        const float* A = acc + ((size_t) yy * W + xx) * 8;
        const float w = __ldg(A + 4);
        return __fmul_rn(__fdiv_rn(__ldg(A + 3), __fadd_rn(w, 0.0000001f)), w > 0.0f ? 1.0f : 0.0f);
    };
    const int groups = (gridDim.x * blockDim.x) >> 4;
    for (int hidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; hidx < n; hidx += groups) {
        const int pix = holes[hidx];
        const int x = pix % W, y = pix / W;
        float ffx = (float) x, ffy = (float) y, tfx = (float) x, tfy = (float) y;
        int ifx = 0, ify = 0, itx = 0, ity = 0;
        bool found = true;
        do {                                                          // common.py:188-196
            ffx = __fsub_rn(ffx, dx); ifx = (int) roundf(ffx);
            ffy = __fsub_rn(ffy, dy); ify = (int) roundf(ffy);
            if ((ifx < 0) | (ifx >= W)) break;
            if ((ify < 0) | (ify >= H)) break;
            if (__ldg(mask + (size_t) ify * W + ifx)) break;
        } while (true);
        if ((ifx < 0) | (ifx >= W) | (ify < 0) | (ify >= H)) found = false;
        if (found) {
            do {                                                      // common.py:199-207
                tfx = __fadd_rn(tfx, dx); itx = (int) roundf(tfx);
                tfy = __fadd_rn(tfy, dy); ity = (int) roundf(tfy);
                if ((itx < 0) | (itx >= W)) break;
                if ((ity < 0) | (ity >= H)) break;
                if (__ldg(mask + (size_t) ity * W + itx)) break;
            } while (true);
            if ((itx < 0) | (itx >= W) | (ity < 0) | (ity >= H)) found = false;
        }
        float dist = 1000000.0f;                                      // fltShortest init :165; a direction only wins with dist < 1e6
        int fill = -1;
        if (found) {
            dist = sqrtf(__fadd_rn(powf((float) (itx - ifx), 2), powf((float) (ity - ify), 2)));     // :210
            fill = ify * W + ifx;
            if (depthv(ify, ifx) < depthv(ity, itx)) fill = ity * W + itx;                           // :216-219
            if (!(1000000.0f > dist)) fill = -1;
        }
        // serial rule: a later direction replaces the best only if strictly shorter -> min over (dist, direction index)
        int best = d;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(gmask, dist, o);
            const int ob = __shfl_xor_sync(gmask, best, o), of = __shfl_xor_sync(gmask, fill, o);
            if (od < dist || (od == dist && ob < best)) { dist = od; best = ob; fill = of; }
        }
        if (d == 0 && fill >= 0) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(acc + (size_t) fill * 8));
            const float dd = __fadd_rn(__ldg(acc + (size_t) fill * 8 + 4), 0.0000001f);
            frame[(size_t) pix * 3 + 0] = pack1(__fdiv_rn(q.x, dd));
            frame[(size_t) pix * 3 + 1] = pack1(__fdiv_rn(q.y, dd));
            frame[(size_t) pix * 3 + 2] = pack1(__fdiv_rn(q.z, dd));
            if (depth_out) depth_out[pix] = __fdiv_rn(q.w, dd);
        }
    }
}

int make_crop(int Ho, int Wo, int pw, int ph, double cx, double cy, CropParams& p) {
    const int H = Ho, W = Wo;
    if (pw <= 0 || ph <= 0) return CSB_ERR_INVALID;
    float fcx = (float) cx, fcy = (float) cy;
    fcx -= (pw - 1) * 0.5f;
    fcy -= (ph - 1) * 0.5f;
    p.ipx = (int) floor(fcx);
    p.ipy = (int) floor(fcy);
    float a = fcx - p.ipx, b = fcy - p.ipy;
    p.a11 = (int) lrint((double) ((1.f - a) * (1.f - b)) * 65536.0);
    p.a12 = (int) lrint((double) (a * (1.f - b)) * 65536.0);
    p.a21 = (int) lrint((double) ((1.f - a) * b) * 65536.0);
    p.a22 = (int) lrint((double) (a * b) * 65536.0);
    p.pw = pw;
    p.ph = ph;
    p.sx = 1.0 / ((double) W / pw);
    p.sy = 1.0 / ((double) H / ph);
    return CSB_OK;
}

}  // namespace

extern "C" int csb_frame_pack_u8(const float* render, int H, int W, uint8_t* frame, void* stream) {
    CSB_REQUIRE(render && frame, "null pointer");
    CSB_REQUIRE(H > 0 && W > 0, "bad shape");
    k_pack_u8<<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(render, (long long) H * W, frame);
    return csb::launched("k_pack_u8", (cudaStream_t) stream);
}

extern "C" int csb_frame_crop_resize(const uint8_t* frame, int H, int W, int pw, int ph, double cx, double cy, uint8_t* out, void* stream) {
    CSB_REQUIRE(frame && out && frame != out, "null or aliased pointer");
    CSB_REQUIRE(H > 0 && W > 0, "bad shape");
    CropParams p;
    CSB_REQUIRE(make_crop(H, W, pw, ph, cx, cy, p) == CSB_OK, "bad crop size");
    if (p.a12 == 0 && p.a21 == 0 && p.a22 == 0 && p.a11 == 65536) k_crop_resize<true><<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(frame, H, W, p, H, W, out);
    else k_crop_resize<false><<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(frame, H, W, p, H, W, out);
    return csb::launched("k_crop_resize", (cudaStream_t) stream);
}

extern "C" int csb_resize_u8c3(const uint8_t* src, int H, int W, uint8_t* dst, int Ho, int Wo, void* stream) {
    CSB_REQUIRE(src && dst && src != dst && H > 0 && W > 0 && Ho > 0 && Wo > 0, "bad arguments");
    if (W == 2 * Wo && H == 2 * Ho) {
        k_resize_half_u8c3<<<csb::wave_grid((long long) Ho * Wo * 3, 256, 8), 256, 0, (cudaStream_t) stream>>>(src, W, Ho, Wo, dst);
        return csb::launched("k_resize_half_u8c3", (cudaStream_t) stream);
    }
    CropParams p;   // full-frame 'crop' at integer offset 0: getRectSubPix degenerates to a copy, leaving cv2.resize(INTER_LINEAR)
    CSB_REQUIRE(make_crop(Ho, Wo, W, H, (W - 1) * 0.5, (H - 1) * 0.5, p) == CSB_OK, "bad size");
    if (p.a12 == 0 && p.a21 == 0 && p.a22 == 0 && p.a11 == 65536) k_crop_resize<true><<<csb::wave_grid((long long) Ho * Wo, 256, 8), 256, 0, (cudaStream_t) stream>>>(src, H, W, p, Ho, Wo, dst);
    else k_crop_resize<false><<<csb::wave_grid((long long) Ho * Wo, 256, 8), 256, 0, (cudaStream_t) stream>>>(src, H, W, p, Ho, Wo, dst);
    return csb::launched("k_crop_resize", (cudaStream_t) stream);
}

extern "C" int csb_kenburns_frame(const float* points, const float* data, int N, int H, int W, double focal, double baseline,
                                  const float* shift, const float* shift_dev, int pw, int ph, double cx, double cy, int32_t* zkey,
                                  float* zee, float* acc,
                                  uint8_t* packed, uint8_t* out, float* depth_out, void* stream) {
    CSB_REQUIRE(((points && data) || N == 0) && zkey && zee && acc && packed, "null pointer");
    CSB_REQUIRE(N >= 0 && H > 0 && W > 0 && (long long) H * W >= 8, "bad shape");
    CSB_REQUIRE(((uintptr_t) acc & 15) == 0, "acc must be 16-byte aligned");
    CropParams p;
    CSB_REQUIRE(make_crop(H, W, pw, ph, cx, cy, p) == CSB_OK, "bad crop size");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb_render_accumulate(points, data, 1, N, 4, H, W, focal, baseline, shift, shift_dev, zkey, zee, acc, st));
    // scratch reuse: the z-buffers are dead after the splat -> zkey holds the hole list, zee the byte mask + the hole counter
    uint8_t* mask = reinterpret_cast<uint8_t*>(zee);
    int* nholes = reinterpret_cast<int*>(zee) + (size_t) H * W / 2 + 1;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(nholes, 0, sizeof(int), st), "memset nholes"));
    csb::memset_done(st);
    k_norm_pack_mark<<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, st>>>(acc, (long long) H * W, packed, depth_out, mask, zkey, nholes);
    CSB_TRY(csb::launched("k_norm_pack_mark", st));
    k_fill_holes<<<csb::num_sms() * 8, 256, 0, st>>>(acc, mask, zkey, nholes, H, W, packed, depth_out);
    CSB_TRY(csb::launched("k_fill_holes", st));
    if (!out) return CSB_OK;                       // caller post-processes `packed` (bokeh) and crops with csb_frame_crop_resize
    if (p.a12 == 0 && p.a21 == 0 && p.a22 == 0 && p.a11 == 65536) k_crop_resize<true><<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, st>>>(packed, H, W, p, H, W, out);
    else k_crop_resize<false><<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, st>>>(packed, H, W, p, H, W, out);
    return csb::launched("k_crop_resize", st);
}

// The reference's whole frame loop (kenburns_effect.py:1015-1072) in ONE call: F frames of csb_kenburns_frame with per-frame camera shifts, written
// to out [F,H,W,3] on the device; if `host_out` (pinned) and `copy_stream` are given, frame f is copied to the host on `copy_stream` as soon as it
// is finished (one reusable event orders the copy after the frame), so the D2H of frame f overlaps the render of frame f+1 and the Python side
// issues one call instead of 6 launches + 1 copy per frame.
extern "C" int csb_kenburns_frames(const float* points, const float* data, int N, int H, int W, double focal, double baseline, const float* shifts, int F,
                                   int pw, int ph, double cx, double cy, int32_t* zkey, float* zee, float* acc, uint8_t* packed, uint8_t* out,
                                   uint8_t* host_out, void* copy_stream, void* stream) {
    CSB_REQUIRE(shifts && out && F > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream, cs = (cudaStream_t) copy_stream;
    const size_t frame_bytes = (size_t) H * W * 3;
    cudaEvent_t ev = nullptr;
    if (host_out) {
        CSB_REQUIRE(copy_stream != nullptr && copy_stream != stream, "host_out needs a separate copy stream");
        CSB_TRY(csb::cuda_ok(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate"));
    }
    int rc = CSB_OK;
    for (int f = 0; f < F && rc == CSB_OK; ++f) {
        rc = csb_kenburns_frame(points, data, N, H, W, focal, baseline, shifts + 3 * f, nullptr, pw, ph, cx, cy, zkey, zee, acc, packed,
                                out + (size_t) f * frame_bytes, nullptr, stream);
        if (rc == CSB_OK && host_out) {
            cudaEventRecord(ev, st);
            cudaStreamWaitEvent(cs, ev, 0);
            rc = csb::cuda_ok(cudaMemcpyAsync(host_out + (size_t) f * frame_bytes, out + (size_t) f * frame_bytes, frame_bytes, cudaMemcpyDeviceToHost, cs), "D2H");
        }
    }
    if (ev) {
        if (rc == CSB_OK) {                        // the caller's stream continues only after the last copy (it may reuse `out`)
            cudaEventRecord(ev, cs);
            cudaStreamWaitEvent(st, ev, 0);
        }
        cudaEventDestroy(ev);
    }
    return rc;
}
