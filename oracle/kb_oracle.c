/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's 3D-Ken-Burns
 * point-cloud kernels and their elementwise neighbours.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product path never does.
 *
 * Each function follows the reference line by line (paths relative to /root/reference):
 *   orc_render_*           anime_3dkenburns/models/utils.py:56-315   (three cupy kernels + host tail)
 *   orc_fill_disocclusion  anime_3dkenburns/common.py:145-247
 *   orc_process_shift      anime_3dkenburns/common.py:76-81          (tensor part; scalar part is host Python)
 *   orc_depth_to_points    anime_3dkenburns/models/utils.py:43-50
 *   orc_laplacian/median   anime_3dkenburns/models/utils.py:9-40
 *   orc_disparity_to_cloud anime_3dkenburns/kenburns_effect.py:928-937
 *   orc_frame_*            anime_3dkenburns/kenburns_effect.py:1040,1069-1070 (numpy astype + OpenCV u8 paths)
 *
 * Floating-point fidelity: compiled with -ffp-contract=off.  Where the SASS of the reference
 * kernels (built from the reference strings by oracle/build_ref_kernels.py) shows a contraction
 * or a double-precision evaluation, it is written out explicitly here (fmaf / double casts):
 *   - intersection = fmaf(-x, dist, x)                  (FFMA in kernel_pointrender_update{Zee,Output})
 *   - z < 0.001, |den| < 0.001, outX/outY, fltError, zee+1.0 are evaluated in double
 *     (bare literals in the kernel strings are C doubles)
 * Parity status: pinned against the reference kernels themselves, compiled unmodified from
 * /root/reference and executed on the B200 (tests/test_ref_kernels_gpu.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* CUDA float -> int conversion (cvt.rzi.s32.f32): NaN -> 0, saturating -- the reference kernels run on the GPU, so `(int) floor(NaN)`
 * is 0 there, not C's undefined behaviour. */
static inline int f2i(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int) v;
}

typedef struct { int ok; float ox, oy, err; int x0, y0; float wnw, wne, wsw, wse; } proj_t;

/* models/utils.py:76-135 (identical in :229-266) */
static proj_t project(float x, float y, float z, int H, int W, double focal, double baseline) {
    proj_t p; memset(&p, 0, sizeof p);
    if ((double) z < 0.001) return p;                               /* :82 */
    float ffocal = (float) focal;
    float s = (0.0f - x) * 0.0f + (0.0f - y) * 0.0f;                /* dot() x/y terms: +-0 (NaN-propagating) */
    float num = (ffocal - z) + s;                                   /* :86 */
    float den = (0.0f - z) + s;                                     /* :87 */
    float dist = num / den;                                         /* :88 */
    if (fabs((double) den) < 0.001) return p;                       /* :90 */
    float ix = fmaf(0.0f - x, dist, x);                             /* :94 (contracted by NVRTC/nvcc) */
    float iy = fmaf(0.0f - y, dist, y);
    p.ox = (float) (((double) ix + (0.5 * W)) - 0.5);               /* :96 */
    p.oy = (float) (((double) iy + (0.5 * H)) - 0.5);               /* :97 */
    p.err = (float) (1000000.0 - ((focal * baseline) / ((double) z + 0.0000001)));  /* :99 */
    p.x0 = f2i(floorf(p.ox)); p.y0 = f2i(floorf(p.oy));           /* :101-102 */
    int sex = p.x0 + 1, sey = p.y0 + 1;
    p.wnw = ((float) sex - p.ox) * ((float) sey - p.oy);            /* :110-113 */
    p.wne = (p.ox - (float) p.x0) * ((float) sey - p.oy);
    p.wsw = ((float) sex - p.ox) * (p.oy - (float) p.y0);
    p.wse = (p.ox - (float) p.x0) * (p.oy - (float) p.y0);
    p.ok = 1;
    return p;
}

static inline int inb(int x, int y, int H, int W) { return x >= 0 && x < W && y >= 0 && y < H; }

/* kernel_pointrender_updateZee, models/utils.py:63-149.  zee pre-filled with 1e6 (:59). */
ORC_API void orc_render_zpass(const float* pts, int B, int N, int H, int W, double focal, double baseline, float* zee) {
    for (long i = 0; i < (long) B * H * W; ++i) zee[i] = 1000000.0f;
    for (int b = 0; b < B; ++b) for (int n = 0; n < N; ++n) {
        const float* P = pts + (long) b * 3 * N;
        proj_t p = project(P[n], P[N + n], P[2 * N + n], H, W, focal, baseline);
        if (!p.ok) continue;
        int tx, ty;                                                 /* :115-135 first max in NW,NE,SW,SE order */
        if (p.wnw >= p.wne && p.wnw >= p.wsw && p.wnw >= p.wse) { tx = p.x0; ty = p.y0; }
        else if (p.wne >= p.wnw && p.wne >= p.wsw && p.wne >= p.wse) { tx = p.x0 + 1; ty = p.y0; }
        else if (p.wsw >= p.wnw && p.wsw >= p.wne && p.wsw >= p.wse) { tx = p.x0; ty = p.y0 + 1; }
        else if (p.wse >= p.wnw && p.wse >= p.wne && p.wse >= p.wsw) { tx = p.x0 + 1; ty = p.y0 + 1; }
        else continue;                                              /* NaN weights: no branch taken */
        if (!inb(tx, ty, H, W)) continue;
        float* z = zee + ((long) b * H + ty) * W + tx;
        if (*z > p.err) *z = p.err;                                 /* float atomicMin, cupy_utils.py:21-29 */
    }
}

/* kernel_pointrender_updateDegrid, models/utils.py:152-212.  The reference updates in place (racy,
 * order-dependent); the oracle is the out-of-place (Jacobi) outcome -- one legal schedule: every
 * thread reads before any thread writes. */
ORC_API void orc_render_degrid(const float* zin, int B, int H, int W, float* zout) {
    static const int ox[4] = { 1, 0, 1, 1 }, oy[4] = { 0, 1, 1, -1 };
    for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const float* Z = zin + (long) b * H * W;
        float c = Z[(long) y * W + x];
        int cnt = 0; float sum = 0.0f;
        for (int k = 0; k < 4; ++k) {
            int x1 = x + ox[k], y1 = y + oy[k], x2 = x - ox[k], y2 = y - oy[k];
            if (!inb(x1, y1, H, W) || !inb(x2, y2, H, W)) continue;
            float a = Z[(long) y1 * W + x1], d = Z[(long) y2 * W + x2];
            if ((double) c >= (double) a + 1.0 && (double) c >= (double) d + 1.0) { cnt += 2; sum += a; sum += d; }
        }
        float o = c;
        if (cnt > 0) o = fminf(c, sum / (float) cnt);
        zout[((long) b * H + y) * W + x] = o;
    }
}

/* kernel_pointrender_updateOutput, models/utils.py:215-313; acc is [B, C+1, H, W], zero-filled, the last
 * channel accumulates the bare weights (tenData gets a ones channel, :57).  Points are visited in index
 * order (one legal atomicAdd order). */
ORC_API void orc_render_splat(const float* pts, const float* data, const float* zee, int B, int N, int C, int H, int W,
                              double focal, double baseline, float* acc) {
    long HW = (long) H * W;
    memset(acc, 0, sizeof(float) * B * (C + 1) * HW);
    for (int b = 0; b < B; ++b) for (int n = 0; n < N; ++n) {
        const float* P = pts + (long) b * 3 * N;
        const float* D = data + (long) b * C * N;
        const float* Z = zee + (long) b * HW;
        float* A = acc + (long) b * (C + 1) * HW;
        proj_t p = project(P[n], P[N + n], P[2 * N + n], H, W, focal, baseline);
        if (!p.ok) continue;
        int cx[4] = { p.x0, p.x0 + 1, p.x0, p.x0 + 1 }, cy[4] = { p.y0, p.y0, p.y0 + 1, p.y0 + 1 };
        float cw[4] = { p.wnw, p.wne, p.wsw, p.wse };
        for (int k = 0; k < 4; ++k) {
            if (!inb(cx[k], cy[k], H, W)) continue;
            long o = (long) cy[k] * W + cx[k];
            if (!((double) p.err <= (double) Z[o] + 1.0)) continue;  /* :269 */
            for (int c = 0; c < C; ++c) A[c * HW + o] += D[(long) c * N + n] * cw[k];
            A[C * HW + o] += 1.0f * cw[k];
        }
    }
}

/* host tail models/utils.py:315: render = acc[:C] / (acc[C] + 1e-7), existing = acc[C] */
ORC_API void orc_render_normalise(const float* acc, int B, int C, int H, int W, float* render, float* existing) {
    long HW = (long) H * W;
    for (int b = 0; b < B; ++b) for (long o = 0; o < HW; ++o) {
        float w = acc[((long) b * (C + 1) + C) * HW + o];
        existing[(long) b * HW + o] = w;
        float d = w + 0.0000001f;
        for (int c = 0; c < C; ++c) render[((long) b * C + c) * HW + o] = acc[((long) b * (C + 1) + c) * HW + o] / d;
    }
}

ORC_API void orc_render_pointcloud(const float* pts, const float* data, int B, int N, int C, int H, int W,
                                   double focal, double baseline, float* render, float* existing,
                                   float* zee_pre, float* zee_post) {
    long HW = (long) H * W;
    float* z0 = zee_pre ? zee_pre : (float*) malloc(sizeof(float) * B * HW);
    float* z1 = zee_post ? zee_post : (float*) malloc(sizeof(float) * B * HW);
    float* acc = (float*) malloc(sizeof(float) * B * (C + 1) * HW);
    orc_render_zpass(pts, B, N, H, W, focal, baseline, z0);
    orc_render_degrid(z0, B, H, W, z1);
    orc_render_splat(pts, data, z1, B, N, C, H, W, focal, baseline, acc);
    orc_render_normalise(acc, B, C, H, W, render, existing);
    free(acc);
    if (!zee_pre) free(z0);
    if (!zee_post) free(z1);
}

/* Autozoom score, common.py:126-128: count of pixels with tenExisting > 0 -- integer exact. */
ORC_API long orc_count_positive(const float* v, long n) {
    long c = 0;
    for (long i = 0; i < n; ++i) c += v[i] > 0.0f;
    return c;
}

/* kernel_discfill_updateOutput, common.py:149-245 */
ORC_API void orc_fill_disocclusion(const float* in, const float* depth, int B, int C, int H, int W, float* out) {
    long HW = (long) H * W;
    memcpy(out, in, sizeof(float) * B * C * HW);                    /* tenOutput = tenInput.clone() :146 */
    float dirx[16] = { -1, 0, 1, 1, -1, 1, 2, 2, -2, -1, 1, 2, 3, 3, 3, 3 };
    float diry[16] = { 1, 1, 1, 0, 2, 2, 1, -1, 3, 3, 3, 3, 2, 1, -1, -2 };
    for (int d = 0; d < 16; ++d) {                                  /* :174-179 */
        float nrm = sqrtf((dirx[d] * dirx[d]) + (diry[d] * diry[d]));
        dirx[d] /= nrm; diry[d] /= nrm;
    }
    for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const float* Dp = depth + (long) b * HW;
        if (Dp[(long) y * W + x] > 0.0f) continue;                  /* :163 */
        float shortest = 1000000.0f; int fx = -1, fy = -1;
        for (int d = 0; d < 16; ++d) {
            float ffx = (float) x, ffy = (float) y, tfx = (float) x, tfy = (float) y;
            int ifx = 0, ify = 0, itx = 0, ity = 0;
            do {                                                    /* :188-196 */
                ffx -= dirx[d]; ifx = f2i(roundf(ffx));
                ffy -= diry[d]; ify = f2i(roundf(ffy));
                if (ifx < 0 || ifx >= W) break;
                if (ify < 0 || ify >= H) break;
                if (Dp[(long) ify * W + ifx] > 0.0f) break;
            } while (1);
            if (ifx < 0 || ifx >= W) continue;
            if (ify < 0 || ify >= H) continue;
            do {                                                    /* :199-207 */
                tfx += dirx[d]; itx = f2i(roundf(tfx));
                tfy += diry[d]; ity = f2i(roundf(tfy));
                if (itx < 0 || itx >= W) break;
                if (ity < 0 || ity >= H) break;
                if (Dp[(long) ity * W + itx] > 0.0f) break;
            } while (1);
            if (itx < 0 || itx >= W) continue;
            if (ity < 0 || ity >= H) continue;
            float ddx = (float) (itx - ifx), ddy = (float) (ity - ify);
            float dist = sqrtf(ddx * ddx + ddy * ddy);              /* :210 powf(int,2) is exact for |int| < 4096 */
            if (shortest > dist) {
                fx = ifx; fy = ify;
                if (Dp[(long) ify * W + ifx] < Dp[(long) ity * W + itx]) { fx = itx; fy = ity; }
                shortest = dist;
            }
        }
        if (fx == -1 || fy == -1) continue;
        for (int c = 0; c < C; ++c)
            out[((long) b * C + c) * HW + (long) y * W + x] = in[((long) b * C + c) * HW + (long) fy * W + fx];
    }
}

/* tensor part of process_shift, common.py:76-81 (torch fp32 elementwise, no fusion) */
ORC_API void orc_process_shift(const float* pts, int B, int N, const float* shift3, float* out) {
    for (int b = 0; b < B; ++b) for (int n = 0; n < N; ++n) {
        const float* P = pts + (long) b * 3 * N; float* O = out + (long) b * 3 * N;
        float z = P[2 * N + n];
        float r = z / (z + 0.0000001f);
        O[n] = P[n] * r + shift3[0];
        O[N + n] = P[N + n] * r + shift3[1];
        O[2 * N + n] = z + shift3[2];
    }
}

/* depth_to_points, models/utils.py:43-50: linspace(-W/2+.5, W/2-.5, W) has unit step, so entry i is exactly
 * i - W/2 + 0.5; it is scaled by float32(1/focal) and then multiplied by depth (two fp32 products). */
ORC_API void orc_depth_to_points(const float* depth, int B, int H, int W, double focal, float* pts) {
    long HW = (long) H * W; float inv = (float) (1.0 / focal);
    for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        long o = (long) y * W + x; float d = depth[(long) b * HW + o];
        float hx = ((float) x + (-0.5f * W + 0.5f)) * inv, vy = ((float) y + (-0.5f * H + 0.5f)) * inv;
        pts[((long) b * 3 + 0) * HW + o] = d * hx;
        pts[((long) b * 3 + 1) * HW + o] = d * vy;
        pts[((long) b * 3 + 2) * HW + o] = d;
    }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflecti(int v, int n) { if (v < 0) v = -v; if (v >= n) v = 2 * (n - 1) - v; return v; }

/* spatial_filter 'laplacian', models/utils.py:12-23: replicate pad 1, cross-correlation with the reference's
 * (non-standard) stencil w[0][1]=w[0][2]=w[1][0]=w[2][0]=-1, w[1][1]=4; taps summed in row-major order. */
ORC_API void orc_laplacian(const float* in, int B, int C, int H, int W, float* out) {
    for (long bc = 0; bc < (long) B * C; ++bc) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const float* I = in + bc * H * W;
#define PX(dy, dx) I[(long) clampi(y + (dy), 0, H - 1) * W + clampi(x + (dx), 0, W - 1)]
        float acc = 0.0f;
        acc += -1.0f * PX(-1, 0);
        acc += -1.0f * PX(-1, 1);
        acc += -1.0f * PX(0, -1);
        acc += 4.0f * PX(0, 0);
        acc += -1.0f * PX(1, -1);
#undef PX
        out[bc * H * W + (long) y * W + x] = acc;
    }
}

static int cmpf(const void* a, const void* b) { float x = *(const float*) a, y = *(const float*) b; return (x > y) - (x < y); }

/* spatial_filter 'median-3' / 'median-5', models/utils.py:25-36: reflect pad, k*k window, torch.median = lower median */
ORC_API void orc_median(const float* in, int B, int C, int H, int W, int k, float* out) {
    int r = k / 2; float win[25];
    for (long bc = 0; bc < (long) B * C; ++bc) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        const float* I = in + bc * H * W; int m = 0;
        for (int dy = -r; dy <= r; ++dy) for (int dx = -r; dx <= r; ++dx)
            win[m++] = I[(long) reflecti(y + dy, H) * W + reflecti(x + dx, W)];
        qsort(win, m, sizeof(float), cmpf);
        out[bc * H * W + (long) y * W + x] = win[(m - 1) / 2];
    }
}

/* kenburns_effect.py:928-937.  In: raw disparity [H,W] (B=1, :40).  Out: disparity (scaled), depth, valid, points
 * (valid-masked), unaltered points, scalars[0..7] = dispmin, dispmax, depthmin, depthmax, minx, miny, maxx, maxy
 * (cv2.minMaxLoc of depth[128:-128,128:-128]: first occurrence in row-major order). */
ORC_API void orc_disparity_to_cloud(const float* raw, int H, int W, double focal, double baseline, float* disp, float* depth,
                                    float* valid, float* pts, float* unalt, double* scalars) {
    long HW = (long) H * W;
    float mx = raw[0];
    for (long i = 1; i < HW; ++i) mx = raw[i] > mx ? raw[i] : mx;
    float fb = (float) baseline, ffb = (float) (focal * baseline);
    for (long i = 0; i < HW; ++i) disp[i] = raw[i] / mx * fb;                       /* :928 */
    for (long i = 0; i < HW; ++i) depth[i] = ffb / (disp[i] + 0.00001f);            /* :929 */
    float mx2 = disp[0], mn2 = disp[0];
    for (long i = 1; i < HW; ++i) { mx2 = disp[i] > mx2 ? disp[i] : mx2; mn2 = disp[i] < mn2 ? disp[i] : mn2; }
    float* nd = (float*) malloc(sizeof(float) * HW); float* lap = (float*) malloc(sizeof(float) * HW);
    for (long i = 0; i < HW; ++i) nd[i] = disp[i] / mx2;                            /* :931 */
    orc_laplacian(nd, 1, 1, H, W, lap);
    for (long i = 0; i < HW; ++i) valid[i] = fabsf(lap[i]) < 0.03f ? 1.0f : 0.0f;
    for (long i = 0; i < HW; ++i) nd[i] = depth[i] * valid[i];
    orc_depth_to_points(nd, 1, H, W, focal, pts);                                   /* :932 */
    orc_depth_to_points(depth, 1, H, W, focal, unalt);                              /* :933 */
    free(nd); free(lap);
    scalars[0] = mn2; scalars[1] = mx2;
    if (H > 256 && W > 256) {
        float dmin = depth[128L * W + 128], dmax = dmin; int ix = 0, iy = 0, ax = 0, ay = 0;
        for (int y = 128; y < H - 128; ++y) for (int x = 128; x < W - 128; ++x) {
            float v = depth[(long) y * W + x];
            if (v < dmin) { dmin = v; ix = x - 128; iy = y - 128; }
            if (v > dmax) { dmax = v; ax = x - 128; ay = y - 128; }
        }
        scalars[2] = dmin; scalars[3] = dmax; scalars[4] = ix; scalars[5] = iy; scalars[6] = ax; scalars[7] = ay;
    }
}

/* kenburns_effect.py:1040: (render[0,0:3].transpose(1,2,0) * 255.0).clip(0,255).astype(uint8) -> HWC, truncation */
ORC_API void orc_frame_pack_u8(const float* render, int Cs, int H, int W, uint8_t* out) {
    long HW = (long) H * W; (void) Cs;
    for (long o = 0; o < HW; ++o) for (int c = 0; c < 3; ++c) {
        float v = render[c * HW + o] * 255.0f;
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        out[o * 3 + c] = (v != v) ? 0 : (uint8_t) v;
    }
}

/* cv2.getRectSubPix on 8UC3 -> 8UC3 (OpenCV imgproc/src/samplers.cpp getRectSubPix_Cn_<uchar,uchar,int,scale_fixpt,cast_8u>):
 * 16-bit fixed-point bilinear, replicate border, result (t + 2^15) >> 16.  kenburns_effect.py:1069. */
ORC_API void orc_getrectsubpix_u8c3(const uint8_t* src, int H, int W, int ph, int pw, double cx, double cy, uint8_t* dst) {
    float fcx = (float) cx, fcy = (float) cy;
    fcx -= (pw - 1) * 0.5f; fcy -= (ph - 1) * 0.5f;
    int ipx = (int) floor(fcx), ipy = (int) floor(fcy);
    float a = fcx - ipx, b = fcy - ipy;
    int a11 = (int) lrint((double) ((1.f - a) * (1.f - b)) * 65536.0), a12 = (int) lrint((double) (a * (1.f - b)) * 65536.0);
    int a21 = (int) lrint((double) ((1.f - a) * b) * 65536.0), a22 = (int) lrint((double) (a * b) * 65536.0);
    for (int y = 0; y < ph; ++y) for (int x = 0; x < pw; ++x) {
        int y0 = clampi(ipy + y, 0, H - 1), y1 = clampi(ipy + y + 1, 0, H - 1);
        int x0 = clampi(ipx + x, 0, W - 1), x1 = clampi(ipx + x + 1, 0, W - 1);
        for (int c = 0; c < 3; ++c) {
            int t = src[((long) y0 * W + x0) * 3 + c] * a11 + src[((long) y0 * W + x1) * 3 + c] * a12 +
                    src[((long) y1 * W + x0) * 3 + c] * a21 + src[((long) y1 * W + x1) * 3 + c] * a22;
            dst[((long) y * pw + x) * 3 + c] = (uint8_t) ((t + (1 << 15)) >> 16);
        }
    }
}

/* cv2.resize(..., INTER_LINEAR) on 8UC3 (OpenCV imgproc/src/resize.cpp: 11-bit fixed-point coefficients,
 * HResizeLinear then VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>).  kenburns_effect.py:1070. */
ORC_API void orc_resize_linear_u8c3(const uint8_t* src, int sh, int sw, int dh, int dw, uint8_t* dst) {
    double sx = 1.0 / ((double) dw / sw), sy = 1.0 / ((double) dh / sh);   /* scale = 1/inv_scale as in resize.cpp */
    int* xo = (int*) malloc(sizeof(int) * dw); short* xa = (short*) malloc(sizeof(short) * dw * 2);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float) ((dx + 0.5) * sx - 0.5);
        int ix = (int) floorf(fx); fx -= ix;
        if (ix < 0) { fx = 0; ix = 0; }
        if (ix >= sw - 1) { fx = 0; ix = sw - 1; }
        xo[dx] = ix;
        xa[dx * 2] = (short) lrintf((1.f - fx) * 2048.f); xa[dx * 2 + 1] = (short) lrintf(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float) ((dy + 0.5) * sy - 0.5);
        int iy = (int) floorf(fy); fy -= iy;
        int y0 = clampi(iy, 0, sh - 1), y1 = clampi(iy + 1, 0, sh - 1);
        short b0 = (short) lrintf((1.f - fy) * 2048.f), b1 = (short) lrintf(fy * 2048.f);
        for (int dx = 0; dx < dw; ++dx) {
            int x0 = xo[dx], x1 = x0 + 1 < sw ? x0 + 1 : x0;
            for (int c = 0; c < 3; ++c) {
                int s0 = src[((long) y0 * sw + x0) * 3 + c] * xa[dx * 2] + src[((long) y0 * sw + x1) * 3 + c] * xa[dx * 2 + 1];
                int s1 = src[((long) y1 * sw + x0) * 3 + c] * xa[dx * 2] + src[((long) y1 * sw + x1) * 3 + c] * xa[dx * 2 + 1];
                int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
                dst[((long) dy * dw + dx) * 3 + c] = (uint8_t) (v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
    free(xo); free(xa);
}

/* kernel_bokeh -- utils/effects.py:16-72, one directional gather pass of bokeh_filter_cupy (effects.py:12-84).  `img` and `blurred` are
 * 3*n floats: the reference hands the kernel a channel-PLANAR buffer (np2flatten_tensor, effects.py:87-98) but the kernel addresses it as
 * (y*w+x)*3+c (effects.py:36-38,58) -- restated as is.  `color += img*w_` is an FFMA in the NVRTC build (-fmad=true default): fmaf here.
 * int(round(float)) = roundf (half away from zero) then float->int. */
ORC_API void orc_bokeh_pass(int n, int h, int w, int nsamples, float dx, float dy, const float* img, const float* depth, float* blurred) {
    const int im_size = h < w ? h : w, sample_offset = nsamples / 2;
    for (long idx = 0; idx < (long) n * 3; ++idx) {
        const int smp = (int) (idx / 3), c = (int) (idx % 3);
        const int y = (smp / w) % h, x = smp % w;
        const long fxy = (long) y * w + x, fid = fxy * 3 + c;
        const float d = depth[fxy];
        const float _dx = dx * d, _dy = dy * d;
        float weight = 0.f, color = 0.f;
        for (int s = 0; s < nsamples; ++s) {
            const int sp = (s - sample_offset) * im_size;
            const int x_ = x + f2i(roundf(_dx * (float) sp));
            const int y_ = y + f2i(roundf(_dy * (float) sp));
            if ((x_ >= w) | (y_ >= h) | (x_ < 0) | (y_ < 0)) continue;
            const long f2 = (long) y_ * w + x_;
            const float w_ = depth[f2];
            weight += w_;
            color = fmaf(img[f2 * 3 + c], w_, color);
        }
        blurred[fid] = (weight != 0.f) ? color / weight : img[fid];
    }
}

/* cv2.resize(src 8UC1, (dw, dh), interpolation=INTER_AREA) when UPSCALING (dw >= sw and dh >= sh) -- the LeReS tail, kenburns_effect.py:572-577.
 * OpenCV imgproc/src/resize.cpp: "true area interpolation is only implemented for scale >= 1; otherwise it is emulated using some variant of
 * bilinear": INTER_LINEAR's fixed-point kernel (11-bit coefficients) with area-mode source positions
 *     sx = floor(dx * scale);  fx = (dx + 1) - (sx + 1) * inv_scale;  fx = fx <= 0 ? 0 : fx - floor(fx). */
ORC_API void orc_resize_area_up_u8c1(const uint8_t* src, int sh, int sw, int dh, int dw, uint8_t* dst) {
    const double inv_x = (double) dw / sw, inv_y = (double) dh / sh;
    const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    int* xo = (int*) malloc(sizeof(int) * dw); short* xa = (short*) malloc(sizeof(short) * dw * 2);
    for (int dx = 0; dx < dw; ++dx) {
        int sx = (int) floor(dx * scale_x);
        float fx = (float) ((dx + 1) - (sx + 1) * inv_x);
        fx = fx <= 0 ? 0.f : fx - floorf(fx);
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xo[dx] = sx;
        xa[dx * 2] = (short) lrintf((1.f - fx) * 2048.f); xa[dx * 2 + 1] = (short) lrintf(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        int sy = (int) floor(dy * scale_y);
        float fy = (float) ((dy + 1) - (sy + 1) * inv_y);
        fy = fy <= 0 ? 0.f : fy - floorf(fy);
        const int y0 = clampi(sy, 0, sh - 1), y1 = clampi(sy + 1, 0, sh - 1);
        const short b0 = (short) lrintf((1.f - fy) * 2048.f), b1 = (short) lrintf(fy * 2048.f);
        for (int dx = 0; dx < dw; ++dx) {
            const int x0 = xo[dx], x1 = x0 + 1 < sw ? x0 + 1 : x0;
            const int s0 = src[(long) y0 * sw + x0] * xa[dx * 2] + src[(long) y0 * sw + x1] * xa[dx * 2 + 1];
            const int s1 = src[(long) y1 * sw + x0] * xa[dx * 2] + src[(long) y1 * sw + x1] * xa[dx * 2 + 1];
            const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            dst[(long) dy * dw + dx] = (uint8_t) (v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xo); free(xa);
}

/* cv2.resize(src 8UC1, (dw, dh), interpolation=INTER_LANCZOS4) -- the LeReS tail when the estimator output has more rows than the frame,
 * kenburns_effect.py:573-575.  OpenCV imgproc/src/resize.cpp, in its own two-pass structure: per destination column / row the source offset
 * sx = cvFloor(fx), fx = (float) ((dx + 0.5) * scale - 0.5) and eight coefficients interpolateLanczos4(fx - sx) (imgwarp.cpp) rounded to
 * short x 2048 (saturate_cast<short>: cvRound); HResizeLanczos4<uchar,int,short> writes int rows (border taps clamped to the row),
 * VResizeLanczos4 combines eight rows (row index clipped to the image) in int32 and FixedPtCast<int,uchar,22> rounds: (v + 2^21) >> 22, saturated. */
static void orc_lanczos4(float x, float* coeffs) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    float sum = 0;
    double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; i++) {
        float y0_ = (x + 3 - i);
        if (fabs(y0_) >= 1e-6f) {
            double y = -y0_ * 3.1415926535897932384626433832795 * 0.25;
            coeffs[i] = (float) ((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            coeffs[i] = 1e30f;
        }
        sum += coeffs[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) coeffs[i] *= sum;
}

static short orc_sat_short(float v) {
    long r = lrintf(v);
    return (short) (r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
}

ORC_API void orc_resize_lanczos4_u8c1(const uint8_t* src, int sh, int sw, int dh, int dw, uint8_t* dst) {
    const double scale_x = 1.0 / ((double) dw / sw), scale_y = 1.0 / ((double) dh / sh);
    int* xofs = (int*) malloc(sizeof(int) * dw);
    short* alpha = (short*) malloc(sizeof(short) * dw * 8);
    int* rows = (int*) malloc(sizeof(int) * (size_t) sh * dw);          /* the horizontal pass of every source row */
    float cb[8];
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float) ((dx + 0.5) * scale_x - 0.5);
        int sx = (int) floorf(fx);
        fx -= sx;
        xofs[dx] = sx;
        orc_lanczos4(fx, cb);
        for (int k = 0; k < 8; ++k) alpha[dx * 8 + k] = orc_sat_short(cb[k] * 2048.f);
    }
    for (int y = 0; y < sh; ++y)
        for (int dx = 0; dx < dw; ++dx) {
            int v = 0;
            for (int k = 0; k < 8; ++k) v += src[(long) y * sw + clampi(xofs[dx] - 3 + k, 0, sw - 1)] * alpha[dx * 8 + k];
            rows[(long) y * dw + dx] = v;
        }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float) ((dy + 0.5) * scale_y - 0.5);
        int sy = (int) floorf(fy);
        fy -= sy;
        short beta[8];
        orc_lanczos4(fy, cb);
        for (int k = 0; k < 8; ++k) beta[k] = orc_sat_short(cb[k] * 2048.f);
        for (int dx = 0; dx < dw; ++dx) {
            unsigned v = 0;                                              /* int32 arithmetic as OpenCV's (unsigned here: defined wrap-around) */
            for (int k = 0; k < 8; ++k) v += (unsigned) (rows[(long) clampi(sy - 3 + k, 0, sh - 1) * dw + dx] * (int) beta[k]);
            int r = ((int) (v + (1u << 21))) >> 22;
            dst[(long) dy * dw + dx] = (uint8_t) (r < 0 ? 0 : (r > 255 ? 255 : r));
        }
    }
    free(xofs); free(alpha); free(rows);
}
