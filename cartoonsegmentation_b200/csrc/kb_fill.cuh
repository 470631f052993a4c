// Ray-march search of fill_disocclusion (anime_3dkenburns/common.py:165-235), shared by the standalone kernel
// (kb_fill.cu) and the fused frame kernel (kb_frame.cu).
#pragma once
#include "common.cuh"

namespace csbfill {

struct Dirs {
    float x[16], y[16];
};

__device__ __forceinline__ Dirs make_dirs() {
    Dirs d = {{-1, 0, 1, 1, -1, 1, 2, 2, -2, -1, 1, 2, 3, 3, 3, 3}, {1, 1, 1, 0, 2, 2, 1, -1, 3, 3, 3, 3, 2, 1, -1, -2}};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float n = sqrtf((d.x[i] * d.x[i]) + (d.y[i] * d.y[i]));
        d.x[i] = __fdiv_rn(d.x[i], n);
        d.y[i] = __fdiv_rn(d.y[i], n);
    }
    return d;
}

// Returns the fill source pixel of hole (x,y) or -1.  `valid(yy,xx)` and `depthv(yy,xx)` abstract the depth plane so the
// fused frame kernel (kb_frame.cu) can evaluate them from the render accumulator.
template <class ValidF, class DepthF>
__device__ __forceinline__ long long find_fill(int x, int y, int H, int W, ValidF valid, DepthF depthv) {
    const Dirs dir = make_dirs();
    float shortest = 1000000.0f;
    int fillx = -1, filly = -1;
#pragma unroll 1
    for (int d = 0; d < 16; ++d) {
        const float dx = dir.x[d], dy = dir.y[d];
        float ffx = (float) x, ffy = (float) y, tfx = (float) x, tfy = (float) y;
        int ifx = 0, ify = 0, itx = 0, ity = 0;
        do {                                                                   // :188-196
            ffx = __fsub_rn(ffx, dx); ifx = (int) roundf(ffx);
            ffy = __fsub_rn(ffy, dy); ify = (int) roundf(ffy);
            if ((ifx < 0) | (ifx >= W)) break;
            if ((ify < 0) | (ify >= H)) break;
            if (valid(ify, ifx)) break;
        } while (true);
        if ((ifx < 0) | (ifx >= W)) continue;
        if ((ify < 0) | (ify >= H)) continue;
        do {                                                                   // :199-207
            tfx = __fadd_rn(tfx, dx); itx = (int) roundf(tfx);
            tfy = __fadd_rn(tfy, dy); ity = (int) roundf(tfy);
            if ((itx < 0) | (itx >= W)) break;
            if ((ity < 0) | (ity >= H)) break;
            if (valid(ity, itx)) break;
        } while (true);
        if ((itx < 0) | (itx >= W)) continue;
        if ((ity < 0) | (ity >= H)) continue;
        float dist = sqrtf(__fadd_rn(powf((float) (itx - ifx), 2), powf((float) (ity - ify), 2)));   // :210
        if (shortest > dist) {
            fillx = ifx; filly = ify;
            if (depthv(ify, ifx) < depthv(ity, itx)) { fillx = itx; filly = ity; }
            shortest = dist;
        }
    }
    if (fillx == -1 || filly == -1) return -1;
    return (long long) filly * W + fillx;
}

}  // namespace csbfill
