"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's depth-of-field chain.

Follows, in order:
  * `colorize(value, cmap='gray_r')[..., 0]`      depth_modules/zoedepth/utils/misc.py:97-150 (called at kenburns_effect.py:1044)
  * the focal-plane rule                           anime_3dkenburns/kenburns_effect.py:1045-1066
  * `bokeh_blur(..., use_cuda=True)`               utils/effects.py:143-182, `bokeh_filter_cupy` :12-84, `np2flatten_tensor` :87-98,
                                                   `ftensor2img` :100-104; the gather itself is `orc_bokeh_pass` in kb_oracle.c

Third-party arithmetic on this path, pinned by the reference's conda_env.yaml and ABSENT from /root/reference (restated from the published
sources, so this part of the parity is unpinned ‡): numpy==1.26.2 `np.percentile` (lib/function_base.py `_quantile`, `_lerp`, method 'linear') and
its value-based scalar casting (a float64 scalar combined with a float32 array is cast to float32 first; the image's numpy 2.x does NOT do
that, hence the explicit casts below), matplotlib==3.9.2 `Colormap.__call__(bytes=True)` and `_create_lookup_table` for 'gray_r'.
The gather kernel is pinned: `oracle/build_ref_kernels.py` compiles the unmodified reference string and tests/test_bokeh_gpu.py compares all
three (reference kernel, this file, product) on the B200.
"""
import ctypes as C
import math

import numpy as np

from . import kb_oracle

F32 = np.float32


def pow_f32(x, e):
    """np.power(float32 array, scalar) of the reference runs the platform's float32 powf (glibc: correctly rounded; numpy's SVML builds: within
    1 ulp).  The unblurred pixels of bokeh_blur evaluate (v/255)^13^(1/13)*255, which lands ON an integer before the uint8 truncation, so the last
    ulp decides the output byte: the oracle (and the product) use the correctly rounded float32 power, computed through float64."""
    return np.power(np.asarray(x, dtype=F32).astype(np.float64), float(F32(e))).astype(F32)


def percentile_linear(values_f32, q):
    """np.percentile(float32 1-D array, q) of numpy 1.26: returns a python float (float64 arithmetic on float32 order statistics)."""
    a = np.sort(np.asarray(values_f32, dtype=F32).ravel())
    n = a.size
    qq = q / 100.0                                                  # np.true_divide(q, 100)
    virt = n * qq + (1.0 + qq * (1.0 - 1.0 - 1.0)) - 1.0            # _compute_virtual_index(n, quantiles, alpha=1, beta=1)
    prev = math.floor(virt)
    nxt = prev + 1
    prev_i, nxt_i = min(max(prev, 0), n - 1), min(max(nxt, 0), n - 1)
    gamma = virt - prev
    lo, hi = a[prev_i], a[nxt_i]
    diff = F32(hi - lo)                                             # subtract(b, a) on float32 operands
    return float(hi) - float(diff) * (1.0 - gamma) if gamma >= 0.5 else float(lo) + float(diff) * gamma     # _lerp


def gray_r_bytes():
    """(lut * 255).astype(uint8), lut = matplotlib.colors._create_lookup_table(256, [(0., 1, 1), (1., 0, 0)], gamma=1) -- channel 0 of 'gray_r'."""
    N = 256
    x = np.array([0.0, 1.0]) * (N - 1)
    y0 = np.array([1.0, 0.0])
    y1 = np.array([1.0, 0.0])
    xind = (N - 1) * np.linspace(0, 1, N) ** 1.0
    ind = np.searchsorted(x, xind)[1:-1]
    distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.concatenate([[y1[0]], distance * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]])
    return (np.clip(lut, 0.0, 1.0) * 255).astype(np.uint8)


def colorize_gray_r(value):
    """misc.py:97-150 with vmin=vmax=None, cmap='gray_r', channel 0."""
    value = np.asarray(value, dtype=F32).squeeze().copy()
    invalid = value == -99
    mask = ~invalid
    vmin = percentile_linear(value[mask], 2)
    vmax = percentile_linear(value[mask], 85)
    if vmin != vmax:
        value = (value - F32(vmin)) / F32(vmax - vmin)              # numpy 1.x: float64 scalars are cast to the array's float32
    else:
        value = value * F32(0.)
    value[invalid] = np.nan
    xa = value * F32(256)                                           # Colormap.__call__: xa *= self.N
    bad = np.isnan(xa)
    with np.errstate(invalid='ignore'):
        idx = np.where(xa == 256, 255, xa)
        under, over = idx < 0, idx >= 256
        idx = np.where(bad, 0, idx).astype(np.int64)
    idx[under] = 0                                                  # _i_under colour == lut[0]
    idx[over] = 255                                                 # _i_over colour == lut[N-1]
    out = gray_r_bytes()[np.clip(idx, 0, 255)]
    out[bad] = 0                                                    # _i_bad = (0,0,0,0)
    out[invalid] = 128                                              # background_color
    return out


def focal_plane_range(depth8, masks):
    """kenburns_effect.py:1045-1059"""
    start, end = 0, 255
    if masks is not None and len(masks) > 0:
        end = -1
        for m in masks:
            sel = depth8[np.asarray(m, dtype=bool)]
            dm = np.median(sel) if sel.size else float('nan')
            if dm > end:
                end = dm
        start = 255 if abs(255 - end) > abs(0 - end) else 0
    return float(start), float(end)


def focal_plane(fltStep, dof_speed, start, end):
    focal_int = 1 / (1 + np.exp((0.5 - fltStep) * dof_speed))       # :1065
    return float(focal_int * end + (1 - focal_int) * start)         # :1066


def bokeh_pass(img_flat, depth_flat, dx, dy, h, w, nsamples):
    img = np.ascontiguousarray(img_flat, dtype=F32).ravel()
    dep = np.ascontiguousarray(depth_flat, dtype=F32).ravel()
    out = np.empty_like(img)
    lib = kb_oracle.lib()
    lib.orc_bokeh_pass(C.c_int(h * w), C.c_int(h), C.c_int(w), C.c_int(nsamples), C.c_float(F32(dx)), C.c_float(F32(dy)),
                       img.ctypes.data_as(C.c_void_p), dep.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def blur_radius(depth8, focal_plane_value, depth_factor):
    """effects.py:146-154,163-164 -> float32 [H,W]"""
    depth = np.asarray(depth8).astype(F32)
    depth = depth.max() - np.abs(depth - F32(focal_plane_value))
    if depth_factor != 1:
        depth = pow_f32(depth, depth_factor)
    depth = depth - depth.min()
    with np.errstate(invalid='ignore', divide='ignore'):
        depth = depth.astype(F32) / depth.max()
    depth = F32(1) - depth
    return depth * F32(0.0005)


def bokeh_blur(img_u8, depth8, num_samples=32, lightness_factor=10, depth_factor=2, focal_plane=None, return_stages=False):
    """effects.py:143-182 with use_cuda=True."""
    img = np.ascontiguousarray(img_u8)
    h, w = img.shape[:2]
    depth = blur_radius(depth8, focal_plane, depth_factor)
    imgf = img.astype(F32) / F32(255)
    hl = pow_f32(imgf, lightness_factor)
    planar = np.ascontiguousarray(hl.transpose(2, 0, 1)).reshape(-1)        # np2flatten_tensor: [1,3,HW]
    dflat = depth.reshape(-1)
    PI = math.pi
    v = bokeh_pass(planar, dflat, 0, 1, h, w, num_samples)
    dg = bokeh_pass(v, dflat, math.cos(-PI / 6), math.sin(-PI / 6), h, w, num_samples)
    rh = bokeh_pass(dg, dflat, math.cos(-PI * 5 / 6), math.sin(-PI * 5 / 6), h, w, num_samples)
    bl = ((dg + rh) / F32(2)).reshape(3, h * w).transpose(1, 0).reshape(h, w, 3)        # ftensor2img
    with np.errstate(invalid='ignore'):
        res = pow_f32(bl, F32(1 / lightness_factor))
        out = (res * F32(255)).astype(np.uint8)
    if return_stages:
        return out, dict(radius=depth, vertical=v, diag=dg, rhom=rh, pre_u8=res * F32(255))
    return out
