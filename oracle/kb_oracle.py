"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of oracle/kb_oracle.c (the CPU restatement of the
reference's Ken-Burns point-cloud kernels; see the C file's header for file:line citations).

`build()` compiles the C file with gcc (`-O2 -ffp-contract=off`: no implicit FMA, the explicit ones mirror the
reference SASS).  Function names and argument order mirror the reference Python ops
(`render_pointcloud`, `fill_disocclusion`, `process_shift`, `depth_to_points`, `spatial_filter`).
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkb_oracle.so")
_SRC = os.path.join(_HERE, "kb_oracle.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-shared",
                               "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_count_positive.restype = C.c_long
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def render_pointcloud(tenInput, tenData, intWidth, intHeight, fltFocal, fltBaseline, return_zee=False):
    """models/utils.py:56-315 -> (render[B,C,H,W], existing[B,1,H,W])"""
    pts, pp = _f(tenInput); dat, dp = _f(tenData)
    B, _, N = pts.shape; Cc = dat.shape[1]
    render = np.empty((B, Cc, intHeight, intWidth), np.float32); existing = np.empty((B, 1, intHeight, intWidth), np.float32)
    z0 = np.empty((B, 1, intHeight, intWidth), np.float32); z1 = np.empty_like(z0)
    lib().orc_render_pointcloud(pp, dp, B, N, Cc, intHeight, intWidth, C.c_double(fltFocal), C.c_double(fltBaseline),
                                render.ctypes.data_as(C.c_void_p), existing.ctypes.data_as(C.c_void_p),
                                z0.ctypes.data_as(C.c_void_p), z1.ctypes.data_as(C.c_void_p))
    if return_zee:
        return render, existing, z0, z1
    return render, existing


def render_zpass(tenInput, intWidth, intHeight, fltFocal, fltBaseline):
    pts, pp = _f(tenInput); B, _, N = pts.shape
    z = np.empty((B, 1, intHeight, intWidth), np.float32)
    lib().orc_render_zpass(pp, B, N, intHeight, intWidth, C.c_double(fltFocal), C.c_double(fltBaseline), z.ctypes.data_as(C.c_void_p))
    return z


def render_degrid(zee):
    z, zp = _f(zee); B, _, H, W = z.shape
    o = np.empty_like(z)
    lib().orc_render_degrid(zp, B, H, W, o.ctypes.data_as(C.c_void_p))
    return o


def count_positive(a):
    a, ap = _f(a)
    return int(lib().orc_count_positive(ap, C.c_long(a.size)))


def fill_disocclusion(tenInput, tenDepth):
    """common.py:145-247"""
    x, xp = _f(tenInput); d, dp = _f(tenDepth)
    B, Cc, H, W = x.shape
    out = np.empty_like(x)
    lib().orc_fill_disocclusion(xp, dp, B, Cc, H, W, out.ctypes.data_as(C.c_void_p))
    return out


def shift_scalars(objSettings, objCommon):
    """Host scalar part of process_shift, common.py:60-72 (Python double arithmetic) -> (sx, sy, sz)."""
    dr = objCommon['objDepthrange']
    fltClosestDepth = dr[0] + (objSettings['fltDepthTo'] - objSettings['fltDepthFrom'])
    fu, fv = dr[2][0], dr[2][1]
    tu, tv = fu + objSettings['fltShiftU'], fv + objSettings['fltShiftV']
    W, H, f = objCommon['intWidth'], objCommon['intHeight'], objCommon['fltFocal']
    fx = ((fu - (W / 2.0)) * fltClosestDepth) / f; fy = ((fv - (H / 2.0)) * fltClosestDepth) / f
    tx = ((tu - (W / 2.0)) * fltClosestDepth) / f; ty = ((tv - (H / 2.0)) * fltClosestDepth) / f
    return fx - tx, fy - ty, objSettings['fltDepthTo'] - objSettings['fltDepthFrom']


def process_shift(objSettings, objCommon):
    """common.py:59-83 -> (tenPoints, tenShift[1,3,1])"""
    s = np.array(shift_scalars(objSettings, objCommon), np.float32)
    pts, pp = _f(objSettings['tenPoints']); B, _, N = pts.shape
    out = np.empty_like(pts)
    lib().orc_process_shift(pp, B, N, s.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out, s.reshape(1, 3, 1)


def depth_to_points(tenDepth, fltFocal):
    d, dp = _f(tenDepth); B, _, H, W = d.shape
    out = np.empty((B, 3, H, W), np.float32)
    lib().orc_depth_to_points(dp, B, H, W, C.c_double(fltFocal), out.ctypes.data_as(C.c_void_p))
    return out


def spatial_filter(tenInput, strType):
    x, xp = _f(tenInput); B, Cc, H, W = x.shape
    out = np.empty_like(x)
    if strType == 'laplacian':
        lib().orc_laplacian(xp, B, Cc, H, W, out.ctypes.data_as(C.c_void_p))
    elif strType in ('median-3', 'median-5'):
        lib().orc_median(xp, B, Cc, H, W, int(strType[-1]), out.ctypes.data_as(C.c_void_p))
    else:
        return None
    return out


def disparity_to_cloud(raw_disparity, fltFocal, fltBaseline):
    """kenburns_effect.py:928-937 -> dict(disparity, depth, valid, points, unaltered, dispmin, dispmax, depthrange)"""
    r, rp = _f(raw_disparity); H, W = r.shape[-2:]
    disp = np.empty((1, 1, H, W), np.float32); depth = np.empty_like(disp); valid = np.empty_like(disp)
    pts = np.empty((1, 3, H, W), np.float32); un = np.empty_like(pts); sc = np.zeros(8, np.float64)
    lib().orc_disparity_to_cloud(rp, H, W, C.c_double(fltFocal), C.c_double(fltBaseline), *[a.ctypes.data_as(C.c_void_p) for a in (disp, depth, valid, pts, un, sc)])
    return dict(disparity=disp, depth=depth, valid=valid, points=pts, unaltered=un, dispmin=float(sc[0]), dispmax=float(sc[1]),
                depthrange=(float(sc[2]), float(sc[3]), (int(sc[4]), int(sc[5])), (int(sc[6]), int(sc[7]))))


def frame_pack_u8(render):
    r, rp = _f(render); Cs, H, W = r.shape[-3:]
    out = np.empty((H, W, 3), np.uint8)
    lib().orc_frame_pack_u8(rp, Cs, H, W, out.ctypes.data_as(C.c_void_p))
    return out


def get_rect_sub_pix(frame, patchSize, center):
    src = np.ascontiguousarray(frame, np.uint8); H, W = src.shape[:2]; pw, ph = patchSize
    out = np.empty((ph, pw, 3), np.uint8)
    lib().orc_getrectsubpix_u8c3(src.ctypes.data_as(C.c_void_p), H, W, ph, pw, C.c_double(center[0]), C.c_double(center[1]), out.ctypes.data_as(C.c_void_p))
    return out


def resize_linear(frame, dsize):
    src = np.ascontiguousarray(frame, np.uint8); sh, sw = src.shape[:2]; dw, dh = dsize
    out = np.empty((dh, dw, 3), np.uint8)
    lib().orc_resize_linear_u8c3(src.ctypes.data_as(C.c_void_p), sh, sw, dh, dw, out.ctypes.data_as(C.c_void_p))
    return out
