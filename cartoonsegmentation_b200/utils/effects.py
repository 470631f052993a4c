"""Bokeh depth-of-field on the B200 (reference `utils/effects.py:12-182`, frame-loop glue `kenburns_effect.py:1042-1067`,
`colorize` `depth_modules/zoedepth/utils/misc.py:97-150`).

The reference moves every frame to the host (numpy pow / percentile / matplotlib LUT / np.median), uploads two planes for three cupy launches and
downloads the result.  Here the whole chain stays on the device behind three C-ABI calls (`csb_depth_colorize_u8`, `csb_focal_plane_range`,
`csb_bokeh_blur`); the host only builds two 256-entry constant tables with numpy, by the same expressions the reference evaluates per pixel.
"""
import ctypes as C
import math

import numpy as np
import torch

from .._lib import check, lib, ptr, stream


def gray_r_bytes():
    """Channel 0 of matplotlib's 'gray_r' colormap as `Colormap.__call__(bytes=True)` sees it: `(lut * 255).astype(uint8)` with
    lut = `_create_lookup_table(256, [(0, 1, 1), (1, 0, 0)])` (matplotlib 3.9 `colors.py`; matplotlib is not in this image, the construction is
    restated): x = [0, 255]; xind = 255 * linspace(0, 1, 256); lut[1:-1] = (xind / 255) * (0 - 1) + 1; ends 1 and 0; clip to [0, 1]."""
    N = 256
    xind = (N - 1) * np.linspace(0, 1, N)
    distance = (xind[1:-1] - 0.0) / (255.0 - 0.0)
    lut = np.concatenate([[1.0], distance * (0.0 - 1.0) + 1.0, [0.0]])
    return (np.clip(lut, 0.0, 1.0) * 255).astype(np.uint8)


def highlight_table(lightness_factor):
    """`np.power(img.astype(float32) / 255, lightness_factor)` (effects.py:156-157) for the 256 possible pixel values.  numpy's float32 power
    is the platform's powf (glibc: correctly rounded; SVML builds: within 1 ulp); the table holds the correctly rounded value (via float64) so
    that it does not depend on the host."""
    x = np.arange(256).astype(np.float32) / np.float32(255)
    return np.power(x.astype(np.float64), float(np.float32(lightness_factor))).astype(np.float32)


class BokehScratch:
    def __init__(self, H, W, K, device, lightness_factor=13):
        lib().csb_bokeh_workspace_bytes.restype = C.c_size_t
        self.key = (H, W, K, lightness_factor)
        self.ws = torch.empty(int(lib().csb_bokeh_workspace_bytes(H, W, K)), device=device, dtype=torch.uint8)
        self.lut = torch.from_numpy(gray_r_bytes()).to(device)
        self.hl = torch.from_numpy(highlight_table(lightness_factor)).to(device)
        self.inv_light = float(np.float32(1 / lightness_factor))                   # np.power(float32 array, python float) runs in float32
        self.range = torch.zeros(2, device=device, dtype=torch.float64)
        self.depth8 = torch.empty((H, W), device=device, dtype=torch.uint8)


def colorize_gray_r(depth, scratch, out=None):
    """`colorize(depth, cmap='gray_r')[..., 0]` -> [H,W] u8 on the device."""
    H, W = depth.shape[-2:]
    d = depth.reshape(H, W).contiguous().float()
    out = scratch.depth8 if out is None else out
    check(lib().csb_depth_colorize_u8(ptr(d), H, W, ptr(scratch.lut), ptr(out), ptr(scratch.ws), stream()), "csb_depth_colorize_u8")
    return out


def focal_plane_range(depth8, masks, scratch):
    """(focalplane_start, focalplane_end) of kenburns_effect.py:1045-1059 as two device doubles in `scratch.range`."""
    H, W = depth8.shape
    K = 0 if masks is None else int(masks.shape[0])
    m = None
    if K:
        m = masks.reshape(K, H, W).contiguous()
        m = m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)
    check(lib().csb_focal_plane_range(ptr(depth8), ptr(m), K, H, W, ptr(scratch.range), ptr(scratch.ws), stream()), "csb_focal_plane_range")
    return scratch.range


def bokeh_blur(img, depth, num_samples=32, lightness_factor=10, depth_factor=2, use_cuda=True, focal_plane=None, scratch=None, focal_int=None,
               out=None):
    """Reference signature (effects.py:143).  img [H,W,3] u8 and depth [H,W] (8-bit values) as device tensors or numpy arrays; returns the same
    kind.  Only the reference's use_cuda=True arithmetic exists here (its numba path indexes the image differently, effects.py:101-138).
    `focal_int` (with `scratch.range` filled by `focal_plane_range`) evaluates the focal plane on the device instead of taking a host float."""
    if not use_cuda:
        raise NotImplementedError("the CPU (numba) variant of bokeh_blur is outside the B200 path")
    if focal_plane is None and focal_int is None:
        raise NotImplementedError("bokeh_blur without a focal plane is not used by the pipeline (kenburns_effect.py:1067 always passes one)")
    as_numpy = isinstance(img, np.ndarray)
    dev = torch.device('cuda') if as_numpy else img.device
    frame = torch.from_numpy(np.ascontiguousarray(img)).to(dev) if as_numpy else img.contiguous()
    d8 = torch.from_numpy(np.ascontiguousarray(depth)).to(dev) if isinstance(depth, np.ndarray) else depth
    d8 = d8.to(torch.uint8).contiguous()
    H, W = frame.shape[:2]
    if scratch is None or scratch.key[:2] != (H, W) or scratch.key[3] != lightness_factor:
        scratch = BokehScratch(H, W, 0, dev, lightness_factor)
    out = torch.empty_like(frame) if out is None else out
    check(lib().csb_bokeh_blur(ptr(frame), ptr(d8), H, W, int(num_samples), ptr(scratch.hl), C.c_float(scratch.inv_light),
                               ptr(scratch.range) if focal_int is not None else None, C.c_double(0.0 if focal_int is None else focal_int),
                               C.c_double(0.0 if focal_plane is None else float(focal_plane)), int(depth_factor), ptr(out), ptr(scratch.ws), stream()),
          "csb_bokeh_blur")
    return out.cpu().numpy() if as_numpy else out


def focal_interp(fltStep, dof_speed):
    """kenburns_effect.py:1065"""
    return float(1 / (1 + np.exp((0.5 - fltStep) * dof_speed)))
