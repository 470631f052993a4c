"""B200 implementations of the reference ops in `anime_3dkenburns/models/utils.py` (same names, same arguments).

  spatial_filter     reference :9-40
  depth_to_points    reference :43-50
  render_pointcloud  reference :56-315  (three cupy kernels -> libcsb200 csb_pointcloud_render)

Tensors must be CUDA fp32; the work is enqueued on torch's current stream through the C ABI (include/csb200.h).
"""
import ctypes as C

import torch

from ..._lib import check, lib, ptr, stream, f3

_KIND = {'laplacian': 0, 'median-3': 3, 'median-5': 5}


def _f32(t):
    return t.contiguous().float()


def spatial_filter(tenInput, strType):
    if strType not in _KIND:
        return None                                  # reference returns None for unknown types (:10,:40)
    x = _f32(tenInput)
    B, Cc, H, W = x.shape
    out = torch.empty_like(x)
    check(lib().csb_spatial_filter(ptr(x), B, Cc, H, W, _KIND[strType], ptr(out), stream()), "csb_spatial_filter")
    return out


def depth_to_points(tenDepth, fltFocal):
    d = _f32(tenDepth)
    B, _, H, W = d.shape
    out = torch.empty((B, 3, H, W), device=d.device, dtype=torch.float32)
    check(lib().csb_depth_to_points(ptr(d), B, H, W, C.c_double(fltFocal), ptr(out), stream()), "csb_depth_to_points")
    return out


def split_shift(tenShift):
    """-> (host float[3] or None, device pointer or None)"""
    if tenShift is None:
        return None, None
    if torch.is_tensor(tenShift) and tenShift.is_cuda:
        t = tenShift.contiguous().float()
        assert t.numel() == 3
        return None, ptr(t)
    return f3([float(v) for v in torch.as_tensor(tenShift).flatten().tolist()]), None


class RenderScratch:
    """Caller-owned scratch of csb_pointcloud_render, reusable across frames of the same shape."""

    def __init__(self, B, C_, H, W, device):
        self.key = (B, C_, H, W, str(device))
        cp = lib().csb_render_acc_channels(C_)
        self.zkey = torch.empty((B, H, W), device=device, dtype=torch.int32)
        self.zee = torch.empty((B, 1, H, W), device=device, dtype=torch.float32)
        self.acc = torch.empty((B, H, W, cp), device=device, dtype=torch.float32)


def render_pointcloud(tenInput, tenData, intWidth, intHeight, fltFocal, fltBaseline, tenShift=None, scratch=None):
    """-> (render [B,C,H,W], existing [B,1,H,W]).  `tenShift` optionally folds process_shift in: 3 host floats, or a
    CUDA tensor of 3 floats (read on the device, no host sync)."""
    pts, dat = _f32(tenInput), _f32(tenData)
    B, three, N = pts.shape
    assert three == 3 and dat.shape[0] == B and dat.shape[2] == N
    Cc = dat.shape[1]
    if scratch is None or scratch.key != (B, Cc, intHeight, intWidth, str(pts.device)):
        scratch = RenderScratch(B, Cc, intHeight, intWidth, pts.device)
    render = torch.empty((B, Cc, intHeight, intWidth), device=pts.device, dtype=torch.float32)
    existing = torch.empty((B, 1, intHeight, intWidth), device=pts.device, dtype=torch.float32)
    sh, sh_dev = split_shift(tenShift)
    check(lib().csb_pointcloud_render(ptr(pts), ptr(dat), B, N, Cc, intHeight, intWidth, C.c_double(fltFocal), C.c_double(fltBaseline),
                                      sh, sh_dev, ptr(scratch.zkey), ptr(scratch.zee), ptr(scratch.acc), ptr(render), ptr(existing), stream()),
          "csb_pointcloud_render")
    return render, existing


def render_zpass(tenInput, intWidth, intHeight, fltFocal, fltBaseline):
    """kernel_pointrender_updateZee alone -> zee [B,1,H,W] before degrid (test hook)."""
    pts = _f32(tenInput)
    B, _, N = pts.shape
    zkey = torch.empty((B, intHeight, intWidth), device=pts.device, dtype=torch.int32)
    check(lib().csb_pointcloud_zpass(ptr(pts), B, N, intHeight, intWidth, C.c_double(fltFocal), C.c_double(fltBaseline), None, ptr(zkey), stream()),
          "csb_pointcloud_zpass")
    k = zkey
    bits = k ^ ((k >> 31) & 0x7fffffff)
    return torch.minimum(bits.view(torch.float32), torch.tensor(1000000.0, device=pts.device)).view(B, 1, intHeight, intWidth), zkey


def render_degrid(zkey):
    B, H, W = zkey.shape
    zee = torch.empty((B, 1, H, W), device=zkey.device, dtype=torch.float32)
    check(lib().csb_pointcloud_degrid(ptr(zkey), B, H, W, ptr(zee), stream()), "csb_pointcloud_degrid")
    return zee


def tensor_stats(x):
    """{mean, population std, max} of a contiguous fp32 CUDA tensor on the device -> float32 [3] (csb_tensor_stats; no host read)."""
    x = _f32(x)
    scratch = torch.empty(3, device=x.device, dtype=torch.float64)
    out = torch.empty(3, device=x.device, dtype=torch.float32)
    check(lib().csb_tensor_stats(ptr(x), C.c_longlong(x.numel()), ptr(scratch), ptr(out), stream()), "csb_tensor_stats")
    return out


def pack_norm16(a, stats_a, b=None, stats_b=None, eps=0.0000001):
    """[1,ca,H,W] (+ [1,cb,H,W]) fp32 -> [1,H,W,16] fp16 = [(a - mean) / (std + eps) | (b - mean) / (std + eps) | zeros] (csb_pack_norm16)."""
    a = _f32(a)
    _, ca, H, W = a.shape
    cb = 0
    if b is not None:
        b = _f32(b)
        cb = b.shape[1]
    out = torch.empty((1, H, W, 16), device=a.device, dtype=torch.float16)
    check(lib().csb_pack_norm16(ptr(a), ca, ptr(stats_a), ptr(b), cb, ptr(stats_b), C.c_longlong(H * W), C.c_float(eps), ptr(out), stream()), "csb_pack_norm16")
    return out


def net_output(a_nhwc, b_nhwc, stats, post, eps=0.0000001):
    """(a [+ b]) [1,H,W,C] fp32 -> [1,C,H,W] fp32 = post(x * (std + eps) + mean); post: 0 none, 1 clip to [0,1], 2 threshold at 0 (csb_net_output)."""
    _, H, W, Cc = a_nhwc.shape
    out = torch.empty((1, Cc, H, W), device=a_nhwc.device, dtype=torch.float32)
    check(lib().csb_net_output(ptr(a_nhwc), ptr(b_nhwc), Cc, C.c_longlong(H * W), ptr(stats), C.c_float(eps), post, ptr(out), stream()), "csb_net_output")
    return out
