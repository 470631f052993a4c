"""Test-only driver for the UNMODIFIED reference kernels compiled by oracle/build_ref_kernels.py
(oracle/_ref/libref_kernels.so).  Reproduces the reference host code around the launches
(anime_3dkenburns/models/utils.py:56-62,315 and common.py:146-148) with torch tensors on the GPU."""
import ctypes as C
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_kernels.so")
RENDER_SHAPES = [(64, 96, 64 * 96, 3), (128, 160, 128 * 160 + 3000, 4), (256, 256, 256 * 256, 4), (256, 256, 256 * 256 + 5000, 4),
                 (1024, 1024, 1024 * 1024, 4), (1024, 1024, 1024 * 1024, 3)]
FILL_SHAPES = [(64, 96, 4), (128, 160, 4), (256, 256, 4), (1024, 1024, 4)]


def available():
    return os.path.exists(_SO)


_lib = None


def _launch(entry, n, tensors):
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    getattr(_lib, "launch_" + entry)(C.c_int(n), arr, C.c_void_p(torch.cuda.current_stream().cuda_stream))


def render_pointcloud(tenInput, tenData, intWidth, intHeight, stages=False):
    H, W, N, Cc = intHeight, intWidth, tenInput.shape[2], tenData.shape[1]
    assert (H, W, N, Cc) in RENDER_SHAPES and tenInput.shape[0] == 1
    sfx = f"H{H}_W{W}_N{N}_C{Cc}"
    tenInput = tenInput.contiguous()
    tenData = torch.cat([tenData, tenData.new_ones([1, 1, N])], 1).contiguous()
    tenZee = tenInput.new_zeros([1, 1, H, W]).fill_(1000000.0)
    tenOutput = tenInput.new_zeros([1, Cc + 1, H, W])
    _launch(f"kernel_pointrender_updateZee_{sfx}", N, [tenInput, tenData, tenZee])
    zee_pre = tenZee.clone()
    _launch(f"kernel_pointrender_updateDegrid_{sfx}", tenZee.nelement(), [tenInput, tenData, tenZee])
    _launch(f"kernel_pointrender_updateOutput_{sfx}", N, [tenInput, tenData, tenZee, tenOutput])
    render, existing = tenOutput[:, :-1] / (tenOutput[:, -1:] + 0.0000001), tenOutput[:, -1:].detach().clone()
    if stages:
        return render, existing, zee_pre, tenZee
    return render, existing


def fill_disocclusion(tenInput, tenDepth):
    _, Cc, H, W = tenInput.shape
    assert (H, W, Cc) in FILL_SHAPES
    tenInput, tenDepth = tenInput.contiguous(), tenDepth.contiguous()
    tenOutput = tenInput.clone()
    _launch(f"kernel_discfill_updateOutput_H{H}_W{W}_C{Cc}", H * W, [tenInput, tenDepth, tenOutput])
    return tenOutput


def bokeh_filter(img, depth, dx, dy, im_h, im_w, num_samples=32):
    """utils/effects.py:12-84 (bokeh_filter_cupy) around the unmodified kernel_bokeh: img [1,3,HW], depth [1,1,HW] fp32 on the GPU."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
    img, depth = img.contiguous(), depth.contiguous()
    blurred = img.clone()
    arr = (C.c_void_p * 3)(img.data_ptr(), depth.data_ptr(), blurred.data_ptr())
    _lib.launch_kernel_bokeh(C.c_int(im_h * im_w), C.c_int(im_h), C.c_int(im_w), C.c_int(num_samples), C.c_float(dx), C.c_float(dy), arr,
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return blurred
