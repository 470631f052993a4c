"""B200 implementation of the reference's `anime_3dkenburns/kenburns_effect.py` hot path.

Functional layer (this part of the file): the tensor math of `generate_kenburns_config` (reference :928-947) and the body
of the `process_kenburns` frame loop (reference :1028-1040, 1069-1070) on the GPU through the C ABI.
"""
import ctypes as C

import numpy as np
import torch

from .._lib import check, lib, ptr, stream, f3
from .models.utils import RenderScratch, _f32, split_shift


def disparity_to_cloud(tenDisparity, fltFocal, fltBaseline, image_u8=None):
    """reference :928-937 fused: raw disparity [1,1,H,W] -> dict(disparity, depth, valid, points [1,3,H,W], unaltered,
    dispmin, dispmax, depthrange=(min, max, (x,y) of min, (x,y) of max) on depth[128:-128,128:-128]).  One D2H of 8 floats."""
    raw = _f32(tenDisparity)
    assert raw.shape[0] == 1                                   # reference asserts batch 1 (:40)
    H, W = raw.shape[-2:]
    dev = raw.device
    out = {k: torch.empty((1, c, H, W), device=dev, dtype=torch.float32)
           for k, c in (('disparity', 1), ('depth', 1), ('valid', 1), ('points', 3), ('unaltered', 3))}
    scalars = torch.empty(8, device=dev, dtype=torch.float32)
    scratch = torch.empty(64, device=dev, dtype=torch.int64)
    if image_u8 is not None:      # optional fused payload: BGR/255 planar ++ depth, [1,4,H*W]
        assert image_u8.dtype == torch.uint8 and tuple(image_u8.shape) == (H, W, 3)
        out['data'] = torch.empty((1, 4, H * W), device=dev, dtype=torch.float32)
    check(lib().csb_disparity_to_cloud(ptr(raw), H, W, C.c_double(fltFocal), C.c_double(fltBaseline), ptr(out['disparity']), ptr(out['depth']),
                                       ptr(out['valid']), ptr(out['points']), ptr(out['unaltered']), ptr(scalars), ptr(scratch),
                                       ptr(image_u8), ptr(out.get('data')), stream()),
          "csb_disparity_to_cloud")
    sc = scalars.cpu().numpy()
    out['dispmin'], out['dispmax'] = float(sc[0]), float(sc[1])
    out['depthrange'] = (float(sc[2]), float(sc[3]), (int(sc[4]), int(sc[5])), (int(sc[6]), int(sc[7])))
    out['scalars'] = scalars
    return out


def shift_from_scalars(scalars, intWidth, intHeight, fltFocal, fltShiftU, fltShiftV, depth_ratio, out=None):
    """process_shift's scalar math (reference common.py:60-72) on the device -> CUDA tensor of 3 floats; no host sync."""
    if out is None:
        out = torch.empty(3, device=scalars.device, dtype=torch.float32)
    check(lib().csb_shift_from_scalars(ptr(scalars), intWidth, intHeight, C.c_double(fltFocal), C.c_double(fltShiftU), C.c_double(fltShiftV),
                                       C.c_double(depth_ratio), ptr(out), stream()), "csb_shift_from_scalars")
    return out


def frame_pack_u8(tenRender):
    """reference :1040 -- (render[0,0:3].transpose(1,2,0) * 255).clip(0,255).astype(uint8), on the device -> [H,W,3] u8"""
    r = _f32(tenRender)
    r = r[0] if r.dim() == 4 else r
    H, W = r.shape[-2:]
    frame = torch.empty((H, W, 3), device=r.device, dtype=torch.uint8)
    check(lib().csb_frame_pack_u8(ptr(r), H, W, ptr(frame), stream()), "csb_frame_pack_u8")
    return frame


def frame_crop_resize(frame, pw, ph, cx, cy):
    """reference :1069-1070 -- cv2.getRectSubPix(patchSize=(pw,ph), center=(cx,cy)) + cv2.resize(INTER_LINEAR) to (W,H)"""
    H, W = frame.shape[:2]
    out = torch.empty_like(frame)
    check(lib().csb_frame_crop_resize(ptr(frame), H, W, int(pw), int(ph), C.c_double(cx), C.c_double(cy), ptr(out), stream()),
          "csb_frame_crop_resize")
    return out


class FrameScratch(RenderScratch):
    def __init__(self, H, W, device):
        super().__init__(1, 4, H, W, device)
        self.packed = torch.empty((H, W, 3), device=device, dtype=torch.uint8)


def kenburns_frame(tenPoints, tenData, intWidth, intHeight, fltFocal, fltBaseline, shift, pw, ph, cx, cy, scratch=None, out=None,
                   want_depth=False):
    """One output frame of the reference loop (:1028-1040,1069-1070): shift + render(C=4) + fill + pack + crop/resize.
    -> (frame [H,W,3] u8 on the device, depth [H,W] or None)"""
    pts, dat = _f32(tenPoints), _f32(tenData)
    assert pts.shape[0] == 1 and dat.shape[1] == 4
    N = pts.shape[2]
    H, W = intHeight, intWidth
    if scratch is None:
        scratch = FrameScratch(H, W, pts.device)
    if out is None:
        out = torch.empty((H, W, 3), device=pts.device, dtype=torch.uint8)
    depth = torch.empty((H, W), device=pts.device, dtype=torch.float32) if want_depth else None
    sh, sh_dev = split_shift(shift)
    check(lib().csb_kenburns_frame(ptr(pts), ptr(dat), N, H, W, C.c_double(fltFocal), C.c_double(fltBaseline),
                                   sh, sh_dev, int(pw), int(ph), C.c_double(cx), C.c_double(cy),
                                   ptr(scratch.zkey), ptr(scratch.zee), ptr(scratch.acc), ptr(scratch.packed), ptr(out), ptr(depth), stream()),
          "csb_kenburns_frame")
    return out, depth
