"""B200 implementations of the reference ops in `anime_3dkenburns/common.py` (same names, same arguments).

  process_shift      reference :59-83
  process_autozoom   reference :86-141  (256 renders + 256-512 host syncs -> one batched coverage call, one D2H)
  fill_disocclusion  reference :145-247
"""
import ctypes as C

import numpy as np
import torch

from .._lib import check, lib, ptr, stream, f3
from .models.utils import _f32


def shift_scalars(objSettings, objCommon):
    """Host scalar part of process_shift (reference :60-72, Python double arithmetic) -> (fltShiftX, fltShiftY, fltShiftZ)."""
    dr = objCommon['objDepthrange']
    fltClosestDepth = dr[0] + (objSettings['fltDepthTo'] - objSettings['fltDepthFrom'])
    fltClosestFromU, fltClosestFromV = dr[2][0], dr[2][1]
    fltClosestToU = fltClosestFromU + objSettings['fltShiftU']
    fltClosestToV = fltClosestFromV + objSettings['fltShiftV']
    W, H, f = objCommon['intWidth'], objCommon['intHeight'], objCommon['fltFocal']
    fltClosestFromX = ((fltClosestFromU - (W / 2.0)) * fltClosestDepth) / f
    fltClosestFromY = ((fltClosestFromV - (H / 2.0)) * fltClosestDepth) / f
    fltClosestToX = ((fltClosestToU - (W / 2.0)) * fltClosestDepth) / f
    fltClosestToY = ((fltClosestToV - (H / 2.0)) * fltClosestDepth) / f
    return fltClosestFromX - fltClosestToX, fltClosestFromY - fltClosestToY, objSettings['fltDepthTo'] - objSettings['fltDepthFrom']


def process_shift(objSettings, objCommon):
    """-> (tenPoints [B,3,N], tenShift [1,3,1])"""
    s = np.array(shift_scalars(objSettings, objCommon), np.float32)
    pts = _f32(objSettings['tenPoints'])
    B, _, N = pts.shape
    out = torch.empty_like(pts)
    check(lib().csb_points_shift(ptr(pts), B, N, f3(s), ptr(out), stream()), "csb_points_shift")
    return out, torch.from_numpy(s).view(1, 3, 1).to(pts.device)


def autozoom_candidates(objSettings, objCommon):
    """The in-frame candidate (ΔU, ΔV) list of process_autozoom in the reference's scan order (:99-114)."""
    npyShiftU = np.linspace(-objSettings['fltShift'], objSettings['fltShift'], 16)[None, :].repeat(16, 0)
    npyShiftV = np.linspace(-objSettings['fltShift'], objSettings['fltShift'], 16)[:, None].repeat(16, 1)
    fltCropWidth = objSettings['objFrom']['intCropWidth'] / objSettings['fltZoom']
    fltCropHeight = objSettings['objFrom']['intCropHeight'] / objSettings['fltZoom']
    cands = []
    for intU in range(16):
        for intV in range(16):
            fltShiftU = npyShiftU[intU, intV].item()
            fltShiftV = npyShiftV[intU, intV].item()
            if objSettings['objFrom']['fltCenterU'] + fltShiftU < fltCropWidth / 2.0:
                continue
            elif objSettings['objFrom']['fltCenterU'] + fltShiftU > objCommon['intWidth'] - (fltCropWidth / 2.0):
                continue
            elif objSettings['objFrom']['fltCenterV'] + fltShiftV < fltCropHeight / 2.0:
                continue
            elif objSettings['objFrom']['fltCenterV'] + fltShiftV > objCommon['intHeight'] - (fltCropHeight / 2.0):
                continue
            cands.append((fltShiftU, fltShiftV))
    return cands, fltCropWidth


def autozoom_coverage(tenPoints, shifts, intWidth, intHeight, fltFocal, fltBaseline):
    """Coverage count (tenExisting > 0).sum() of the render for each 3-float shift -> int32 tensor [S] on the device."""
    pts = _f32(tenPoints)
    assert pts.shape[0] == 1
    N = pts.shape[2]
    S = len(shifts)
    dev = pts.device
    zkey = torch.empty((S, intHeight, intWidth), device=dev, dtype=torch.int32)
    zee = torch.empty((S, intHeight, intWidth), device=dev, dtype=torch.float32)
    cover = torch.empty((S, intHeight, intWidth), device=dev, dtype=torch.uint8)
    counts = torch.empty((S,), device=dev, dtype=torch.int32)
    sh = np.ascontiguousarray(np.asarray(shifts, np.float32).reshape(S, 3))
    check(lib().csb_autozoom_coverage(ptr(pts), N, intHeight, intWidth, C.c_double(fltFocal), C.c_double(fltBaseline),
                                      sh.ctypes.data_as(C.c_void_p), S, ptr(zkey), ptr(zee), ptr(cover), ptr(counts), stream()),
          "csb_autozoom_coverage")
    return counts


def process_autozoom(objSettings, objCommon):
    cands, fltCropWidth = autozoom_candidates(objSettings, objCommon)
    fltDepthFrom = objCommon['objDepthrange'][0]
    fltDepthTo = objCommon['objDepthrange'][0] * (fltCropWidth / objSettings['objFrom']['intCropWidth'])
    fltBestU = fltBestV = None
    if cands:
        shifts = [np.array(shift_scalars({'fltShiftU': u, 'fltShiftV': v, 'fltDepthFrom': fltDepthFrom, 'fltDepthTo': fltDepthTo}, objCommon),
                           np.float32) for (u, v) in cands]
        counts = autozoom_coverage(objCommon['tenRawPoints'], shifts, objCommon['intWidth'], objCommon['intHeight'],
                                   objCommon['fltFocal'], objCommon['fltBaseline']).cpu().numpy()      # the one host read
        fltBest = 0.0
        for (u, v), c in zip(cands, counts.tolist()):
            if fltBest < float(c):                     # strict '<': first maximum in scan order wins (:128)
                fltBest, fltBestU, fltBestV = float(c), u, v
    return {
        'fltCenterU': objSettings['objFrom']['fltCenterU'] + fltBestU,
        'fltCenterV': objSettings['objFrom']['fltCenterV'] + fltBestV,
        'intCropWidth': int(round(objSettings['objFrom']['intCropWidth'] / objSettings['fltZoom'])),
        'intCropHeight': int(round(objSettings['objFrom']['intCropHeight'] / objSettings['fltZoom']))
    }


def fill_disocclusion(tenInput, tenDepth):
    x, d = _f32(tenInput), _f32(tenDepth)
    B, Cc, H, W = x.shape
    out = torch.empty_like(x)
    check(lib().csb_disocclusion_fill(ptr(x), ptr(d), B, Cc, H, W, ptr(out), stream()), "csb_disocclusion_fill")
    return out
