"""B200 implementation of the reference's `anime_3dkenburns/kenburns_effect.py` hot path.

Functional layer (this part of the file): the tensor math of `generate_kenburns_config` (reference :928-947) and the body
of the `process_kenburns` frame loop (reference :1028-1040, 1069-1070) on the GPU through the C ABI.
"""
import ctypes as C
import os

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
    elif out is False:                      # stop after `scratch.packed` (the bokeh stage crops later)
        out = None
    depth = torch.empty((H, W), device=pts.device, dtype=torch.float32) if want_depth else None
    sh, sh_dev = split_shift(shift)
    check(lib().csb_kenburns_frame(ptr(pts), ptr(dat), N, H, W, C.c_double(fltFocal), C.c_double(fltBaseline),
                                   sh, sh_dev, int(pw), int(ph), C.c_double(cx), C.c_double(cy),
                                   ptr(scratch.zkey), ptr(scratch.zee), ptr(scratch.acc), ptr(scratch.packed), ptr(out), ptr(depth), stream()),
          "csb_kenburns_frame")
    return out, depth


# ================================================================================================================
# Call surface of the reference: KenBurnsConfig (reference :207-366), build_kenburns_cfg (:369-374), KenBurnsPipeline (:392-1081)
# ================================================================================================================
import math
from copy import copy as _shallow_copy
from dataclasses import dataclass, field, fields
from typing import Any, Optional, Union

from ..animeinsseg import AnimeInsSeg, AnimeInstances
from ..utils import effects as fx
from .common import process_autozoom, shift_scalars

_ALIASES = {'fltFocal': 'focal', 'fltBaseline': 'baseline', 'intWidth': 'int_width', 'intHeight': 'int_height', 'fltDispmin': 'disparity_min',
            'fltDispmax': 'disparity_max', 'objDepthrange': 'depth_range', 'tenRawImage': 'tensor_raw_image', 'tenRawDisparity': 'raw_disparity',
            'tenRawDepth': 'raw_depth', 'tenRawPoints': 'raw_point', 'tenRawUnaltered': 'raw_unaltered', 'tenInpaImage': 'inpainted_img',
            'tenInpaDisparity': 'inpainted_disparity', 'tenInpaDepth': 'inpainted_depth', 'tenInpaPoints': 'inpainted_points'}


@dataclass
class KenBurnsConfig:
    """Same fields and dict-style aliases (the upstream `objCommon` names) as the reference dataclass, kenburns_effect.py:207-366."""
    detector: str = 'animeinsseg'
    det_ckpt: str = 'models/AnimeInstanceSegmentation/rtmdetl_e60.ckpt'
    det_size: int = 640
    scale_depth: bool = False
    depth_field: bool = False
    mask_refine_kwargs: dict = field(default_factory=dict)
    marigold_kwargs: dict = field(default_factory=dict)
    pred_score_thr: float = 0.3
    depth_est: str = 'zoe'
    depth_est_device: str = ''
    depth_refinement: str = 'default'
    depthest_use_medium: bool = False
    inpaint_type: str = 'default'
    num_frame: int = 75
    playback: bool = True
    auto_zoom: bool = True
    focal: float = 1024 / 2.0
    baseline: float = 40.0
    dof_speed: float = 50.
    depth_factor: int = 1
    lightness_factor: int = 13
    max_size: int = 720
    int_height: int = 1024
    int_width: int = 1024
    default_depth_refine: bool = False
    refine_crf: bool = True
    depth_est_size: int = 640
    sd_img2img_url: str = 'http://127.0.0.1:7860/sdapi/v1/img2img'
    ldm_inpaint_options: dict = field(default_factory=dict)
    ldm_inpaint_size: int = 0
    instances: AnimeInstances = None
    # non-init state (class attributes in the reference)
    disparity_min = 0
    disparity_max = 0
    depth_range = None
    tensor_raw_image = None
    original_img_nparray = None
    raw_disparity = None
    raw_depth = None
    raw_point = None
    raw_unaltered = None
    inpainted_img = None
    inpainted_disparity = None
    inpainted_depth = None
    inpainted_points = None
    bg_prompt = None
    save_path = r''
    stage_depth_coarse = None
    stage_depth_adjusted = None
    stage_depth_final = None

    def __post_init__(self):
        # the reference keeps these as shared class-level lists that grow forever (:284-285); per-config here
        self.stage_inpainted_imgs = []
        self.stage_inpainted_masks = []

    def __getitem__(self, item: str):
        return getattr(self, _ALIASES.get(item, item))

    def __setitem__(self, item, value):
        setattr(self, _ALIASES.get(item, item), value)

    def copy(self):
        """Reference: deepcopy (:365).  Tensors are immutable inputs of the pipeline, so a shallow field copy is equivalent and
        avoids cloning hundreds of MB of device memory."""
        c = _shallow_copy(self)
        c.stage_inpainted_imgs, c.stage_inpainted_masks = list(self.stage_inpainted_imgs), list(self.stage_inpainted_masks)
        return c


def build_kenburns_cfg(tgt_cfg: Union[str, dict]):
    if isinstance(tgt_cfg, str):
        import yaml                                   # the reference uses OmegaConf.load (:370); plain YAML is what the shipped configs are
        with open(tgt_cfg) as f:
            tgt_cfg = dict(yaml.safe_load(f))
    fieldSet = {f.name for f in fields(KenBurnsConfig) if f.init}
    return KenBurnsConfig(**{k: v for k, v in tgt_cfg.items() if k in fieldSet})


def scaledown_maxsize(img: np.ndarray, max_size: int, divisior: int = None):
    """utils/io_utils.py:254-274: only ever shrinks; Python banker's round(); cv2 INTER_LINEAR (host side, as in the reference)."""
    import cv2
    im_h, im_w = img.shape[:2]
    ori_h, ori_w = im_h, im_w
    resize_ratio = max_size / max(im_h, im_w)
    if max_size < max(im_h, im_w):
        im_h, im_w = int(round(im_h * resize_ratio)), int(round(im_w * resize_ratio))
    if divisior is not None:
        im_w, im_h = int(round(im_w / divisior) * divisior), int(round(im_h / divisior) * divisior)
    if im_w != ori_w or im_h != ori_h:
        img = cv2.resize(img, (im_w, im_h), interpolation=cv2.INTER_LINEAR)
    return img


def depth_adjustment_animesseg(instances: AnimeInstances, tenDisparity, tenImage, use_medium=False):
    """reference :39-91 -- per instance, in order: flatten the disparity under the mask to the maximum found in the bottom 3% of its rows
    (use_medium: to the lower median of the masked positive disparities, :80).  All variants run on the device without a host sync (the reference
    does ~8 ATen kernels and 5 `.item()` syncs per instance): csb_depth_adjust_batch (one cooperative launch) / csb_depth_adjust_median (exact radix
    select); when disparity and image differ in size (:50-56, :86-90) the map goes through csb_resample_f32 (bilinear, align_corners=False) both ways."""
    assert tenDisparity.shape[0] == 1
    if not tenDisparity.is_cuda:
        raise RuntimeError("depth_adjustment_animesseg: CUDA tensors only (no CPU path)")
    from ..engine import resample_f32
    h, w = tenDisparity.shape[2:]
    H, W = tenImage.shape[2:]
    out = tenDisparity.contiguous().float()
    out = resample_f32(out.view(1, h, w), H, W, False).view(1, 1, H, W) if (h, w) != (H, W) else out.clone()
    if instances is not None and not instances.is_empty:
        masks = instances.masks.contiguous()
        K = masks.shape[0]
        assert masks.shape[1:] == (H, W)
        if use_medium:
            state = torch.empty(1040 // 4, device=out.device, dtype=torch.int32)
            check(lib().csb_depth_adjust_median(ptr(out), ptr(masks.view(torch.uint8)), K, H, W, ptr(state), stream()), "csb_depth_adjust_median")
        else:
            depth_adjust_batch(out.view(1, H, W), masks.view(1, K, H, W), torch.tensor([K], device=out.device, dtype=torch.int32))
    if (h, w) != (H, W):
        out = resample_f32(out.view(1, H, W), h, w, False).view(1, 1, h, w)
    return out


def depth_adjust_batch(disparity, masks, num):
    """csb_depth_adjust_batch: disparity [N,H,W] fp32 (in place), masks [N,Kmax,H,W] bool, num [N] int32 on the device -- one cooperative launch."""
    N, Kmax, H, W = masks.shape
    lib().csb_depth_adjust_state_words.restype = C.c_longlong
    state = torch.empty(int(lib().csb_depth_adjust_state_words(N, Kmax, H, W)), device=disparity.device, dtype=torch.int32)
    check(lib().csb_depth_adjust_batch(ptr(disparity), ptr(masks.view(torch.uint8)), ptr(num), N, Kmax, H, W, ptr(state), stream()), "csb_depth_adjust_batch")
    return disparity


class KenBurnsPipeline:
    """reference :392.  Same public methods; every tensor op on the hot path goes through the C ABI (include/csb200.h)."""

    def __init__(self, cfg: Union[KenBurnsConfig, str, dict] = None, device: str = None) -> None:
        if cfg is None:
            cfg = KenBurnsConfig()
        elif isinstance(cfg, (str, dict)):
            cfg = build_kenburns_cfg(cfg)
        elif not isinstance(cfg, KenBurnsConfig):
            raise NotImplementedError
        self.cfg = cfg
        self.device = torch.device('cuda' if device is None else device)
        self.animeinsseg = None
        self.depth_model = None          # callable: (img_bgr_u8 ndarray, img_tensor [1,3,H,W]) -> disparity [1,1,H,W] on the device
        self.kenburns_inpaintnet = None
        self.depth_refinenet = None
        self.inpaint_type = 'default'
        self._frame_scratch = None
        self.set_detector(cfg.detector)
        self.set_depth_estimation(cfg.depth_est)
        if self.cfg.default_depth_refine:                                     # reference :421-422
            self.set_depth_refinement(cfg.depth_refinement)
        self.set_inpainting(cfg.inpaint_type)

    # ---- component selection (reference :427-560)
    def set_detector(self, detector: str, det_ckpt=None):
        if detector != 'animeinsseg':
            raise NotImplementedError(f"detector '{detector}' is outside the hot path (SURVEY.md §2 row 7: sam / maskrcnn are out of scope)")
        if self.animeinsseg is None:
            import os
            ck = det_ckpt if det_ckpt is not None else (self.cfg.det_ckpt if isinstance(self.cfg.det_ckpt, str) and os.path.exists(self.cfg.det_ckpt) else None)
            self.animeinsseg = AnimeInsSeg(ck, default_det_size=self.cfg.det_size, device=self.device, refine_kwargs={'refine_method': 'none'})

    def set_depth_estimation(self, depth_est: str, ckpt=None):
        """reference :529-560.  'leres' (ResNeXt-101 32x8d + decoder) and 'zoe' (DPT-BEiT-L + metric bins head, flip + pad augmentation) run on
        the tcgen05 engine; 'marigold' (a diffusion model, SURVEY.md §2 out of scope) and 'default' raise at use time unless `depth_model` is
        set by the caller ('external')."""
        self.cfg.depth_est = depth_est
        if depth_est not in ('zoe', 'leres', 'marigold', 'default', 'external'):
            raise NotImplementedError(depth_est)
        if depth_est == 'leres':
            from ..depth_modules.leres import LeReS
            if getattr(self, 'leres', None) is None:
                sd = None
                if ckpt is not None:                                         # res101.pth: {'depth_model': state_dict} (leres/__init__.py:84-89)
                    from ..utils.checkpoints import leres_state_dict
                    sd = leres_state_dict(ckpt)
                self.leres = LeReS(sd, self.device)
            self.depth_model = lambda img, img_tensor: self._depth_est_leres(img_tensor, img)
        elif depth_est == 'zoe':
            from ..depth_modules.zoedepth import ZoeDepth
            if getattr(self, 'depth_zoe', None) is None:
                sd = None
                if ckpt is not None:                                         # ZoeD_M12_N.pt: {'model': state_dict} (model_io.py:27-52)
                    from ..utils.checkpoints import zoe_state_dict
                    sd = zoe_state_dict(ckpt)
                self.depth_zoe = ZoeDepth(sd, self.device, img_size=[672, 672])               # reference :543
            self.depth_model = lambda img, img_tensor: self._depth_est_zoe(img_tensor, img)

    def _depth_est_zoe(self, img_tensor, img, *args, **kwargs):
        """reference :812-818: `depth_zoe.infer(img_tensor, with_flip_aug=True, pad_input=True)` -> disparity [1,1,H,W].  `img` is the BGR uint8
        frame whose /255 the reference passes as `img_tensor` (without the RGB reorder ZoeDepth expects -- kept)."""
        u8 = img if torch.is_tensor(img) else torch.from_numpy(np.ascontiguousarray(img)).to(self.device)
        depth = self.depth_zoe.infer(u8, pad_input=True, with_flip_aug=True)
        return self.depth_zoe.disparity(depth, self.cfg.focal, self.cfg.baseline)[None, None]

    # ---- reference :563-581
    def _leres_post(self, depth_logits: np.ndarray, ori_hw):
        """apply_leres quantisation (leres/__init__.py:117-140) + resize back (:572-577), host numpy/OpenCV exactly as the reference."""
        import cv2
        from ..depth_modules.leres import quantise_depth
        depth = quantise_depth(depth_logits)
        k = depth.shape[0] / ori_hw[0]
        depth = cv2.resize(depth, (ori_hw[1], ori_hw[0]), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)
        return depth.astype(np.float32)

    def _depth_est_leres(self, img_tensor, img, *args, **kwargs):
        return self._depth_est_leres_batch([img])[0]

    def _depth_est_leres_batch(self, imgs, imgs_dev=None):
        """LeReS for a list of same-size BGR uint8 images -> list of disparity tensors [1,1,H,W] on the device."""
        return self.leres_finish(self.leres_enqueue(imgs, imgs_dev))

    def leres_enqueue(self, imgs, imgs_dev=None):
        """Phase 1 (asynchronous): scaledown_maxsize + network forward + D2H of the logits into pinned memory, all enqueued on the current
        stream.  Returns a handle; the caller may enqueue other GPU work (e.g. the detector) before calling leres_finish, so that the
        reference's host-side tail overlaps with it."""
        ori_h, ori_w = (imgs[0].shape[:2] if imgs is not None else imgs_dev.shape[1:3])
        if imgs_dev is not None:            # device-resident inputs: scaledown_maxsize's cv2.resize(INTER_LINEAR) on the device, bit-exact
            small_hw = scaledown_maxsize(np.empty((ori_h, ori_w, 1), np.uint8), self.cfg.depth_est_size, 32).shape[:2]
            if small_hw != (ori_h, ori_w):
                small = torch.empty((imgs_dev.shape[0], small_hw[0], small_hw[1], 3), device=self.device, dtype=torch.uint8)
                for i in range(imgs_dev.shape[0]):
                    check(lib().csb_resize_u8c3(ptr(imgs_dev[i]), ori_h, ori_w, ptr(small[i]), small_hw[0], small_hw[1], stream()), "csb_resize_u8c3")
            else:
                small = imgs_dev
        else:
            small = torch.from_numpy(np.stack([scaledown_maxsize(im, self.cfg.depth_est_size, 32) for im in imgs])).to(self.device)
        logits = self.leres.forward(small)                                     # [N,h,w] fp32
        n, h, w = logits.shape
        if (h > ori_h or (ori_h >= h and ori_w >= w)) and not getattr(self, 'leres_host_tail', False):
            # the reference's numpy/OpenCV tail (16 -> 8 bit quantisation, INTER_AREA / INTER_LANCZOS4 resize back) on the device, bit-exact (csrc/leres_tail.cu):
            # no D2H, no host work, the stream never waits for the CPU
            mm = torch.empty(2 * n, device=self.device, dtype=torch.int32)
            q8 = torch.empty((n, h, w), device=self.device, dtype=torch.uint8)
            out = torch.empty((n, 1, ori_h, ori_w), device=self.device, dtype=torch.float32)
            check(lib().csb_leres_depth_tail(ptr(logits), n, h, w, ori_h, ori_w, ptr(mm), ptr(q8), ptr(out), stream()), "csb_leres_depth_tail")
            scr = torch.empty(n, device=self.device, dtype=torch.int32)                                    # :577 depth[depth == 0] = depth[depth > 0].min(), per image
            check(lib().csb_zero_to_min_positive(ptr(out), n, C.c_longlong(ori_h * ori_w), ptr(scr), stream()), "csb_zero_to_min_positive")
            return (None, (ori_h, ori_w), n, out)
        key = tuple(logits.shape)
        if getattr(self, '_leres_pin', None) is None or tuple(self._leres_pin.shape) != key:
            self._leres_pin = torch.empty(key, dtype=torch.float32).pin_memory()
        self._leres_pin.copy_(logits, non_blocking=True)                       # one D2H for the batch
        ev = torch.cuda.Event()
        ev.record()
        return (ev, (ori_h, ori_w), logits.shape[0], None)

    def leres_finish_batch(self, handle):
        """leres_finish as ONE tensor [N,1,H,W] (no per-image views / re-stacking)"""
        if handle[0] is None:
            return handle[3]
        return torch.cat(self.leres_finish(handle))

    def leres_finish(self, handle):
        """Phase 2: wait for the logits, run the reference's host-side tail (16->8 bit quantisation, OpenCV resize; a thread pool -- OpenCV and
        numpy release the GIL), upload the disparities."""
        from concurrent.futures import ThreadPoolExecutor
        if handle[0] is None:                                                  # finished on the device by leres_enqueue
            return [handle[3][i:i + 1] for i in range(handle[2])]
        ev, (ori_h, ori_w), n = handle[:3]
        ev.synchronize()
        logits = self._leres_pin.numpy()
        if getattr(self, '_pool', None) is None:
            import os
            self._pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
        depth = list(self._pool.map(lambda d: self._leres_post(d, (ori_h, ori_w)), [logits[i] for i in range(n)]))
        if getattr(self, '_disp_pin', None) is None or tuple(self._disp_pin.shape) != (n, 1, ori_h, ori_w):
            self._disp_pin = torch.empty((n, 1, ori_h, ori_w), dtype=torch.float32).pin_memory()
        for i, d in enumerate(depth):
            self._disp_pin[i, 0] = torch.from_numpy(d)
        out = self._disp_pin.to(self.device, non_blocking=True)               # [N,1,H,W]
        # depth[depth == 0] = depth[depth > 0].min()  (:577), per image, without a host sync
        scr = torch.empty(n, device=self.device, dtype=torch.int32)
        check(lib().csb_zero_to_min_positive(ptr(out), n, C.c_longlong(ori_h * ori_w), ptr(scr), stream()), "csb_zero_to_min_positive")
        return [out[i:i + 1] for i in range(n)]

    def set_inpainting(self, inpainting: str, ckpt=None):
        """reference :427-440.  'default' = the point-cloud Inpaint GridNet on the tcgen05 engine; 'ldm' / 'patchmatch' (stable-diffusion webui,
        external PatchMatch .so) are out of scope."""
        if inpainting not in ('default',):
            raise NotImplementedError(f"inpaint_type '{inpainting}' is an optional external back-end outside the hot path")
        self.inpaint_type = inpainting
        if self.kenburns_inpaintnet is None:
            from .models.pointcloud_inpainting import Inpaint
            from ..utils.checkpoints import plain_state_dict
            sd = plain_state_dict(ckpt) if ckpt is not None else None             # models/__init__.py:16-20
            self.kenburns_inpaintnet = Inpaint(sd, self.device)

    def set_depth_refinement(self, depth_refinement: str, ckpt=None):
        """reference :820-826: only 'default' (the `Refine` net, `models/disparity_refinement.py`) exists upstream."""
        if depth_refinement != 'default':
            raise NotImplementedError(f'Invalid depth refinement: {depth_refinement}')
        if getattr(self, 'depth_refinenet', None) is None:
            from .models.disparity_refinement import Refine
            from ..utils.checkpoints import plain_state_dict
            sd = plain_state_dict(ckpt) if ckpt is not None else None             # models/__init__.py:7-11
            self.depth_refinenet = Refine(sd, self.device)
        self._refine_depth = lambda img, disparity: self.depth_refinenet.forward(img, disparity)

    def refine_depth(self, img: torch.Tensor, disparity: torch.Tensor):
        """reference :828-829"""
        return self._refine_depth(img, disparity)

    def set_config(self, cfg: KenBurnsConfig):
        self.cfg = cfg

    def update_config_param(self, cfg_key: str, cfg_value: Any):
        self.cfg[cfg_key] = cfg_value

    # ---- segmentation (reference :862-872)
    def run_instance_segmentation(self, img: np.ndarray, scale_down_to_maxsize=True):
        if scale_down_to_maxsize:
            img = scaledown_maxsize(img, self.cfg.max_size)
        kw = self.cfg.mask_refine_kwargs if self.cfg.mask_refine_kwargs else {'refine_method': 'none'}       # `{}` selects 'none' in the reference (Appendix C.8)
        instances = self.animeinsseg.infer(img, self.cfg.pred_score_thr, kw, output_type='tensor')
        return instances, img

    # ---- depth (reference :583-635)
    def infer_disparity(self, img: np.ndarray, instances: AnimeInstances, img_tensor=None, kcfg: KenBurnsConfig = None, verbose=False):
        kcfg = self.cfg if kcfg is None else kcfg
        if img_tensor is None:
            img_tensor = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)[None].astype(np.float32) * (1.0 / 255.0))).to(self.device)
        if self.depth_model is None:
            raise NotImplementedError(f"depth estimator '{self.cfg.depth_est}' (SURVEY.md §8a rows B1-B6) is not built yet; set pipeline.depth_model "
                                      "to a callable or pass disparity= to generate_kenburns_config")
        disparity = self.depth_model(img, img_tensor)
        disparity = depth_adjustment_animesseg(instances, disparity, img_tensor, use_medium=self.cfg.depthest_use_medium)      # :604
        if self.cfg.default_depth_refine:                                                                                       # :619-620
            disparity = self.refine_depth(img_tensor, disparity)
        elif self.cfg.refine_crf and not getattr(KenBurnsPipeline, '_warned_crf', False):                                     # :621-622
            import warnings
            KenBurnsPipeline._warned_crf = True
            warnings.warn("KenBurnsConfig.refine_crf=True (the dataclass default; configs/3dkenburns.yaml sets it False): the reference would run its CPU "
                          "KMeans + floodFill + dense-CRF depth refinement here (kenburns_effect.py:636-809), which is outside this package's hot path "
                          "(SURVEY.md §8f rank 4) -- the instance-adjusted disparity is used unrefined.  Set refine_crf=False to silence this.")
        return disparity

    # ---- reference :898-951
    def generate_kenburns_config(self, img: np.ndarray, instances: Optional[AnimeInstances] = None, verbose: bool = False, savep=None, disparity=None):
        """`disparity` (extension): a precomputed raw disparity [1,1,H,W]; skips the depth estimator (still instance-adjusted)."""
        if isinstance(img, str):
            import cv2
            img = cv2.imread(img)
        with torch.no_grad():
            if instances is None:
                instances, _ = self.run_instance_segmentation(img, scale_down_to_maxsize=False)
            img = scaledown_maxsize(img, self.cfg.max_size)
            instances.resize(img.shape[0], img.shape[1])
            self.cfg.int_height, self.cfg.int_width = img.shape[:2]
            img_u8 = torch.from_numpy(np.ascontiguousarray(img)).to(self.device)
            cfg: KenBurnsConfig = self.cfg.copy()
            if disparity is None:
                disparity = self.infer_disparity(img, instances, None, kcfg=cfg)
            else:
                img_tensor = (img_u8.permute(2, 0, 1)[None].float() * (1.0 / 255.0))
                disparity = depth_adjustment_animesseg(instances, disparity.to(self.device).float(), img_tensor, use_medium=self.cfg.depthest_use_medium)
            if img.shape[0] <= 256 or img.shape[1] <= 256:            # the reference dies in cv2.minMaxLoc on the empty crop (:937)
                raise ValueError(f"image {img.shape[1]}x{img.shape[0]}: the depth centre crop [128:-128, 128:-128] is empty (needs > 256 x 256)")
            c = disparity_to_cloud(disparity, cfg.focal, cfg.baseline, image_u8=img_u8)                  # :928-937 fused, one D2H of 8 floats
            cfg['fltDispmin'], cfg['fltDispmax'], cfg['objDepthrange'] = c['dispmin'], c['dispmax'], c['depthrange']
            H, W = img.shape[:2]
            cfg['tenRawImage'] = c['data'][:, :3].view(1, 3, H, W)
            cfg['tenRawDisparity'], cfg['tenRawDepth'] = c['disparity'], c['depth']
            cfg['tenRawPoints'], cfg['tenRawUnaltered'] = c['points'].view(1, 3, -1), c['unaltered'].view(1, 3, -1)
            cfg.inpainted_img = cfg['tenRawImage'].view(1, 3, -1)
            cfg['tenInpaDisparity'] = cfg['tenRawDisparity'].view(1, 1, -1)
            cfg['tenInpaDepth'] = cfg['tenRawDepth'].view(1, 1, -1)
            cfg['tenInpaPoints'] = cfg['tenRawPoints'].view(1, 3, -1)
            cfg._render_data = c['data']                       # [1,4,N] = image ++ depth, the frame loop's payload (:1036)
            cfg._render_key = (cfg.inpainted_img.data_ptr(), cfg['tenInpaDepth'].data_ptr())     # the tensors the payload was built from
            cfg.instances = instances
            cfg.original_img_nparray = img
            return cfg

    # ---- reference :953-977
    def autozoom(self, cfg: KenBurnsConfig, verbose: bool = False, inpaint: bool = True):
        with torch.no_grad():
            objFrom = {'fltCenterU': cfg.int_width / 2.0, 'fltCenterV': cfg.int_height / 2.0,
                       'intCropWidth': int(math.floor(0.97 * cfg.int_width)), 'intCropHeight': int(math.floor(0.97 * cfg.int_height))}
            objTo = process_autozoom({'fltShift': 100.0, 'fltZoom': 1.25, 'objFrom': objFrom}, cfg)
            npy_frame_list, _ = self.process_kenburns({'fltSteps': np.linspace(0.0, 1.0, cfg.num_frame).tolist(), 'objFrom': objFrom, 'objTo': objTo,
                                                       'boolInpaint': True}, cfg, inpaint, verbose)
            return npy_frame_list

    def inpaint(self, tenShift, tenPoints, objCommon, verbose=False):
        """reference :441-512 ('default' branch): run the Inpaint net for the shifted view, lift its disparity to points, and append the pixels that
        were holes in that view (tenExisting == 0) to the growing point cloud.  (The
        `stage_inpainted_*` preview images of the reference, which need a D2H per call, are not produced.)"""
        from .models.utils import depth_to_points, spatial_filter
        sh = torch.as_tensor(tenShift, dtype=torch.float32).flatten()
        o = self.kenburns_inpaintnet.forward(objCommon['tenRawImage'], objCommon['tenRawDisparity'], sh, objCommon, None)
        H_, W_ = o['tenExisting'].shape[-2:]
        focal, baseline = objCommon['fltFocal'], objCommon['fltBaseline']
        depth = (focal * baseline) / (o['tenDisparity'] + 0.0000001)                                              # :454
        valid = (spatial_filter(o['tenDisparity'] / o['tenDisparity'].max(), 'laplacian').abs() < 0.03).float()  # :455
        points = depth_to_points(depth * valid, focal).view(1, 3, -1) - sh.view(1, 3, 1).to(self.device)          # :456-458
        # :462-512 -- append the pixels that were holes in this view (tenExisting == 0) to the cloud: order-preserving stream compaction on the device
        # (csrc/kb_compact.cu), one 4-byte host read for the new size (the reference's boolean-mask gathers synchronise too)
        P = H_ * W_
        existing = _f32(o['tenExisting']).reshape(-1)
        lib().csb_cloud_append_scratch_ints.restype = C.c_longlong
        scratch = torch.empty(int(lib().csb_cloud_append_scratch_ints(C.c_longlong(P))), device=self.device, dtype=torch.int32)
        total = torch.empty(1, device=self.device, dtype=torch.int32)
        check(lib().csb_cloud_append_count(ptr(existing), C.c_longlong(P), ptr(scratch), ptr(total), stream()), "csb_cloud_append_count")
        srcs = [(_f32(o['tenImage']).reshape(3, P), 3), (_f32(o['tenDisparity']).reshape(1, P), 1), (_f32(depth).reshape(1, P), 1), (_f32(points).reshape(3, P), 3)]
        olds = [_f32(objCommon.inpainted_img).reshape(3, -1), _f32(objCommon['tenInpaDisparity']).reshape(1, -1), _f32(objCommon['tenInpaDepth']).reshape(1, -1),
                _f32(objCommon['tenInpaPoints']).reshape(3, -1)]
        n_old = olds[0].shape[1]
        n_new = n_old + int(total.item())
        news = [torch.empty((1, c, n_new), device=self.device, dtype=torch.float32) for _, c in srcs]
        plane = lambda t, c, n: [t.data_ptr() + 4 * n * k for k in range(c)]
        sp = [a for (t, c) in srcs for a in plane(t, c, P)]
        op = [a for t, (_, c) in zip(olds, srcs) for a in plane(t, c, n_old)]
        dp = [a for t, (_, c) in zip(news, srcs) for a in plane(t, c, n_new)]
        arr = lambda v: (C.c_void_p * len(v))(*v)
        check(lib().csb_cloud_append(ptr(existing), C.c_longlong(P), ptr(scratch), arr(sp), arr(op), arr(dp), len(sp), C.c_longlong(n_old), stream()),
              "csb_cloud_append")
        objCommon.inpainted_img, objCommon['tenInpaDisparity'], objCommon['tenInpaDepth'], objCommon['tenInpaPoints'] = news
        return o

    # ---- batched per-frame path (extension; the reference is a batch-1 loop over exactly these calls)
    def render_frame_batch(self, imgs, shift_u: float = 40.0, shift_v: float = -25.0, depth_ratio: float = 0.8, crop_frac: float = 0.97,
                           raw_disparity=None, segment: bool = True, warp: bool = True, out_host=None, det_sub: int = None, zoe_sub: int = 16):
        """One Ken-Burns frame per input frame, for a batch of equally sized frames -- per frame exactly the reference's sequence
        `AnimeInsSeg.infer` body (:862-872, refine off) -> `_depth_est` (:563-581 / :812-818) -> `depth_adjustment_animesseg` (:604) ->
        disparity -> cloud (:928-937) -> `process_shift` at camera offset (shift_u, shift_v) with the closest depth scaled by `depth_ratio`
        (common.py:59-83) -> render + fill + uint8 pack (:1028-1040) -> centre crop `crop_frac` + resize (:1069-1070) -- with the networks run over
        the whole batch and no host synchronisation besides the instance counts.

        imgs: uint8 [B,H,W,3] BGR -- a CUDA tensor (inputs resident in HBM) or a (pinned) host tensor / ndarray, uploaded here.
        raw_disparity: optional [B,H,W] fp32 CUDA tensor replacing the depth estimator.  out_host: optional pinned uint8 [B,H,W,3]; every
        finished frame is copied into it and the call returns after the last copy.
        -> dict(frames=[B,H,W,3] uint8 CUDA tensor, num_instances=list[int] or None)"""
        cfg = self.cfg
        dev = self.device
        if isinstance(imgs, np.ndarray):
            imgs = torch.from_numpy(imgs)
        B, H, W = imgs.shape[:3]
        st = getattr(self, '_rfb', None)
        if st is None or st['key'] != (B, H, W):
            st = self._rfb = {
                'key': (B, H, W), 'stage': torch.empty((B, H, W, 3), device=dev, dtype=torch.uint8), 'scratch': FrameScratch(H, W, dev),
                'clouds': [{k: torch.empty((1, c, H, W), device=dev) for k, c in (('disparity', 1), ('depth', 1), ('valid', 1), ('points', 3), ('unaltered', 3))}
                           for _ in range(2)],
                'data': [torch.empty((1, 4, H * W), device=dev) for _ in range(2)], 'scalars': torch.empty(8, device=dev),
                'd2c': torch.empty(64, device=dev, dtype=torch.int64), 'shift': torch.empty(3, device=dev),
                'out': torch.empty((B, H, W, 3), device=dev, dtype=torch.uint8)}
        if det_sub is None:                                          # measured (gpurun r2c23): e2e 85.2 / 85.7 / 90.0 ms per 32 frames with sub-batches of 32 / 16 / 8
            det_sub = int(os.environ.get('CSB_E2E_DET_SUB', '32'))
        h2d_done = {}
        if imgs.is_cuda:
            batch = imgs
        else:                                                        # H2D of this call's inputs from pinned memory, on the copy stream, one event per
            batch = st['stage']                                      # detector sub-batch: only the first chunk's copy is exposed
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream())           # the staging buffer's previous consumers have finished
            with torch.cuda.stream(self._copy_stream):
                for s0 in range(0, B, det_sub if segment else B):
                    s1 = min(B, s0 + (det_sub if segment else B))
                    batch[s0:s1].copy_(imgs[s0:s1], non_blocking=True)
                    h2d_done[s0] = torch.cuda.Event()
                    h2d_done[s0].record(self._copy_stream)
            if not segment:
                torch.cuda.current_stream().wait_event(h2d_done.pop(0))
        cd = C.c_double
        focal, baseline = float(cfg.focal), float(cfg.baseline)
        # ---- segmentation (A1-A9): detector forward + post-process per sub-batch (each waits only for its own frames' H2D)
        masks, nums_dev = [], []
        if segment:
            from ..animeinsseg import rtmdet_postprocess
            seg = self.animeinsseg
            test_cfg = seg.model.bbox_head.test_cfg
            for s0 in range(0, B, det_sub):
                if s0 in h2d_done:
                    torch.cuda.current_stream().wait_event(h2d_done[s0])
                cls, reg, ker, mf = seg.model.net.forward(batch[s0:s0 + det_sub])
                o = rtmdet_postprocess(cls, reg, ker, mf, (H, W), test_cfg)
                masks.append(o['masks']); nums_dev.append(o['num'])
        # ---- depth over the whole batch: LeReS (tail on the device) or ZoeDepth in sub-batches
        disp = raw_disparity
        if disp is None:
            if cfg.depth_est == 'leres':
                disp = self.leres_finish_batch(self.leres_enqueue(None, imgs_dev=batch))
            elif cfg.depth_est == 'zoe':
                disp = torch.empty((B, H, W), device=dev, dtype=torch.float32)
                for s0 in range(0, B, zoe_sub):
                    d = self.depth_zoe.infer_batch(batch[s0:s0 + zoe_sub])
                    for i in range(d.shape[0]):                               # :815-817 is per image (min positive)
                        self.depth_zoe.disparity(d[i], focal, baseline, out=disp[s0 + i])
            else:
                raise NotImplementedError(f"render_frame_batch: depth_est '{cfg.depth_est}' (pass raw_disparity=)")
        disp = disp.reshape(B, H, W)
        nums = None
        if segment:
            nums = [int(v) for t in nums_dev for v in t.cpu().tolist()]          # the one host read: instance counts for the caller
            for j, s0 in enumerate(range(0, B, det_sub)):                      # instance-guided depth flattening (C2), in place
                depth_adjust_batch(disp[s0:s0 + det_sub], masks[j], nums_dev[j])
        if not warp:
            return dict(frames=None, num_instances=nums, disparity=disp)
        pw, ph = int(math.floor(crop_frac * W)), int(math.floor(crop_frac * H))
        fs = st['scratch']
        copy_stream = None
        if out_host is not None:
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            copy_stream = self._copy_stream
        for b in range(B):
            c, data = st['clouds'][b & 1], st['data'][b & 1]
            check(lib().csb_disparity_to_cloud(ptr(disp[b]), H, W, cd(focal), cd(baseline), ptr(c['disparity']), ptr(c['depth']), ptr(c['valid']),
                                               ptr(c['points']), ptr(c['unaltered']), ptr(st['scalars']), ptr(st['d2c']), ptr(batch[b]), ptr(data), stream()),
                  "csb_disparity_to_cloud")
            check(lib().csb_shift_from_scalars(ptr(st['scalars']), W, H, cd(focal), cd(shift_u), cd(shift_v), cd(depth_ratio), ptr(st['shift']), stream()),
                  "csb_shift_from_scalars")
            check(lib().csb_kenburns_frame(ptr(c['points']), ptr(data), H * W, H, W, cd(focal), cd(baseline), None, ptr(st['shift']), pw, ph,
                                           cd(W / 2.0), cd(H / 2.0), ptr(fs.zkey), ptr(fs.zee), ptr(fs.acc), ptr(fs.packed), ptr(st['out'][b]), None, stream()),
                  "csb_kenburns_frame")
            if copy_stream is not None:                                        # D2H of frame b overlaps the render of frame b+1
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    out_host[b].copy_(st['out'][b], non_blocking=True)
        if copy_stream is not None:
            copy_stream.synchronize()                                          # the caller receives this batch's frames
        return dict(frames=st['out'], num_instances=nums, disparity=disp)

    # ---- reference :979-1081
    def process_kenburns(self, objSettings, objCommon: KenBurnsConfig, inpaint: bool = True, verbose: bool = False):
        with torch.no_grad():
            W, H = objCommon['intWidth'], objCommon['intHeight']
            oF, oT = objSettings['objFrom'], objSettings['objTo']

            def camera(fltStep):
                fltFrom = 1.0 - fltStep
                fltTo = 1.0 - fltFrom
                fltShiftU = ((fltFrom * oF['fltCenterU']) + (fltTo * oT['fltCenterU'])) - (W / 2.0)
                fltShiftV = ((fltFrom * oF['fltCenterV']) + (fltTo * oT['fltCenterV'])) - (H / 2.0)
                fltCropWidth = (fltFrom * oF['intCropWidth']) + (fltTo * oT['intCropWidth'])
                fltDepthFrom = objCommon['objDepthrange'][0]
                fltDepthTo = objCommon['objDepthrange'][0] * (fltCropWidth / max(oF['intCropWidth'], oT['intCropWidth']))
                return {'fltShiftU': fltShiftU, 'fltShiftV': fltShiftV, 'fltDepthFrom': fltDepthFrom, 'fltDepthTo': fltDepthTo}
            if inpaint:
                objCommon.inpainted_img = objCommon['tenRawImage'].view(1, 3, -1)
                objCommon['tenInpaDisparity'] = objCommon['tenRawDisparity'].view(1, 1, -1)
                objCommon['tenInpaDepth'] = objCommon['tenRawDepth'].view(1, 1, -1)
                objCommon['tenInpaPoints'] = objCommon['tenRawPoints'].view(1, 3, -1)
                for fltStep in [0.0, 1.0]:
                    sh = np.array(shift_scalars(camera(fltStep), objCommon), np.float32)
                    self.inpaint(1.1 * sh, None, objCommon, verbose)
            pts = objCommon['tenInpaPoints']
            N = pts.shape[2]
            # the cached payload is only valid for the very tensors it was built from (a caller may have replaced inpainted_img / tenInpaDepth by
            # same-sized ones, e.g. a repainted image): identity check, else rebuild as the reference does every frame (:1036)
            data = getattr(objCommon, '_render_data', None)
            key = (objCommon.inpainted_img.data_ptr(), objCommon['tenInpaDepth'].data_ptr())
            if data is None or data.shape[2] != N or getattr(objCommon, '_render_key', None) != key:
                data = torch.cat([objCommon.inpainted_img, objCommon['tenInpaDepth']], 1).view(1, 4, -1).contiguous()       # :1036
            pw, ph = max(oF['intCropWidth'], oT['intCropWidth']), max(oF['intCropHeight'], oT['intCropHeight'])
            steps = objSettings['fltSteps']
            if self._frame_scratch is None or self._frame_scratch.key[2:4] != (H, W):
                self._frame_scratch = FrameScratch(H, W, pts.device)
            out_dev = torch.empty((len(steps), H, W, 3), device=pts.device, dtype=torch.uint8)
            out_host = torch.empty((len(steps), H, W, 3), dtype=torch.uint8, pin_memory=True)     # straight from torch's pinned-block cache (no pageable staging copy)
            bokeh = None
            if objCommon.depth_field:                                              # :1042-1067, all on the device (utils/effects.py)
                ins = objCommon.instances
                masks = None if ins is None or ins.is_empty else ins.masks
                bokeh = fx.BokehScratch(H, W, 0 if masks is None else int(masks.shape[0]), pts.device, objCommon.lightness_factor)
            if bokeh is None:                                                      # the reference frame loop, :1015-1072, as ONE library call
                shifts = np.ascontiguousarray(np.stack([np.array(shift_scalars(camera(t), objCommon), np.float32) for t in steps]), np.float32)
                if getattr(self, '_copy_stream', None) is None:
                    self._copy_stream = torch.cuda.Stream(device=pts.device)
                fs = self._frame_scratch
                self._copy_stream.wait_stream(torch.cuda.current_stream())         # out_host / out_dev were allocated on the current stream
                check(lib().csb_kenburns_frames(ptr(_f32(pts)), ptr(_f32(data)), N, H, W, C.c_double(objCommon['fltFocal']), C.c_double(objCommon['fltBaseline']),
                                                shifts.ctypes.data_as(C.POINTER(C.c_float)), len(steps), int(pw), int(ph), C.c_double(W / 2.0), C.c_double(H / 2.0),
                                                ptr(fs.zkey), ptr(fs.zee), ptr(fs.acc), ptr(fs.packed), ptr(out_dev), C.c_void_p(out_host.data_ptr()),
                                                C.c_void_p(self._copy_stream.cuda_stream), stream()), "csb_kenburns_frames")
                out_dev.record_stream(self._copy_stream)
            for i, fltStep in (enumerate(steps) if bokeh is not None else ()):    # depth of field (:1042-1067): the bokeh stage sits between pack and crop
                sh = np.array(shift_scalars(camera(fltStep), objCommon), np.float32)
                _, depth = kenburns_frame(pts, data, W, H, objCommon['fltFocal'], objCommon['fltBaseline'], sh, pw, ph, W / 2.0, H / 2.0,
                                          scratch=self._frame_scratch, out=False, want_depth=True)
                d8 = fx.colorize_gray_r(depth, bokeh)                              # colorize(depth_rendered, cmap='gray_r')[..., 0]
                if i == 0:
                    fx.focal_plane_range(d8, masks, bokeh)                         # focalplane_start / _end, :1045-1059
                blurred = fx.bokeh_blur(self._frame_scratch.packed, d8, 32, objCommon.lightness_factor, objCommon.depth_factor, True,
                                        scratch=bokeh, focal_int=fx.focal_interp(fltStep, objCommon.dof_speed))
                check(lib().csb_frame_crop_resize(ptr(blurred), H, W, int(pw), int(ph), C.c_double(W / 2.0), C.c_double(H / 2.0),
                                                  ptr(out_dev[i]), stream()), "csb_frame_crop_resize")
                out_host[i].copy_(out_dev[i], non_blocking=True)                   # D2H overlaps the next frame's render
            torch.cuda.current_stream().synchronize()
            frames = [out_host[i].numpy() for i in range(len(steps))]
            return [frames, objCommon]


def npyframes2video(npy_frame_list, video_save_path: str, playback: bool = False, fps: int = 25):
    """reference :1086-1090: write the BGR uint8 frames as a 25 fps video; `playback` appends the reversed sequence without its two end frames
    (forward + back loop).  The reference goes through moviepy (libx264, RGB frames); here OpenCV's writer takes the BGR frames directly."""
    import cv2
    sequence = list(npy_frame_list)
    if playback:
        sequence += sequence[::-1][1:-1]
    if not sequence:
        raise ValueError("npyframes2video: empty frame list")
    h, w = sequence[0].shape[:2]
    writer = cv2.VideoWriter(video_save_path, cv2.VideoWriter_fourcc(*'mp4v'), float(fps), (w, h))
    if not writer.isOpened():
        raise RuntimeError(f"npyframes2video: cannot open {video_save_path} for writing")
    for frame in sequence:
        writer.write(np.ascontiguousarray(frame))
    writer.release()
    return len(sequence)
