"""B200 implementation of the reference's `animeinsseg` package surface for the hot path: `AnimeInsSeg.infer` -> `AnimeInstances`
(reference animeinsseg/__init__.py:185-227, 395-504, 704-708).

The mmdet model object of the reference (`self.model`, built from the checkpoint-embedded config, :196-209) is replaced by
`rtmdet.RTMDetIns` (tcgen05 conv engine) + the post-processing kernels of csrc/det_post.cu; `self.model.bbox_head.test_cfg` stays a
mutable dict with the reference's keys (`max_per_img`, `mask_thr_binary`, `score_thr`, `nms_pre`, `nms`, `min_bbox_size`), so
`set_max_instance` / `set_mask_threshold` behave as in the reference.
"""
import ctypes as C
import math
from types import SimpleNamespace
from typing import List, Union

import numpy as np
import torch

from .._lib import check, lib, ptr, stream
from .anime_instances import AnimeInstances
from .rtmdet import RTMDetIns, STRIDES, synthetic_state_dict

__all__ = ["AnimeInsSeg", "AnimeInstances"]


def rtmdet_postprocess(cls, reg, ker, mask_feat, img_hw, test_cfg, ori_hw=None, scale_factor=(1.0, 1.0)):
    """A5-A8 on the device: (cls, reg, ker per level NHWC fp32, mask_feat [N,h,w,8]) ->
    dict(boxes [N,K,4] xyxy in the ORIGINAL image frame, scores [N,K], num [N] int32 (device), masks [N,K,H,W] bool, logits)."""
    N = cls[0].shape[0]
    L = len(cls)
    dev = cls[0].device
    K, nms_pre = int(test_cfg['max_per_img']), int(test_cfg['nms_pre'])
    hs = (C.c_int * L)(*[c.shape[1] for c in cls]); ws = (C.c_int * L)(*[c.shape[2] for c in cls]); st = (C.c_int * L)(*STRIDES[:L])
    arr = lambda ts: (C.c_void_p * L)(*[t.data_ptr() for t in ts])
    cand = torch.empty((N, L * nms_pre, 10), device=dev); cand_count = torch.empty((N * L,), device=dev, dtype=torch.int32)
    boxes = torch.empty((N, K, 4), device=dev); scores = torch.empty((N, K), device=dev); priors = torch.empty((N, K, 4), device=dev)
    kernels = torch.empty((N, K, 169), device=dev); num = torch.empty((N,), device=dev, dtype=torch.int32)
    cls = [t.contiguous() for t in cls]; reg = [t.contiguous() for t in reg]; ker = [t.contiguous() for t in ker]
    check(lib().csb_rtmdet_select(arr(cls), arr(reg), arr(ker), hs, ws, st, L, N, C.c_float(test_cfg['score_thr']), nms_pre,
                                  C.c_float(test_cfg['nms']['iou_threshold']), K, C.c_float(test_cfg['min_bbox_size']), int(img_hw[0]), int(img_hw[1]),
                                  ptr(cand), ptr(cand_count), ptr(boxes), ptr(scores), ptr(priors), ptr(kernels), ptr(num), stream()), "csb_rtmdet_select")
    h, w = mask_feat.shape[1:3]
    ori_hw = img_hw if ori_hw is None else ori_hw
    # mask tail (reference :361-370): x8, then resize to ceil(size * 1/scale_factor), crop to the original image
    rh, rw = math.ceil(h * STRIDES[0] * (1.0 / scale_factor[0])), math.ceil(w * STRIDES[0] * (1.0 / scale_factor[1]))
    logits = torch.empty((N, K, h, w), device=dev); masks = torch.empty((N, K, ori_hw[0], ori_hw[1]), device=dev, dtype=torch.uint8)
    check(lib().csb_rtmdet_masks(ptr(mask_feat.contiguous()), ptr(kernels), ptr(priors), ptr(num), N, K, h, w, STRIDES[0], int(ori_hw[0]), int(ori_hw[1]),
                                 rh, rw, C.c_float(test_cfg['mask_thr_binary']), ptr(logits), ptr(masks), stream()), "csb_rtmdet_masks")
    if scale_factor != (1.0, 1.0):        # rescale boxes to the original frame (mmdet: bboxes /= scale_factor (w,h,w,h))
        boxes = boxes / boxes.new_tensor([scale_factor[0], scale_factor[1], scale_factor[0], scale_factor[1]])
    return dict(boxes=boxes, scores=scores, num=num, masks=masks.view(torch.bool), logits=logits, priors=priors, kernels=kernels)


class AnimeInsSeg:
    """reference animeinsseg/__init__.py:185.  `ckpt`: path to a checkpoint whose 'state_dict' uses mmdet parameter names, a state_dict,
    or None for the seeded synthetic ConvNeXt-B RTMDet-Ins weights (BASELINE.json: 'random-init ConvNeXt-B RTMDet weights')."""

    def __init__(self, ckpt=None, default_det_size: int = 640, device: str = None, refine_kwargs: dict = {'refine_method': 'refinenet_isnet'},
                 tagger_path: str = None, mask_thr=0.3, max_det_batch: int = 32) -> None:
        """Defaults as the reference's (:187-189): ISNet mask refinement is ON unless `refine_kwargs={'refine_method': 'none'}` is passed.
        `max_det_batch` (extension): same-shape images of one `infer` call run through the detector in sub-batches of at most this many
        (activations of 32 x 1024^2 are ~12 GB, the [N,K,H,W] mask buffer 3.4 GB), so a long list costs no more memory than a short one."""
        self.device = torch.device('cuda' if device is None else device)
        if self.device.type != 'cuda':
            raise RuntimeError("cartoonsegmentation_b200 has no CPU path (the reference's device='cpu' mode is the oracle's job)")
        if ckpt is None:
            sd = synthetic_state_dict(0)
        else:                                      # a checkpoint path in the reference's format ({'state_dict', 'meta': {'cfg'}}, :196-208) or a state_dict
            from ..utils.checkpoints import detector_state_dict
            sd = detector_state_dict(ckpt)
        net = RTMDetIns(sd, self.device)
        test_cfg = dict(nms_pre=1000, score_thr=0.05, nms=dict(type='nms', iou_threshold=0.6), max_per_img=100, min_bbox_size=0, mask_thr_binary=0.5)
        self.model = SimpleNamespace(net=net, bbox_head=SimpleNamespace(test_cfg=test_cfg, prior_generator=SimpleNamespace(strides=[(s, s) for s in STRIDES])))
        self.default_det_size = default_det_size
        self.max_det_batch = max(1, int(max_det_batch))
        self.det_size = (default_det_size, default_det_size)
        self.postprocess_refine = None
        self.refinenet = None
        self.refine_size = 720
        self.mask_thr = mask_thr
        if refine_kwargs is not None:
            self.set_refine_method(**refine_kwargs)

    # ---- reference :395-399, :623-637, :704-708
    def set_detect_size(self, det_size):
        self.det_size = (det_size, det_size) if isinstance(det_size, int) else tuple(det_size)

    def set_refine_method(self, refine_method: str = 'none', refine_size: int = 720, refine_ckpt=None):
        """reference :623-633.  'refinenet_isnet' = ISNetDIS(in_ch=4) on the tcgen05 engine (isnet.py); 'animeseg' (the alternative
        anime-seg matting net, AnimeSegmentation.try_load) cannot run in the reference either (see the branch below), so it raises here too."""
        if refine_method == 'none':
            self.postprocess_refine = None
        elif refine_method == 'refinenet_isnet':
            if self.refinenet is None:
                from .isnet import ISNetDIS
                sd = None
                if refine_ckpt is not None:            # models/AnimeInstanceSegmentation/refine_last.ckpt: a plain state_dict (animeseg_refine/__init__.py:161-165)
                    from ..utils.checkpoints import plain_state_dict
                    sd = plain_state_dict(refine_ckpt)
                self.refinenet = ISNetDIS(sd, self.device)
            self.refine_size = refine_size
            self.postprocess_refine = self._postprocess_refine
        elif refine_method == 'animeseg':
            # In the reference this branch is dead on the infer path: `animeseg_refine` (:78-82) reads `det_pred.pred_instances` of an mmdet
            # DetDataSample, but `postprocess_results` (:700-702) hands it the AnimeInstances built by `_det_forward` (:447-462), which has no such
            # attribute -> AttributeError on the first image with a detection.  Error behaviour is kept (an exception at selection time instead).
            raise NotImplementedError("refine method 'animeseg': the reference's own call path for it raises AttributeError "
                                      "(animeseg_refine expects a DetDataSample, infer passes AnimeInstances); use 'refinenet_isnet' or 'none'")
        else:
            raise NotImplementedError(f'Invalid refine method: {refine_method}')

    def _postprocess_refine(self, instances: AnimeInstances, img: np.ndarray, refine_size: int = None, max_refine_batch: int = 16, **kwargs):
        """reference :638-665 + prepare_refine_batch :37-55, entirely on the device: [BGR/255 | mask] at refine_size^2 (shrink-only resize, pad
        bottom/right) -> ISNet d1 -> sigmoid -> un-pad -> bilinear(align_corners=True) to (H,W) -> > self.mask_thr.  Sub-batches of 16
        instead of the reference's 4 (results do not depend on the grouping)."""
        if instances.is_empty:
            return
        from ..anime_3dkenburns.kenburns_effect import scaledown_maxsize
        S = self.refine_size if refine_size is None else refine_size
        was_numpy = instances.is_numpy
        masks = instances.masks if instances.is_tensor else torch.from_numpy(instances.masks)
        masks = masks.to(self.device).contiguous()
        K, H, W = masks.shape
        img_dev = torch.from_numpy(np.ascontiguousarray(img)).to(self.device) if isinstance(img, np.ndarray) else img
        h, w = scaledown_maxsize(np.empty((H, W, 1), np.uint8), S).shape[:2]                  # same rounding as the reference's resize_pad
        if (h, w) != (H, W):
            small = torch.empty((h, w, 3), device=self.device, dtype=torch.uint8)
            check(lib().csb_resize_u8c3(ptr(img_dev), H, W, ptr(small), h, w, stream()), "csb_resize_u8c3")
        else:
            small = img_dev
        out = torch.empty((K, H, W), device=self.device, dtype=torch.uint8)
        for k0 in range(0, K, max_refine_batch):
            k1 = min(K, k0 + max_refine_batch)
            x16 = torch.empty((k1 - k0, S, S, 16), device=self.device, dtype=torch.float16)
            check(lib().csb_refine_prep(ptr(small), h, w, ptr(masks[k0:k1].view(torch.uint8)), k1 - k0, H, W, S, ptr(x16), stream()), "csb_refine_prep")
            d1 = self.refinenet.forward(x16)
            check(lib().csb_refine_post(ptr(d1), k1 - k0, S, h, w, H, W, C.c_float(self.mask_thr), ptr(out[k0:k1]), stream()), "csb_refine_post")
        instances.masks = out.view(torch.bool)
        if was_numpy:
            instances.masks = instances.masks.cpu().numpy()

    def set_mask_threshold(self, mask_thr: float):
        self.model.bbox_head.test_cfg['mask_thr_binary'] = mask_thr

    def set_max_instance(self, num_ins):
        self.model.bbox_head.test_cfg['max_per_img'] = num_ins

    # ---- A1: mmdet Resize(keep_ratio) + Pad(size, 114) on the host (the reference does this in OpenCV on the CPU too)
    def _prepare(self, img: np.ndarray):
        import cv2
        S = self.det_size
        h, w = img.shape[:2]
        sf = min(max(S) / max(h, w), min(S) / min(h, w))
        nw, nh = int(w * sf + 0.5), int(h * sf + 0.5)
        if (nw, nh) != (w, h):
            img = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
        scale_factor = (nw / w, nh / h)
        ph, pw = S[1], S[0]
        if (nh, nw) != (ph, pw):
            img = cv2.copyMakeBorder(img, 0, ph - nh, 0, pw - nw, cv2.BORDER_CONSTANT, value=(114, 114, 114))
        return np.ascontiguousarray(img), scale_factor, (h, w)

    @torch.no_grad()
    def infer(self, imgs: Union[List, str, np.ndarray], pred_score_thr: float = 0.3, refine_kwargs: dict = None, output_type: str = "tensor",
              det_size: int = None, save_dir: str = '', save_visualization: bool = False, save_annotation: str = '', infer_tags: bool = False,
              obj_id_start: int = -1, img_id_start: int = -1, verbose: bool = False, infer_grey: bool = False, save_mask_only: bool = False,
              val_dir=None, max_instances: int = 100, **kwargs):
        if det_size is not None:
            self.set_detect_size(det_size)
        if refine_kwargs is not None:
            self.set_refine_method(**refine_kwargs)
        self.set_max_instance(max_instances)
        if save_annotation or save_visualization or infer_tags:
            raise NotImplementedError("annotation export / visualisation / tagging are outside the hot path (SURVEY.md §2 row 1)")
        assert output_type in {'tensor', 'numpy'}
        return_list = isinstance(imgs, list)
        if not return_list:
            imgs = [imgs]
        loaded = []
        for im in imgs:
            if isinstance(im, str):
                import cv2
                im = cv2.imread(im)
            loaded.append(im)
        preds = [None] * len(loaded)
        # images that share (shape after Resize/Pad, scale) go through the detector as ONE batch (the reference loops batch-1, :485)
        groups = {}
        for i, im in enumerate(loaded):
            arr, sf, ori = self._prepare(im)
            groups.setdefault((arr.shape, sf, ori), []).append((i, arr))
        for (shape, sf, ori), items in groups.items():
            for s0 in range(0, len(items), self.max_det_batch):           # bounded sub-batches: memory does not grow with the list length
                sub = items[s0:s0 + self.max_det_batch]
                batch = torch.from_numpy(np.stack([a for _, a in sub])).to(self.device, non_blocking=True)
                for j, inst in enumerate(self._det_forward(batch, sf, ori, pred_score_thr)):
                    preds[sub[j][0]] = inst
        for inst, im in zip(preds, loaded):
            if self.postprocess_refine is not None:
                self.postprocess_refine(inst, im)
            if output_type == 'numpy':
                inst.to_numpy()
        return preds if return_list else preds[0]

    def _det_forward(self, batch_u8, scale_factor, ori_hw, pred_score_thr: float = 0.3):
        """reference :447-462 for a batch: detector + post-process, then score filter, int32 truncation, xyxy -> xywh."""
        cls, reg, ker, mask_feat = self.model.net.forward(batch_u8)
        out = rtmdet_postprocess(cls, reg, ker, mask_feat, batch_u8.shape[1:3], self.model.bbox_head.test_cfg, ori_hw, scale_factor)
        nums = out['num'].cpu().tolist()                                    # the one host read per batch
        res = []
        for n, k in enumerate(nums):
            scores = out['scores'][n, :k]
            keep = scores > pred_score_thr
            if int(keep.sum()) < 1:
                res.append(AnimeInstances())
                continue
            bboxes = out['boxes'][n, :k][keep].to(torch.int32)
            bboxes[:, 2:] -= bboxes[:, :2]
            res.append(AnimeInstances(out['masks'][n, :k][keep], bboxes, scores[keep]))
        return res
