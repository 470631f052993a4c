"""`AnimeInstances` -- result container of `AnimeInsSeg.infer`, mirroring the reference's animeinsseg/anime_instances.py:31-298
(masks [K,H,W] bool, bboxes [K,4] xywh int, scores [K]).  Drawing helpers of the reference (:128-203) are out of scope."""
from typing import List, Union

import numpy as np
import torch


class AnimeInstances:

    def __init__(self, masks: Union[np.ndarray, torch.Tensor] = None, bboxes: Union[np.ndarray, torch.Tensor] = None,
                 scores: Union[np.ndarray, torch.Tensor] = None, tags: List[str] = None, character_tags: List[str] = None) -> None:
        self.masks = masks
        self.tags = tags
        self.bboxes = bboxes
        if scores is None:
            scores = [1.] * len(self)
            if self.is_numpy:
                scores = np.array(scores)
            elif self.is_tensor:
                scores = torch.tensor(scores)
        self.scores = scores
        if tags is None:
            self.tags = [''] * len(self)
            self.character_tags = [''] * len(self)
        else:
            self.tags = tags
            self.character_tags = character_tags

    @property
    def is_cuda(self):
        return isinstance(self.masks, torch.Tensor) and self.masks.is_cuda

    @property
    def is_tensor(self):
        return False if self.is_empty else isinstance(self.masks, torch.Tensor)

    @property
    def is_numpy(self):
        return True if self.is_empty else isinstance(self.masks, np.ndarray)

    @property
    def is_empty(self):
        return self.masks is None or len(self.masks) == 0

    def remove_duplicated(self, overlap: float = 0.8):
        """Drop instances that are mostly covered by larger ones (reference :84-127): visit the masks from the largest to the smallest; an
        instance whose intersection with the union of the instances kept so far exceeds `overlap` of its own area is a duplicate.
        (The reference never adds the LAST visited mask to the union -- it has no successor -- which does not change the outcome.)"""
        n = len(self)
        if n < 2:
            return
        was_numpy = self.is_numpy
        masks = torch.from_numpy(self.masks) if was_numpy else self.masks
        areas = masks.flatten(1).sum(1).to(torch.float32)
        order = torch.argsort(areas, descending=True)
        union = torch.zeros_like(masks[0])
        kept = []
        for rank, idx in enumerate(order.tolist()):
            m = masks[idx]
            if rank > 0 and float((union & m).sum()) / float(areas[idx]) > overlap:
                continue
            kept.append(idx)
            union = union | m
        sel = torch.as_tensor(kept, device=masks.device, dtype=torch.long)
        take = (lambda t: t[sel.cpu().numpy()]) if was_numpy else (lambda t: t[sel.to(t.device)])
        self.masks, self.bboxes, self.scores = take(self.masks), take(self.bboxes), take(self.scores)
        self.tags = [self.tags[i] for i in kept]
        self.character_tags = [self.character_tags[i] for i in kept] if self.character_tags is not None else None

    def cuda(self):
        if self.is_empty:
            return self
        self.masks, self.scores, self.bboxes = self.masks.cuda(), self.scores.cuda(), self.bboxes.cuda()
        return self

    def cpu(self):
        if not self.is_tensor or not self.is_cuda:
            return self
        self.masks, self.scores, self.bboxes = self.masks.cpu(), self.scores.cpu(), self.bboxes.cpu()
        return self

    def to_tensor(self, device: str = 'cpu'):
        if self.is_empty:
            return self
        elif self.is_tensor and self.masks.device == device:
            return self
        if self.is_tensor:
            self.masks, self.bboxes, self.scores = self.masks.to(device), self.bboxes.to(device), self.scores.to(device)
            return self
        self.masks = torch.from_numpy(self.masks).to(device)
        self.bboxes = torch.from_numpy(self.bboxes).to(device)
        self.scores = torch.from_numpy(self.scores).to(device)
        return self

    def to_numpy(self):
        if self.is_numpy:
            return self
        self.masks, self.scores, self.bboxes = self.masks.cpu().numpy(), self.scores.cpu().numpy(), self.bboxes.cpu().numpy()
        return self

    def get_instance(self, ins_idx: int, out_type: str = None, device: str = None):
        mask, bbox, score = self.masks[ins_idx], self.bboxes[ins_idx], self.scores[ins_idx]
        if out_type is not None:
            if out_type == 'numpy' and not self.is_numpy:
                mask, bbox, score = mask.cpu().numpy(), bbox.cpu().numpy(), score.cpu().numpy()
            if out_type == 'tensor' and not self.is_tensor:
                mask, bbox, score = torch.from_numpy(mask), torch.from_numpy(bbox), torch.from_numpy(score)
            if isinstance(mask, torch.Tensor) and device is not None and mask.device != device:
                mask, bbox, score = mask.to(device), bbox.to(device), score.to(device)
        return {'mask': mask, 'tags': self.tags[ins_idx], 'character_tags': self.character_tags[ins_idx], 'bbox': bbox, 'score': score}

    def __len__(self):
        return 0 if self.is_empty else len(self.masks)

    def resize(self, h, w, mode='area'):
        """reference :268-280: masks -> float -> F.interpolate(mode='area') -> > 0.3; bboxes scaled (x by the height ratio, y by the width ratio:
        the reference's quirk, kept) and rounded.  On the device: one kernel (csrc/masks.cu), bit-identical to those torch ops."""
        if self.is_empty:
            return
        if self.is_tensor:
            if mode != 'area' or not self.masks.is_cuda:
                raise NotImplementedError("AnimeInstances.resize: only mode='area' on CUDA tensors (the reference's only call, kenburns_effect.py:918)")
            from .._lib import check, lib, ptr, stream
            import ctypes as C
            K, oh, ow = self.masks.shape
            if (oh, ow) == (int(h), int(w)):
                # identity: 1 x 1 windows give avg == mask, so `> 0.3` returns the masks; scale factors are 1.0, round(int) is the int
                self.bboxes = self.bboxes.to(torch.int32)
                return
            src = self.masks.contiguous().view(torch.uint8)
            out = torch.empty((K, h, w), device=src.device, dtype=torch.uint8)
            bin_ = self.bboxes.to(torch.int32).contiguous()
            bout = torch.empty_like(bin_)
            check(lib().csb_masks_area_resize(ptr(src), K, oh, ow, ptr(out), int(h), int(w), C.c_float(0.3), ptr(bin_), ptr(bout), stream()), "csb_masks_area_resize")
            self.masks, self.bboxes = out.view(torch.bool), bout

    def compose_masks(self, output_type=None):
        """reference :282-298: logical OR over the instances"""
        if self.is_empty:
            return None
        if self.is_numpy:
            mask = self.masks[0] if len(self.masks) == 1 else np.logical_or.reduce(self.masks, 0)
        elif self.masks.is_cuda:
            from .._lib import check, lib, ptr, stream
            import ctypes as C
            K, H, W = self.masks.shape
            src = self.masks.contiguous().view(torch.uint8)
            out = torch.empty((H, W), device=src.device, dtype=torch.uint8)
            check(lib().csb_masks_compose(ptr(src), K, C.c_longlong(H * W), ptr(out), stream()), "csb_masks_compose")
            mask = out.view(torch.bool)
        else:
            mask = self.masks.any(0)
        if output_type is not None:
            if output_type == 'numpy' and not self.is_numpy:
                mask = mask.cpu().numpy()
            if output_type == 'tensor' and not self.is_tensor:
                mask = torch.from_numpy(mask)
        return mask
