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

    def remove_duplicated(self):
        """reference :84-127"""
        num_masks = len(self)
        if num_masks < 2:
            return
        need_cvt = False
        if self.is_numpy:
            need_cvt = True
            self.to_tensor()
        mask_areas = torch.Tensor([mask.sum() for mask in self.masks])
        sids = torch.argsort(mask_areas, descending=True).cpu().numpy().tolist()
        mask_areas = mask_areas[sids]
        masks, bboxes, scores = self.masks[sids], self.bboxes[sids], self.scores[sids]
        tags = [self.tags[sid] for sid in sids]
        canvas = masks[0]
        valid_ids = np.arange(num_masks).tolist()
        for ii, mask in enumerate(masks[1:]):
            mask_id = ii + 1
            and_area = torch.bitwise_and(canvas, mask).sum()
            if and_area / mask_areas[mask_id] > 0.8:
                valid_ids.remove(mask_id)
            elif mask_id != num_masks - 1:
                canvas = torch.bitwise_or(canvas, mask)
        self.masks, self.bboxes, self.scores = masks[valid_ids], bboxes[valid_ids], scores[valid_ids]
        self.tags = [tags[sid] for sid in valid_ids]
        if need_cvt:
            self.to_numpy()

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
        """reference :268-280 (incl. its quirk: x scaled by the height ratio, y by the width ratio)"""
        if self.is_empty:
            return
        if self.is_tensor:
            masks = self.masks.to(torch.float).unsqueeze(1)
            oh, ow = masks.shape[2], masks.shape[3]
            hs, ws = h / oh, w / ow
            bboxes = self.bboxes.float()
            bboxes[:, ::2] *= hs
            bboxes[:, 1::2] *= ws
            self.bboxes = torch.round(bboxes).int()
            masks = torch.nn.functional.interpolate(masks, (h, w), mode=mode)
            self.masks = masks.squeeze(1) > 0.3

    def compose_masks(self, output_type=None):
        if self.is_empty:
            return None
        mask = self.masks[0]
        if len(self.masks) > 1:
            mask = np.logical_or.reduce(self.masks, 0) if self.is_numpy else self.masks.any(0)
        if output_type is not None:
            if output_type == 'numpy' and not self.is_numpy:
                mask = mask.cpu().numpy()
            if output_type == 'tensor' and not self.is_tensor:
                mask = torch.from_numpy(mask)
        return mask
