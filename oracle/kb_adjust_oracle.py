"""TEST INFRASTRUCTURE ONLY -- the reference's own torch formulation of `depth_adjustment_animesseg`
(anime_3dkenburns/kenburns_effect.py:39-91), restated: per instance, in order, flatten the disparity under the mask to the maximum found in the
bottom 3 % of its rows (or, use_medium, to the median of the masked disparities); when the disparity and the image differ in size the map is
bilinearly resized to the image, adjusted, and resized back (:50-56, :86-90).  Runs on whatever device its inputs live on (CPU in the CPU suite).

Parity status: PINNED BY SOURCE -- a line-by-line restatement of in-repo reference code (no third-party semantics involved)."""
import torch


def depth_adjustment_animesseg(masks, tenDisparity, tenImage, use_medium=False):
    """masks: bool [K,H,W] (or None); tenDisparity [1,1,h,w]; tenImage [1,3,H,W] -> adjusted disparity [1,1,h,w]"""
    assert tenDisparity.shape[0] == 1
    tenMasks = [] if masks is None or len(masks) == 0 else [masks[i].float() for i in range(masks.shape[0])]
    resized = tenDisparity.shape[2] != tenImage.shape[2] or tenDisparity.shape[3] != tenImage.shape[3]
    tenAdjusted = torch.nn.functional.interpolate(tenDisparity, size=(tenImage.shape[2], tenImage.shape[3]), mode='bilinear', align_corners=False) \
        if resized else tenDisparity.clone()
    for tenAdjust in tenMasks:
        tenPlane = tenAdjusted * tenAdjust
        if tenPlane.sum().item() == 0:                                             # :68
            continue
        if not use_medium:
            rows = (tenPlane.sum([3], True) > 0.0).flatten().nonzero()             # :74-77
            intTop, intBottom = rows[0].item(), rows[-1].item()
            tenAdjusted = ((1.0 - tenAdjust) * tenAdjusted) + (tenAdjust * tenPlane[:, :, int(round(intTop + (0.97 * (intBottom - intTop)))):, :].max())   # :78
        else:
            tenAdjusted[tenPlane > 0] = tenAdjusted[tenPlane > 0].median()         # :80
    if resized:
        return torch.nn.functional.interpolate(tenAdjusted, size=(tenDisparity.shape[2], tenDisparity.shape[3]), mode='bilinear', align_corners=False)
    return tenAdjusted
