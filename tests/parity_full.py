"""Full-size parity measurement (TEST INFRASTRUCTURE): the CUDA path against the CPU oracle at the sizes BASELINE.json names
(detector 1024^2, LeReS 640^2 net input), returning the MEASURED errors.  `tests/test_parity_full_gpu.py` asserts bounds of at most
2x these measurements; `python tests/parity_full.py out.json` prints / stores them (profiles/r2_parity_full.json is such a run).

Reference behaviour being matched: animeinsseg/__init__.py:447-462 (`_det_forward`), depth_modules/leres/__init__.py:117-140 (16 -> 8 bit).

Tie band.  A mask pixel is `sigmoid(up8(logit)) > 0.5`, i.e. the SIGN of the bilinearly upsampled logit.  The network runs fp16 x fp16 -> fp32
on the tensor cores against the oracle's fp32, so logits carry a measured absolute error e_logit; pixels whose oracle logit lies within the
band |logit| <= band are decided by that rounding, not by the algorithm.  IoU is reported both raw and outside the band.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def rel_rms(a, b):
    a, b = a.double(), b.double()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


def det_parity(size=1024, seed=1234, backbone='convnext_b', bands=(0.0, 0.02, 0.05, 0.1)):
    """AnimeInsSeg.infer (refine off, det_size == image size) on one seeded 1024^2 image vs oracle.det_oracle.infer."""
    from cartoonsegmentation_b200 import animeinsseg as A
    from cartoonsegmentation_b200.animeinsseg import rtmdet
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    from oracle import det_oracle as D
    from torchvision.ops import box_iou
    torch.set_num_threads(os.cpu_count() or 1)
    sd = rtmdet.synthetic_state_dict(0) if backbone == 'convnext_b' else rtmdet.synthetic_state_dict(0, backbone=backbone)
    oracle = D.RTMDetIns(backbone).eval()
    oracle.load_state_dict(sd, strict=False)
    img = smooth_image(size, size, seed=seed)
    t0 = time.time()
    ref = D.infer(oracle, img)
    t_or = time.time() - t0
    seg = A.AnimeInsSeg(sd, default_det_size=size, refine_kwargs={'refine_method': 'none'})
    cls, reg, ker, mf = seg.model.net.forward(torch.from_numpy(img).cuda())
    nchw = lambda t: t.permute(0, 3, 1, 2).float().cpu()
    res = {'size': size, 'backbone': backbone, 'oracle_seconds': round(t_or, 2), 'heads_rel_rms': {}}
    raw = ref['raw']
    for l in range(3):
        res['heads_rel_rms'][f'cls{l}'] = rel_rms(nchw(cls[l]), raw['cls'][l])
        res['heads_rel_rms'][f'reg{l}'] = rel_rms(nchw(reg[l]), raw['reg'][l])
        res['heads_rel_rms'][f'ker{l}'] = rel_rms(nchw(ker[l]), raw['ker'][l])
    res['heads_rel_rms']['mask_feat'] = rel_rms(nchw(mf), raw['mask_feat'])
    res['heads_rel_rms_max'] = max(res['heads_rel_rms'].values())
    # per-location score error (what selection / NMS see)
    sc = torch.cat([nchw(c).flatten().sigmoid() for c in cls]); so = torch.cat([c.flatten().sigmoid() for c in raw['cls']])
    res['score_abs_err_max'] = (sc - so).abs().max().item()
    # ---- end to end
    out = A.rtmdet_postprocess(cls, reg, ker, mf, (size, size), seg.model.bbox_head.test_cfg)
    k = int(out['num'][0])
    keep = (out['scores'][0, :k] > 0.3).cpu()
    ours_boxes = out['boxes'][0, :k].cpu()[keep]; ours_masks = out['masks'][0, :k].cpu()[keep]; ours_logits = out['logits'][0, :k].cpu()[keep]
    ours_scores = out['scores'][0, :k].cpu()[keep]
    okeep = raw['all_scores'] > 0.3
    o_boxes = raw['boxes'][okeep]; o_logits = raw['logits'][okeep]; o_masks = ref['masks']; o_scores = raw['all_scores'][okeep]
    res['instances_ours'], res['instances_oracle'] = int(keep.sum()), int(okeep.sum())
    iou = box_iou(ours_boxes, o_boxes)
    best, idx = iou.max(1)
    matched = best > 0.9
    res['matched'] = int(matched.sum())
    res['matched_frac'] = res['matched'] / max(1, min(res['instances_ours'], res['instances_oracle']))
    res['same_rank_frac'] = float((idx[matched] == torch.arange(len(idx))[matched]).float().mean()) if matched.any() else 0.0
    mi = matched.nonzero().flatten()
    res['matched_box_abs_err_max'] = (ours_boxes[mi] - o_boxes[idx[mi]]).abs().max().item()
    res['matched_score_abs_err_max'] = (ours_scores[mi] - o_scores[idx[mi]]).abs().max().item()
    lg_o = o_logits[idx[mi]]; lg_g = ours_logits[mi]
    res['logit_abs_err_max'] = (lg_g - lg_o).abs().max().item()
    res['logit_abs_err_rms'] = (lg_g - lg_o).pow(2).mean().sqrt().item()
    res['logit_rms'] = lg_o.pow(2).mean().sqrt().item()
    res['logit_rel_rms'] = rel_rms(lg_g, lg_o)
    ious = {b: [] for b in bands}
    differ = 0
    total = 0
    for j in range(len(mi)):
        a, b = ours_masks[mi[j]], o_masks[idx[mi[j]]]
        up = F.interpolate(lg_o[j][None, None], scale_factor=8, mode='bilinear')[0, 0, :size, :size]
        differ += int((a != b).sum()); total += a.numel()
        for band in bands:
            ok = up.abs() > band if band > 0 else torch.ones_like(a)
            inter, union = (a & b & ok).sum().item(), ((a | b) & ok).sum().item()
            ious[band].append(1.0 if union == 0 else inter / union)
    res['mask_pixels_differ_frac'] = differ / max(1, total)
    res['mask_iou'] = {str(b): {'mean': float(np.mean(v)), 'min': float(np.min(v)), 'n_below_0.999': int((np.array(v) < 0.999).sum())} for b, v in ious.items()}
    return res


def leres_parity(size=640, seed=77):
    """LeReS forward at the reference's 640^2 net input (kenburns_effect.py:563-581) vs oracle.leres_oracle, before and after the 16 -> 8 bit tail."""
    from cartoonsegmentation_b200.depth_modules import leres as L
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    from oracle import leres_oracle as LO
    torch.set_num_threads(os.cpu_count() or 1)
    sd = L.synthetic_state_dict(0)
    oracle = LO.RelDepthModel().eval()
    oracle.load_state_dict(sd, strict=False)
    img = smooth_image(size, size, seed=seed)
    t0 = time.time()
    with torch.no_grad():
        ref = oracle.depth_model(LO.preprocess(img))[0, 0].numpy()
    t_or = time.time() - t0
    net = L.LeReS(sd)
    out = net.forward(torch.from_numpy(img).cuda()[None])[0].float().cpu().numpy()
    rng = float(ref.max() - ref.min())
    d = np.abs(out - ref)
    res = {'size': size, 'oracle_seconds': round(t_or, 2),
           'rel_rms': float(np.sqrt((d ** 2).mean()) / np.sqrt((ref ** 2).mean())),
           'centred_rel_rms': float(np.sqrt((((out - out.mean()) - (ref - ref.mean())) ** 2).mean()) / ref.std()),
           'abs_err_max_over_range': float(d.max() / rng), 'abs_err_rms_over_range': float(np.sqrt((d ** 2).mean()) / rng)}
    qa, qb = LO.quantise_depth(out), LO.quantise_depth(ref)
    q = np.abs(qa.astype(int) - qb.astype(int))
    res['q8_max_levels'] = int(q.max()); res['q8_mean_levels'] = float(q.mean()); res['q8_frac_gt1'] = float((q > 1).mean()); res['q8_frac_ne'] = float((q > 0).mean())
    return res


def main():
    only = sys.argv[2] if len(sys.argv) > 2 else ''
    out = {'device': torch.cuda.get_device_name(0), 'det_1024': det_parity(1024)}
    if only != 'det':
        out['leres_640'] = leres_parity(640)
    if only != 'det' and os.environ.get('CSB_PARITY_CSP', '1') != '0':
        out['det_1024_cspnext_l'] = det_parity(1024, backbone='cspnext_l')
    s = json.dumps(out, indent=1)
    print(s)
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
