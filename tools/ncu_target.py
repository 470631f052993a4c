"""Short workload for `ncu --set full` captures (one GPU, few launches): `python tools/ncu_target.py det|warp|leres`.
det  : detector forward + post-process, batch 2 @1024^2 (k_conv_tc, k_dwconv_tile, k_layernorm, k_mask_tail ...)
warp : 4 Ken-Burns frames @1024^2 (k_splat, k_zpass, k_degrid, k_fill_holes, k_crop_resize ...)
leres: LeReS forward, batch 2 @640^2"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200.utils.synthetic import smooth_disparity, smooth_image       # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "det"
if what == "det":
    from cartoonsegmentation_b200.animeinsseg import AnimeInsSeg, rtmdet_postprocess
    seg = AnimeInsSeg(None, default_det_size=1024, refine_kwargs={'refine_method': 'none'})
    imgs = torch.from_numpy(np.stack([smooth_image(1024, 1024, seed=i) for i in range(2)])).cuda()
    for _ in range(2):
        cls, reg, ker, mf = seg.model.net.forward(imgs)
        rtmdet_postprocess(cls, reg, ker, mf, (1024, 1024), seg.model.bbox_head.test_cfg)
elif what == "leres":
    from cartoonsegmentation_b200.depth_modules.leres import LeReS
    net = LeReS(None)
    imgs = torch.from_numpy(np.stack([smooth_image(640, 640, seed=i) for i in range(2)])).cuda()
    for _ in range(2):
        net.forward(imgs)
else:
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb
    H = W = 1024
    img = torch.from_numpy(smooth_image(H, W, seed=1)).cuda()
    c = kb.disparity_to_cloud(torch.from_numpy(smooth_disparity(H, W, seed=2)).cuda(), 512.0, 40.0, image_u8=img)
    for i in range(4):
        sh = kb.shift_from_scalars(c['scalars'], W, H, 512.0, 10.0 * i, -5.0 * i, 0.9)
        kb.kenburns_frame(c['points'].view(1, 3, -1), c['data'], W, H, 512.0, 40.0, sh, 993, 993, 512.0, 512.0)
torch.cuda.synchronize()
print("done", what)
