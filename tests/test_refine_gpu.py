"""`Refine` disparity net on the B200 vs golden outputs of the UNMODIFIED reference module (tests/golden/make_refine_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_refine_vs_reference_golden(built_lib):
    sys.path.insert(0, GOLD)
    import make_refine_golden as mk
    from cartoonsegmentation_b200.anime_3dkenburns.models.disparity_refinement import Refine
    net = Refine(None, 'cuda')
    g = np.load(os.path.join(GOLD, "refine_ref.npz"))
    for i, (hw, dhw) in enumerate(mk.CASES):
        img, disp = mk.inputs(hw, dhw, 40 + i)
        y = net.forward(img.cuda(), disp.cuda())
        ref = g[f"out{i}"]
        assert y.shape == (1, 1) + tuple(hw)
        err = float(np.sqrt(((y[0, 0].cpu().numpy() - ref) ** 2).mean() / (ref ** 2).mean()))
        print("Refine rel RMS vs reference module:", hw, dhw, round(err, 5))
        assert err < 3.5e-4, err                                                    # measured 1.7e-4


def test_pipeline_depth_refine_hook(built_lib):
    """KenBurnsPipeline(default_depth_refine=True): `refine_depth` (kenburns_effect.py:619-620,828-829) runs the net on the raw disparity."""
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    from cartoonsegmentation_b200.anime_3dkenburns.kenburns_effect import KenBurnsConfig, KenBurnsPipeline
    from cartoonsegmentation_b200.utils.synthetic import smooth_disparity, smooth_image
    pipe = KenBurnsPipeline(KenBurnsConfig(det_size=320, max_size=320, num_frame=3, depth_est='external', default_depth_refine=True))
    img, raw = smooth_image(288, 320, seed=5), smooth_disparity(288, 320, seed=6)
    pipe.depth_model = lambda im, t: torch.from_numpy(raw).cuda().reshape(1, 1, 288, 320)
    kcfg = pipe.generate_kenburns_config(img, instances=AnimeInstances())
    d = kcfg['tenRawDisparity']
    assert d.shape == (1, 1, 288, 320) and torch.isfinite(d).all()
    plain = KenBurnsPipeline(KenBurnsConfig(det_size=320, max_size=320, num_frame=3, depth_est='external'))
    plain.depth_model = pipe.depth_model
    d0 = plain.generate_kenburns_config(img, instances=AnimeInstances())['tenRawDisparity']
    assert (d - d0).abs().max() > 1e-3                                          # the refinement changed the map
