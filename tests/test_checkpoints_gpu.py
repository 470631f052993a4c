"""SURVEY.md §8f rank 3 on the device: the networks built from checkpoint FILES in the reference's formats give the same outputs as the same weights
passed as state_dicts (AnimeInsSeg(ckpt) animeinsseg/__init__.py:196-208; set_depth_estimation / set_inpainting / set_depth_refinement /
set_refine_method with the reference's loaders' formats)."""
import numpy as np
import pytest
import torch

from cartoonsegmentation_b200.utils.synthetic import smooth_image

pytestmark = pytest.mark.gpu


def test_animeinsseg_from_mmengine_checkpoint_file(built_lib, tmp_path):
    from cartoonsegmentation_b200.animeinsseg import AnimeInsSeg, rtmdet as R
    from cartoonsegmentation_b200.animeinsseg import isnet as I
    sd = R.synthetic_state_dict(0, backbone='cspnext_l')                      # the layout of the shipped rtmdetl_e60.ckpt
    det = str(tmp_path / "rtmdetl_e60.ckpt")
    extra = {'backbone.stem.0.bn.num_batches_tracked': torch.tensor(1), 'data_preprocessor.mean': torch.zeros(3, 1, 1)}
    torch.save({'meta': {'cfg': "model = dict(type='RTMDet')"}, 'state_dict': {**sd, **extra}}, det)
    ref_sd = I.synthetic_state_dict(0)
    rfile = str(tmp_path / "refine_last.ckpt")
    torch.save(ref_sd, rfile)
    img = smooth_image(320, 320, seed=5)
    a = AnimeInsSeg(det, default_det_size=320, refine_kwargs={'refine_method': 'refinenet_isnet', 'refine_size': 128, 'refine_ckpt': rfile})
    b = AnimeInsSeg(sd, default_det_size=320, refine_kwargs={'refine_method': 'refinenet_isnet', 'refine_size': 128})
    assert a.model.net.cspnext is not None
    ia, ib = a.infer(img, pred_score_thr=0.0), b.infer(img, pred_score_thr=0.0)
    assert len(ia) == len(ib)
    if len(ia):
        assert torch.equal(ia.masks, ib.masks) and torch.equal(ia.bboxes, ib.bboxes) and torch.equal(ia.scores, ib.scores)


def test_pipeline_components_from_checkpoint_files(built_lib, tmp_path):
    from cartoonsegmentation_b200.anime_3dkenburns.kenburns_effect import KenBurnsConfig, KenBurnsPipeline
    from cartoonsegmentation_b200.anime_3dkenburns.models import disparity_refinement as DR, pointcloud_inpainting as PI
    from cartoonsegmentation_b200.depth_modules import leres as L
    lsd = L.synthetic_state_dict(1)
    lfile = str(tmp_path / "res101.pth")
    torch.save({'depth_model': {'module.' + k: v for k, v in lsd.items()}}, lfile)
    isd, rsd = PI.synthetic_state_dict(1), DR.synthetic_state_dict(1)
    ifile, rfile = str(tmp_path / "inpaint.ckpt"), str(tmp_path / "refine.ckpt")
    torch.save(isd, ifile)
    torch.save(rsd, rfile)
    cfg = KenBurnsConfig(det_size=320, max_size=512, num_frame=2, depth_est='external', refine_crf=False)
    pipe = KenBurnsPipeline(cfg)
    pipe.kenburns_inpaintnet = None
    pipe.set_depth_estimation('leres', lfile)
    pipe.set_inpainting('default', ifile)
    pipe.set_depth_refinement('default', rfile)
    img = smooth_image(288, 320, seed=9)
    a = pipe._depth_est_leres_batch([img])[0]
    ref = L.LeReS(lsd)
    pipe2 = KenBurnsPipeline(KenBurnsConfig(det_size=320, max_size=512, num_frame=2, depth_est='external', refine_crf=False))
    pipe2.leres = ref
    b = pipe2._depth_est_leres_batch([img])[0]
    assert torch.equal(a, b)
    x = torch.rand(1, 3, 96, 128, device='cuda')
    d = torch.rand(1, 1, 96, 128, device='cuda') * 30 + 5
    assert torch.equal(pipe.depth_refinenet.forward(x, d), DR.Refine(rsd).forward(x, d))
    w = pipe.kenburns_inpaintnet
    w2 = PI.Inpaint(isd)
    common = {'fltFocal': 512.0, 'fltBaseline': 40.0, 'intWidth': 128, 'intHeight': 96}
    o1 = w.forward(x, d, np.array([3.0, -2.0, 1.0], np.float32), common)
    o2 = w2.forward(x, d, np.array([3.0, -2.0, 1.0], np.float32), common)
    assert (o1['tenImage'] - o2['tenImage']).abs().max().item() < 1e-3          # the context splat is fp32-atomic-order dependent
