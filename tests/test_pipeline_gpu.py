"""GPU: the reference call surface end to end -- KenBurnsPipeline(cfg).generate_kenburns_config(img) -> process_kenburns / autozoom
(reference run_kenburns.py:19-33, naive_interface.py:160-165) -- against the oracle composition of the same steps."""
import numpy as np
import pytest
import torch

from cartoonsegmentation_b200.utils.synthetic import smooth_disparity, smooth_image
from oracle import kb_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe(built_lib):
    from cartoonsegmentation_b200.anime_3dkenburns.kenburns_effect import KenBurnsConfig, KenBurnsPipeline
    cfg = KenBurnsConfig(det_size=320, max_size=320, num_frame=5, depth_est='external')
    return KenBurnsPipeline(cfg)


def test_generate_config_and_frames_vs_oracle(pipe):
    H, W = 288, 320
    img = smooth_image(H, W, seed=5)
    raw = smooth_disparity(H, W, seed=6)
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    kcfg = pipe.generate_kenburns_config(img, instances=AnimeInstances(), disparity=torch.from_numpy(raw).cuda())
    o = orc.disparity_to_cloud(raw, kcfg.focal, kcfg.baseline)
    assert kcfg['objDepthrange'] == o['depthrange'] and kcfg['fltDispmax'] == o['dispmax'] and kcfg['fltDispmin'] == o['dispmin']
    assert np.array_equal(kcfg['tenRawPoints'].cpu().numpy().reshape(o['points'].shape), o['points'])
    assert (kcfg.int_height, kcfg.int_width) == (H, W) and kcfg['tenInpaPoints'].shape == (1, 3, H * W)
    # dict-style aliases of the upstream objCommon (reference :291-363)
    assert kcfg['fltFocal'] == kcfg.focal and kcfg['intWidth'] == W
    objFrom = {'fltCenterU': W / 2.0, 'fltCenterV': H / 2.0, 'intCropWidth': int(np.floor(0.97 * W)), 'intCropHeight': int(np.floor(0.97 * H))}
    objTo = {'fltCenterU': W / 2.0 + 20.0, 'fltCenterV': H / 2.0 - 13.3, 'intCropWidth': int(round(objFrom['intCropWidth'] / 1.25)),
             'intCropHeight': int(round(objFrom['intCropHeight'] / 1.25))}
    steps = np.linspace(0.0, 1.0, 4).tolist()
    frames, _ = pipe.process_kenburns({'fltSteps': steps, 'objFrom': objFrom, 'objTo': objTo, 'boolInpaint': False}, kcfg, inpaint=False)
    assert len(frames) == 4 and frames[0].shape == (H, W, 3) and frames[0].dtype == np.uint8
    img_t = np.ascontiguousarray(img.transpose(2, 0, 1)[None].astype(np.float32) * np.float32(1.0 / 255.0))
    data = np.concatenate([img_t.reshape(1, 3, -1), o['depth'].reshape(1, 1, -1)], 1)
    common = {'objDepthrange': o['depthrange'], 'intWidth': W, 'intHeight': H, 'fltFocal': kcfg.focal, 'fltBaseline': kcfg.baseline}
    for fltStep, frame in zip(steps, frames):
        fltFrom, fltTo = 1.0 - fltStep, 1.0 - (1.0 - fltStep)
        su = ((fltFrom * objFrom['fltCenterU']) + (fltTo * objTo['fltCenterU'])) - (W / 2.0)
        sv = ((fltFrom * objFrom['fltCenterV']) + (fltTo * objTo['fltCenterV'])) - (H / 2.0)
        cw = (fltFrom * objFrom['intCropWidth']) + (fltTo * objTo['intCropWidth'])
        dto = o['depthrange'][0] * (cw / max(objFrom['intCropWidth'], objTo['intCropWidth']))
        p, _ = orc.process_shift({'tenPoints': o['points'].reshape(1, 3, -1), 'fltShiftU': su, 'fltShiftV': sv, 'fltDepthFrom': o['depthrange'][0], 'fltDepthTo': dto}, common)
        r, e = orc.render_pointcloud(p, data, W, H, kcfg.focal, kcfg.baseline)
        f = orc.fill_disocclusion(r, r[:, 3:4] * (e > 0.0))
        expect = orc.resize_linear(orc.get_rect_sub_pix(orc.frame_pack_u8(f[0]), (objFrom['intCropWidth'], objFrom['intCropHeight']), (W / 2.0, H / 2.0)), (W, H))
        diff = np.abs(frame.astype(int) - expect.astype(int))
        assert diff.max() <= 2 and (diff > 0).mean() < 0.08, (fltStep, diff.max(), (diff > 0).mean())     # the reference's own fp32-order envelope
    frames2 = pipe.autozoom(kcfg, inpaint=False)
    assert len(frames2) == kcfg.num_frame
    with pytest.raises(NotImplementedError):
        pipe.autozoom(kcfg)                      # inpaint=True needs the Inpaint net: fails loudly, never silently skipped


def test_pipeline_with_detector_instances(pipe):
    H, W = 320, 320
    img = smooth_image(H, W, seed=9)
    raw = smooth_disparity(H, W, seed=10)
    kcfg = pipe.generate_kenburns_config(img, disparity=torch.from_numpy(raw).cuda())       # runs AnimeInsSeg.infer + depth adjustment
    assert kcfg.instances is not None and kcfg['tenRawDisparity'].shape == (1, 1, H, W)
    assert float(kcfg['tenRawDisparity'].max()) == pytest.approx(kcfg.baseline, rel=1e-6)
