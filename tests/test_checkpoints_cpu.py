"""SURVEY.md §8f rank 3 -- the reference's on-disk checkpoint formats and the video edge, without a GPU: synthetic files written in each
format (animeinsseg/__init__.py:196-208; depth_modules/leres/__init__.py:84-89; zoedepth/models/model_io.py:27-52;
anime_3dkenburns/models/__init__.py:7-20; animeseg_refine/__init__.py:159-165) are read back to the exact tensors, including an mmengine-style
checkpoint that pickles objects of packages which are not installed; `npyframes2video` (kenburns_effect.py:1086-1090) writes the forward + back loop."""
import os
import sys
import types

import numpy as np
import torch

from cartoonsegmentation_b200.utils import checkpoints as CK


def _same(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_mmengine_detector_checkpoint_with_foreign_pickles(tmp_path):
    from cartoonsegmentation_b200.animeinsseg import rtmdet
    sd = {k: v for k, v in list(rtmdet.synthetic_state_dict(0, backbone='cspnext_l').items())[:40]}
    saved = dict(sd)
    saved['backbone.stem.0.bn.num_batches_tracked'] = torch.tensor(7)
    saved['data_preprocessor.mean'] = torch.tensor([103.53, 116.28, 123.675]).view(3, 1, 1)
    saved['ema_backbone_stem_0_conv_weight'] = torch.zeros(2)
    # an object of a package that is not installed (mmengine keeps HistoryBuffer objects in 'message_hub')
    mod = types.ModuleType("mmengine_fake_logging")
    HistoryBuffer = type("HistoryBuffer", (), {"__module__": "mmengine_fake_logging", "__init__": lambda self: setattr(self, 'log', np.arange(3))})
    mod.HistoryBuffer = HistoryBuffer
    sys.modules["mmengine_fake_logging"] = mod
    path = str(tmp_path / "rtmdetl_e60.ckpt")
    torch.save({'meta': {'cfg': "model = dict(type='RTMDet')\n", 'epoch': 60, 'dataset_meta': {'classes': ('object',)}},
                'state_dict': {'module.' + k if i % 2 else k: v for i, (k, v) in enumerate(saved.items())},
                'message_hub': {'log_scalars': {'loss': HistoryBuffer()}}}, path)
    del sys.modules["mmengine_fake_logging"]                   # the reader's process cannot import it
    got = CK.detector_state_dict(path)
    _same(got, sd)
    _same(CK.detector_state_dict(sd), sd)                      # a state_dict passes through


def test_leres_zoe_and_plain_formats(tmp_path):
    g = torch.Generator().manual_seed(0)
    core = {f"depth_model.encoder_modules.encoder.layer1.{i}.conv1.weight": torch.randn(4, 3, generator=g) for i in range(3)}
    p = str(tmp_path / "res101.pth")
    torch.save({'depth_model': {'module.' + k: v for k, v in core.items()}, 'epoch': 3}, p)
    _same(CK.leres_state_dict(p), core)
    torch.save({'depth_model': {k[len('depth_model.'):]: v for k, v in core.items()}}, p)          # keys relative to the inner module
    _same(CK.leres_state_dict(p), core)
    zoe = {"core.core.pretrained.model.cls_token": torch.randn(1, 1, 8, generator=g), "seed_bin_regressor._net.0.weight": torch.randn(4, 4, 1, 1, generator=g)}
    p = str(tmp_path / "ZoeD_M12_N.pt")
    torch.save({'model': {'module.' + k: v for k, v in zoe.items()}}, p)
    _same(CK.zoe_state_dict(p), zoe)
    torch.save(zoe, p)
    _same(CK.zoe_state_dict(p), zoe)
    plain = {"netContext.0.weight": torch.randn(2, 2, 3, 3, generator=g), "netContext.1.weight": torch.randn(2, generator=g)}
    p = str(tmp_path / "kenburns_inpaintnet.ckpt")
    torch.save(plain, p)
    _same(CK.plain_state_dict(p), plain)
    torch.save({'state_dict': plain}, p)
    _same(CK.plain_state_dict(p), plain)


def test_npyframes2video_forward_and_playback(tmp_path):
    import cv2
    import importlib.util
    # kenburns_effect imports the CUDA binding lazily, so the host-side video edge is importable without a GPU
    from cartoonsegmentation_b200.anime_3dkenburns.kenburns_effect import npyframes2video
    frames = [np.full((48, 64, 3), 20 * i, np.uint8) for i in range(6)]
    for playback, expect in ((False, 6), (True, 10)):          # seq + seq[::-1][1:-1]
        path = str(tmp_path / f"out_{playback}.mp4")
        assert npyframes2video(frames, path, playback=playback) == expect
        cap = cv2.VideoCapture(path)
        n = 0
        while True:
            ok, f = cap.read()
            if not ok:
                break
            if n == 0:
                assert f.shape == (48, 64, 3)
            n += 1
        assert n == expect and abs(cap.get(cv2.CAP_PROP_FPS) - 25.0) < 1e-3
