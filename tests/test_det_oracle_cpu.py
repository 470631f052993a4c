"""CPU suite: pins the restated (un-vendored, mmdet/mmcv) pieces of the detector oracle against independent implementations that ARE
installed: torchvision's ConvNeXt (same block math) and torchvision.ops.batched_nms."""
import pytest
import torch

from oracle import det_oracle as D


def test_nms_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(0)
    for n in (1, 50, 400):
        xy = torch.rand(n, 2, generator=g) * 200
        wh = torch.rand(n, 2, generator=g) * 60 + 5
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        keep = D.nms_greedy(boxes, scores, 0.6)
        assert torch.equal(keep, tv.ops.nms(boxes, scores, 0.6))


def test_convnext_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    bb = D.ConvNeXt(depths=(1, 1, 2, 1), dims=(32, 64, 128, 256)).eval()
    for p in bb.parameters():
        torch.nn.init.normal_(p, std=0.2)
    from torchvision.models.convnext import CNBlockConfig, ConvNeXt as TV
    m = TV([CNBlockConfig(32, 64, 1), CNBlockConfig(64, 128, 1), CNBlockConfig(128, 256, 2), CNBlockConfig(256, None, 1)], layer_scale=1.0).eval()
    with torch.no_grad():
        m.features[0][0].load_state_dict(bb.downsample_layers[0][0].state_dict()); m.features[0][1].load_state_dict(bb.downsample_layers[0][1].state_dict())
        for i in range(4):
            if i > 0:
                m.features[2 * i][0].load_state_dict(bb.downsample_layers[i][0].state_dict()); m.features[2 * i][1].load_state_dict(bb.downsample_layers[i][1].state_dict())
            for b, blk in enumerate(m.features[2 * i + 1]):
                o = bb.stages[i][b]
                blk.block[0].load_state_dict(o.depthwise_conv.state_dict()); blk.block[2].load_state_dict(o.norm.state_dict())
                blk.block[3].load_state_dict(o.pointwise_conv1.state_dict()); blk.block[5].load_state_dict(o.pointwise_conv2.state_dict())
                blk.layer_scale.copy_(o.gamma.view(-1, 1, 1))
        x = torch.randn(1, 3, 64, 96)
        outs = bb(x)
        f, feats = x, []
        for i in range(8):
            f = m.features[i](f)
            if i in (3, 5, 7):
                feats.append(f)
        for k, (a, b) in enumerate(zip(outs, feats)):
            assert torch.allclose(a, getattr(bb, f'norm{k + 1}')(b), atol=1e-5)


def test_mask_head_matches_direct_mlp():
    """Grouped-conv formulation of the reference (rtmdet_inshead_custom.py:277-294) == a per-pixel MLP with the parsed weights."""
    g = torch.Generator().manual_seed(1)
    mf = torch.randn(8, 12, 20, generator=g)
    K = 5
    kern = torch.randn(K, 169, generator=g) * 0.3
    pri = torch.tensor([[16., 24., 8., 8.], [64., 32., 16., 16.], [0., 0., 32., 32.], [96., 80., 8., 8.], [152., 88., 16., 16.]])
    out = D.mask_predict_by_feat_single(mf, kern, pri)
    ys, xs = torch.meshgrid(torch.arange(12.) * 8, torch.arange(20.) * 8, indexing='ij')
    for k in range(K):
        rel = torch.stack([(pri[k, 0] - xs) / (pri[k, 2] * 8), (pri[k, 1] - ys) / (pri[k, 2] * 8)], 0)
        x = torch.cat([rel, mf], 0).reshape(10, -1)
        w0, w1, w2 = kern[k, :80].view(8, 10), kern[k, 80:144].view(8, 8), kern[k, 144:152].view(1, 8)
        b0, b1, b2 = kern[k, 152:160], kern[k, 160:168], kern[k, 168:169]
        h = torch.relu(w0 @ x + b0[:, None]); h = torch.relu(w1 @ h + b1[:, None]); o = (w2 @ h + b2[:, None]).view(12, 20)
        assert torch.allclose(out[k], o, atol=1e-5)
    assert D.mask_predict_by_feat_single(mf, kern[:0], pri[:0]).shape == (0, 12, 20)
    assert D.mask_tail(out, (96, 160)).shape == (K, 96, 160)


def test_synthetic_weights_load_and_run():
    from cartoonsegmentation_b200.animeinsseg import rtmdet
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    sd = rtmdet.synthetic_state_dict(0)
    m = D.RTMDetIns().eval()
    missing, unexpected = m.load_state_dict(sd, strict=True)
    res = D.infer(m, smooth_image(128, 128, seed=1))
    assert res['masks'].dtype == torch.bool and res['bboxes'].dtype == torch.int32 and len(res['scores']) > 0
    assert sum(p.numel() for p in m.parameters()) > 100e6        # ConvNeXt-B (87.6M) + neck + head
