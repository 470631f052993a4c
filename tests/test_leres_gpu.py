"""GPU parity of the LeReS depth forward against golden outputs of the UNMODIFIED reference network run on
the CPU with the same seeded weights (tests/golden/make_leres_golden.py), plus the new engine pieces against plain PyTorch."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng(built_lib):
    from cartoonsegmentation_b200 import engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return engine


@pytest.mark.parametrize("C,groups,stride,H,W", [(256, 32, 1, 20, 24), (512, 32, 2, 24, 20), (1024, 32, 1, 10, 12), (2048, 32, 2, 8, 8), (128, 2, 1, 16, 16)])
def test_grouped_conv_vs_torch(eng, C, groups, stride, H, W):
    g = torch.Generator(device='cuda').manual_seed(C + stride)
    x = torch.randn(2, H, W, C, device='cuda', generator=g).half()
    w = torch.randn(C, C // groups, 3, 3, device='cuda', generator=g) / (9 * C // groups) ** 0.5
    b = torch.randn(C, device='cuda', generator=g) * 0.1
    y = eng.conv2d_nhwc(x, eng.pack_grouped_weight(w, groups), b, stride=stride, pad=1, act='relu', groups=groups)
    r = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, stride=stride, padding=1, groups=groups)).permute(0, 2, 3, 1)
    assert y.shape == r.shape
    assert (y.float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3


def test_pool_add_resample_vs_torch(eng):
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(2, 17, 22, 64, device='cuda', generator=g).half()
    r = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(eng.maxpool3s2_nhwc(x).float(), r)
    y = torch.randn_like(x)
    assert (eng.add_nhwc(x, y).float() - (x.float() + y.float())).abs().max().item() < 4e-3
    d = torch.randn(2, 13, 19, device='cuda', generator=g)
    for ac in (True, False):
        o = eng.resample_f32(d, 26, 38, ac)
        rr = F.interpolate(d[:, None], size=(26, 38), mode='bilinear', align_corners=ac)[:, 0]
        assert (o - rr).abs().max().item() < 1e-5
    s1 = torch.randn(1, 16, 16, 64, device='cuda', generator=g).half()              # 1x1 stride-2 conv (ResNeXt downsample)
    w = torch.randn(128, 64, 1, 1, device='cuda', generator=g) / 8
    o = eng.conv2d_nhwc(s1, eng.pack_conv_weight(w), None, stride=2)
    rr = F.conv2d(s1.float().permute(0, 3, 1, 2), w.half().float(), stride=2).permute(0, 2, 3, 1)
    assert (o.float() - rr).abs().max().item() <= 2e-3 * rr.abs().max().item() + 1e-3


@pytest.mark.parametrize("name", ["leres_ref_96x128.npz", "leres_ref_160x128.npz"])
def test_leres_forward_vs_reference_golden(eng, name):
    from cartoonsegmentation_b200.depth_modules import leres as L
    gold = np.load(os.path.join(GOLD, name))
    net = L.LeReS(L.synthetic_state_dict(0))
    out = net.forward(torch.from_numpy(gold['image']).cuda())[0].cpu().numpy()
    ref = gold['depth']
    rel = np.sqrt(((out - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean())
    rel_c = np.sqrt((((out - out.mean()) - (ref - ref.mean())) ** 2).mean()) / ref.std()
    print(f"{name}: relative RMS error {rel:.5f}, centred (what min-max normalisation sees) {rel_c:.5f}")
    assert out.shape == ref.shape and rel < 1e-3 and rel_c < 3.5e-3     # measured 4.4e-4 / 5.7e-4 (centred 1.3e-3 / 1.8e-3); north_star 1e-3
    # after the reference's own 16 -> 8 bit quantisation (apply_leres) the maps agree to one grey level
    qa, qb = L.quantise_depth(out), L.quantise_depth(ref)
    d = np.abs(qa.astype(int) - qb.astype(int))
    print(f"   8-bit depth image: max |diff| {d.max()} levels, mean {d.mean():.3f}")
    assert d.max() <= 1 and d.mean() < 0.17                                      # measured: max 1 level, mean 0.05 / 0.085 (+-1 LSB)


@pytest.mark.parametrize("hw,HW", [((640, 640), (1024, 1024)), ((480, 640), (720, 960)), ((96, 128), (100, 131)), ((64, 96), (64, 96))])
def test_device_tail_equals_host_numpy_opencv(built_lib, hw, HW):
    """csb_leres_depth_tail == the reference's host tail (apply_leres :117-140 + kenburns_effect.py:572-577) run with numpy + OpenCV, exactly."""
    import cv2
    from cartoonsegmentation_b200._lib import check, lib, ptr, stream
    from cartoonsegmentation_b200.depth_modules.leres import quantise_depth
    (h, w), (H, W) = hw, HW
    n = 3
    rng = np.random.default_rng(h + W)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    logits = np.stack([(3 + np.sin(xx / (7.0 + i)) * np.cos(yy / 11.0) * 2 + rng.standard_normal((h, w)) * 0.05).astype(np.float32) for i in range(n)])
    logits[2] = 1.5                                                             # constant map: max - min <= eps -> zeros -> 255 everywhere
    dev = torch.from_numpy(logits).cuda()
    mm = torch.empty(2 * n, device='cuda', dtype=torch.int32)
    q8 = torch.empty((n, h, w), device='cuda', dtype=torch.uint8)
    out = torch.empty((n, H, W), device='cuda', dtype=torch.float32)
    check(lib().csb_leres_depth_tail(ptr(dev), n, h, w, H, W, ptr(mm), ptr(q8), ptr(out), stream()), "csb_leres_depth_tail")
    for i in range(n):
        d8 = quantise_depth(logits[i])
        assert np.array_equal(q8[i].cpu().numpy(), d8)
        ref = cv2.resize(d8, (W, H), interpolation=cv2.INTER_AREA).astype(np.float32)
        assert np.array_equal(out[i].cpu().numpy(), ref)


@pytest.mark.parametrize("hw,HW", [((640, 448), (630, 441)), ((512, 512), (500, 500)), ((96, 128), (90, 131)), ((64, 96), (33, 47)), ((32, 32), (31, 64))])
def test_device_tail_lanczos_branch_equals_opencv(built_lib, hw, HW):
    """kenburns_effect.py:573-575: more estimator rows than frame rows -> cv2.resize(..., INTER_LANCZOS4).  The device branch (k_lt_lanczos, OpenCV's
    8-tap fixed-point arithmetic) == cv2 and == the C restatement, bit for bit, incl. a ringing (0 / 255 stripes) image that saturates."""
    import ctypes as C
    import cv2
    from cartoonsegmentation_b200._lib import check, lib, ptr, stream
    from cartoonsegmentation_b200.depth_modules.leres import quantise_depth
    from oracle import kb_oracle
    (h, w), (H, W) = hw, HW
    n = 3
    rng = np.random.default_rng(h + W)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    logits = np.stack([(3 + np.sin(xx / (7.0 + i)) * np.cos(yy / 11.0) * 2 + rng.standard_normal((h, w)) * 0.05).astype(np.float32) for i in range(n)])
    logits[1] = np.where((yy.astype(int) % 2) == 0, 0.0, 1.0)                    # stripes: every tap overshoots
    logits[2] = rng.random((h, w), dtype=np.float32)
    dev = torch.from_numpy(logits).cuda()
    mm = torch.empty(2 * n, device='cuda', dtype=torch.int32)
    q8 = torch.empty((n, h, w), device='cuda', dtype=torch.uint8)
    out = torch.empty((n, H, W), device='cuda', dtype=torch.float32)
    for _ in range(2):                                                          # second call: cached coefficient tables
        check(lib().csb_leres_depth_tail(ptr(dev), n, h, w, H, W, ptr(mm), ptr(q8), ptr(out), stream()), "csb_leres_depth_tail")
    for i in range(n):
        d8 = quantise_depth(logits[i])
        assert np.array_equal(q8[i].cpu().numpy(), d8)
        ref = cv2.resize(d8, (W, H), interpolation=cv2.INTER_LANCZOS4)
        orc = np.empty((H, W), np.uint8)
        kb_oracle.lib().orc_resize_lanczos4_u8c1(d8.ctypes.data_as(C.c_void_p), h, w, H, W, orc.ctypes.data_as(C.c_void_p))
        assert np.array_equal(orc, ref)
        assert np.array_equal(out[i].cpu().numpy(), ref.astype(np.float32))


def test_pipeline_leres_small_frame_takes_device_lanczos(built_lib):
    """310 x 215 and 310 x 200 frames with depth_est_size 320: scaledown_maxsize rounds the estimator input to 320 x 224 / 320 x 192 -> k > 1 -> LANCZOS4
    on both axes (the width shrinks or grows), on the device"""
    from cartoonsegmentation_b200.anime_3dkenburns.kenburns_effect import KenBurnsConfig, KenBurnsPipeline
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    from cartoonsegmentation_b200 import _lib
    pipe = KenBurnsPipeline(KenBurnsConfig(det_size=320, max_size=1024, num_frame=3, depth_est='leres', depth_est_size=320))
    for (fh, fw) in ((310, 215), (310, 200)):
        imgs = [smooth_image(fh, fw, seed=5 + i) for i in range(2)]
        pipe.leres_host_tail = False
        l0 = _lib.launch_count()
        a = pipe._depth_est_leres_batch(imgs)
        assert _lib.launch_count() > l0
        pipe.leres_host_tail = True
        b = pipe._depth_est_leres_batch(imgs)
        for x, y in zip(a, b):
            assert x.shape == (1, 1, fh, fw) and torch.equal(x, y)


def test_pipeline_leres_device_tail_matches_host_tail(built_lib):
    from cartoonsegmentation_b200.anime_3dkenburns.kenburns_effect import KenBurnsConfig, KenBurnsPipeline
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    pipe = KenBurnsPipeline(KenBurnsConfig(det_size=320, max_size=1024, num_frame=3, depth_est='leres', depth_est_size=320))
    imgs = [smooth_image(512, 640, seed=31 + i) for i in range(2)]
    a = pipe._depth_est_leres_batch(imgs)
    pipe.leres_host_tail = True
    b = pipe._depth_est_leres_batch(imgs)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("CP", [16, 64])
def test_space_to_depth_prep_p2(built_lib, CP):
    """csb_image_prep_s2d_nhwc, P = 2 (the LeReS stem input): channel (dy*2+dx)*3 + c = normalised pixel (2Y+dy, 2X+dx), zero padded to CP; the
    specialised one-thread-per-cell kernel against the torch formula, with and without the R/B swap, even and 4-unaligned widths."""
    from cartoonsegmentation_b200 import engine as E
    mean, std = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
    for (N, H, W), swap in (((2, 12, 20), False), ((1, 6, 10), True), ((3, 64, 62), True)):
        img = torch.randint(0, 256, (N, H, W, 3), device='cuda', dtype=torch.uint8)
        out = E.image_prep_s2d_nhwc(img, mean, std, 2, CP, swap_rb=swap)
        src = img.flip(-1) if swap else img
        ref = (src.float() - torch.tensor(mean, device='cuda')) / torch.tensor(std, device='cuda')
        ref = ref.view(N, H // 2, 2, W // 2, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, W // 2, 12)
        assert out.shape == (N, H // 2, W // 2, CP)
        assert (out[..., :12].float() - ref).abs().max().item() < 2e-3 and not out[..., 12:].any()
