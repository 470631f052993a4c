"""CPU suite: the oracle against the reference's own dependencies (OpenCV for the uint8 tail), internal properties
of the restated kernels, and the committed golden fixtures (tests/golden, produced on the B200 by running the
UNMODIFIED reference kernels -- see tests/golden/make_kb_golden.py)."""
import os

import numpy as np
import pytest

from oracle import kb_oracle as orc
from tests.kb_scene import BASELINE, FOCAL, make_scene

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_u8_tail_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (H, W, ph, pw) in [(1024, 1024, 993, 993), (256, 320, 249, 310), (100, 100, 97, 97), (101, 77, 60, 81), (64, 64, 63, 62)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        a = cv2.getRectSubPix(image=img, patchSize=(pw, ph), center=(W / 2.0, H / 2.0))
        assert np.array_equal(a, orc.get_rect_sub_pix(img, (pw, ph), (W / 2.0, H / 2.0)))
        c = cv2.resize(src=a, dsize=(W, H), fx=0.0, fy=0.0, interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(c, orc.resize_linear(a, (W, H)))


def test_frame_pack_matches_numpy():
    rng = np.random.default_rng(1)
    r = rng.uniform(-0.2, 1.2, (4, 33, 47)).astype(np.float32)
    ref = (r[0:3].transpose(1, 2, 0) * 255.0).clip(0.0, 255.0).astype(np.uint8)      # kenburns_effect.py:1040
    assert np.array_equal(ref, orc.frame_pack_u8(r))


def test_depth_to_points_matches_torch_restatement():
    torch = pytest.importorskip("torch")
    d = torch.rand(2, 1, 37, 53) * 1000 + 1

    def ref(tenDepth, fltFocal):   # the reference body, anime_3dkenburns/models/utils.py:43-50, on CPU tensors
        W, H = tenDepth.shape[3], tenDepth.shape[2]
        hor = torch.linspace((-0.5 * W) + 0.5, (0.5 * W) - 0.5, W).view(1, 1, 1, -1).repeat(tenDepth.shape[0], 1, H, 1) * (1.0 / fltFocal)
        ver = torch.linspace((-0.5 * H) + 0.5, (0.5 * H) - 0.5, H).view(1, 1, -1, 1).repeat(tenDepth.shape[0], 1, 1, W) * (1.0 / fltFocal)
        return torch.cat([tenDepth * hor, tenDepth * ver, tenDepth], 1)
    for f in (512.0, 300.0):
        assert np.array_equal(ref(d, f).numpy(), orc.depth_to_points(d.numpy(), f))


def test_spatial_filter_matches_torch_restatement():
    torch = pytest.importorskip("torch")
    x = torch.rand(1, 2, 23, 31)
    for k, name in ((3, 'median-3'), (5, 'median-5')):
        t = torch.nn.functional.pad(x, [k // 2] * 4, mode='reflect').unfold(2, k, 1).unfold(3, k, 1)
        t = t.contiguous().view(1, 2, 23, 31, k * k).median(-1, False)[0]
        assert np.array_equal(t.numpy(), orc.spatial_filter(x.numpy(), name))
    w = torch.zeros(2, 2, 3, 3)
    for i in range(2):
        w[i, i, 0, 1] = -1.0; w[i, i, 0, 2] = -1.0; w[i, i, 1, 1] = 4.0; w[i, i, 1, 0] = -1.0; w[i, i, 2, 0] = -1.0
    t = torch.nn.functional.conv2d(torch.nn.functional.pad(x, [1, 1, 1, 1], mode='replicate'), w)
    np.testing.assert_allclose(t.numpy(), orc.spatial_filter(x.numpy(), 'laplacian'), rtol=0, atol=2e-6)
    assert orc.spatial_filter(x.numpy(), 'nope') is None


def test_render_identity_reproduces_image():
    """Unshifted cloud: every valid point lands exactly on its own pixel, so render == image there."""
    s = make_scene(64, 96)
    render, existing = orc.render_pointcloud(s['points'], s['data'], 96, 64, FOCAL, BASELINE)
    valid = s['cloud']['valid'][0, 0] > 0
    img = s['data'].reshape(1, 4, 64, 96)
    assert valid.mean() > 0.5
    np.testing.assert_allclose(render[0, :3][:, valid], img[0, :3][:, valid], atol=2e-5)
    assert np.all(existing[0, 0][valid] > 0.99)


def test_fill_disocclusion_properties():
    s = make_scene(64, 96)
    common = s['common']
    pts, _ = orc.process_shift({'tenPoints': s['points'], 'fltShiftU': 12.0, 'fltShiftV': -7.0, 'fltDepthFrom': common['objDepthrange'][0],
                                'fltDepthTo': common['objDepthrange'][0] * 0.9}, common)
    render, existing = orc.render_pointcloud(pts, s['data'], 96, 64, FOCAL, BASELINE)
    depth = render[:, 3:4] * (existing > 0.0)
    holes = depth[0, 0] <= 0
    assert 0 < holes.sum() < holes.size
    out = orc.fill_disocclusion(render, depth)
    assert np.array_equal(out[0][:, ~holes], render[0][:, ~holes])          # valid pixels untouched
    assert np.array_equal(orc.fill_disocclusion(out, out[:, 3:4] * np.float32(1) + (holes * 0)[None, None] + 1.0), out)  # no holes -> identity
    filled = out[0, 3][holes]
    assert (filled > 0).mean() > 0.5                                         # most holes receive a valid source


def test_autozoom_count_matches_render():
    s = make_scene(64, 96)
    _, existing = orc.render_pointcloud(s['points'], s['data'][:, :3], 96, 64, FOCAL, BASELINE)
    assert orc.count_positive(existing) == int((existing > 0).sum())


def test_disparity_to_cloud_minmaxloc_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    s = make_scene(300, 290)
    depth = s['cloud']['depth'][0, 0]
    mn, mx, lmn, lmx = cv2.minMaxLoc(src=depth[128:-128, 128:-128], mask=None)
    dr = s['cloud']['depthrange']
    assert (np.float32(mn), np.float32(mx), tuple(lmn), tuple(lmx)) == (np.float32(dr[0]), np.float32(dr[1]), dr[2], dr[3])


@pytest.mark.parametrize("name", sorted(f for f in (os.listdir(GOLD) if os.path.isdir(GOLD) else []) if f.startswith("kb_ref_") and f.endswith(".npz")))
def test_oracle_matches_reference_kernel_golden(name):
    """Golden vectors = outputs of the UNMODIFIED reference kernels run on a B200 (tests/golden/make_kb_golden.py)."""
    g = np.load(os.path.join(GOLD, name))
    H, W = int(g['H']), int(g['W'])
    pts, data = g['points'], g['data']
    render, existing, z0, z1 = orc.render_pointcloud(pts, data, W, H, FOCAL, BASELINE, return_zee=True)
    assert np.array_equal(z0, g['zee_pre'])                                  # atomicMin is order independent: bit exact
    # in-place degrid race of the reference: allow a few differing pixels
    assert (z1 != g['zee_post']).mean() < 5e-3
    assert (np.abs(existing - g['existing']) > 1e-4).mean() < 5e-3        # only where the racy in-place degrid differs
    bad = np.abs(render - g['render']) > 1e-3
    assert bad.mean() < 5e-3
    filled = orc.fill_disocclusion(g['render'], g['render'][:, 3:4] * (g['existing'] > 0.0))
    assert np.array_equal(filled, g['filled'])


# ---------------------------------------------------------------------------------------------------------------
# C8 bokeh oracle (oracle/bokeh_oracle.py + orc_bokeh_pass)
# ---------------------------------------------------------------------------------------------------------------
def test_bokeh_oracle_percentile_and_lut():
    from oracle import bokeh_oracle as bo
    rng = np.random.default_rng(0)
    for n in (7, 1000, 20481):
        a = (rng.standard_normal(n) * 7).astype(np.float32)
        for q in (2, 85):
            # numpy 2.x interpolates in float32, numpy 1.26 (the reference's pin) in float64: equal to float32 resolution
            assert bo.percentile_linear(a, q) == pytest.approx(float(np.percentile(a.astype(np.float64), q)), rel=1e-6, abs=1e-6)
    lut = bo.gray_r_bytes()
    assert lut[0] == 255 and lut[-1] == 0 and (np.diff(lut.astype(int)) <= 0).all() and np.abs(lut.astype(int) + np.arange(256) - 255).max() <= 1
    from cartoonsegmentation_b200.utils import effects as fx       # host tables of the product are built by the same expressions
    assert np.array_equal(fx.gray_r_bytes(), lut)
    x = np.arange(256).astype(np.float32) / np.float32(255)
    assert np.array_equal(fx.highlight_table(13), bo.pow_f32(x, 13))
    assert np.abs(fx.highlight_table(13) - np.power(x, 13)).max() <= np.spacing(np.float32(1.0))          # numpy's own float32 pow: within an ulp


def test_bokeh_oracle_pass_matches_python_loops():
    """orc_bokeh_pass (C) against a direct Python transcription of kernel_bokeh's index arithmetic (utils/effects.py:16-72) on a tiny image."""
    import math
    from oracle import bokeh_oracle as bo
    h, w, ns = 9, 11, 8
    rng = np.random.default_rng(3)
    img = rng.random(3 * h * w).astype(np.float32)
    dep = (rng.random(h * w) * 0.02).astype(np.float32)
    dx, dy = np.float32(math.cos(-math.pi / 6)), np.float32(math.sin(-math.pi / 6))
    got = bo.bokeh_pass(img, dep, dx, dy, h, w, ns)
    exp = np.empty_like(img)
    rnd = lambda v: int(math.floor(abs(float(v)) + 0.5)) * (1 if v >= 0 else -1)          # roundf: half away from zero
    for idx in range(3 * h * w):
        smp, c = idx // 3, idx % 3
        y, x = (smp // w) % h, smp % w
        d = dep[y * w + x]
        _dx, _dy = np.float32(dx * d), np.float32(dy * d)
        weight, color = np.float32(0), np.float32(0)
        for s_ in range(ns):
            sp = (s_ - ns // 2) * min(h, w)
            x_, y_ = x + rnd(np.float32(_dx * np.float32(sp))), y + rnd(np.float32(_dy * np.float32(sp)))
            if x_ >= w or y_ >= h or x_ < 0 or y_ < 0:
                continue
            w_ = dep[y_ * w + x_]
            weight = np.float32(weight + w_)
            color = np.float32(float(img[(y_ * w + x_) * 3 + c]) * float(w_) + float(color))           # fma: one rounding
        exp[idx] = np.float32(color / weight) if weight != 0 else img[idx]
    assert np.array_equal(got, exp)


def test_bokeh_oracle_chain_runs():
    from oracle import bokeh_oracle as bo
    rng = np.random.default_rng(1)
    d = (rng.random((40, 56)) * 30 + 5).astype(np.float32)
    d8 = bo.colorize_gray_r(d)
    assert d8.dtype == np.uint8 and d8.min() == 0 and d8.max() == 255
    masks = np.zeros((2, 40, 56), bool)
    masks[0, 5:20, 5:30] = True
    start, end = bo.focal_plane_range(d8, masks)
    assert start in (0.0, 255.0) and end == float(np.median(d8[masks[0]]))
    assert bo.focal_plane_range(d8, None) == (0.0, 255.0)
    img = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    out = bo.bokeh_blur(img, d8, 32, 13, 1, bo.focal_plane(0.5, 50.0, start, end))
    assert out.shape == img.shape and out.dtype == np.uint8


def test_area_upscale_restatement_matches_opencv():
    """orc_resize_area_up_u8c1 (cv2.resize INTER_AREA when upscaling = fixed-point bilinear with area-mode source positions) == cv2, bit for bit."""
    import ctypes as C
    import cv2
    from oracle import kb_oracle
    L = kb_oracle.lib()
    rng = np.random.default_rng(0)
    for (sh, sw, dh, dw) in [(640, 640, 1024, 1024), (480, 640, 720, 960), (96, 128, 100, 131), (33, 47, 97, 50), (64, 64, 64, 64)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        dst = np.empty((dh, dw), np.uint8)
        L.orc_resize_area_up_u8c1(src.ctypes.data_as(C.c_void_p), sh, sw, dh, dw, dst.ctypes.data_as(C.c_void_p))
        assert np.array_equal(dst, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA))


def test_lanczos4_restatement_matches_opencv():
    """orc_resize_lanczos4_u8c1 (cv2.resize INTER_LANCZOS4 on 8UC1: 8-tap fixed point, the LeReS tail for frames smaller than the estimator input,
    reference kenburns_effect.py:573-575) == cv2, bit for bit, for shrinking, growing and mixed geometries."""
    import ctypes as C
    import cv2
    from oracle import kb_oracle
    L = kb_oracle.lib()
    rng = np.random.default_rng(1)
    for (sh, sw, dh, dw) in [(640, 448, 630, 441), (512, 512, 500, 500), (96, 128, 90, 131), (64, 96, 33, 47), (40, 40, 39, 40), (32, 32, 31, 64), (17, 9, 16, 5)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        if sh == 512:
            src[::2] = 255; src[1::2] = 0                         # ringing: exercises the saturation of FixedPtCast
        dst = np.empty((dh, dw), np.uint8)
        L.orc_resize_lanczos4_u8c1(src.ctypes.data_as(C.c_void_p), sh, sw, dh, dw, dst.ctypes.data_as(C.c_void_p))
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4)
        assert np.array_equal(dst, ref), (sh, sw, dh, dw, int(np.abs(dst.astype(int) - ref).max()), float((dst != ref).mean()))
