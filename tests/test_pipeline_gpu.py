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
    n0 = kcfg['tenInpaPoints'].shape[2]
    frames3 = pipe.autozoom(kcfg)                # the reference's run_kenburns.py path: autozoom + 2 inpaint passes + num_frame frames
    assert len(frames3) == kcfg.num_frame and frames3[0].shape == (H, W, 3)
    n1 = kcfg['tenInpaPoints'].shape[2]
    assert n1 > n0 and kcfg.inpainted_img.shape[2] == n1 and kcfg['tenInpaDepth'].shape[2] == n1      # the cloud grew by the inpainted holes
    # inpainting can only add points: the frames of the inpainted cloud have no more holes than before at the end pose
    assert (frames3[-1].sum(-1) == 0).sum() <= (frames2[-1].sum(-1) == 0).sum()


def test_pipeline_with_detector_instances(pipe):
    H, W = 320, 320
    img = smooth_image(H, W, seed=9)
    raw = smooth_disparity(H, W, seed=10)
    kcfg = pipe.generate_kenburns_config(img, disparity=torch.from_numpy(raw).cuda())       # runs AnimeInsSeg.infer + depth adjustment
    assert kcfg.instances is not None and kcfg['tenRawDisparity'].shape == (1, 1, H, W)
    assert float(kcfg['tenRawDisparity'].max()) == pytest.approx(kcfg.baseline, rel=1e-6)


def test_depth_adjustment_kernel_matches_reference_formulation(built_lib):
    """csb_depth_adjust_instances == the reference's torch code (kenburns_effect.py:39-91) run on the GPU, bit for bit, with overlapping,
    empty and full-frame masks."""
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb
    from cartoonsegmentation_b200.utils.synthetic import ellipse_masks
    H, W = 200, 260
    disp = torch.from_numpy(smooth_disparity(H, W, seed=3)).cuda()
    masks = torch.from_numpy(ellipse_masks(H, W, k=9, seed=4)).cuda()
    masks[3] = False                         # empty instance -> skipped
    masks[5] = True                          # full-frame instance
    masks[7, H - 1, :] = True                # touches the last row
    inst = AnimeInstances(masks, torch.zeros(9, 4, dtype=torch.int32, device='cuda'), torch.ones(9, device='cuda'))
    img = torch.zeros(1, 3, H, W, device='cuda')
    a = kb.depth_adjustment_animesseg(inst, disp, img)                       # csb_depth_adjust_batch (one cooperative launch)
    from oracle import kb_adjust_oracle as AO                                # the reference's torch formulation, run on the GPU tensors
    b = AO.depth_adjustment_animesseg(inst.masks, disp.clone(), img)
    assert torch.equal(a, b)
    from cartoonsegmentation_b200._lib import check, lib, ptr, stream
    c = disp.clone().contiguous()
    state = torch.empty(4, device='cuda', dtype=torch.int32)                 # csb_depth_adjust_instances (per-instance launches)
    check(lib().csb_depth_adjust_instances(ptr(c), ptr(masks.contiguous().view(torch.uint8)), 9, H, W, ptr(state), stream()))
    assert torch.equal(c, b)
    # batch of 3 images with different instance counts (0, 9, 4)
    d3 = disp[0].repeat(3, 1, 1).contiguous()
    m3 = masks[None].repeat(3, 1, 1, 1).contiguous()
    kb.depth_adjust_batch(d3, m3, torch.tensor([0, 9, 4], device='cuda', dtype=torch.int32))
    assert torch.equal(d3[0], disp[0, 0]) and torch.equal(d3[1], b[0, 0])
    b4 = AO.depth_adjustment_animesseg(masks[:4], disp.clone(), img)
    assert torch.equal(d3[2], b4[0, 0])
    assert torch.equal(kb.depth_adjustment_animesseg(AnimeInstances(), disp, img), disp)


def test_inpaint_net_vs_reference_golden(built_lib):
    """Inpaint.forward (tcgen05 engine + context render kernels) against the UNMODIFIED reference module run on the CPU with the same seeded
    weights (tests/golden/make_inpaint_golden.py)."""
    import os
    from cartoonsegmentation_b200.anime_3dkenburns.models import pointcloud_inpainting as P
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "inpaint_ref_96x128.npz"))
    H, W = g['disparity'].shape[-2:]
    net = P.Inpaint(P.synthetic_state_dict(0))
    img = torch.from_numpy(np.ascontiguousarray(g['image'].transpose(2, 0, 1)[None].astype(np.float32) * (1.0 / 255.0))).cuda()
    o = net.forward(img, torch.from_numpy(g['disparity']).cuda(), g['shift'], {'fltFocal': 512.0, 'fltBaseline': 40.0, 'intWidth': W, 'intHeight': H})
    ex = o['tenExisting'].cpu().numpy()
    assert (ex != g['tenExisting']).mean() < 2e-3                                   # coverage + median-5: discrete, order-noise only at the z-test margin
    for k in ('tenImage', 'tenDisparity'):
        a, b = o[k].cpu().numpy(), g[k]
        rel = np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean())
        print(f"{k}: relative RMS error {rel:.5f}, max abs {np.abs(a - b).max():.4f}")
        assert a.shape == b.shape and rel < 8.5e-4                                 # measured 3.4e-4 (tenImage) / 4.1e-4 (tenDisparity)


@pytest.mark.parametrize("H,W,K,mode", [(200, 260, 9, 'positive'), (200, 260, 9, 'zeros'), (96, 130, 5, 'positive'), (256, 512, 100, 'positive'), (128, 256, 40, 'negative')])
def test_depth_adjust_batch_paths_match_reference_formulation(built_lib, H, W, K, mode):
    """csb_depth_adjust_batch: the order-free path (positive disparities, W % 4 == 0, K <= 128) and the device-selected sequential fallback
    (zeros / negatives under a mask, other widths) are both bit-identical to the reference's torch loop (kenburns_effect.py:39-91), with heavily
    overlapping masks, empty instances and different instance counts per image."""
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb
    rng = np.random.default_rng(H + K)
    disp = torch.from_numpy(smooth_disparity(H, W, seed=3)).cuda().reshape(1, 1, H, W).clone()
    if mode == 'zeros':
        disp[0, 0, 40:90, 30:200] = 0.0
    if mode == 'negative':
        disp[0, 0, 10:60, 100:180] *= -1.0
    yy, xx = np.mgrid[0:H, 0:W]
    masks = np.zeros((K, H, W), bool)
    for k in range(K):                                   # big overlapping ellipses + a few thin / empty ones
        cy, cx, ry, rx = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(4, H / 1.5), rng.uniform(4, W / 1.5)
        masks[k] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
    masks[K // 2] = False
    masks[K - 1, H - 1, :] = True
    m = torch.from_numpy(masks).cuda()
    img = torch.zeros(1, 3, H, W, device='cuda')
    counts = [K, 0, max(1, K // 3)]
    d3 = disp[0].repeat(3, 1, 1).contiguous()
    kb.depth_adjust_batch(d3, m[None].repeat(3, 1, 1, 1).contiguous(), torch.tensor(counts, device='cuda', dtype=torch.int32))
    for i, c in enumerate(counts):
        inst = AnimeInstances(m[:c], torch.zeros(c, 4, dtype=torch.int32, device='cuda'), torch.ones(c, device='cuda')) if c else AnimeInstances()
        from oracle import kb_adjust_oracle as AO
        ref = AO.depth_adjustment_animesseg(None if inst.is_empty else inst.masks, disp.clone(), img)
        assert torch.equal(d3[i], ref[0, 0]), (mode, i, float((d3[i] - ref[0, 0]).abs().max()))


def test_depth_adjustment_median_and_resized_variants(built_lib):
    """use_medium=True (kenburns_effect.py:80, exact lower median by radix select) and the resized form (:50-56, :86-90) against the reference's
    torch formulation (oracle/kb_adjust_oracle.py) on the same tensors."""
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb
    from cartoonsegmentation_b200.utils.synthetic import ellipse_masks
    from oracle import kb_adjust_oracle as AO
    H, W = 200, 260
    disp = torch.from_numpy(smooth_disparity(H, W, seed=5)).cuda()
    masks = torch.from_numpy(ellipse_masks(H, W, k=9, seed=6)).cuda()
    masks[3] = False
    masks[5] = True
    inst = AnimeInstances(masks, torch.zeros(9, 4, dtype=torch.int32, device='cuda'), torch.ones(9, device='cuda'))
    img = torch.zeros(1, 3, H, W, device='cuda')
    a = kb.depth_adjustment_animesseg(inst, disp, img, use_medium=True)
    b = AO.depth_adjustment_animesseg(masks, disp.clone(), img, use_medium=True)
    assert torch.equal(a, b)                                                  # selection, not arithmetic: exact
    small = torch.nn.functional.interpolate(disp, size=(H // 2, W // 2 + 3), mode='bilinear', align_corners=False).contiguous()
    for um in (False, True):
        a = kb.depth_adjustment_animesseg(inst, small, img, use_medium=um)
        b = AO.depth_adjustment_animesseg(masks, small.clone(), img, use_medium=um)
        assert a.shape == b.shape == small.shape
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item()      # two bilinear resamplings: fp32 rounding order only


@pytest.mark.gpu
def test_net_io_kernels_equal_the_torch_ops_of_the_reference(built_lib):
    """csrc/kb_netio.cu against the eager torch expressions of pointcloud_inpainting.py:117-131, 190-200 and disparity_refinement.py:99-100, 128-135:
    statistics to fp32 rounding; every elementwise kernel BIT-EXACT when fed the same statistics."""
    import ctypes as C
    from cartoonsegmentation_b200._lib import check, lib, ptr, stream
    from cartoonsegmentation_b200.anime_3dkenburns.models import utils as U
    H, W, focal, baseline = 104, 136, 512.0, 40.0
    g = torch.Generator(device='cuda').manual_seed(5)
    img = torch.rand((1, 3, H, W), device='cuda', generator=g)
    disp = (torch.from_numpy(smooth_disparity(H, W, seed=3)).cuda().reshape(1, 1, H, W) * 0.7 + 0.05).contiguous()
    # ---- statistics
    for t in (img, disp):
        st = U.tensor_stats(t).cpu()
        ref = torch.stack([t.mean(), t.std(unbiased=False), t.max()]).cpu()
        assert (st - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() and st[2] == ref[2]
    st_i = torch.stack([img.mean(), img.std(unbiased=False), img.max()]).contiguous()            # torch's own statistics -> exact comparisons below
    st_d = torch.stack([disp.mean(), disp.std(unbiased=False), disp.max()]).contiguous()
    # ---- normalise + pack
    x16 = U.pack_norm16(img, st_i, disp, st_d)
    ref16 = torch.zeros((1, H, W, 16), device='cuda', dtype=torch.float16)
    ref16[0, ..., :3] = ((img - st_i[0]) / (st_i[1] + 0.0000001))[0].permute(1, 2, 0)
    ref16[0, ..., 3] = ((disp - st_d[0]) / (st_d[1] + 0.0000001))[0, 0]
    assert torch.equal(x16, ref16)
    xd = U.pack_norm16(disp, st_d)
    assert torch.equal(xd[..., 0], ref16[..., 3]) and not xd[..., 1:].any()
    # ---- payload
    ctx = torch.randn((1, H, W, 64), device='cuda', generator=g).half()
    payload = torch.empty((H * W, 72), device='cuda', dtype=torch.float16)
    check(lib().csb_inpaint_payload(ptr(x16), ptr(ctx), C.c_longlong(H * W), ptr(payload), stream()), "csb_inpaint_payload")
    assert torch.equal(payload[:, :4], x16[0, ..., :4].reshape(-1, 4)) and torch.equal(payload[:, 4:68], ctx.view(-1, 64)) and not payload[:, 68:].any()
    # ---- valid-masked cloud of the raw frame
    pts = torch.empty((1, 3, H * W), device='cuda')
    check(lib().csb_inpaint_points(ptr(disp), H, W, C.c_double(focal), C.c_double(baseline), ptr(st_d), ptr(pts), stream()), "csb_inpaint_points")
    tenDepth = (focal * baseline) / (disp + 0.0000001)
    tenValid = (U.spatial_filter(disp / disp.max(), 'laplacian').abs() < 0.03).float()
    ref_pts = U.depth_to_points(tenDepth * tenValid, focal).view(1, 3, -1)
    assert 0.05 < tenValid.mean().item() < 1.0 and torch.equal(pts, ref_pts)
    # ---- output: de-normalise, clip / threshold, NHWC -> NCHW
    a, b = torch.randn((1, H, W, 3), device='cuda', generator=g), torch.randn((1, H, W, 3), device='cuda', generator=g)
    out = U.net_output(a, b, st_i, 1)
    assert torch.equal(out, ((a + b).permute(0, 3, 1, 2) * (st_i[1] + 0.0000001) + st_i[0]).clip(0.0, 1.0))
    d1 = torch.randn((1, H, W, 1), device='cuda', generator=g)
    out = U.net_output(d1, None, st_d, 2)
    assert torch.equal(out, torch.nn.functional.threshold(d1[..., 0][:, None] * (st_d[1] + 0.0000001) + st_d[0], 0.0, 0.0))
