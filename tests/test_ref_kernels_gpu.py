"""GPU parity against the UNMODIFIED reference kernels (anime_3dkenburns/models/utils.py:63-313, common.py:149-245),
compiled from /root/reference by oracle/build_ref_kernels.py into oracle/_ref/libref_kernels.so and executed on the B200.
This is what pins the oracle and the product kernels to the reference itself."""
import numpy as np
import pytest
import torch

from oracle import kb_oracle as orc
from tests import ref_kernels as ref
from tests.kb_scene import BASELINE, FOCAL, make_scene

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_kernels.so not built")]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _scene(H, W, C, extra, seed=0):
    s = make_scene(H, W, seed=seed, extra_points=extra)
    c = s['common']
    st = {'tenPoints': s['points'], 'fltShiftU': 10.0, 'fltShiftV': -6.0, 'fltDepthFrom': c['objDepthrange'][0], 'fltDepthTo': c['objDepthrange'][0] * 0.9}
    pts, _ = orc.process_shift(st, c)
    return s, pts, s['data'][:, :C]


@pytest.mark.parametrize("H,W,C,extra", [(64, 96, 3, 0), (256, 256, 4, 0), (256, 256, 4, 5000)])
def test_oracle_and_product_vs_reference_render(built_lib, H, W, C, extra):
    from cartoonsegmentation_b200.anime_3dkenburns.models import utils
    s, pts, data = _scene(H, W, C, extra)
    r_ref, e_ref, z0_ref, z1_ref = ref.render_pointcloud(cu(pts), cu(data), W, H, stages=True)
    r_o, e_o, z0_o, z1_o = orc.render_pointcloud(pts, data, W, H, FOCAL, BASELINE, return_zee=True)
    z0_g, zkey = utils.render_zpass(cu(pts), W, H, FOCAL, BASELINE)
    # z-pass: bit exact, three ways
    assert np.array_equal(z0_ref.cpu().numpy(), z0_o)
    assert torch.equal(z0_ref, z0_g)
    # degrid: reference is in place (racy, order dependent); ours is the Jacobi schedule -> identical except where updated
    # neighbours interact (measured on the B200: 0.14% of pixels)
    z1_g = utils.render_degrid(zkey)
    assert float((z1_ref != z1_g).float().mean()) < 5e-3
    r_g, e_g = utils.render_pointcloud(cu(pts), cu(data), W, H, FOCAL, BASELINE)
    same_z = (z1_ref == z1_g)
    frac_bad = float(((r_ref - r_g).abs() > 1e-3).float().mean())
    assert frac_bad < 5e-3
    assert float(((e_ref - e_g).abs() > 1e-4).float().mean()) < 5e-3
    assert abs(int((e_ref > 0).sum()) - int((e_g > 0).sum())) <= max(2, int(2e-4 * H * W))
    # oracle vs reference the same way
    assert float((np.abs(r_ref.cpu().numpy() - r_o) > 1e-3).mean()) < 5e-3


@pytest.mark.parametrize("H,W", [(64, 96), (256, 256)])
def test_oracle_and_product_vs_reference_fill(built_lib, H, W):
    from cartoonsegmentation_b200.anime_3dkenburns import common
    s, pts, data = _scene(H, W, 4, 0, seed=2)
    r_o, e_o = orc.render_pointcloud(pts, s['data'], W, H, FOCAL, BASELINE)
    depth = r_o[:, 3:4] * (e_o > 0.0)
    f_ref = ref.fill_disocclusion(cu(r_o), cu(depth))
    f_g = common.fill_disocclusion(cu(r_o), cu(depth))
    assert torch.equal(f_ref, f_g)                                                       # bit exact incl. tie-breaks
    assert np.array_equal(f_ref.cpu().numpy(), orc.fill_disocclusion(r_o, depth))


def test_full_size_vs_reference(built_lib):
    """1024x1024 (BASELINE size): product kernels against the reference kernels on the GPU."""
    from cartoonsegmentation_b200.anime_3dkenburns import common
    from cartoonsegmentation_b200.anime_3dkenburns.models import utils
    H = W = 1024
    s = make_scene(H, W, seed=1)
    c = s['common']
    st = {'tenPoints': cu(s['points']), 'fltShiftU': 40.0, 'fltShiftV': -25.0, 'fltDepthFrom': c['objDepthrange'][0], 'fltDepthTo': c['objDepthrange'][0] * 0.8}
    pts, _ = common.process_shift(st, c)
    data = cu(s['data'])
    r_ref, e_ref, z0_ref, z1_ref = ref.render_pointcloud(pts, data, W, H, stages=True)
    z0_g, zkey = utils.render_zpass(pts, W, H, FOCAL, BASELINE)
    assert torch.equal(z0_ref, z0_g)
    r_g, e_g = utils.render_pointcloud(pts, data, W, H, FOCAL, BASELINE)
    assert float(((r_ref - r_g).abs() > 1e-3).float().mean()) < 5e-3
    d_ref = r_ref[:, 3:4] * (e_ref > 0).float()
    assert torch.equal(ref.fill_disocclusion(r_ref, d_ref), common.fill_disocclusion(r_ref, d_ref))
    r3_ref, e3_ref = ref.render_pointcloud(pts, data[:, :3].contiguous(), W, H)
    cnt = common.autozoom_coverage(pts, [np.zeros(3, np.float32)], W, H, FOCAL, BASELINE)
    assert abs(int(cnt[0]) - int((e3_ref > 0).sum())) <= 200                               # degrid race only
