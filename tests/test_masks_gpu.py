"""AnimeInstances.resize / compose_masks on the device against the torch ops the reference runs
(animeinsseg/anime_instances.py:268-298), bit-exact; and the `depth[depth == 0] = depth[depth > 0].min()` kernel (kenburns_effect.py:577)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ellipse_masks(K, H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((K, H, W), bool)
    for k in range(K):
        cy, cx, ry, rx = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, H / 2), rng.uniform(3, W / 2)
        m[k] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1
    m[K - 1] = False                                                # an empty mask
    return m


@pytest.mark.parametrize("H0,W0,H,W", [(1024, 1024, 768, 768), (300, 420, 200, 280), (200, 280, 300, 420), (97, 131, 64, 64), (64, 64, 64, 64), (50, 70, 125, 175)])
def test_resize_equals_reference_torch_ops(built_lib, H0, W0, H, W):
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    K = 7
    m = torch.from_numpy(_ellipse_masks(K, H0, W0, H0 + W)).cuda()
    boxes = torch.randint(0, 900, (K, 4), dtype=torch.int32, device='cuda')
    inst = AnimeInstances(m.clone(), boxes.clone(), torch.rand(K, device='cuda'))
    inst.resize(H, W)
    # the reference's body, verbatim semantics (anime_instances.py:268-280)
    masks = m.to(torch.float).unsqueeze(1)
    hs, ws = H / H0, W / W0
    b = boxes.float()
    b[:, ::2] *= hs
    b[:, 1::2] *= ws
    ref_b = torch.round(b).int()
    ref_m = F.interpolate(masks, (H, W), mode='area').squeeze(1) > 0.3
    assert inst.masks.dtype == torch.bool and torch.equal(inst.masks, ref_m)
    assert inst.bboxes.dtype == torch.int32 and torch.equal(inst.bboxes, ref_b)


@pytest.mark.parametrize("K,H,W", [(1, 64, 64), (9, 100, 112), (100, 256, 256), (3, 33, 35)])
def test_compose_equals_any(built_lib, K, H, W):
    from cartoonsegmentation_b200.animeinsseg import AnimeInstances
    m = torch.from_numpy(_ellipse_masks(K, H, W, K)).cuda()
    inst = AnimeInstances(m, torch.zeros((K, 4), dtype=torch.int32, device='cuda'), torch.ones(K, device='cuda'))
    c = inst.compose_masks()
    assert c.dtype == torch.bool and torch.equal(c, m.any(0))
    assert np.array_equal(inst.compose_masks(output_type='numpy'), m.any(0).cpu().numpy())


def test_zero_to_min_positive(built_lib):
    from cartoonsegmentation_b200._lib import check, lib, ptr, stream
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randint(0, 256, (3, 97, 131), device='cuda', generator=g).float()
    x[1] = torch.clamp(x[1], min=17)                                # no zeros in image 1
    x[2, 5:9] = 0
    ref = x.clone()
    for i in range(3):
        ref[i][ref[i] == 0] = ref[i][ref[i] > 0].min()
    scr = torch.empty(3, device='cuda', dtype=torch.int32)
    check(lib().csb_zero_to_min_positive(ptr(x), 3, C.c_longlong(97 * 131), ptr(scr), stream()), "csb_zero_to_min_positive")
    assert torch.equal(x, ref)
