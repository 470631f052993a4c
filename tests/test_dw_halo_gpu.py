"""The TMA halo-tile depthwise kernel (csrc/dw_halo.cu) against round 1's register-tiled kernel (bit-exact: same arithmetic order) and against
torch's depthwise conv2d (fp32), on full / ragged / tiny images, channel-slice inputs and outputs, K = 7 with LayerNorm statistics and K = 5 + SiLU."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _both(fn):
    os.environ['CSB_DW_HALO'] = '1'
    a = fn()
    os.environ['CSB_DW_HALO'] = '0'
    b = fn()
    os.environ.pop('CSB_DW_HALO')
    return a, b


@pytest.mark.parametrize("N,H,W,C", [(2, 64, 64, 512), (1, 33, 47, 128), (2, 16, 16, 1024), (3, 5, 70, 64), (1, 128, 96, 256)])
def test_dw7_stats_halo_equals_tile_and_torch(built_lib, N, H, W, C):
    from cartoonsegmentation_b200 import engine as E
    g = torch.Generator(device='cuda').manual_seed(N * H + C)
    x = torch.randn(N, H, W, C, device='cuda', generator=g).half()
    w = torch.randn(7, 7, C, device='cuda', generator=g) / 7
    b = torch.randn(C, device='cuda', generator=g) * 0.1
    (y1, s1), (y0, s0) = _both(lambda: E.dwconv_stats_nhwc(x, w, b))
    torch.cuda.synchronize()
    assert torch.equal(y1, y0) and torch.allclose(s1, s0, rtol=2e-6, atol=1e-5)      # outputs bit-exact; the statistics are summed in another order
    r = F.conv2d(x.float().permute(0, 3, 1, 2), w.permute(2, 0, 1)[:, None], b, padding=3, groups=C).permute(0, 2, 3, 1)
    assert (y1.float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3
    yf = y1.float().view(N * H * W, C // 64, 64)
    assert torch.allclose(s1[..., 0], yf.sum(-1), rtol=1e-4, atol=1e-3) and torch.allclose(s1[..., 1], (yf * yf).sum(-1), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("K,act", [(5, 'silu'), (7, None), (5, None)])
def test_dw_slices_halo_equals_tile(built_lib, K, act):
    """x is channels [64, 192) of a 256-channel tensor, y goes to channels [128, 256) of a 320-channel tensor (concat-free CSP blocks)."""
    from cartoonsegmentation_b200 import engine as E
    g = torch.Generator(device='cuda').manual_seed(K)
    N, H, W, C = 2, 40, 52, 128
    xb = torch.randn(N, H, W, 256, device='cuda', generator=g).half()
    w = torch.randn(K, K, C, device='cuda', generator=g) / K
    b = torch.randn(C, device='cuda', generator=g) * 0.1

    def run():
        out = torch.zeros(N, H, W, 320, device='cuda', dtype=torch.float16)
        E.dwconv_nhwc(xb, w, b, act=act, out=out, xoff=64, yoff=128, channels=C)
        return out
    a, c = _both(run)
    torch.cuda.synchronize()
    assert torch.equal(a, c) and a[..., :128].abs().max().item() == 0 and a[..., 256:].abs().max().item() == 0
    r = F.conv2d(xb[..., 64:192].float().permute(0, 3, 1, 2), w.permute(2, 0, 1)[:, None], b, padding=K // 2, groups=C)
    r = (F.silu(r) if act == 'silu' else r).permute(0, 2, 3, 1)
    assert (a[..., 128:256].float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3
