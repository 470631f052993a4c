"""GPU numerics of the halo-tile conv kernel (csrc/tc_halo.cu: one activation halo per 64-channel chunk, taps through shifted tcgen05 descriptors;
grouped and depthwise convolutions as block-diagonal MMAs) against plain PyTorch fp32 on fp16-rounded operands.
Tolerance: |err| <= 2e-3 * max|ref| + 1e-3 (fp16 output rounding), as tests/test_conv_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built_lib):
    from cartoonsegmentation_b200 import engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return engine


def _act(y, act, slope=None):
    if act == 'relu': return F.relu(y)
    if act == 'silu': return F.silu(y)
    if act == 'prelu': return F.prelu(y, slope)
    return y


DENSE = [
    # N, H, W, Cin, Cout, k, pad, dil, act
    (1, 16, 8, 64, 64, 3, 1, 1, None),            # exactly one tile
    (2, 45, 45, 64, 64, 3, 1, 1, 'relu'),         # ISNet RSU (resident weights), ragged tiles in both directions
    (1, 40, 24, 128, 64, 3, 1, 1, 'relu'),        # two 64-channel chunks, streamed weights
    (1, 23, 23, 64, 32, 3, 2, 2, 'relu'),         # dilation 2
    (1, 23, 23, 128, 16, 3, 4, 4, 'relu'),        # dilation 4, 16 output channels
    (1, 36, 20, 64, 1, 3, 1, 1, None),            # single output channel (side / depth head), fp32 output
    (1, 32, 32, 64, 96, 3, 1, 1, 'prelu'),        # PReLU slopes, Cout not a multiple of 32
    (1, 20, 20, 64, 48, 5, 2, 1, 'silu'),         # 5x5
    (2, 18, 26, 64, 16, 7, 3, 1, None),           # 7x7: the widest shift (6 pixels)
    (1, 30, 30, 256, 128, 3, 1, 1, 'relu'),       # four chunks, N = 128
    (6, 120, 120, 128, 64, 3, 1, 1, 'relu'),      # ~9 tiles per CTA, two chunks, resident weights squeezed to 2 halo stages (one issuer)
    (6, 120, 120, 64, 64, 3, 1, 1, 'relu'),       # ~9 tiles per CTA, three issuing warps
    (4, 100, 100, 256, 64, 3, 1, 1, 'relu'),      # ~5 tiles per CTA, streamed weights through the ring (36 tiles of 8 KiB per m-tile)
    (4, 90, 90, 64, 32, 3, 2, 2, 'relu'),         # dilation 2: 40 KiB halos, two issuers
]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,pad,dil,act", DENSE)
def test_halo_dense_vs_torch(eng, mode, N, H, W, Cin, Cout, k, pad, dil, act):
    """mode 1 = descriptor base offset 0 (the product path), mode 2 = base offset (start >> 7) & 7: for a start address that is not aligned to the
    1 KiB swizzle pattern the B200 accepts exactly one of the two (measured: mode 1, i.e. the swizzle XOR uses absolute shared-memory address bits);
    the mode-2 run is reported, not asserted."""
    from cartoonsegmentation_b200._lib import lib
    g = torch.Generator(device='cuda').manual_seed(H * 131 + Cin + Cout + k)
    x = torch.randn(N, H, W, Cin + 64, device='cuda', generator=g).half()
    w = torch.randn(Cout, Cin, k, k, device='cuda', generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device='cuda', generator=g) * 0.1
    slope = torch.rand(Cout, device='cuda', generator=g) * 0.5 if act == 'prelu' else None
    f32 = Cout == 1
    prev = lib().csb_conv_set_halo_mode(mode)
    try:
        y = eng.conv2d_halo_nhwc(x, eng.pack_conv_weight(w), b, pad=pad, dil=dil, act=act, act_param=slope, in_coff=64, out_f32=f32)
    finally:
        lib().csb_conv_set_halo_mode(prev)
    r = _act(F.conv2d(x[..., 64:].float().permute(0, 3, 1, 2), w.half().float(), b, padding=pad, dilation=dil), act, slope).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    err = (y.float() - r).abs().max().item()
    tol = 2e-3 * r.abs().max().item() + 1e-3
    ok = y.shape == r.shape and err <= tol
    if mode == 2:
        pytest.skip(f"diagnostic mode (non-zero base offset): {'matches' if ok else 'differs'} (max err {err:.4g})")
    assert ok, f"max err {err} > {tol}"


def test_halo_residual_and_slices(eng):
    g = torch.Generator(device='cuda').manual_seed(11)
    x = torch.randn(2, 24, 24, 64, device='cuda', generator=g).half()
    w = torch.randn(64, 64, 3, 3, device='cuda', generator=g) / (64 * 9) ** 0.5
    b = torch.randn(64, device='cuda', generator=g) * 0.1
    res = torch.randn(2, 24, 24, 64, device='cuda', generator=g).half()
    out = torch.zeros(2, 24, 24, 128, device='cuda', dtype=torch.float16)
    eng.conv2d_halo_nhwc(x, eng.pack_conv_weight(w), b, pad=1, act='relu', residual=res, res_mode=2, out=out, out_coff=64)
    r = (F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, padding=1)) + res.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert (out[..., 64:].float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3
    assert out[..., :64].abs().max().item() == 0
    y = eng.conv2d_halo_nhwc(x, eng.pack_conv_weight(w), b, pad=1, act='relu', residual=res, res_mode=1)
    r = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, padding=1) + res.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert (y.float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3


def test_conv2d_nhwc_routes_thin_layers_to_the_halo_kernel(eng):
    """csb_conv2d_nhwc sends eligible dense shapes (stride 1, Cout <= 128) to the halo kernel: same results as the per-tap kernel."""
    from cartoonsegmentation_b200._lib import lib
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(2, 40, 36, 128, device='cuda', generator=g).half()
    w = eng.pack_conv_weight(torch.randn(64, 128, 3, 3, device='cuda', generator=g) / (128 * 9) ** 0.5)
    b = torch.randn(64, device='cuda', generator=g) * 0.1
    outs = {}
    prev = lib().csb_conv_set_halo_mode(1)
    try:
        for mode in (0, 1):
            lib().csb_conv_set_halo_mode(mode)
            l0 = lib().csb_launch_count()
            outs[mode] = eng.conv2d_nhwc(x, w, b, pad=1, act='relu').clone()
            assert lib().csb_launch_count() == l0 + 1
    finally:
        lib().csb_conv_set_halo_mode(prev)
    d = (outs[0].float() - outs[1].float()).abs().max().item()
    assert d <= 2e-3 * outs[0].float().abs().max().item() + 1e-3, d


@pytest.mark.parametrize("N,H,W,Cc,groups,k,dil", [(1, 40, 40, 1024, 32, 3, 1), (2, 20, 28, 512, 32, 3, 1), (1, 33, 17, 256, 32, 3, 1), (1, 16, 16, 2048, 32, 3, 1),
                                                   (1, 24, 24, 128, 2, 3, 2)])
def test_halo_grouped_vs_torch(eng, N, H, W, Cc, groups, k, dil):
    """ResNeXt-101 32x8d grouped 3x3 (8 / 16 / 32 / 64 channels per group) as diagonal-block MMAs."""
    g = torch.Generator(device='cuda').manual_seed(Cc + H)
    cpg = Cc // groups
    x = torch.randn(N, H, W, Cc, device='cuda', generator=g).half()
    w = torch.randn(Cc, cpg, k, k, device='cuda', generator=g) / (cpg * k * k) ** 0.5
    b = torch.randn(Cc, device='cuda', generator=g) * 0.1
    y = eng.conv2d_halo_nhwc(x, eng.pack_grouped_weight_compact(w, groups), b, pad=dil * (k // 2), dil=dil, act='relu', groups=groups)
    r = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, padding=dil * (k // 2), dilation=dil, groups=groups)).permute(0, 2, 3, 1)
    assert (y.float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3


@pytest.mark.parametrize("N,H,W,Cc,k,act", [(2, 18, 23, 128, 7, None), (1, 32, 32, 256, 7, None), (1, 9, 40, 512, 7, None), (1, 40, 40, 128, 5, 'silu')])
def test_halo_depthwise_and_stats_vs_torch(eng, N, H, W, Cc, k, act):
    """depthwise KxK as groups == C on the tensor cores (16 x 16 diagonal blocks, fp16 weights) + the LayerNorm partial statistics."""
    g = torch.Generator(device='cuda').manual_seed(Cc + H + k)
    x = torch.randn(N, H, W, Cc, device='cuda', generator=g).half()
    dw = torch.randn(k, k, Cc, device='cuda', generator=g) / k
    b = torch.randn(Cc, device='cuda', generator=g) * 0.1
    y, st = eng.conv2d_halo_nhwc(x, eng.pack_dw_weight_compact(dw), b, pad=k // 2, act=act, groups=Cc, stats=True)
    r = _act(F.conv2d(x.float().permute(0, 3, 1, 2), dw.half().float().permute(2, 0, 1)[:, None], b, padding=k // 2, groups=Cc), act).permute(0, 2, 3, 1)
    assert (y.float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3
    yf = y.float().reshape(-1, Cc // 64, 64)
    assert torch.allclose(st[..., 0], yf.sum(2), rtol=1e-4, atol=1e-2) and torch.allclose(st[..., 1], (yf * yf).sum(2), rtol=1e-4, atol=1e-2)
