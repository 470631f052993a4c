"""GPU numerics of the tcgen05 implicit-GEMM conv engine against a plain PyTorch fp32 reference of the same op
(inputs/weights rounded to fp16 first, so the only differences are fp32 accumulation order and the fp16 output rounding).
Tolerance: |err| <= 2e-3 * max|ref| + 1e-3 (fp16 output has 2^-11 relative rounding)."""
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


def ref_conv(x_nhwc, w_oihw, bias, stride, pad, dil, act, residual=None, res_mode=0, slope=None):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w_oihw.half().float(), bias, stride=stride, padding=pad, dilation=dil)
    if residual is not None and res_mode == 1:
        y = y + residual.float().permute(0, 3, 1, 2)
    if act == 'relu': y = F.relu(y)
    elif act == 'silu': y = F.silu(y)
    elif act == 'gelu': y = F.gelu(y)
    elif act == 'prelu': y = F.prelu(y, slope)
    elif act == 'sigmoid': y = torch.sigmoid(y)
    if residual is not None and res_mode == 2:
        y = y + residual.float().permute(0, 3, 1, 2)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, dil, act
    (1, 16, 16, 64, 64, 1, 1, 0, 1, None),          # smallest GEMM: one tile, one k-block
    (2, 24, 40, 128, 256, 1, 1, 0, 1, 'silu'),      # 1x1, flat mode, N tile 256
    (1, 32, 32, 64, 128, 3, 1, 1, 1, 'relu'),       # 3x3 s1: TMA zero padding on all borders
    (2, 20, 28, 128, 96, 3, 1, 1, 1, 'silu'),       # ragged spatial tiles, Cout not a multiple of 32
    (1, 33, 47, 64, 64, 3, 2, 1, 1, 'relu'),        # stride 2 through elementStrides, odd sizes
    (1, 32, 32, 256, 512, 3, 1, 1, 1, None),        # two N tiles, 36 k-blocks
    (1, 24, 24, 64, 64, 3, 1, 2, 2, 'relu'),        # dilation 2 (ISNet RSU)
    (1, 16, 16, 512, 2048, 1, 1, 0, 1, 'gelu'),     # ConvNeXt MLP up-projection
    (1, 40, 40, 32, 32, 3, 1, 1, 1, 'prelu'),       # 32-channel layers (Inpaint net): 64 B swizzle
    (1, 40, 40, 16, 48, 3, 1, 1, 1, None),          # 16-channel k-block: 32 B swizzle
    (1, 64, 64, 256, 169, 1, 1, 0, 1, None),        # rtm_kernel head: 169 outputs, scalar store path
    (1, 128, 128, 256, 256, 3, 1, 1, 1, 'silu'),    # the hottest detector shape (Appendix B)
    (3, 8, 8, 64, 16, 7, 1, 3, 1, None),            # 7x7, tiny maps, batch 3
    (1, 64, 64, 64, 64, 4, 4, 0, 1, None),          # 4x4 stride 4 (patchify)
    (1, 32, 32, 128, 256, 2, 2, 0, 1, None),        # 2x2 stride 2 (ConvNeXt downsample)
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad,dil,act", CASES)
def test_conv_vs_torch(eng, N, H, W, Cin, Cout, k, stride, pad, dil, act):
    g = torch.Generator(device='cuda').manual_seed(H * 1000 + Cin + Cout)
    x = torch.randn(N, H, W, Cin, device='cuda', generator=g).half()
    w = (torch.randn(Cout, Cin, k, k, device='cuda', generator=g) / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, device='cuda', generator=g) * 0.1
    slope = torch.rand(Cout, device='cuda', generator=g) * 0.5 if act == 'prelu' else None
    y = eng.conv2d_nhwc(x, eng.pack_conv_weight(w), b, stride=stride, pad=pad, dil=dil, act=act, act_param=slope)
    r = ref_conv(x, w, b, stride, pad, dil, act, slope=slope)
    torch.cuda.synchronize()
    err = (y.float() - r).abs().max().item()
    tol = 2e-3 * r.abs().max().item() + 1e-3
    assert y.shape == r.shape
    assert err <= tol, f"max err {err} > {tol}"


def test_conv_residual_slices_and_f32(eng):
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(2, 24, 24, 192, device='cuda', generator=g).half()
    w = torch.randn(64, 128, 3, 3, device='cuda', generator=g) / (128 * 9) ** 0.5
    b = torch.randn(64, device='cuda', generator=g) * 0.1
    res = torch.randn(2, 24, 24, 64, device='cuda', generator=g).half()
    wp = eng.pack_conv_weight(w)
    # input = channel slice [64:192) of x, output written into channels [32:96) of a 128-channel tensor (concat fusion)
    out = torch.zeros(2, 24, 24, 128, device='cuda', dtype=torch.float16)
    eng.conv2d_nhwc(x, wp, b, pad=1, act='silu', residual=res, res_mode=2, out=out, out_coff=32, in_coff=64)
    r = ref_conv(x[..., 64:], w, b, 1, 1, 1, 'silu', residual=res, res_mode=2)
    assert (out[..., 32:96].float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3
    assert out[..., :32].abs().max().item() == 0 and out[..., 96:].abs().max().item() == 0
    # residual before activation (ResNeXt bottleneck), fp32 output
    y = eng.conv2d_nhwc(x, wp, b, pad=1, act='relu', residual=res, res_mode=1, in_coff=64, out_f32=True)
    r = ref_conv(x[..., 64:], w, b, 1, 1, 1, 'relu', residual=res, res_mode=1)
    assert y.dtype == torch.float32 and (y - r).abs().max().item() <= 5e-4 * r.abs().max().item() + 1e-4
    # bf16
    xb, wb = x.bfloat16(), eng.pack_conv_weight(w, torch.bfloat16)
    yb = eng.conv2d_nhwc(xb, wb, b, pad=1, in_coff=64, out_f32=True)
    rb = F.conv2d(xb[..., 64:].float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=1).permute(0, 2, 3, 1)
    assert (yb - rb).abs().max().item() <= 5e-4 * rb.abs().max().item() + 1e-4


def test_conv_rejects_bad_arguments(eng):
    x = torch.zeros(1, 8, 8, 24, device='cuda', dtype=torch.float16)
    w = torch.zeros(16, 3, 3, 24, device='cuda', dtype=torch.float16)
    with pytest.raises(Exception):
        eng.conv2d_nhwc(x, w, None, pad=1)           # Cin = 24 is not a multiple of 16


@pytest.mark.parametrize("N,H,W,C,mean_shift", [(2, 18, 23, 128, 0.0), (1, 16, 16, 256, 3.0), (2, 9, 40, 512, -1.5), (1, 7, 7, 1024, 0.5)])
def test_convnext_block_head_ln_folded(eng, N, H, W, C, mean_shift):
    """depthwise 7x7 -> LayerNorm -> Linear(C, 4C) -> GELU (mmpretrain ConvNeXtBlock) with the LayerNorm folded into the GEMM
    epilogue (csb_dwconv_stats_nhwc + csb_conv2d_ln_nhwc) against plain PyTorch fp32 and against the unfused kernels.  `mean_shift` moves
    the per-pixel mean away from zero (the folded form subtracts mean * colsum in fp32)."""
    g = torch.Generator(device='cuda').manual_seed(C + H)
    x = (torch.randn(N, H, W, C, device='cuda', generator=g) + mean_shift).half()
    dw = torch.randn(7, 7, C, device='cuda', generator=g) / 7.0
    dwb = torch.randn(C, device='cuda', generator=g) * 0.1 + mean_shift
    gamma = torch.rand(C, device='cuda', generator=g) + 0.5
    beta = torch.randn(C, device='cuda', generator=g) * 0.2
    w1 = torch.randn(4 * C, C, device='cuda', generator=g) / C ** 0.5
    b1 = torch.randn(4 * C, device='cuda', generator=g) * 0.1
    # fp32 reference on the fp16-rounded depthwise output (what both GPU paths normalise)
    u_ref = F.conv2d(x.float().permute(0, 3, 1, 2), dw.permute(2, 0, 1)[:, None], dwb, padding=3, groups=C).permute(0, 2, 3, 1)
    u16 = u_ref.half().float()
    ref = F.gelu(F.linear(F.layer_norm(u16, (C,), gamma, beta, 1e-6), w1, b1))
    u, stats = eng.dwconv_stats_nhwc(x, dw.contiguous(), dwb)
    assert (u.float() - u_ref).abs().max() <= 2e-3 * u_ref.abs().max() + 1e-3
    s = stats.sum(1)
    uf = u.float().reshape(-1, C)
    assert torch.allclose(s[:, 0], uf.sum(1), rtol=1e-4, atol=1e-2) and torch.allclose(s[:, 1], (uf * uf).sum(1), rtol=1e-4, atol=1e-2)
    wl, bl, cs = eng.fold_layernorm(w1, b1, gamma, beta)
    y = eng.conv2d_ln_nhwc(u, stats, wl, bl, cs, eps=1e-6, act='gelu').float()
    tol = 3e-3 * ref.abs().max() + 2e-3
    assert (y - ref).abs().max() <= tol, ((y - ref).abs().max().item(), tol.item())
    # the unfused kernels (dwconv + k_layernorm, then the plain GEMM) agree to the same tolerance
    u2 = eng.dwconv_nhwc(x, dw.contiguous(), dwb, ln=(gamma.contiguous(), beta.contiguous()), eps=1e-6)
    y2 = eng.conv2d_nhwc(u2, eng.pack_conv_weight(w1[:, :, None, None]), b1, act='gelu').float()
    assert (y2 - ref).abs().max() <= tol
    assert ((y - ref) ** 2).mean().sqrt() <= 1.5 * ((y2 - ref) ** 2).mean().sqrt() + 1e-4          # folding does not cost accuracy


PAIR_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, act: shapes that take the CTA-pair path (tcgen05 cta_group::2) in mode 2
    (2, 24, 40, 128, 256, 1, 1, 0, 'silu'),       # flat, 15 m-tiles: the last pair has a peer tile past the end
    (1, 64, 64, 512, 2048, 1, 1, 0, 'gelu'),      # ConvNeXt fc1: 8 n-tiles, 32 m-tiles
    (1, 32, 32, 2048, 512, 1, 1, 0, None),        # ConvNeXt fc2: 32 k-blocks
    (2, 36, 28, 256, 256, 3, 1, 1, 'relu'),       # 3x3 with ragged spatial tiles, 2 images
    (1, 33, 47, 64, 64, 3, 2, 1, 'relu'),         # stride 2, block_n 64 (pairs only in mode 2)
    (3, 20, 20, 128, 96, 3, 1, 1, 'silu'),        # block_n 96: a 48-row weight half per CTA
    (1, 16, 16, 1024, 4096, 1, 1, 0, 'gelu'),     # more n-tiles than m-tiles
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad,act", PAIR_CASES)
def test_conv_cta_pair_matches_single_cta(eng, N, H, W, Cin, Cout, k, stride, pad, act):
    """tcgen05 cta_group::2 (two SMs share one weight tile) against the single-CTA kernel: the k-order of the accumulation is the same, so
    the results are bit-identical; both are also checked against PyTorch fp32."""
    from cartoonsegmentation_b200._lib import lib
    g = torch.Generator(device='cuda').manual_seed(H * 77 + Cin + Cout)
    x = torch.randn(N, H, W, Cin, device='cuda', generator=g).half()
    w = (torch.randn(Cout, Cin, k, k, device='cuda', generator=g) / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, device='cuda', generator=g) * 0.1
    res = torch.randn(N, (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1, Cout, device='cuda', generator=g).half()
    wp = eng.pack_conv_weight(w)
    outs = {}
    prev = lib().csb_conv_set_pair_mode(0)
    try:
        for mode in (0, 2):
            lib().csb_conv_set_pair_mode(mode)
            outs[mode] = eng.conv2d_nhwc(x, wp, b, stride=stride, pad=pad, act=act, residual=res, res_mode=2).clone()
    finally:
        lib().csb_conv_set_pair_mode(prev)
    torch.cuda.synchronize()
    r = ref_conv(x, w, b, stride, pad, 1, act, residual=res, res_mode=2)
    assert (outs[0].float() - r).abs().max().item() <= 2e-3 * r.abs().max().item() + 1e-3
    assert torch.equal(outs[0], outs[2]), f"pair path differs: max {(outs[0].float() - outs[2].float()).abs().max().item()}"


@pytest.mark.parametrize("N,H,W,C,dt", [(2, 40, 48, 128, 'f16'), (1, 33, 47, 128, 'f16'), (3, 24, 40, 256, 'f16'), (1, 67, 129, 256, 'f16'), (8, 64, 64, 128, 'f16'),
                                        (1, 33, 47, 128, 'bf16'), (2, 24, 40, 256, 'bf16')])
def test_fused_convnext_mlp_is_bit_identical_to_the_two_launch_path(eng, N, H, W, C, dt):
    """k_mlp_tc (LN -> fc1 -> GELU -> fc2 -> + residual, hidden activations kept on the SM) against csb_conv2d_ln_nhwc + csb_conv2d_nhwc: same k-order
    of both accumulations and the same fp16 rounding of the hidden activations -> bit-identical, incl. ragged last tiles, multi-tile CTAs and the
    in-place (out == residual) form the backbone uses; the two-launch result is also checked against PyTorch fp32."""
    g = torch.Generator(device='cuda').manual_seed(C + H)
    Hd = 4 * C
    tdt = torch.float16 if dt == 'f16' else torch.bfloat16
    x = (torch.randn(N, H, W, C, device='cuda', generator=g) * 1.5 + 0.3).to(tdt)
    t = torch.randn(N, H, W, C, device='cuda', generator=g).to(tdt)
    w1 = torch.randn(Hd, C, device='cuda', generator=g) / C ** 0.5
    b1 = torch.randn(Hd, device='cuda', generator=g) * 0.1
    gamma, beta = torch.rand(C, device='cuda', generator=g) + 0.5, torch.randn(C, device='cuda', generator=g) * 0.1
    w2 = torch.randn(C, Hd, device='cuda', generator=g) / Hd ** 0.5
    b2 = torch.randn(C, device='cuda', generator=g) * 0.1
    wf, bf, colsum = eng.fold_layernorm(w1, b1, gamma, beta, tdt)
    w2p = eng.pack_conv_weight(w2.reshape(C, Hd, 1, 1), tdt)
    stats = torch.stack([torch.stack([x[..., c:c + 64].float().sum(-1), x[..., c:c + 64].float().pow(2).sum(-1)], -1) for c in range(0, C, 64)], -2)
    stats = stats.reshape(-1, C // 64, 2).contiguous()
    h = eng.conv2d_ln_nhwc(x, stats, wf, bf, colsum, act='gelu')
    two = eng.conv2d_nhwc(h, w2p, b2, residual=t, res_mode=2)
    one = eng.convnext_mlp_nhwc(x, stats, wf, bf, colsum, w2p, b2, t)
    torch.cuda.synchronize()
    assert torch.equal(one, two), f"fused MLP differs: max {(one.float() - two.float()).abs().max().item()}, {(one != two).float().mean().item():.4f} of elements"
    t2 = t.clone()
    eng.convnext_mlp_nhwc(x, stats, wf, bf, colsum, w2p, b2, t2, out=t2)                         # in place
    assert torch.equal(t2, two)
    xn = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-6)
    ref = t.float() + torch.nn.functional.gelu(xn @ w1.t() + b1) @ w2.t() + b2
    tol = 4e-3 if dt == 'f16' else 3e-2
    assert (two.float() - ref).abs().max().item() <= tol * ref.abs().max().item() + tol
