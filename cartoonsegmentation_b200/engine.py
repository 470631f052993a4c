"""Host side of the tcgen05 conv/GEMM engine (csrc/tc_conv.cu, include/csb200.h csb_conv2d_nhwc).

Activations are NHWC fp16 (or bf16) CUDA tensors; weights are packed once to [Cout][R][S][Cin] (K-major rows) with
BatchNorm folded in.  Every call enqueues one persistent tcgen05 kernel on torch's current stream.  No CPU fallback.
"""
import ctypes as C

import torch

from ._lib import check, lib, ptr, stream

ACT = {None: 0, 'none': 0, 'relu': 1, 'silu': 2, 'gelu': 3, 'prelu': 4, 'sigmoid': 5, 'softplus': 6, 'hardsigmoid': 7}


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("N", "Hin", "Win", "Cin", "in_ld", "in_coff", "Cout", "R", "S", "stride", "pad", "dil",
                                       "out_ld", "out_coff", "act", "res_mode", "res_ld", "res_coff", "dtype", "groups")]


def pack_conv_weight(w, dtype=torch.float16, cin_pad=None):
    """torch conv weight [Cout, Cin, R, S] -> [Cout, R, S, Cin(_pad)] K-major, contiguous, on the same device."""
    Cout, Cin, R, S = w.shape
    wp = w.permute(0, 2, 3, 1)
    if cin_pad is not None and cin_pad != Cin:
        wp = torch.nn.functional.pad(wp, (0, cin_pad - Cin))
    return wp.contiguous().to(dtype)


def out_hw(Hin, Win, R, S, stride, pad, dil):
    return (Hin + 2 * pad - dil * (R - 1) - 1) // stride + 1, (Win + 2 * pad - dil * (S - 1) - 1) // stride + 1


def pack_grouped_weight(w, groups, dtype=torch.float16):
    """torch grouped conv weight [Cout, Cin/groups, R, S] (Cin == Cout) -> [Cout, R, S, 64]: for each output channel the weights over the
    64-channel input slice that contains its group (zeros outside the group) -- the block-diagonal form csb_conv2d_nhwc(groups>1) expects."""
    Cout, cpg, R, S = w.shape
    assert Cout % 64 == 0 and 64 % cpg == 0 and Cout // groups == cpg
    out = torch.zeros((Cout, R, S, 64), device=w.device, dtype=torch.float32)
    co = torch.arange(Cout, device=w.device)
    base = ((co % 64) // cpg) * cpg                       # first input channel of the group inside the 64-slice
    idx = base[:, None] + torch.arange(cpg, device=w.device)[None]          # [Cout, cpg]
    out.view(Cout, R * S, 64).scatter_(2, idx[:, None, :].expand(Cout, R * S, cpg), w.permute(0, 2, 3, 1).reshape(Cout, R * S, cpg).float())
    return out.contiguous().to(dtype)


def pack_grouped_weight_compact(w, groups, dtype=torch.float16):
    """torch grouped conv weight [Cout, Cin/groups, R, S] (Cin == Cout) -> [Cout, R, S, gk], gk = max(16, Cin/groups): per output channel the
    gk input channels of the gk-aligned block that holds its group (zeros outside the group when a block holds several groups) -- the compact
    block-diagonal form of csb_conv2d_halo_nhwc(groups > 1)."""
    Cout, cpg, R, S = w.shape
    assert Cout // groups == cpg and (cpg >= 16 and 64 % cpg == 0 or 16 % cpg == 0)
    gk = max(16, cpg)
    if cpg >= 16:
        return w.permute(0, 2, 3, 1).contiguous().to(dtype)
    out = torch.zeros((Cout, R * S, gk), device=w.device, dtype=torch.float32)
    co = torch.arange(Cout, device=w.device)
    base = ((co // cpg) * cpg) % gk                                           # first input channel of the group inside its 16-channel block
    idx = base[:, None] + torch.arange(cpg, device=w.device)[None]          # [Cout, cpg]
    out.scatter_(2, idx[:, None, :].expand(Cout, R * S, cpg), w.permute(0, 2, 3, 1).reshape(Cout, R * S, cpg).float())
    return out.view(Cout, R, S, gk).contiguous().to(dtype)


def pack_dw_weight_compact(w, dtype=torch.float16):
    """depthwise weight [K, K, C] (the layout of dwconv_nhwc) -> compact block-diagonal [C, K, K, 16] for csb_conv2d_halo_nhwc(groups = C)."""
    K, _, Cc = w.shape
    return pack_grouped_weight_compact(w.permute(2, 0, 1)[:, None].contiguous(), Cc, dtype)


def halo_supported(N, H, W, Cin, Cx, in_coff, Cout, R, S, stride, pad, dil, groups=1):
    d = ConvDesc(N, H, W, Cin, Cx, in_coff, Cout, R, S, stride, pad, dil, Cout, 0, 0, 0, 0, 0, 0, groups)
    return bool(lib().csb_conv_halo_supported(C.byref(d)))


def conv2d_halo_nhwc(x, w, bias=None, pad=0, dil=1, act=None, act_param=None, residual=None, res_mode=0, out=None, out_coff=0, in_coff=0,
                     res_coff=0, out_f32=False, groups=1, stats=False):
    """Stride-1 RxS conv on the halo-tile kernel (csrc/tc_halo.cu).  groups <= 1: w packed [Cout,R,S,Cin]; groups > 1: w from
    pack_grouped_weight_compact / pack_dw_weight_compact ([Cout,R,S,gk]), Cin == Cout.  stats=True (grouped): also returns the per-pixel,
    per-64-channel (sum, sum of squares) of the rounded outputs for conv2d_ln_nhwc."""
    N, H, W, Cx = x.shape
    Cout, R, S, Kw = w.shape
    Cin = Cout if groups > 1 else Kw
    Ho, Wo = out_hw(H, W, R, S, 1, pad, dil)
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    if residual is not None and res_mode == 0:
        res_mode = 1
    d = ConvDesc(N, H, W, Cin, Cx, in_coff, Cout, R, S, 1, pad, dil, out.shape[3], out_coff, ACT[act], res_mode,
                 residual.shape[3] if residual is not None else 0, res_coff, 1 if x.dtype == torch.bfloat16 else 0, groups)
    st = torch.empty((N * Ho * Wo, Cout // 64, 2), device=x.device, dtype=torch.float32) if stats else None
    is_f32 = out.dtype == torch.float32
    check(lib().csb_conv2d_halo_nhwc(C.byref(d), ptr(x), ptr(w), ptr(bias), ptr(act_param), ptr(residual), None if is_f32 else ptr(out),
                                     ptr(out) if is_f32 else None, ptr(st), stream()), "csb_conv2d_halo_nhwc")
    return (out, st) if stats else out


def conv2d_nhwc(x, w, bias=None, stride=1, pad=0, dil=1, act=None, act_param=None, residual=None, res_mode=0, out=None, out_coff=0,
                in_coff=0, cin=None, res_coff=0, out_f32=False, groups=1):
    """x: [N,H,W,Cx] NHWC fp16/bf16 (channels in_coff..in_coff+cin used); w: packed [Cout,R,S,Cin]; -> y [N,Ho,Wo,Cout] (or `out`
    written at channel offset out_coff).  residual: NHWC tensor added before (res_mode=1) / after (res_mode=2) the activation."""
    N, H, W, Cx = x.shape
    Cout, R, S, Cin = w.shape
    if groups > 1:
        Cin = Cout
    cin = Cin if cin is None else cin
    assert cin == Cin and x.dtype == w.dtype and x.dtype in (torch.float16, torch.bfloat16)
    Ho, Wo = out_hw(H, W, R, S, stride, pad, dil)
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    assert out.shape[:3] == (N, Ho, Wo)
    if residual is not None and res_mode == 0:
        res_mode = 1
    d = ConvDesc(N, H, W, Cin, Cx, in_coff, Cout, R, S, stride, pad, dil, out.shape[3], out_coff, ACT[act], res_mode,
                 residual.shape[3] if residual is not None else 0, res_coff, 1 if x.dtype == torch.bfloat16 else 0, groups)
    is_f32 = out.dtype == torch.float32
    check(lib().csb_conv2d_nhwc(C.byref(d), ptr(x), ptr(w), ptr(bias), ptr(act_param), ptr(residual), None if is_f32 else ptr(out),
                                ptr(out) if is_f32 else None, stream()), "csb_conv2d_nhwc")
    return out


def fold_layernorm(w, b, gamma, beta, dtype=torch.float16):
    """LayerNorm(gamma, beta) followed by Linear(w [Cout,Cin], b) == rstd * (x W'^T - mean * colsum) + b' on the un-normalised x:
    -> (packed W' [Cout,1,1,Cin], b' fp32, colsum fp32 of the ROUNDED W' so that the mean term cancels exactly in the epilogue)."""
    w, b, gamma, beta = w.float(), b.float(), gamma.float(), beta.float()
    wp = (w * gamma[None, :]).to(dtype)
    return wp.reshape(w.shape[0], 1, 1, w.shape[1]).contiguous(), (b + w @ beta).contiguous(), wp.float().sum(1).contiguous()


def dwconv_stats_nhwc(x, w, bias=None, out=None, xoff=0, yoff=0, channels=None):
    """Depthwise 7x7 + bias -> (y, stats [N*H*W, C/64, 2] fp32 partial (sum, sumsq) per pixel and 64-channel chunk) for conv2d_ln_nhwc."""
    N, H, W, ldx = x.shape
    Cc = w.shape[2] if channels is None else channels
    if out is None:
        out = torch.empty((N, H, W, Cc), device=x.device, dtype=x.dtype)
    stats = torch.empty((N * H * W, Cc // 64, 2), device=x.device, dtype=torch.float32)
    check(lib().csb_dwconv_stats_nhwc(ptr(x), ldx, xoff, ptr(w), ptr(bias), N, H, W, Cc, w.shape[0], ptr(out), out.shape[3], yoff, ptr(stats), stream()),
          "csb_dwconv_stats_nhwc")
    return out, stats


def convnext_mlp_supported(C_, hidden):
    return bool(lib().csb_convnext_mlp_supported(int(C_), int(hidden)))


def convnext_mlp_nhwc(x, stats, w1, b1, colsum, w2, b2, residual, eps=1e-6, out=None):
    """residual + W2 GELU(LayerNorm(x) W1^T + b1) + b2 in one launch (csb_convnext_mlp_nhwc); LayerNorm folded as for conv2d_ln_nhwc,
    layer scale folded into w2 / b2.  x, residual: [N,H,W,C]; w1 [4C,1,1,C], w2 [C,1,1,4C] packed; out may be `residual` (in place)."""
    N, H, W, Cx = x.shape
    hidden = w1.shape[0]
    assert w1.shape[3] == Cx and w2.shape[0] == Cx and w2.shape[3] == hidden and x.dtype == w1.dtype == w2.dtype
    if out is None:
        out = torch.empty((N, H, W, Cx), device=x.device, dtype=x.dtype)
    check(lib().csb_convnext_mlp_nhwc(ptr(x), Cx, 0, C.c_longlong(N * H * W), Cx, hidden, ptr(w1), ptr(b1), ptr(colsum), ptr(stats), _cf(eps), ptr(w2), ptr(b2),
                                      ptr(residual), residual.shape[3] if residual is not None else 0, 0, ptr(out), out.shape[3], 0,
                                      1 if x.dtype == torch.bfloat16 else 0, stream()), "csb_convnext_mlp_nhwc")
    return out


def conv2d_ln_nhwc(x, stats, w, bias, colsum, eps=1e-6, act=None, residual=None, res_mode=0, out=None, out_coff=0):
    """1x1 conv of LayerNorm(x) with the LayerNorm folded into the epilogue (see fold_layernorm); x is the un-normalised NHWC tensor."""
    N, H, W, Cx = x.shape
    Cout, R, S, Cin = w.shape
    assert (R, S) == (1, 1) and Cin == Cx and x.dtype == w.dtype
    if out is None:
        out = torch.empty((N, H, W, Cout), device=x.device, dtype=x.dtype)
    if residual is not None and res_mode == 0:
        res_mode = 1
    d = ConvDesc(N, H, W, Cin, Cx, 0, Cout, 1, 1, 1, 0, 1, out.shape[3], out_coff, ACT[act], res_mode,
                 residual.shape[3] if residual is not None else 0, 0, 1 if x.dtype == torch.bfloat16 else 0, 1)
    check(lib().csb_conv2d_ln_nhwc(C.byref(d), ptr(x), ptr(w), ptr(bias), ptr(colsum), ptr(stats), _cf(eps), ptr(residual), ptr(out), stream()),
          "csb_conv2d_ln_nhwc")
    return out


def dwconv_nhwc(x, w, bias=None, ln=None, eps=1e-6, act=None, out=None, xoff=0, yoff=0, channels=None):
    """Depthwise KxK (stride 1, pad K/2) + bias [+ LayerNorm over C (ln=(gamma, beta))] [+ act].  x NHWC fp16, w [K,K,C] fp32."""
    N, H, W, ldx = x.shape
    K = w.shape[0]
    Cc = w.shape[2] if channels is None else channels
    if out is None:
        out = torch.empty((N, H, W, Cc), device=x.device, dtype=x.dtype)
    check(lib().csb_dwconv_nhwc(ptr(x), ldx, xoff, ptr(w), ptr(bias), ptr(ln[0]) if ln else None, ptr(ln[1]) if ln else None, _cf(eps), ACT[act], N, H, W, Cc, K, ptr(out), out.shape[3], yoff, stream()), "csb_dwconv_nhwc")
    return out


def _cf(v):
    return C.c_float(float(v))


def layernorm_nhwc(x, gamma, beta, eps=1e-6, out=None, xoff=0, yoff=0, channels=None):
    ldx = x.shape[-1]
    C_ = gamma.numel() if channels is None else channels
    npix = x.numel() // ldx
    if out is None:
        out = torch.empty(x.shape[:-1] + (C_,), device=x.device, dtype=x.dtype)
    check(lib().csb_layernorm_nhwc(ptr(x), ldx, xoff, ptr(gamma), ptr(beta), _cf(eps), C.c_longlong(npix), C_, ptr(out), out.shape[-1], yoff, stream()),
          "csb_layernorm_nhwc")
    return out


def resample_nhwc(x, Ho, Wo, mode, out=None, xoff=0, yoff=0, channels=None):
    """mode: 'nearest' | 'bilinear' (align_corners=False) | 'bilinear_ac' (align_corners=True)"""
    N, Hi, Wi, ldx = x.shape
    C_ = ldx if channels is None else channels
    if out is None:
        out = torch.empty((N, Ho, Wo, C_), device=x.device, dtype=x.dtype)
    m = {'nearest': 0, 'bilinear': 1, 'bilinear_ac': 2}[mode]
    check(lib().csb_resample_nhwc(ptr(x), ldx, xoff, N, Hi, Wi, C_, Ho, Wo, m, ptr(out), out.shape[3], yoff, stream()), "csb_resample_nhwc")
    return out


def image_prep_nhwc(img_u8, mean, std, swap_rb=False, CP=16):
    """uint8 [N,H,W,3] (or [H,W,3]) -> fp16 NHWC [N,H,W,CP], (x-mean)/std, zero padded channels."""
    if img_u8.dim() == 3:
        img_u8 = img_u8[None]
    N, H, W, _ = img_u8.shape
    out = torch.empty((N, H, W, CP), device=img_u8.device, dtype=torch.float16)
    m = (C.c_float * 3)(*[float(v) for v in mean]); s = (C.c_float * 3)(*[float(v) for v in std])
    check(lib().csb_image_prep_nhwc(ptr(img_u8), C.c_longlong(N * H * W), m, s, int(swap_rb), CP, ptr(out), stream()), "csb_image_prep_nhwc")
    return out


def image_prep_s2d_nhwc(img_u8, mean, std, P, CP, swap_rb=False):
    """uint8 [N,H,W,3] -> fp16 [N,H/P,W/P,CP] space-to-depth (channel (r*P+s)*3+c), (x-mean)/std, zero padded: input of a patchify stem as a 1x1 GEMM."""
    if img_u8.dim() == 3:
        img_u8 = img_u8[None]
    N, H, W, _ = img_u8.shape
    out = torch.empty((N, H // P, W // P, CP), device=img_u8.device, dtype=torch.float16)
    m = (C.c_float * 3)(*[float(v) for v in mean]); s = (C.c_float * 3)(*[float(v) for v in std])
    check(lib().csb_image_prep_s2d_nhwc(ptr(img_u8), N, H, W, P, m, s, int(swap_rb), CP, ptr(out), stream()), "csb_image_prep_s2d_nhwc")
    return out


def maxpool3s2_nhwc(x):
    N, H, W, Cc = x.shape
    out = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc), device=x.device, dtype=x.dtype)
    check(lib().csb_maxpool_nhwc(ptr(x), N, H, W, Cc, ptr(out), stream()), "csb_maxpool_nhwc")
    return out


def add_nhwc(a, b, out=None):
    Cc = a.shape[-1]
    npix = a.numel() // Cc
    if out is None:
        out = torch.empty_like(a)
    check(lib().csb_add_nhwc(ptr(a), Cc, 0, ptr(b), b.shape[-1], 0, C.c_longlong(npix), Cc, ptr(out), out.shape[-1], 0, stream()), "csb_add_nhwc")
    return out


def resample_f32(x, Ho, Wo, align_corners):
    """x [N,Hi,Wi] fp32 -> [N,Ho,Wo] bilinear"""
    N, Hi, Wi = x.shape
    out = torch.empty((N, Ho, Wo), device=x.device, dtype=torch.float32)
    check(lib().csb_resample_f32(ptr(x), N, Hi, Wi, Ho, Wo, int(align_corners), ptr(out), stream()), "csb_resample_f32")
    return out


def pool_out(H, K, stride, pad, ceil_mode):
    o = (H + 2 * pad - K + stride - 1) // stride + 1 if ceil_mode else (H + 2 * pad - K) // stride + 1
    if ceil_mode and (o - 1) * stride >= H + pad:
        o -= 1
    return o


def maxpool2d_nhwc(x, K=2, stride=2, pad=0, ceil_mode=False, xoff=0, channels=None, out=None, yoff=0):
    N, H, W, ldx = x.shape
    Cc = ldx if channels is None else channels
    Ho, Wo = pool_out(H, K, stride, pad, ceil_mode), pool_out(W, K, stride, pad, ceil_mode)
    if out is None:
        out = torch.empty((N, Ho, Wo, Cc), device=x.device, dtype=x.dtype)
    check(lib().csb_maxpool2d_nhwc(ptr(x), ldx, xoff, N, H, W, Cc, K, stride, pad, int(ceil_mode), ptr(out), out.shape[3], yoff, stream()), "csb_maxpool2d_nhwc")
    return out
