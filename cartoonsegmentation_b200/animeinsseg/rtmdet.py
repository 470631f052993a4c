"""RTMDet-Ins detector forward on the B200 engine: ConvNeXt-B backbone, CSPNeXtPAFPN neck,
RTMDetInsSepBNHead + MaskFeatModule -- the network `AnimeInsSeg.model` holds in the reference
(built from the checkpoint's mmdet config, animeinsseg/__init__.py:196-209; architecture restated in SURVEY.md Appendix A.2-A.6).

Everything runs NHWC fp16 with fp32 accumulation:
  * every conv / linear is one `csb_conv2d_nhwc` launch (tcgen05 implicit GEMM) with BatchNorm, layer-scale and the `* stride`
    of the regression branch folded into the packed weights, and bias + SiLU/GELU/ReLU + residual fused into its epilogue;
  * sibling convs that read the same tensor are fused along Cout (CSPLayer main+short, the three head towers' first conv);
  * concats never copy: producers write straight into channel slices of the concat buffer (out_coff), consumers read slices (in_coff);
  * depthwise 7x7 (+ LayerNorm), depthwise 5x5 + BN + SiLU, LayerNorm2d, nearest / bilinear resampling are the HBM-bound kernels of
    csrc/nn_elem.cu; the LayerNorm of every ConvNeXt block is folded into the fc1 GEMM (the depthwise kernel emits per-pixel partial
    statistics, the GEMM epilogue applies rstd * (acc - mean * colsum) + bias'), so no LayerNorm pass touches the activations.

Parameters come in as a state_dict keyed by mmdet's parameter names (`backbone.*`, `neck.*`, `bbox_head.*`), so a real checkpoint
drops in; `synthetic_state_dict` generates the seeded variance-preserving weights BASELINE.json asks for.
"""
import math

import torch

from .. import engine as E
from .._lib import check, lib, ptr, stream

MEAN_BGR, STD_BGR = (103.53, 116.28, 123.675), (57.375, 57.12, 58.395)      # rtmdet base cfg, SURVEY Appendix A.1
DEPTHS, DIMS = (3, 3, 27, 3), (128, 256, 512, 1024)
STRIDES = (8, 16, 32)
NUM_GEN_PARAMS = 169
import os as _os
# Depthwise convs as diagonal-block MMAs on the halo-tile kernel (csrc/tc_halo.cu).  Measured on B200 (tools/halo_bench.py): the neck's 5x5 + SiLU
# gains 1.3x over the FFMA kernel (491 vs 377 G outputs/s), the backbone's 7x7 loses (246 vs 358: 49 taps x 4 MMAs of N = 16 per 64-channel
# tile are bound by MMA issue and the re-read of the A tile per tap), so the 7x7 stays on k_dwconv_halo unless CSB_DW_TC=1.
DW_TC = _os.environ.get("CSB_DW_TC", "0") != "0"
DW5_TC = _os.environ.get("CSB_DW5_TC", "1") != "0"
FOLD_LN = _os.environ.get("CSB_FOLD_LN", "1") != "0"        # ConvNeXt block: LayerNorm applied in the fc1 GEMM epilogue (engine.fold_layernorm)


# ------------------------------------------------------------------------------------------------ parameter inventory
def _convmodule(specs, name, cin, cout, k, groups=1, kind='conv_act'):
    specs.append((f"{name}.conv.weight", (cout, cin // groups, k, k), kind))
    for p, kind in (('weight', 'bn_w'), ('bias', 'bn_b'), ('running_mean', 'bn_m'), ('running_var', 'bn_v')):
        specs.append((f"{name}.bn.{p}", (cout,), kind))


def _csplayer(specs, name, cin, cout, n, identity=False):
    """identity: the blocks add their input (backbone stages 1-3).  Their last conv then gets a small synthetic gain, so that the residual stream keeps
    an O(1) variance through the 3-6 blocks of a stage (x + f(x) with a full-gain f doubles it per block: the detector's mask logits reached 1e8)."""
    mid = cout // 2
    _convmodule(specs, f"{name}.main_conv", cin, mid, 1)
    _convmodule(specs, f"{name}.short_conv", cin, mid, 1)
    _convmodule(specs, f"{name}.final_conv", 2 * mid, cout, 1)
    for b in range(n):
        _convmodule(specs, f"{name}.blocks.{b}.conv1", mid, mid, 3)
        _convmodule(specs, f"{name}.blocks.{b}.conv2.depthwise_conv", mid, mid, 5, groups=mid)
        _convmodule(specs, f"{name}.blocks.{b}.conv2.pointwise_conv", mid, mid, 1, kind='conv_act:0.6' if identity else 'conv_act')


CSPNEXT_L = dict(stem=(32, 32, 64), stages=((64, 128, 3, True, False), (128, 256, 6, True, False), (256, 512, 6, True, False), (512, 1024, 3, False, True)))


def _cspnext_specs(s):
    """mmdet `CSPNeXt(arch='P5', deepen_factor=1, widen_factor=1, expand_ratio=.5, channel_attention=True)` -- the backbone of the shipped
    rtmdetl_e60.ckpt: stem + stage1..4 = [conv 3x3 s2, (SPPBottleneck), CSPLayer + ChannelAttention]."""
    c0, c1, c2 = CSPNEXT_L['stem']
    _convmodule(s, "backbone.stem.0", 3, c0, 3)
    _convmodule(s, "backbone.stem.1", c0, c1, 3)
    _convmodule(s, "backbone.stem.2", c1, c2, 3)
    for i, (cin, cout, n, _identity, spp) in enumerate(CSPNEXT_L['stages'], 1):
        _convmodule(s, f"backbone.stage{i}.0", cin, cout, 3)
        j = 1
        if spp:
            _convmodule(s, f"backbone.stage{i}.1.conv1", cout, cout // 2, 1)
            _convmodule(s, f"backbone.stage{i}.1.conv2", cout * 2, cout, 1)
            j = 2
        _csplayer(s, f"backbone.stage{i}.{j}", cout, cout, n, identity=_identity)
        s += [(f"backbone.stage{i}.{j}.attention.fc.weight", (cout, cout, 1, 1), 'conv_lin'), (f"backbone.stage{i}.{j}.attention.fc.bias", (cout,), 'bias')]


def param_specs(backbone='convnext_b'):
    """[(name, shape, kind)] of the RTMDet-Ins detector (ConvNeXt-B or CSPNeXt-L backbone), mmdet parameter names."""
    s = []
    if backbone == 'cspnext_l':
        _cspnext_specs(s)
        return s + _neck_head_specs()
    assert backbone == 'convnext_b', backbone
    s += [("backbone.downsample_layers.0.0.weight", (DIMS[0], 3, 4, 4), 'conv_lin'), ("backbone.downsample_layers.0.0.bias", (DIMS[0],), 'bias'),
          ("backbone.downsample_layers.0.1.weight", (DIMS[0],), 'ln_w'), ("backbone.downsample_layers.0.1.bias", (DIMS[0],), 'ln_b')]
    for i in range(1, 4):
        s += [(f"backbone.downsample_layers.{i}.0.weight", (DIMS[i - 1],), 'ln_w'), (f"backbone.downsample_layers.{i}.0.bias", (DIMS[i - 1],), 'ln_b'),
              (f"backbone.downsample_layers.{i}.1.weight", (DIMS[i], DIMS[i - 1], 2, 2), 'conv_lin'), (f"backbone.downsample_layers.{i}.1.bias", (DIMS[i],), 'bias')]
    for i in range(4):
        d = DIMS[i]
        for b in range(DEPTHS[i]):
            p = f"backbone.stages.{i}.{b}"
            s += [(f"{p}.depthwise_conv.weight", (d, 1, 7, 7), 'conv_lin'), (f"{p}.depthwise_conv.bias", (d,), 'bias'),
                  (f"{p}.norm.weight", (d,), 'ln_w'), (f"{p}.norm.bias", (d,), 'ln_b'),
                  (f"{p}.pointwise_conv1.weight", (4 * d, d), 'conv_act'), (f"{p}.pointwise_conv1.bias", (4 * d,), 'bias'),
                  (f"{p}.pointwise_conv2.weight", (d, 4 * d), 'conv_res'), (f"{p}.pointwise_conv2.bias", (d,), 'bias'),
                  (f"{p}.gamma", (d,), 'gamma')]
    for i in (1, 2, 3):
        s += [(f"backbone.norm{i}.weight", (DIMS[i],), 'ln_w'), (f"backbone.norm{i}.bias", (DIMS[i],), 'ln_b')]
    return s + _neck_head_specs()


def _neck_head_specs():
    s = []
    c = DIMS[1:]
    _convmodule(s, "neck.reduce_layers.0", c[2], c[1], 1)
    _convmodule(s, "neck.reduce_layers.1", c[1], c[0], 1)
    _csplayer(s, "neck.top_down_blocks.0", 2 * c[1], c[1], 3)
    _csplayer(s, "neck.top_down_blocks.1", 2 * c[0], c[0], 3)
    _convmodule(s, "neck.downsamples.0", c[0], c[0], 3)
    _convmodule(s, "neck.downsamples.1", c[1], c[1], 3)
    _csplayer(s, "neck.bottom_up_blocks.0", 2 * c[0], c[1], 3)
    _csplayer(s, "neck.bottom_up_blocks.1", 2 * c[1], c[2], 3)
    for i in range(3):
        _convmodule(s, f"neck.out_convs.{i}", c[i], 256, 3)
    for tower in ("cls_convs", "reg_convs", "kernel_convs"):
        for lvl in range(3):
            for i in range(2):
                _convmodule(s, f"bbox_head.{tower}.{lvl}.{i}", 256, 256, 3)
    for lvl in range(3):
        s += [(f"bbox_head.rtm_cls.{lvl}.weight", (1, 256, 1, 1), 'conv_lin:25'), (f"bbox_head.rtm_cls.{lvl}.bias", (1,), 'cls_bias'),
              (f"bbox_head.rtm_reg.{lvl}.weight", (4, 256, 1, 1), 'conv_lin'), (f"bbox_head.rtm_reg.{lvl}.bias", (4,), 'reg_bias'),
              (f"bbox_head.rtm_kernel.{lvl}.weight", (NUM_GEN_PARAMS, 256, 1, 1), 'conv_lin:4'), (f"bbox_head.rtm_kernel.{lvl}.bias", (NUM_GEN_PARAMS,), 'bias')]
    s += [("bbox_head.mask_head.fusion_conv.weight", (256, 768, 1, 1), 'conv_lin'), ("bbox_head.mask_head.fusion_conv.bias", (256,), 'bias')]
    for i in range(4):
        _convmodule(s, f"bbox_head.mask_head.stacked_convs.{i}", 256, 256, 3)
    s += [("bbox_head.mask_head.projection.weight", (8, 256, 1, 1), 'conv_lin:4'), ("bbox_head.mask_head.projection.bias", (8,), 'bias')]
    return s


def synthetic_state_dict(seed=0, cls_bias=-5.0, backbone='convnext_b'):
    """Seeded variance-preserving weights: conv std sqrt(2/fan_in) before SiLU/GELU/ReLU, 1/sqrt(fan_in) before linear
    outputs, biases U(-0.1,0.1), BN gamma U(0.8,1.2) beta U(-0.1,0.1) mean N(0,0.1) var U(0.8,1.2), LN gamma U(0.8,1.2), layer-scale 1.0.
    (The reference's 'random init' is PyTorch's default because init_weights() is never called, animeinsseg/__init__.py:204-209; that
    init collapses activations -- documented deviation.)  share_conv: tower conv weights are generated once and shared across levels."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def u(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo
    for name, shape, kind in param_specs(backbone):
        gain = 1.0
        if ':' in kind:                       # 'conv_lin:25' -> output gain, so that scores / mask logits are well spread (not clustered at 0)
            kind, gain = kind.split(':')[0], float(kind.split(':')[1])
        if kind in ('conv_act', 'conv_lin', 'conv_res'):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            std = math.sqrt(2.0 / fan_in) if kind == 'conv_act' else (0.5 / math.sqrt(fan_in) if kind == 'conv_res' else 1.0 / math.sqrt(fan_in))
            sd[name] = torch.randn(shape, generator=g) * (std * gain)
        elif kind in ('bias', 'bn_b', 'ln_b'):
            sd[name] = u(shape, -0.1, 0.1)
        elif kind in ('bn_w', 'ln_w', 'bn_v'):
            sd[name] = u(shape, 0.8, 1.2)
        elif kind == 'bn_m':
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif kind == 'gamma':
            sd[name] = torch.ones(shape)
        elif kind == 'cls_bias':
            sd[name] = torch.full(shape, float(cls_bias))
        elif kind == 'reg_bias':
            sd[name] = u(shape, 1.0, 3.0)           # ltrb distances of a few strides -> boxes of useful size
        else:
            raise ValueError(kind)
    for tower in ("cls_convs", "reg_convs", "kernel_convs"):        # share_conv=True
        for lvl in (1, 2):
            for i in range(2):
                sd[f"bbox_head.{tower}.{lvl}.{i}.conv.weight"] = sd[f"bbox_head.{tower}.0.{i}.conv.weight"]
    return sd


# ------------------------------------------------------------------------------------------------ weight folding / packing
def _fold_bn(sd, name, eps):
    w = sd[f"{name}.conv.weight"].float()
    g, b, m, v = (sd[f"{name}.bn.{k}"].float() for k in ("weight", "bias", "running_mean", "running_var"))
    s = g / torch.sqrt(v + eps)
    return w * s.view(-1, 1, 1, 1), b - m * s


class _Conv:
    """One packed conv: weight [Cout,R,S,Cin] fp16 + fp32 bias, with its launch attributes."""

    def __init__(self, w, b, dev, stride=1, pad=0, act=None, cin_pad=None):
        self.w = E.pack_conv_weight(w.to(dev), torch.float16, cin_pad)
        self.b = None if b is None else b.float().contiguous().to(dev)
        self.stride, self.pad, self.act = stride, pad, act

    def __call__(self, x, **kw):
        return E.conv2d_nhwc(x, self.w, self.b, stride=self.stride, pad=self.pad, act=self.act, **kw)


class _CSP:
    """mmdet CSPLayer (CSPNeXtBlocks): add_identity=False / no attention in the neck, add_identity per stage + ChannelAttention in the backbone."""

    def __init__(self, sd, name, dev, eps, add_identity=False):
        wm, bm = _fold_bn(sd, f"{name}.main_conv", eps)
        ws, bs = _fold_bn(sd, f"{name}.short_conv", eps)
        self.mid = wm.shape[0]
        self.add_identity = add_identity
        self.ms = _Conv(torch.cat([wm, ws], 0), torch.cat([bm, bs], 0), dev, act='silu')          # main + short fused along Cout
        self.final = _Conv(*_fold_bn(sd, f"{name}.final_conv", eps), dev, act='silu')
        self.att = None
        if f"{name}.attention.fc.weight" in sd:                                                    # x * hardsigmoid(fc(GAP(x)))
            self.att = _Conv(sd[f"{name}.attention.fc.weight"], sd[f"{name}.attention.fc.bias"], dev, act='hardsigmoid')
        self.blocks = []
        b = 0
        while f"{name}.blocks.{b}.conv1.conv.weight" in sd:
            c1 = _Conv(*_fold_bn(sd, f"{name}.blocks.{b}.conv1", eps), dev, pad=1, act='silu')
            wd, bd = _fold_bn(sd, f"{name}.blocks.{b}.conv2.depthwise_conv", eps)
            dw = (wd[:, 0].permute(1, 2, 0).contiguous().to(dev), bd.contiguous().to(dev))            # [K,K,C] fp32
            if DW5_TC and dw[0].shape[2] % 64 == 0:
                dw = dw + (E.pack_dw_weight_compact(dw[0]),)
            pw = _Conv(*_fold_bn(sd, f"{name}.blocks.{b}.conv2.pointwise_conv", eps), dev, act='silu')
            self.blocks.append((c1, dw, pw))
            b += 1

    def __call__(self, x, out=None, out_coff=0):
        y = self.ms(x)                                        # [.., 2*mid] = [main | short]
        for c1, dw, pw in self.blocks:                        # CSPNeXtBlock
            a = E.conv2d_nhwc(y, c1.w, c1.b, pad=1, act='silu', in_coff=0)
            if len(dw) == 3:
                d = E.conv2d_halo_nhwc(a, dw[2], dw[1], pad=dw[0].shape[0] // 2, act='silu', groups=a.shape[3])
            else:
                d = E.dwconv_nhwc(a, dw[0], dw[1], act='silu')
            if self.add_identity:                             # out + identity: the block input is the main half itself (read, then overwritten per element)
                E.conv2d_nhwc(d, pw.w, pw.b, act='silu', out=y, out_coff=0, residual=y, res_mode=2, res_coff=0)
            else:
                E.conv2d_nhwc(d, pw.w, pw.b, act='silu', out=y, out_coff=0)      # back into the main half
        if self.att is not None:
            N, H, W, Cc = y.shape
            acc = torch.empty((N, Cc), device=y.device, dtype=torch.float32)
            pooled = torch.empty((N, 1, 1, Cc), device=y.device, dtype=y.dtype)
            check(lib().csb_gap_nhwc(ptr(y), Cc, 0, N, H, W, Cc, ptr(acc), ptr(pooled), stream()), "csb_gap_nhwc")
            gate = self.att(pooled)
            check(lib().csb_scale_channels_nhwc(ptr(y), Cc, 0, N, H, W, Cc, ptr(gate), stream()), "csb_scale_channels_nhwc")
        return E.conv2d_nhwc(y, self.final.w, self.final.b, act='silu', out=out, out_coff=out_coff)


class _CSPNeXt:
    """CSPNeXt-L backbone on the engine.  __call__(x16 [N,H,W,16], cat_slices {stage index 2|3|4: (buffer, channel offset)})."""

    def __init__(self, sd, dev, eps):
        cm = lambda n, **kw: _Conv(*_fold_bn(sd, n, eps), dev, act='silu', **kw)
        w0, b0 = _fold_bn(sd, "backbone.stem.0", eps)
        self.stem = [_Conv(w0, b0, dev, stride=2, pad=1, act='silu', cin_pad=16), cm("backbone.stem.1", pad=1), cm("backbone.stem.2", pad=1)]
        self.stages = []
        for i, (_cin, cout, _n, identity, spp) in enumerate(CSPNEXT_L['stages'], 1):
            st = dict(down=cm(f"backbone.stage{i}.0", stride=2, pad=1), spp=None)
            j = 1
            if spp:
                st['spp'] = (cm(f"backbone.stage{i}.1.conv1"), cm(f"backbone.stage{i}.1.conv2"), cout // 2)
                j = 2
            st['csp'] = _CSP(sd, f"backbone.stage{i}.{j}", dev, eps, add_identity=identity)
            self.stages.append(st)

    def __call__(self, x16, cat_slices):
        t, toff = x16, 0
        for c in self.stem:
            t = c(t)
        for i, st in enumerate(self.stages, 1):
            t = st['down'](t, in_coff=toff)                    # the previous stage may have written into a slice of a concat buffer
            if st['spp'] is not None:                          # SPPBottleneck: cat(x, pool5, pool9, pool13); pool9 = pool5 o pool5, pool13 = pool5^3 (exact for max)
                c1, c2, mid = st['spp']
                N, H, W, _ = t.shape
                cat = torch.empty((N, H, W, 4 * mid), device=t.device, dtype=t.dtype)
                c1(t, out=cat, out_coff=0)
                for k in range(3):
                    E.maxpool2d_nhwc(cat, 5, 1, 2, False, xoff=k * mid, channels=mid, out=cat, yoff=(k + 1) * mid)
                t = c2(cat)
            buf, toff = cat_slices.get(i, (None, 0))
            t = st['csp'](t, out=buf, out_coff=toff)
        return t


class RTMDetIns:
    """B200 forward of the detector.  `forward(img_u8)` -> (cls, reg, ker: per-level NHWC fp32; mask_feat [N,h,w,8] fp32)."""

    def __init__(self, state_dict, device='cuda', bn_eps=1e-5):
        sd, dev = state_dict, torch.device(device)
        self.dev = dev
        f32 = lambda t: t.float().contiguous().to(dev)
        self.cspnext = _CSPNeXt(sd, dev, bn_eps) if "backbone.stem.0.conv.weight" in sd else None          # rtmdetl_e60.ckpt layout (Appendix A.3)
        if self.cspnext is None:
            self._init_convnext(sd, dev, f32)
        self._init_neck_head(sd, dev, bn_eps)

    def _init_convnext(self, sd, dev, f32):
        # ---- backbone (Appendix A.4)
        # 4x4 stride-4 patchify as a 1x1 GEMM over the space-to-depth input written by csb_image_prep_s2d_nhwc: weight [Cout, (r*4+s)*3+c] padded to 64
        w0 = sd["backbone.downsample_layers.0.0.weight"].float()
        self.stem = _Conv(w0.permute(0, 2, 3, 1).reshape(w0.shape[0], 48, 1, 1), sd["backbone.downsample_layers.0.0.bias"], dev, cin_pad=64)
        self.stem_ln = (f32(sd["backbone.downsample_layers.0.1.weight"]), f32(sd["backbone.downsample_layers.0.1.bias"]))
        self.down = [None]
        for i in range(1, 4):
            self.down.append(((f32(sd[f"backbone.downsample_layers.{i}.0.weight"]), f32(sd[f"backbone.downsample_layers.{i}.0.bias"])),
                              _Conv(sd[f"backbone.downsample_layers.{i}.1.weight"], sd[f"backbone.downsample_layers.{i}.1.bias"], dev, stride=2)))
        self.blocks = []
        for i in range(4):
            st = []
            for b in range(DEPTHS[i]):
                p = f"backbone.stages.{i}.{b}"
                gamma = sd[f"{p}.gamma"].float()
                dw = f32(sd[f"{p}.depthwise_conv.weight"][:, 0].permute(1, 2, 0))                  # [7,7,C]
                # LayerNorm folded into fc1 (csb_conv2d_ln_nhwc): W' = W diag(gamma), b' = b + W beta, column sums of the rounded W'
                wl, bl, cs = E.fold_layernorm(sd[f"{p}.pointwise_conv1.weight"], sd[f"{p}.pointwise_conv1.bias"], sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"])
                st.append(dict(dw=dw, dw_tc=E.pack_dw_weight_compact(dw), fc1_ln=(wl.to(dev), bl.to(dev), cs.to(dev)), dwb=f32(sd[f"{p}.depthwise_conv.bias"]), ln=(f32(sd[f"{p}.norm.weight"]), f32(sd[f"{p}.norm.bias"])),
                               fc1=_Conv(sd[f"{p}.pointwise_conv1.weight"][:, :, None, None], sd[f"{p}.pointwise_conv1.bias"], dev, act='gelu'),
                               # layer scale folded: gamma * (W2 h + b2)
                               fc2=_Conv(sd[f"{p}.pointwise_conv2.weight"][:, :, None, None] * gamma.view(-1, 1, 1, 1), sd[f"{p}.pointwise_conv2.bias"] * gamma, dev)))
            self.blocks.append(st)
        self.out_ln = {i: (f32(sd[f"backbone.norm{i}.weight"]), f32(sd[f"backbone.norm{i}.bias"])) for i in (1, 2, 3)}

    def _init_neck_head(self, sd, dev, bn_eps):
        # ---- neck (Appendix A.5)
        cm = lambda n, **kw: _Conv(*_fold_bn(sd, n, bn_eps), dev, act='silu', **kw)
        self.reduce = [cm("neck.reduce_layers.0"), cm("neck.reduce_layers.1")]
        self.top_down = [_CSP(sd, "neck.top_down_blocks.0", dev, bn_eps), _CSP(sd, "neck.top_down_blocks.1", dev, bn_eps)]
        self.downs = [cm("neck.downsamples.0", stride=2, pad=1), cm("neck.downsamples.1", stride=2, pad=1)]
        self.bottom_up = [_CSP(sd, "neck.bottom_up_blocks.0", dev, bn_eps), _CSP(sd, "neck.bottom_up_blocks.1", dev, bn_eps)]
        self.out_convs = [cm(f"neck.out_convs.{i}", pad=1) for i in range(3)]
        # ---- head (Appendix A.6): first convs of the three towers fused along Cout per level (BN is per level)
        self.tower0, self.tower1, self.preds = [], [], []
        for lvl, stride in enumerate(STRIDES):
            ws, bs = zip(*[_fold_bn(sd, f"bbox_head.{t}.{lvl}.0", bn_eps) for t in ("cls_convs", "reg_convs", "kernel_convs")])
            self.tower0.append(_Conv(torch.cat(ws, 0), torch.cat(bs, 0), dev, pad=1, act='silu'))
            self.tower1.append([cm(f"bbox_head.{t}.{lvl}.1", pad=1) for t in ("cls_convs", "reg_convs", "kernel_convs")])
            self.preds.append((_Conv(sd[f"bbox_head.rtm_cls.{lvl}.weight"], sd[f"bbox_head.rtm_cls.{lvl}.bias"], dev),
                               # relu(x) * stride == relu(stride * x): fold the stride into weight and bias
                               _Conv(sd[f"bbox_head.rtm_reg.{lvl}.weight"] * stride, sd[f"bbox_head.rtm_reg.{lvl}.bias"] * stride, dev, act='relu'),
                               _Conv(sd[f"bbox_head.rtm_kernel.{lvl}.weight"], sd[f"bbox_head.rtm_kernel.{lvl}.bias"], dev)))
        self.fusion = _Conv(sd["bbox_head.mask_head.fusion_conv.weight"], sd["bbox_head.mask_head.fusion_conv.bias"], dev)
        self.mask_convs = [cm(f"bbox_head.mask_head.stacked_convs.{i}", pad=1) for i in range(4)]
        self.projection = _Conv(sd["bbox_head.mask_head.projection.weight"], sd["bbox_head.mask_head.projection.bias"], dev)

    # -------------------------------------------------------------------------------------------- forward
    def backbone(self, x16, cat_slices):
        """x16: [N,H,W,16] fp16 normalised image.  cat_slices: {stage: (buffer, channel offset)} where norm{stage} output is written."""
        t = self.stem(x16)
        t = E.layernorm_nhwc(t, *self.stem_ln)
        for i in range(4):
            if i > 0:
                ln, conv = self.down[i]
                t = conv(E.layernorm_nhwc(t, *ln))
            for blk in self.blocks[i]:
                if FOLD_LN and t.shape[3] % 64 == 0:
                    if DW_TC:       # tensor-core depthwise (fp16 taps, fp32 accumulate) emitting the LayerNorm partial sums
                        u, stats = E.conv2d_halo_nhwc(t, blk['dw_tc'], blk['dwb'], pad=3, groups=t.shape[3], stats=True)
                    else:
                        u, stats = E.dwconv_stats_nhwc(t, blk['dw'], blk['dwb'])
                    if E.convnext_mlp_supported(t.shape[3], blk['fc2'].w.shape[3]):
                        # stages 1-2: LN -> fc1 -> GELU -> fc2 -> + residual in one launch, the 4C intermediate never reaches HBM (k_mlp_tc)
                        E.convnext_mlp_nhwc(u, stats, *blk['fc1_ln'], blk['fc2'].w, blk['fc2'].b, t, eps=1e-6, out=t)
                        continue
                    h = E.conv2d_ln_nhwc(u, stats, *blk['fc1_ln'], eps=1e-6, act='gelu')
                else:
                    u = E.dwconv_nhwc(t, blk['dw'], blk['dwb'], ln=blk['ln'], eps=1e-6)
                    h = blk['fc1'](u)
                E.conv2d_nhwc(h, blk['fc2'].w, blk['fc2'].b, residual=t, res_mode=2, out=t)       # in place: t = t + gamma*(W2 h + b2)
            if i in cat_slices:
                buf, off = cat_slices[i]
                E.layernorm_nhwc(t, *self.out_ln[i], out=buf, yoff=off)
        return t

    def forward(self, img_u8):
        """[N,H,W,3] uint8 BGR -> (cls, reg, ker per level NHWC fp32, mask_feat).  Batches of up to 4 images (launch-bound: ~200 launches of a few us
        each) are replayed from a CUDA graph per input shape (utils/graphs.py); the returned tensors are then the graph's static outputs, valid until
        the next forward of the same shape."""
        if img_u8.dim() == 3:
            img_u8 = img_u8[None]
        if getattr(self, '_graphed', None) is None:
            from ..utils.graphs import GraphedForward
            self._graphed = GraphedForward(self._forward_impl)
        return self._graphed(img_u8.contiguous())

    def _forward_impl(self, img_u8):
        N, H, W, _ = img_u8.shape
        assert H % 32 == 0 and W % 32 == 0, "detector input must be padded to a multiple of 32 (Pad to det_size)"
        dev, f16 = img_u8.device, torch.float16
        h3, w3, h4, w4, h5, w5 = H // 8, W // 8, H // 16, W // 16, H // 32, W // 32
        if self.cspnext is not None:
            x16 = E.image_prep_nhwc(img_u8, MEAN_BGR, STD_BGR, swap_rb=False, CP=16)
        else:
            x16 = E.image_prep_s2d_nhwc(img_u8, MEAN_BGR, STD_BGR, 4, 64)                  # [N,H/4,W/4,64]: the ConvNeXt stem's GEMM input
        cat3 = torch.empty((N, h3, w3, 512), device=dev, dtype=f16)      # [up(p4) | c3]
        cat4 = torch.empty((N, h4, w4, 1024), device=dev, dtype=f16)     # [up(p5) | c4]
        c5 = torch.empty((N, h5, w5, 1024), device=dev, dtype=f16)
        catB4 = torch.empty((N, h4, w4, 512), device=dev, dtype=f16)     # [down(o3) | p4]
        catB5 = torch.empty((N, h5, w5, 1024), device=dev, dtype=f16)    # [down(o4) | p5]
        catM = torch.empty((N, h3, w3, 768), device=dev, dtype=f16)      # [P3 | up(P4) | up(P5)]
        if self.cspnext is not None:
            self.cspnext(x16, {2: (cat3, 256), 3: (cat4, 512), 4: (c5, 0)})
        else:
            self.backbone(x16, {1: (cat3, 256), 2: (cat4, 512), 3: (c5, 0)})
        # ---- neck
        self.reduce[0](c5, out=catB5, out_coff=512)                                           # p5
        E.resample_nhwc(catB5, h4, w4, 'nearest', out=cat4, xoff=512, yoff=0, channels=512)
        t4 = self.top_down[0](cat4)
        self.reduce[1](t4, out=catB4, out_coff=256)                                           # p4
        E.resample_nhwc(catB4, h3, w3, 'nearest', out=cat3, xoff=256, yoff=0, channels=256)
        o3 = self.top_down[1](cat3)
        self.downs[0](o3, out=catB4, out_coff=0)
        o4 = self.bottom_up[0](catB4)
        self.downs[1](o4, out=catB5, out_coff=0)
        o5 = self.bottom_up[1](catB5)
        self.out_convs[0](o3, out=catM, out_coff=0)                                           # P3 lives in the mask-feat concat buffer
        P4 = self.out_convs[1](o4)
        P5 = self.out_convs[2](o5)
        E.resample_nhwc(P4, h3, w3, 'bilinear', out=catM, yoff=256)
        E.resample_nhwc(P5, h3, w3, 'bilinear', out=catM, yoff=512)
        # ---- head
        cls, reg, ker = [], [], []
        for lvl, feat in enumerate((catM, P4, P5)):
            T = E.conv2d_nhwc(feat, self.tower0[lvl].w, self.tower0[lvl].b, pad=1, act='silu', in_coff=0)      # [.., 768] = [cls | reg | kernel]
            outs = []
            for j, (t1, pred) in enumerate(zip(self.tower1[lvl], self.preds[lvl])):
                f = E.conv2d_nhwc(T, t1.w, t1.b, pad=1, act='silu', in_coff=256 * j)
                outs.append(E.conv2d_nhwc(f, pred.w, pred.b, act=pred.act, out_f32=True))
            cls.append(outs[0]); reg.append(outs[1]); ker.append(outs[2])
        m = self.fusion(catM)
        for c in self.mask_convs:
            m = c(m)
        mask_feat = E.conv2d_nhwc(m, self.projection.w, self.projection.b, out_f32=True)
        return cls, reg, ker, mask_feat
