"""ZoeDepth metric head on the B200 engine.

Reference: `depth_modules/zoedepth/models/zoedepth/zoedepth_v1.py:124-202` with `config_zoedepth.json:2-19` (n_bins 64, softplus bin centres,
attractors [16,8,4,1], attractor_type inv, kind mean, min_temp 0.0212, max_temp 50): `conv2` on the bottleneck, `SeedBinRegressorUnnormed`
(`layers/localbins_layers.py:71-96`), `Projector`s (:99-117), four `AttractorLayerUnnormed` (`layers/attractor.py:139-208`),
`ConditionalLogBinomial` (`layers/dist_layers.py:72-121`), expectation over the bin centres.

All 1x1 convs run on the tcgen05 engine; the elementwise stages are the kernels of csrc/zoe_head.cu.  Inputs are what `MidasCore.forward`
returns (`base_models/midas.py:258-276`): the relative depth and six feature maps; parameters use the reference's names.

Rows B1-B3 -- `BeitDPT` (the MiDaS v3.1 `DPT_BEiT_L_384` the reference pulls through torch.hub at `base_models/midas.py:341`; NOT vendored in the
reference, so it is restated from the published timm BEiT + MiDaS DPT code and cross-checked against transformers' independent port,
tests/golden/make_zoe_dpt_golden.py) and `ZoeDepth` (`DepthModel.infer` with reflect-pad + flip augmentation, `depth_model.py:57-129`, and
`PrepForMidas`, `midas.py:164-186`).  Every Linear / Conv is a tcgen05 conv launch; attention is csrc/zoe_attn.cu.
"""
import ctypes as C
import math

import numpy as np
import torch

import os

from .. import engine as E
from .._lib import check, lib, ptr, stream

ATTN_TC = os.environ.get("CSB_ATTN_TC", "1") != "0"          # tcgen05 attention (zoe_attn_tc.cu); 0 = the mma.sync kernel (zoe_attn.cu)

N_BINS, EMB, N_ATTR = 64, 128, (16, 8, 4, 1)
MIN_TEMP, MAX_TEMP, P_EPS = 0.0212, 50.0, 1e-4
ATTRACTOR_ALPHA = 300.0          # the function default the reference actually applies (attractor.py:45,195), not config's 1000


def param_specs():
    s = [("conv2.weight", (256, 256, 1, 1), 'lin'), ("conv2.bias", (256,), 'bias')]

    def mlp(name, cin, mid, cout, last='lin'):
        s.extend([(f"{name}._net.0.weight", (mid, cin, 1, 1), 'act'), (f"{name}._net.0.bias", (mid,), 'bias'),
                  (f"{name}._net.2.weight", (cout, mid, 1, 1), last), (f"{name}._net.2.bias", (cout,), 'bias')])
    mlp("seed_bin_regressor", 256, 256, N_BINS)
    mlp("seed_projector", 256, 128, EMB)
    for i in range(4):
        mlp(f"projectors.{i}", 256, 128, EMB)
    for i, na in enumerate(N_ATTR):
        mlp(f"attractors.{i}", EMB, 128, na)
    s += [("conditional_log_binomial.mlp.0.weight", (80, 161, 1, 1), 'act'), ("conditional_log_binomial.mlp.0.bias", (80,), 'bias'),
          ("conditional_log_binomial.mlp.2.weight", (4, 80, 1, 1), 'lin'), ("conditional_log_binomial.mlp.2.bias", (4,), 'bias')]
    return s


def synthetic_state_dict(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in param_specs():
        if kind in ('lin', 'act'):
            fan_in = shape[1]
            sd[name] = torch.randn(shape, generator=g) * (math.sqrt(2.0 / fan_in) if kind == 'act' else 1.0 / math.sqrt(fan_in))
        else:
            sd[name] = torch.rand(shape, generator=g) * 0.2 - 0.1
    return sd


class _MLP:
    def __init__(self, sd, name, dev, last_act=None, cin_pad=None):
        self.c0 = (E.pack_conv_weight(sd[f"{name}.0.weight"].to(dev), torch.float16, cin_pad), sd[f"{name}.0.bias"].float().contiguous().to(dev))
        self.c1 = (E.pack_conv_weight(sd[f"{name}.2.weight"].to(dev)), sd[f"{name}.2.bias"].float().contiguous().to(dev))
        self.last_act = last_act

    def __call__(self, x, mid_act='relu', out_f32=False):
        h = E.conv2d_nhwc(x, self.c0[0], self.c0[1], act=mid_act)
        return E.conv2d_nhwc(h, self.c1[0], self.c1[1], act=self.last_act, out_f32=out_f32)


class ZoeHead:
    """forward(rel_depth [N,Hr,Wr] fp32, outconv [N,H,W,32] fp16, btlnck [N,h,w,256] fp16, blocks: 4 x [N,h_i,w_i,256] fp16) -> metric depth [N,H,W] fp32"""

    def __init__(self, state_dict=None, device='cuda'):
        sd = synthetic_state_dict(0) if state_dict is None else state_dict
        dev = self.dev = torch.device(device)
        self.conv2 = (E.pack_conv_weight(sd["conv2.weight"].to(dev)), sd["conv2.bias"].float().contiguous().to(dev))
        self.seed = _MLP(sd, "seed_bin_regressor._net", dev, last_act='softplus')
        self.seed_proj = _MLP(sd, "seed_projector._net", dev)
        self.proj = [_MLP(sd, f"projectors.{i}._net", dev) for i in range(4)]
        self.attr = [_MLP(sd, f"attractors.{i}._net", dev, last_act='softplus') for i in range(4)]
        # csb_zoe_cond_input writes [feat 0..31 | embedding 32..159 | rel_depth 160] (the reference concatenates [feat, rel, embedding]): permute the
        # input channels of the first 1x1 conv to that order
        clb = dict(sd)
        w0 = sd["conditional_log_binomial.mlp.0.weight"]
        clb["conditional_log_binomial.mlp.0.weight"] = torch.cat([w0[:, :32], w0[:, 33:161], w0[:, 32:33]], 1)
        self.clb = _MLP(clb, "conditional_log_binomial.mlp", dev, last_act='softplus', cin_pad=176)

    def forward(self, rel_depth, outconv, btlnck, blocks):
        N = btlnck.shape[0]
        dev = btlnck.device
        x = E.conv2d_nhwc(btlnck, self.conv2[0], self.conv2[1])                               # conv2 :151
        b_prev = self.seed(x, out_f32=True)                                                   # seed bin centres [N,h,w,64] fp32 :153,159
        prev_emb = self.seed_proj(x)                                                          # :161
        b_emb = None
        for proj, attr, na, feat in zip(self.proj, self.attr, N_ATTR, blocks):                # :164-170
            b_emb = proj(feat)
            h, w = feat.shape[1:3]
            up = E.resample_nhwc(prev_emb, h, w, 'bilinear_ac')                               # attractor.py:176-180
            A = attr(E.add_nhwc(b_emb, up), out_f32=True)                                     # [N,h,w,na] fp32 (softplus)
            b_new = torch.empty((N, h, w, N_BINS), device=dev, dtype=torch.float32)
            check(lib().csb_zoe_attractor(ptr(A), na, ptr(b_prev), b_prev.shape[1], b_prev.shape[2], N, h, w, N_BINS, C.c_float(ATTRACTOR_ALPHA), ptr(b_new), stream()),
                  "csb_zoe_attractor")
            b_prev, prev_emb = b_new, b_emb
        H, W = outconv.shape[1:3]
        cond = torch.empty((N, H, W, 176), device=dev, dtype=torch.float16)                   # cat([last, rel_cond, b_embedding]) :176-184
        check(lib().csb_zoe_cond_input(ptr(outconv), ptr(rel_depth.contiguous()), rel_depth.shape[1], rel_depth.shape[2], ptr(b_emb), b_emb.shape[1], b_emb.shape[2],
                                       N, H, W, ptr(cond), stream()), "csb_zoe_cond_input")
        pt = self.clb(cond, mid_act='gelu', out_f32=True)                                     # [N,H,W,4] fp32 (softplus)
        depth = torch.empty((N, H, W), device=dev, dtype=torch.float32)
        check(lib().csb_zoe_logbinom_depth(ptr(pt), ptr(b_prev), b_prev.shape[1], b_prev.shape[2], N, H, W, N_BINS, C.c_float(P_EPS), C.c_float(MIN_TEMP),
                                           C.c_float(MAX_TEMP), ptr(depth), stream()), "csb_zoe_logbinom_depth")
        return depth


# ================================================================================================================
# B3: MiDaS DPT_BEiT_L_384 = timm beit_large_patch16_384 (24 blocks, 1024 wide, 16 heads, per-block relative position bias, layer scale) +
# DPT reassemble ('project' readout, hooks after blocks 5/11/17/23) + 4 FeatureFusionBlock_custom + output_conv.
# Parameter names are those of the reference checkpoint ZoeD_M12_N.pt below its `core.core.` prefix.
# ================================================================================================================
BEIT = dict(depth=24, dim=1024, heads=16, mlp=4096, patch=16, window=24, hooks=(5, 11, 17, 23))
DPT_FEATURES = (256, 512, 1024, 1024)


def dpt_param_specs():
    D, M = BEIT['dim'], BEIT['mlp']
    nrel = (2 * BEIT['window'] - 1) ** 2 + 3
    s = [("pretrained.model.cls_token", (1, 1, D), 'tok'), ("pretrained.model.patch_embed.proj.weight", (D, 3, 16, 16), 'lin'),
         ("pretrained.model.patch_embed.proj.bias", (D,), 'bias')]
    for i in range(BEIT['depth']):
        b = f"pretrained.model.blocks.{i}"
        s += [(f"{b}.norm1.weight", (D,), 'ln_w'), (f"{b}.norm1.bias", (D,), 'bias'), (f"{b}.attn.qkv.weight", (3 * D, D), 'lin'),
              (f"{b}.attn.q_bias", (D,), 'bias'), (f"{b}.attn.v_bias", (D,), 'bias'), (f"{b}.attn.relative_position_bias_table", (nrel, BEIT['heads']), 'relpos'),
              (f"{b}.attn.proj.weight", (D, D), 'lin'), (f"{b}.attn.proj.bias", (D,), 'bias'), (f"{b}.gamma_1", (D,), 'gamma'),
              (f"{b}.norm2.weight", (D,), 'ln_w'), (f"{b}.norm2.bias", (D,), 'bias'), (f"{b}.mlp.fc1.weight", (M, D), 'act'), (f"{b}.mlp.fc1.bias", (M,), 'bias'),
              (f"{b}.mlp.fc2.weight", (D, M), 'lin'), (f"{b}.mlp.fc2.bias", (D,), 'bias'), (f"{b}.gamma_2", (D,), 'gamma')]
    for k, (f, factor) in enumerate(zip(DPT_FEATURES, (4, 2, 1, 0)), 1):
        a = f"pretrained.act_postprocess{k}"
        s += [(f"{a}.0.project.0.weight", (D, 2 * D), 'act'), (f"{a}.0.project.0.bias", (D,), 'bias'), (f"{a}.3.weight", (f, D, 1, 1), 'lin'), (f"{a}.3.bias", (f,), 'bias')]
        if factor > 1:
            s += [(f"{a}.4.weight", (f, f, factor, factor), 'convT'), (f"{a}.4.bias", (f,), 'bias')]
        elif factor == 0:
            s += [(f"{a}.4.weight", (f, f, 3, 3), 'lin'), (f"{a}.4.bias", (f,), 'bias')]
        s += [(f"scratch.layer{k}_rn.weight", (256, f, 3, 3), 'lin')]
    for k in range(1, 5):
        r = f"scratch.refinenet{k}"
        s += [(f"{r}.out_conv.weight", (256, 256, 1, 1), 'lin'), (f"{r}.out_conv.bias", (256,), 'bias')]
        for u in ("resConfUnit1", "resConfUnit2"):
            for c in ("conv1", "conv2"):
                s += [(f"{r}.{u}.{c}.weight", (256, 256, 3, 3), 'act'), (f"{r}.{u}.{c}.bias", (256,), 'bias')]
    s += [("scratch.output_conv.0.weight", (128, 256, 3, 3), 'lin'), ("scratch.output_conv.0.bias", (128,), 'bias'),
          ("scratch.output_conv.2.weight", (32, 128, 3, 3), 'act'), ("scratch.output_conv.2.bias", (32,), 'bias'),
          ("scratch.output_conv.4.weight", (1, 32, 1, 1), 'act'), ("scratch.output_conv.4.bias", (1,), 'bias')]
    return s


def dpt_synthetic_state_dict(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in dpt_param_specs():
        if kind in ('lin', 'act', 'convT'):
            fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 and kind != 'convT' else 1)
            sd[name] = torch.randn(shape, generator=g) * (math.sqrt(2.0 / fan_in) if kind == 'act' else 1.0 / math.sqrt(fan_in))
        elif kind == 'ln_w':
            sd[name] = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif kind == 'gamma':
            sd[name] = torch.rand(shape, generator=g) * 0.2 + 0.1
        elif kind == 'relpos':
            sd[name] = torch.randn(shape, generator=g) * 0.5
        elif kind == 'tok':
            sd[name] = torch.randn(shape, generator=g) * 0.5
        else:
            sd[name] = torch.rand(shape, generator=g) * 0.2 - 0.1
    return sd


def relative_position_bias(table, old_window, new_window):
    """MiDaS v3.1 `backbones/beit.py:_get_rel_pos_bias` + `gen_relative_position_index`: bilinear interpolation of the (2W-1)^2 table to the
    new window, then the (area+1)^2 index gather -> [heads, T, T] fp32 (T = area + 1, index 0 = cls)."""
    oh, ow = 2 * old_window[0] - 1, 2 * old_window[1] - 1
    nh, nw = 2 * new_window[0] - 1, 2 * new_window[1] - 1
    n_old = oh * ow + 3
    sub = table[:n_old - 3].reshape(1, ow, oh, -1).permute(0, 3, 1, 2)
    sub = torch.nn.functional.interpolate(sub.float(), size=(nh, nw), mode="bilinear")
    new_table = torch.cat([sub.permute(0, 2, 3, 1).reshape(nh * nw, -1), table[n_old - 3:].float()])
    n_new = nh * nw + 3
    area = new_window[0] * new_window[1]
    coords = torch.stack(torch.meshgrid(torch.arange(new_window[0]), torch.arange(new_window[1]), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += new_window[0] - 1
    rel[:, :, 1] += new_window[1] - 1
    rel[:, :, 0] *= 2 * new_window[1] - 1
    idx = torch.zeros((area + 1,) * 2, dtype=rel.dtype)
    idx[1:, 1:] = rel.sum(-1)
    idx[0, 0:] = n_new - 3
    idx[0:, 0] = n_new - 2
    idx[0, 0] = n_new - 1
    return new_table[idx.view(-1).to(new_table.device)].view(area + 1, area + 1, -1).permute(2, 0, 1).contiguous()


def _lin(w, b, dev, scale=None):
    """nn.Linear [out,in] (optionally scaled per output row: layer scale folded) -> packed 1x1 conv weight + fp32 bias."""
    w = w.float()
    b = None if b is None else b.float()
    if scale is not None:
        w = w * scale.float()[:, None]
        b = None if b is None else b * scale.float()
    return E.pack_conv_weight(w.reshape(w.shape[0], w.shape[1], 1, 1).to(dev)), None if b is None else b.contiguous().to(dev)


def _conv(sd, name, dev, bias=True):
    return E.pack_conv_weight(sd[f"{name}.weight"].to(dev)), sd[f"{name}.bias"].float().contiguous().to(dev) if bias else None


class BeitDPT:
    """forward(patches [B,hp,wp,768] fp16 from csb_zoe_prep) -> (rel_depth [B,16hp,16wp] fp32, outconv [B,16hp,16wp,32], l4_rn, [r4, r3, r2, r1])
    -- what `MidasCore.forward(x, return_rel_depth=True)` returns (midas.py:258-276) for layer_names ('out_conv','l4_rn','r4','r3','r2','r1')."""

    def __init__(self, state_dict=None, device='cuda', eps=1e-6):
        sd = dpt_synthetic_state_dict(0) if state_dict is None else state_dict
        dev = self.dev = torch.device(device)
        self.eps = eps
        f32 = lambda t: t.float().contiguous().to(dev)
        D = BEIT['dim']
        pw = sd["pretrained.model.patch_embed.proj.weight"].float()                        # [D,3,16,16] -> [(r*16+s)*3+c]
        self.patch = (E.pack_conv_weight(pw.permute(0, 2, 3, 1).reshape(D, 768, 1, 1).to(dev)), f32(sd["pretrained.model.patch_embed.proj.bias"]))
        self.cls = sd["pretrained.model.cls_token"].reshape(D).to(dev).half().contiguous()
        self.blocks = []
        self.tables = []
        for i in range(BEIT['depth']):
            b = f"pretrained.model.blocks.{i}"
            qkv_b = torch.cat([sd[f"{b}.attn.q_bias"].float(), torch.zeros(D), sd[f"{b}.attn.v_bias"].float()])       # timm: k has no bias
            self.blocks.append(dict(
                n1=(f32(sd[f"{b}.norm1.weight"]), f32(sd[f"{b}.norm1.bias"])), qkv=_lin(sd[f"{b}.attn.qkv.weight"], qkv_b, dev),
                proj=_lin(sd[f"{b}.attn.proj.weight"], sd[f"{b}.attn.proj.bias"], dev, sd[f"{b}.gamma_1"]),
                n2=(f32(sd[f"{b}.norm2.weight"]), f32(sd[f"{b}.norm2.bias"])), fc1=_lin(sd[f"{b}.mlp.fc1.weight"], sd[f"{b}.mlp.fc1.bias"], dev),
                fc2=_lin(sd[f"{b}.mlp.fc2.weight"], sd[f"{b}.mlp.fc2.bias"], dev, sd[f"{b}.gamma_2"])))
            self.tables.append(sd[f"{b}.attn.relative_position_bias_table"].float())
        self._bias_cache = {}
        self.post = []
        for k, (f, factor) in enumerate(zip(DPT_FEATURES, (4, 2, 1, 0)), 1):
            a = f"pretrained.act_postprocess{k}"
            e = dict(readout=_lin(sd[f"{a}.0.project.0.weight"], sd[f"{a}.0.project.0.bias"], dev), proj=_conv(sd, f"{a}.3", dev), factor=factor, f=f)
            if factor > 1:                       # ConvTranspose2d(f, f, k, stride=k): weight [in, out, i, j] -> 1x1 conv to (i*k+j)*f + out
                wt = sd[f"{a}.4.weight"].float()
                e['up'] = (E.pack_conv_weight(wt.permute(2, 3, 1, 0).reshape(factor * factor * f, f, 1, 1).to(dev)), f32(sd[f"{a}.4.bias"].float().repeat(factor * factor)))
            elif factor == 0:
                e['down'] = _conv(sd, f"{a}.4", dev)
            e['rn'] = _conv(sd, f"scratch.layer{k}_rn", dev, bias=False)
            self.post.append(e)
        self.refine = {}
        for k in range(1, 5):
            r = f"scratch.refinenet{k}"
            self.refine[k] = dict(out=_conv(sd, f"{r}.out_conv", dev), u1=(_conv(sd, f"{r}.resConfUnit1.conv1", dev), _conv(sd, f"{r}.resConfUnit1.conv2", dev)),
                                  u2=(_conv(sd, f"{r}.resConfUnit2.conv1", dev), _conv(sd, f"{r}.resConfUnit2.conv2", dev)))
        self.oc0, self.oc2 = _conv(sd, "scratch.output_conv.0", dev), _conv(sd, "scratch.output_conv.2", dev)
        w4 = torch.zeros(16, 32, 1, 1)                                                   # Cout 1 -> padded to 16 rows for the engine
        w4[0] = sd["scratch.output_conv.4.weight"].float()[0]
        self.oc4 = (E.pack_conv_weight(w4.to(dev)), f32(torch.cat([sd["scratch.output_conv.4.bias"].float(), torch.zeros(15)])))
        self.zero256 = torch.zeros(256, device=dev)

    def _bias(self, hp, wp):
        """per-block additive attention bias [16, Tp, Tp] fp16, rows/cols padded to a multiple of 128 (padding columns = -60000)."""
        key = (hp, wp)
        if key not in self._bias_cache:
            T = hp * wp + 1
            Tp = (T + 127) // 128 * 128
            out = []
            for tab in self.tables:
                rb = relative_position_bias(tab.to(self.dev), (BEIT['window'],) * 2, (hp, wp))
                full = torch.full((BEIT['heads'], Tp, Tp), -60000.0, device=self.dev, dtype=torch.float16)
                full[:, :T, :T] = rb.half()
                out.append(full)
            self._bias_cache = {key: (out, T, Tp)}                                       # one window size is kept (315 MB for 24 x 24)
        return self._bias_cache[key]

    def _rcu(self, x, unit):
        """ResidualConvUnit_custom (MiDaS blocks.py): conv2(relu(conv1(relu(x)))) + x"""
        N, H, W, Cc = x.shape
        a = torch.empty_like(x)
        check(lib().csb_prelu_nhwc(ptr(x), Cc, 0, ptr(self.zero256), C.c_longlong(N * H * W), Cc, ptr(a), Cc, 0, stream()), "csb_prelu_nhwc")
        h = E.conv2d_nhwc(a, unit[0][0], unit[0][1], pad=1, act='relu')
        return E.conv2d_nhwc(h, unit[1][0], unit[1][1], pad=1, residual=x, res_mode=2)

    def _fusion(self, k, path, skip=None):
        """FeatureFusionBlock_custom (MiDaS blocks.py): [path + rcu1(skip)] -> rcu2 -> x2 bilinear (align_corners=True) -> out_conv"""
        r = self.refine[k]
        out = path if skip is None else E.add_nhwc(path, self._rcu(skip, r['u1']))
        out = self._rcu(out, r['u2'])
        out = E.resample_nhwc(out, out.shape[1] * 2, out.shape[2] * 2, 'bilinear_ac')
        return E.conv2d_nhwc(out, r['out'][0], r['out'][1])

    def forward(self, patches):
        B, hp, wp, _ = patches.shape
        D, heads = BEIT['dim'], BEIT['heads']
        P = hp * wp
        biases, T, Tp = self._bias(hp, wp)
        emb = E.conv2d_nhwc(patches, self.patch[0], self.patch[1])                                 # patch_embed.proj as a 768 -> 1024 GEMM
        x = torch.empty((B, 1, T, D), device=self.dev, dtype=torch.float16)
        check(lib().csb_tokens_assemble(ptr(emb), ptr(self.cls), B, P, D, ptr(x), stream()), "csb_tokens_assemble")
        hooked = []
        use_tc = ATTN_TC
        if use_tc:
            lib().csb_attention_tc_scratch_bytes.restype = C.c_longlong
            vt = torch.empty(int(lib().csb_attention_tc_scratch_bytes(B, T, heads)), device=self.dev, dtype=torch.uint8)
        for i, blk in enumerate(self.blocks):
            h = E.layernorm_nhwc(x, blk['n1'][0], blk['n1'][1], self.eps)
            qkv = E.conv2d_nhwc(h, blk['qkv'][0], blk['qkv'][1])
            att = torch.empty((B, 1, T, D), device=self.dev, dtype=torch.float16)
            if use_tc:      # tcgen05 kernel (zoe_attn_tc.cu)
                check(lib().csb_attention_bias_tc(ptr(qkv), B, T, heads, D // heads, ptr(biases[i]), Tp, C.c_float((D // heads) ** -0.5), ptr(vt), ptr(att),
                                                  stream()), "csb_attention_bias_tc")
            else:           # mma.sync kernel (zoe_attn.cu), kept for A/B runs: CSB_ATTN_TC=0
                check(lib().csb_attention_bias(ptr(qkv), B, T, heads, D // heads, ptr(biases[i]), Tp, C.c_float((D // heads) ** -0.5), ptr(att), stream()),
                      "csb_attention_bias")
            x = E.conv2d_nhwc(att, blk['proj'][0], blk['proj'][1], residual=x, res_mode=2)          # x + gamma_1 * proj(attn)
            h = E.layernorm_nhwc(x, blk['n2'][0], blk['n2'][1], self.eps)
            h = E.conv2d_nhwc(h, blk['fc1'][0], blk['fc1'][1], act='gelu')
            x = E.conv2d_nhwc(h, blk['fc2'][0], blk['fc2'][1], residual=x, res_mode=2)              # x + gamma_2 * mlp
            if i in BEIT['hooks']:
                hooked.append(x)
        feats = []
        for e, tok in zip(self.post, hooked):                                                      # act_postprocess1..4 + scratch.layerN_rn
            cat = torch.empty((B, hp, wp, 2 * D), device=self.dev, dtype=torch.float16)
            check(lib().csb_readout_concat(ptr(tok), B, P, D, ptr(cat), stream()), "csb_readout_concat")
            f = E.conv2d_nhwc(cat, e['readout'][0], e['readout'][1], act='gelu')                    # ProjectReadout
            f = E.conv2d_nhwc(f, e['proj'][0], e['proj'][1])
            if e['factor'] > 1:
                k = e['factor']
                wide = E.conv2d_nhwc(f, e['up'][0], e['up'][1])
                f = torch.empty((B, hp * k, wp * k, e['f']), device=self.dev, dtype=torch.float16)
                check(lib().csb_pixel_shuffle_nhwc(ptr(wide), B, hp, wp, k, e['f'], ptr(f), stream()), "csb_pixel_shuffle_nhwc")
            elif e['factor'] == 0:
                f = E.conv2d_nhwc(f, e['down'][0], e['down'][1], stride=2, pad=1)
            feats.append(E.conv2d_nhwc(f, e['rn'][0], None, pad=1))
        l1, l2, l3, l4 = feats
        p4 = self._fusion(4, l4)
        p3 = self._fusion(3, p4, l3)
        p2 = self._fusion(2, p3, l2)
        p1 = self._fusion(1, p2, l1)
        o = E.conv2d_nhwc(p1, self.oc0[0], self.oc0[1], pad=1)                                     # scratch.output_conv
        o = E.resample_nhwc(o, o.shape[1] * 2, o.shape[2] * 2, 'bilinear_ac')
        outconv = E.conv2d_nhwc(o, self.oc2[0], self.oc2[1], pad=1, act='relu')                    # hooked 'out_conv' activation (after the ReLU, midas.py:294-296)
        rel = E.conv2d_nhwc(outconv, self.oc4[0], self.oc4[1], act='relu', out_f32=True)[..., 0].contiguous()
        return rel, outconv, l4, [p4, p3, p2, p1]


def midas_net_size(h, w, net=384, multiple=32):
    """`Resize(net_w, net_h, keep_aspect_ratio=True, ensure_multiple_of=32, resize_method='minimal').get_size` (midas.py:104-148); `net` is the
    config's `img_size` = int or (net_h, net_w) (midas.py:177-179)"""
    net_h, net_w = (net, net) if isinstance(net, int) else (int(net[0]), int(net[1]))
    sh, sw = net_h / h, net_w / w
    if abs(1 - sw) < abs(1 - sh):
        sh = sw
    else:
        sw = sh
    rnd = lambda v: int(np.round(v / multiple) * multiple)
    return rnd(sh * h), rnd(sw * w)


class ZoeDepth:
    """`ZoeDepth.infer(x, pad_input=True, with_flip_aug=True)` of the reference (depth_model.py:114-129) for one [H,W,3] uint8 image on the device
    (the reference passes BGR/255 without reordering, kenburns_effect.py:813 -- kept): -> metric depth [H,W] fp32."""

    def __init__(self, state_dict=None, device='cuda', img_size=384):
        """`img_size`: the MidasCore input resolution (midas.py:336-339; 384 when the config has none).  The reference's Ken-Burns pipeline
        loads ZoeDepth with img_size=[672, 672] (kenburns_effect.py:543); `load_zoe` alone defaults to [512, 672] (depth_modules/__init__.py:40)."""
        dev = self.dev = torch.device(device)
        self.img_size = img_size
        core = head = None
        if state_dict is not None:
            core = {k[len("core.core."):]: v for k, v in state_dict.items() if k.startswith("core.core.")}
            head = {k: v for k, v in state_dict.items() if not k.startswith("core.")}
        self.core = BeitDPT(core, dev)
        self.head = ZoeHead(head, dev)
        self._scratch = torch.zeros(1, device=dev, dtype=torch.int32)

    def infer(self, img_u8, pad_input=True, with_flip_aug=True):
        H, W = img_u8.shape[:2]
        ph = int(np.sqrt(H / 2) * 3) if pad_input else 0                                           # depth_model.py:81-82 (fh = fw = 3)
        pw = int(np.sqrt(W / 2) * 3) if pad_input else 0
        Hn, Wn = midas_net_size(H + 2 * ph, W + 2 * pw, self.img_size)
        nb = 2 if with_flip_aug else 1
        patches = torch.empty((nb, Hn // 16, Wn // 16, 768), device=self.dev, dtype=torch.float16)
        check(lib().csb_zoe_prep(ptr(img_u8.contiguous()), H, W, ph, pw, Hn, Wn, int(with_flip_aug), ptr(patches), stream()), "csb_zoe_prep")
        rel, outconv, btl, blocks = self.core.forward(patches)
        metric = self.head.forward(rel, outconv, btl, blocks)                                      # [nb, Hn, Wn] fp32
        out = torch.empty((H, W), device=self.dev, dtype=torch.float32)
        check(lib().csb_zoe_finish(ptr(metric), int(with_flip_aug), Hn, Wn, H, W, ph, pw, ptr(out), stream()), "csb_zoe_finish")
        return out

    def infer_batch(self, imgs_u8, pad_input=True, with_flip_aug=True):
        """Same as `infer` for a stack [B,H,W,3] of equally sized images: one encoder / head pass over all 2B net inputs -> [B,H,W] fp32.
        (The reference is strictly batch 1; results per image are identical to `infer`.)"""
        Bn, H, W = imgs_u8.shape[:3]
        ph = int(np.sqrt(H / 2) * 3) if pad_input else 0
        pw = int(np.sqrt(W / 2) * 3) if pad_input else 0
        Hn, Wn = midas_net_size(H + 2 * ph, W + 2 * pw, self.img_size)
        nb = 2 if with_flip_aug else 1
        imgs_u8 = imgs_u8.contiguous()
        patches = torch.empty((Bn * nb, Hn // 16, Wn // 16, 768), device=self.dev, dtype=torch.float16)
        for i in range(Bn):
            check(lib().csb_zoe_prep(ptr(imgs_u8[i]), H, W, ph, pw, Hn, Wn, int(with_flip_aug), ptr(patches[i * nb:]), stream()), "csb_zoe_prep")
        rel, outconv, btl, blocks = self.core.forward(patches)
        metric = self.head.forward(rel, outconv, btl, blocks)
        out = torch.empty((Bn, H, W), device=self.dev, dtype=torch.float32)
        for i in range(Bn):
            check(lib().csb_zoe_finish(ptr(metric[i * nb:]), int(with_flip_aug), Hn, Wn, H, W, ph, pw, ptr(out[i]), stream()), "csb_zoe_finish")
        return out

    def disparity(self, depth, focal, baseline, out=None):
        """`_depth_est_zoe` tail (kenburns_effect.py:815-817)"""
        out = torch.empty_like(depth) if out is None else out
        check(lib().csb_zoe_disparity(ptr(depth), C.c_longlong(depth.numel()), C.c_double(focal), C.c_double(baseline), ptr(out), ptr(self._scratch), stream()),
              "csb_zoe_disparity")
        return out
