"""ZoeDepth metric head on the B200 engine.

Reference: `depth_modules/zoedepth/models/zoedepth/zoedepth_v1.py:124-202` with `config_zoedepth.json:2-19` (n_bins 64, softplus bin centres,
attractors [16,8,4,1], attractor_type inv, kind mean, min_temp 0.0212, max_temp 50): `conv2` on the bottleneck, `SeedBinRegressorUnnormed`
(`layers/localbins_layers.py:71-96`), `Projector`s (:99-117), four `AttractorLayerUnnormed` (`layers/attractor.py:139-208`),
`ConditionalLogBinomial` (`layers/dist_layers.py:72-121`), expectation over the bin centres.

All 1x1 convs run on the tcgen05 engine; the elementwise stages are the kernels of csrc/zoe_head.cu.  Inputs are what `MidasCore.forward`
returns (`base_models/midas.py:258-276`): the relative depth and six feature maps.  The DPT-BEiT-L encoder itself (torch.hub MiDaS, not vendored)
is NOT built yet (row B3), so this head is exercised with externally supplied features; parameters use the reference's names.
"""
import ctypes as C
import math

import torch

from .. import engine as E
from .._lib import check, lib, ptr, stream

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
        self.clb = _MLP(sd, "conditional_log_binomial.mlp", dev, last_act='softplus', cin_pad=176)

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
