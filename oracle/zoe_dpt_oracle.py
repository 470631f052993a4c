"""TEST INFRASTRUCTURE ONLY -- fp32 torch restatement of the MiDaS v3.1 `DPT_BEiT_L_384` forward that the reference loads through torch.hub
(depth_modules/zoedepth/models/base_models/midas.py:341; hooks :289-311; returned tensors :258-276).

Not in /root/reference (third-party: intel-isl/MiDaS v3.1 `midas/dpt_depth.py`, `midas/blocks.py`, `midas/backbones/beit.py` on top of
timm==0.6.7 `models/beit.py`, pinned by conda_env.yaml:346) -- restated from the published code:
  * BEiT block:  x += gamma_1 * proj(softmax(q k^T / sqrt(d) + rel_pos_bias) v),  x += gamma_2 * fc2(gelu(fc1(norm2(x)))), LayerNorm eps 1e-6,
    qkv bias = (q_bias, 0, v_bias); per-block bias table, bilinearly resized to the actual window (MiDaS `_get_rel_pos_bias`)
  * reassemble:  ProjectReadout (Linear(2D, D) + GELU on [token | cls]), 1x1 conv, ConvTranspose 4 / 2 / identity / 3x3 stride-2 conv
  * scratch:     layerN_rn 3x3 (no bias), FeatureFusionBlock_custom x4 (ResidualConvUnit_custom, x2 bilinear align_corners=True, 1x1 out_conv),
                 output_conv = conv3x3(256,128), x2 bilinear (align_corners=True), conv3x3(128,32), ReLU, conv1x1(32,1), ReLU
Pinned against transformers' independent port of the same network (tests/golden/make_zoe_dpt_golden.py; tests/test_oracle_nets_cpu.py).
Parameters: the reference checkpoint's names below `core.core.`.
"""
import torch
import torch.nn.functional as F

from cartoonsegmentation_b200.depth_modules.zoedepth import BEIT, relative_position_bias


def forward(sd, x, eps=1e-6):
    """x [B,3,Hn,Wn] fp32 (already normalised) -> dict(rel [B,Hn,Wn], outconv [B,32,Hn,Wn], btl, fused [r4,r3,r2,r1], tokens [4 x B,T,D])"""
    B, _, Hn, Wn = x.shape
    hp, wp = Hn // 16, Wn // 16
    D, heads = BEIT['dim'], BEIT['heads']
    p = lambda k: sd[k].float()
    t = F.conv2d(x, p("pretrained.model.patch_embed.proj.weight"), p("pretrained.model.patch_embed.proj.bias"), stride=16).flatten(2).transpose(1, 2)
    t = torch.cat([p("pretrained.model.cls_token").expand(B, -1, -1), t], 1)
    T = t.shape[1]
    hooked = []
    for i in range(BEIT['depth']):
        b = f"pretrained.model.blocks.{i}"
        h = F.layer_norm(t, (D,), p(f"{b}.norm1.weight"), p(f"{b}.norm1.bias"), eps)
        qkv_bias = torch.cat([p(f"{b}.attn.q_bias"), torch.zeros(D), p(f"{b}.attn.v_bias")])
        qkv = F.linear(h, p(f"{b}.attn.qkv.weight"), qkv_bias).reshape(B, T, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (D // heads) ** -0.5, qkv[1], qkv[2]
        att = q @ k.transpose(-2, -1) + relative_position_bias(p(f"{b}.attn.relative_position_bias_table"), (BEIT['window'],) * 2, (hp, wp))[None]
        a = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, T, D)
        t = t + p(f"{b}.gamma_1") * F.linear(a, p(f"{b}.attn.proj.weight"), p(f"{b}.attn.proj.bias"))
        h = F.layer_norm(t, (D,), p(f"{b}.norm2.weight"), p(f"{b}.norm2.bias"), eps)
        h = F.linear(F.gelu(F.linear(h, p(f"{b}.mlp.fc1.weight"), p(f"{b}.mlp.fc1.bias"))), p(f"{b}.mlp.fc2.weight"), p(f"{b}.mlp.fc2.bias"))
        t = t + p(f"{b}.gamma_2") * h
        if i in BEIT['hooks']:
            hooked.append(t)
    feats = []
    for k, tok in enumerate(hooked, 1):
        a = f"pretrained.act_postprocess{k}"
        cat = torch.cat([tok[:, 1:], tok[:, :1].expand(-1, T - 1, -1)], -1)
        f = F.gelu(F.linear(cat, p(f"{a}.0.project.0.weight"), p(f"{a}.0.project.0.bias"))).transpose(1, 2).reshape(B, D, hp, wp)
        f = F.conv2d(f, p(f"{a}.3.weight"), p(f"{a}.3.bias"))
        if k == 1:
            f = F.conv_transpose2d(f, p(f"{a}.4.weight"), p(f"{a}.4.bias"), stride=4)
        elif k == 2:
            f = F.conv_transpose2d(f, p(f"{a}.4.weight"), p(f"{a}.4.bias"), stride=2)
        elif k == 4:
            f = F.conv2d(f, p(f"{a}.4.weight"), p(f"{a}.4.bias"), stride=2, padding=1)
        feats.append(F.conv2d(f, p(f"scratch.layer{k}_rn.weight"), None, padding=1))

    def rcu(x_, name):
        o = F.conv2d(F.relu(x_), p(f"{name}.conv1.weight"), p(f"{name}.conv1.bias"), padding=1)
        o = F.conv2d(F.relu(o), p(f"{name}.conv2.weight"), p(f"{name}.conv2.bias"), padding=1)
        return o + x_

    def fusion(k, path, skip=None):
        r = f"scratch.refinenet{k}"
        o = path if skip is None else path + rcu(skip, f"{r}.resConfUnit1")
        o = rcu(o, f"{r}.resConfUnit2")
        o = F.interpolate(o, scale_factor=2, mode="bilinear", align_corners=True)
        return F.conv2d(o, p(f"{r}.out_conv.weight"), p(f"{r}.out_conv.bias"))
    l1, l2, l3, l4 = feats
    p4 = fusion(4, l4)
    p3 = fusion(3, p4, l3)
    p2 = fusion(2, p3, l2)
    p1 = fusion(1, p2, l1)
    o = F.conv2d(p1, p("scratch.output_conv.0.weight"), p("scratch.output_conv.0.bias"), padding=1)
    o = F.interpolate(o, scale_factor=2, mode="bilinear", align_corners=True)
    outconv = F.relu(F.conv2d(o, p("scratch.output_conv.2.weight"), p("scratch.output_conv.2.bias"), padding=1))
    rel = F.relu(F.conv2d(outconv, p("scratch.output_conv.4.weight"), p("scratch.output_conv.4.bias")))[:, 0]
    return dict(rel=rel, outconv=outconv, btl=l4, fused=[p4, p3, p2, p1], tokens=hooked)
