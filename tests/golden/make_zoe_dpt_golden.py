"""Golden vectors for the DPT-BEiT-L encoder of ZoeDepth.

The reference obtains this network from torch.hub (`intel-isl/MiDaS`, `DPT_BEiT_L_384`, base_models/midas.py:341); neither the hub code nor timm is in
/root/reference or in this image, so the product restates it.  The independent implementation available offline is transformers' port
(`ZoeDepthForDepthEstimation`: BeitBackbone + ZoeDepthNeck + ZoeDepthRelativeDepthEstimationHead), which was validated by its authors against the
original checkpoints.  This script instantiates it with the DPT_BEiT_L_384 geometry, fills every parameter from a seeded generator (the default
init leaves the relative-position tables at zero), runs it in fp32 on the CPU and stores sub-sampled outputs.  The GPU test rebuilds the same
parameters from the same seed (`build_hf_model`), converts them to the reference checkpoint's (MiDaS/timm) names with `hf_to_midas` and feeds them to
the product.  Run:  python tests/golden/make_zoe_dpt_golden.py
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(384, 384), (256, 320)]                 # native 24 x 24 window, and a 16 x 20 window (bias-table interpolation, non-square)


def build_hf_model(seed=1234):
    from transformers import BeitConfig, ZoeDepthConfig, ZoeDepthForDepthEstimation
    bc = BeitConfig(image_size=384, patch_size=16, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                    use_relative_position_bias=True, use_shared_relative_position_bias=False, use_absolute_position_embeddings=False,
                    layer_scale_init_value=0.1, drop_path_rate=0.0, layer_norm_eps=1e-6, out_features=["stage6", "stage12", "stage18", "stage24"],
                    reshape_hidden_states=False, use_mask_token=False)
    cfg = ZoeDepthConfig(backbone_config=bc, neck_hidden_sizes=[256, 512, 1024, 1024], reassemble_factors=[4, 2, 1, 0.5], fusion_hidden_size=256,
                         readout_type="project", num_relative_features=32, use_batch_norm_in_fusion_residual=False, add_projection=False)
    with torch.device("cpu"):
        model = ZoeDepthForDepthEstimation(cfg).eval().float()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith("metric_head"):
                continue
            if "relative_position_bias_table" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif "cls_token" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif "lambda_" in name:
                p.copy_(torch.rand(p.shape, generator=g) * 0.2 + 0.1)
            elif "layernorm" in name and name.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) * 0.4 + 0.8)
            elif p.dim() >= 2:
                is_t = "resize" in name and p.dim() == 4 and ("layers.0" in name or "layers.1" in name)         # ConvTranspose2d [in, out, k, k]
                fan_in = p.shape[0] if is_t else p[0].numel()
                gain = math.sqrt(2.0) if any(k in name for k in ("intermediate", "readout", "convolution", "conv2")) else 1.0
                p.copy_(torch.randn(p.shape, generator=g) * (gain / math.sqrt(fan_in)))
            else:
                p.copy_(torch.rand(p.shape, generator=g) * 0.2 - 0.1)
    return model


def hf_to_midas(sd):
    """transformers parameter names -> the names of the reference checkpoint below `core.core.` (MiDaS DPT + timm BEiT)."""
    out = {}
    g = lambda k: sd[k].detach().clone()
    out["pretrained.model.cls_token"] = g("backbone.embeddings.cls_token")
    out["pretrained.model.patch_embed.proj.weight"] = g("backbone.embeddings.patch_embeddings.projection.weight")
    out["pretrained.model.patch_embed.proj.bias"] = g("backbone.embeddings.patch_embeddings.projection.bias")
    for i in range(24):
        h, m = f"backbone.encoder.layer.{i}", f"pretrained.model.blocks.{i}"
        a = f"{h}.attention.attention"
        out[f"{m}.attn.qkv.weight"] = torch.cat([g(f"{a}.query.weight"), g(f"{a}.key.weight"), g(f"{a}.value.weight")], 0)
        out[f"{m}.attn.q_bias"], out[f"{m}.attn.v_bias"] = g(f"{a}.query.bias"), g(f"{a}.value.bias")
        out[f"{m}.attn.relative_position_bias_table"] = g(f"{a}.relative_position_bias.relative_position_bias_table")
        for src, dst in (("attention.output.dense", "attn.proj"), ("intermediate.dense", "mlp.fc1"), ("output.dense", "mlp.fc2"),
                         ("layernorm_before", "norm1"), ("layernorm_after", "norm2")):
            out[f"{m}.{dst}.weight"], out[f"{m}.{dst}.bias"] = g(f"{h}.{src}.weight"), g(f"{h}.{src}.bias")
        out[f"{m}.gamma_1"], out[f"{m}.gamma_2"] = g(f"{h}.lambda_1"), g(f"{h}.lambda_2")
    for k in range(4):
        a = f"pretrained.act_postprocess{k + 1}"
        out[f"{a}.0.project.0.weight"] = g(f"neck.reassemble_stage.readout_projects.{k}.0.weight")
        out[f"{a}.0.project.0.bias"] = g(f"neck.reassemble_stage.readout_projects.{k}.0.bias")
        out[f"{a}.3.weight"], out[f"{a}.3.bias"] = g(f"neck.reassemble_stage.layers.{k}.projection.weight"), g(f"neck.reassemble_stage.layers.{k}.projection.bias")
        if k != 2:
            out[f"{a}.4.weight"], out[f"{a}.4.bias"] = g(f"neck.reassemble_stage.layers.{k}.resize.weight"), g(f"neck.reassemble_stage.layers.{k}.resize.bias")
        out[f"scratch.layer{k + 1}_rn.weight"] = g(f"neck.convs.{k}.weight")
    for j in range(4):                                       # fusion layer j is applied j-th, i.e. refinenet(4 - j)
        f, r = f"neck.fusion_stage.layers.{j}", f"scratch.refinenet{4 - j}"
        out[f"{r}.out_conv.weight"], out[f"{r}.out_conv.bias"] = g(f"{f}.projection.weight"), g(f"{f}.projection.bias")
        for u in (1, 2):
            for c in (1, 2):
                out[f"{r}.resConfUnit{u}.conv{c}.weight"] = g(f"{f}.residual_layer{u}.convolution{c}.weight")
                out[f"{r}.resConfUnit{u}.conv{c}.bias"] = g(f"{f}.residual_layer{u}.convolution{c}.bias")
    for hf, idx in (("conv1", 0), ("conv2", 2), ("conv3", 4)):
        out[f"scratch.output_conv.{idx}.weight"], out[f"scratch.output_conv.{idx}.bias"] = g(f"relative_head.{hf}.weight"), g(f"relative_head.{hf}.bias")
    return out


def net_input(Hn, Wn, seed=7):
    """a normalised-image-like input: smooth field + noise in about [-1, 1], NCHW fp32"""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(Hn, dtype=torch.float32), torch.arange(Wn, dtype=torch.float32), indexing="ij")
    base = torch.stack([torch.sin(xx / 37.0 + c) * torch.cos(yy / 29.0 - c) for c in range(3)])
    return (0.7 * base + 0.2 * torch.randn(3, Hn, Wn, generator=g))[None].contiguous()


def run_hf(model, x):
    with torch.no_grad():
        hs = model.backbone.forward_with_filtered_kwargs(x).feature_maps
        hp, wp = x.shape[2] // 16, x.shape[3] // 16
        fused, bottleneck = model.neck(list(hs), hp, wp)
        rel, feat = model.relative_head(fused)
    return hs, fused, bottleneck, rel, feat


def main():
    torch.set_num_threads(os.cpu_count())
    model = build_hf_model()
    for Hn, Wn in CASES:
        x = net_input(Hn, Wn)
        hs, fused, btl, rel, feat = run_hf(model, x)
        nhwc = lambda t: t[0].permute(1, 2, 0).contiguous()
        out = dict(rel=rel[0].numpy().astype(np.float32), outconv=nhwc(feat)[::4, ::4].numpy().astype(np.float16), btl=nhwc(btl).numpy().astype(np.float16))
        for k, t in enumerate(hs):
            out[f"tok{k}"] = t[0, ::8].numpy().astype(np.float16)                      # every 8th token of the hooked blocks
        for k, t in enumerate(fused):                                                  # r4, r3, r2, r1
            st = max(1, t.shape[2] // 24)
            out[f"fused{k}"] = nhwc(t)[::st, ::st].numpy().astype(np.float16)
        path = os.path.join(HERE, f"zoe_dpt_ref_{Hn}x{Wn}.npz")
        np.savez_compressed(path, **out)
        print(path, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB", "rel range", float(rel.min()), float(rel.max()))


if __name__ == "__main__":
    sys.exit(main())
