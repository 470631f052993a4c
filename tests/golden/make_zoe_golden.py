"""Generates tests/golden/zoe_head_ref_*.npz with the UNMODIFIED reference ZoeDepth head (depth_modules/zoedepth/models/zoedepth/zoedepth_v1.py + layers/*,
loaded by path from /root/reference) driven by a fake core that returns seeded feature maps instead of the torch.hub MiDaS encoder, with
cartoonsegmentation_b200.depth_modules.zoedepth.synthetic_state_dict(0).  Run in the build container."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("CSB_REFERENCE", "/root/reference")
ZM = os.path.join(REF, "depth_modules", "zoedepth", "models")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference_zoe():
    for pkg, path in (("depth_modules", None), ("depth_modules.zoedepth", None), ("depth_modules.zoedepth.models", ZM), ("depth_modules.zoedepth.models.layers", os.path.join(ZM, "layers")),
                      ("depth_modules.zoedepth.models.base_models", None), ("depth_modules.zoedepth.models.zoedepth", None)):
        m = types.ModuleType(pkg)
        m.__path__ = [path] if path else []
        sys.modules[pkg] = m
    _load("depth_modules.zoedepth.models.depth_model", os.path.join(ZM, "depth_model.py"))
    midas = types.ModuleType("depth_modules.zoedepth.models.base_models.midas")
    midas.MidasCore = object
    sys.modules["depth_modules.zoedepth.models.base_models.midas"] = midas
    mio = types.ModuleType("depth_modules.zoedepth.models.model_io")
    mio.load_state_from_resource = lambda *a, **k: None
    sys.modules["depth_modules.zoedepth.models.model_io"] = mio
    for n in ("attractor", "dist_layers", "localbins_layers"):
        _load(f"depth_modules.zoedepth.models.layers.{n}", os.path.join(ZM, "layers", n + ".py"))
    return _load("depth_modules.zoedepth.models.zoedepth.zoedepth_v1", os.path.join(ZM, "zoedepth", "zoedepth_v1.py"))


class FakeCore(nn.Module):
    output_channels = [256] * 5

    def __init__(self, rel, feats):
        super().__init__()
        self.rel, self.feats = rel, feats

    def forward(self, x, denorm=False, return_rel_depth=True):
        return self.rel, self.feats


def make_features(H, W, seed):
    g = torch.Generator().manual_seed(seed)

    def smooth(c, h, w, scale):
        t = torch.randn(1, c, max(2, h // 4), max(2, w // 4), generator=g)
        return (torch.nn.functional.interpolate(t, size=(h, w), mode='bilinear', align_corners=True) * scale).half().float()
    rel = torch.rand(1, H, W, generator=g) * 5 + 1
    outconv = smooth(32, H, W, 1.0).relu()
    sizes = [(H // 32, W // 32), (H // 16, W // 16), (H // 8, W // 8), (H // 4, W // 4), (H // 2, W // 2)]
    feats = [outconv] + [smooth(256, h, w, 1.0) for (h, w) in sizes]
    return rel, feats


def main():
    from cartoonsegmentation_b200.depth_modules import zoedepth as Z
    ref = load_reference_zoe()
    for (H, W, seed) in [(64, 96, 0)]:
        rel, feats = make_features(H, W, seed)
        model = ref.ZoeDepth(FakeCore(rel, feats), n_bins=64, bin_centers_type="softplus", bin_embedding_dim=128, n_attractors=[16, 8, 4, 1], attractor_alpha=1000,
                             attractor_gamma=2, attractor_kind='mean', attractor_type='inv', min_temp=0.0212, max_temp=50.0, train_midas=False, inverse_midas=False).eval()
        sd = Z.synthetic_state_dict(0)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith(('conditional_log_binomial.log_binomial_transform', 'core')) for k in missing), (missing, unexpected)
        with torch.no_grad():
            out = model(torch.zeros(1, 3, H, W))['metric_depth'][0, 0].numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"zoe_head_ref_{H}x{W}.npz"), rel=rel.numpy(), metric_depth=out, seed=seed,
                            **{f"feat{i}": f.numpy().astype(np.float16) for i, f in enumerate(feats)})
        print(H, W, "metric depth mean/std/min/max", out.mean(), out.std(), out.min(), out.max())


if __name__ == "__main__":
    main()
