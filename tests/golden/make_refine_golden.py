"""Golden outputs of the UNMODIFIED reference `Refine` module (anime_3dkenburns/models/disparity_refinement.py:84-135), loaded by path in the build
container (where /root/reference exists) with the product's seeded synthetic state_dict.  Run: python tests/golden/make_refine_golden.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get("CSB_REFERENCE", "/root/reference")
CASES = [((96, 128), (96, 128)), ((96, 128), (24, 32)), ((100, 132), (100, 132))]      # same-res (kenburns_effect.py:620), quarter-res (depthestim.py:68), odd sizes


def inputs(hw, dhw, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(hw[0], dtype=torch.float32), torch.arange(hw[1], dtype=torch.float32), indexing="ij")
    img = torch.stack([0.5 + 0.4 * torch.sin(xx / 9.0 + c) * torch.cos(yy / 7.0 - c) for c in range(3)])[None] + 0.05 * torch.randn(1, 3, *hw, generator=g)
    yy, xx = torch.meshgrid(torch.arange(dhw[0], dtype=torch.float32), torch.arange(dhw[1], dtype=torch.float32), indexing="ij")
    disp = (20 + 10 * torch.sin(xx / (dhw[1] / 6.0)) * torch.cos(yy / (dhw[0] / 4.0)))[None, None] + 0.3 * torch.randn(1, 1, *dhw, generator=g)
    return img.clamp(0, 1).contiguous(), disp.contiguous()


def main():
    from cartoonsegmentation_b200.anime_3dkenburns.models.disparity_refinement import synthetic_state_dict
    spec = importlib.util.spec_from_file_location("ref_disparity_refinement", os.path.join(REF, "anime_3dkenburns", "models", "disparity_refinement.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    net = mod.Refine().eval()
    net.load_state_dict(synthetic_state_dict(0), strict=True)
    out = {}
    for i, (hw, dhw) in enumerate(CASES):
        img, disp = inputs(hw, dhw, 40 + i)
        with torch.no_grad():
            y = net(img, disp)
        out[f"out{i}"] = y[0, 0].numpy().astype(np.float32)
        print(i, hw, dhw, y.shape, float(y.min()), float(y.max()))
    np.savez_compressed(os.path.join(HERE, "refine_ref.npz"), **out)


if __name__ == "__main__":
    main()
