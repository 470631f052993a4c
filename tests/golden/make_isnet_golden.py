"""Generates tests/golden/isnet_ref_*.npz with the UNMODIFIED reference ISNetDIS(in_ch=4) (animeinsseg/models/animeseg_refine/isnet.py, loaded by
path from /root/reference) on the CPU with cartoonsegmentation_b200.animeinsseg.isnet.synthetic_state_dict(0).  Run in the build container."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("CSB_REFERENCE", "/root/reference")


def load_reference_isnet():
    spec = importlib.util.spec_from_file_location("ref_isnet", os.path.join(REF, "animeinsseg", "models", "animeseg_refine", "isnet.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_isnet"] = m
    spec.loader.exec_module(m)
    return m


def make_input(H, W, seed):
    from cartoonsegmentation_b200.utils.synthetic import ellipse_masks, smooth_image
    img = smooth_image(H, W, seed=60 + seed)
    mask = ellipse_masks(H, W, k=1, seed=70 + seed)[0]
    x = np.concatenate([img.transpose(2, 0, 1).astype(np.float32) / 255.0, mask[None].astype(np.float32)], 0)      # prepare_refine_batch :37-55
    return img, mask, x


def main():
    from cartoonsegmentation_b200.animeinsseg import isnet as I
    net = load_reference_isnet().ISNetDIS(in_ch=4).eval()
    net.load_state_dict(I.synthetic_state_dict(0), strict=True)
    for (H, W, seed) in [(96, 128, 0), (144, 112, 1)]:
        img, mask, x = make_input(H, W, seed)
        with torch.no_grad():
            d1 = net(torch.from_numpy(x)[None])[0][0][0, 0].numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"isnet_ref_{H}x{W}.npz"), image=img, mask=mask, d1=d1)
        print(H, W, "d1 mean/std", d1.mean(), d1.std())


if __name__ == "__main__":
    main()
