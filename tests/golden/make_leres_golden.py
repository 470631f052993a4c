"""Generates tests/golden/leres_ref_*.npz by running the UNMODIFIED reference LeReS network (depth_modules/leres/leres/Resnext_torch.py +
network_auxi.py, loaded by path from /root/reference) on the CPU with the seeded synthetic weights of
cartoonsegmentation_b200.depth_modules.leres.synthetic_state_dict(0).  Run in the build container (the reference is not on the GPU box):

    python tests/golden/make_leres_golden.py

The GPU test compares the B200 forward against these outputs."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("CSB_REFERENCE", "/root/reference")


def load_reference_leres():
    pkg = types.ModuleType("refleres")
    pkg.__path__ = [os.path.join(REF, "depth_modules", "leres", "leres")]
    sys.modules["refleres"] = pkg
    mods = {}
    for name in ("Resnet", "Resnext_torch", "network_auxi"):
        spec = importlib.util.spec_from_file_location(f"refleres.{name}", os.path.join(pkg.__path__[0], name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refleres.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def main():
    from cartoonsegmentation_b200.depth_modules import leres as L
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    mods = load_reference_leres()
    enc = mods["network_auxi"].resnext101_stride32x8d().eval()            # DepthNet wrapping Resnext_torch.resnext101_32x8d
    dec = mods["network_auxi"].Decoder().eval()
    sd = L.synthetic_state_dict(0)
    enc.load_state_dict({k[len("depth_model.encoder_modules."):]: v for k, v in sd.items() if k.startswith("depth_model.encoder_modules.")}, strict=False)
    missing = [k for k in enc.state_dict() if "num_batches_tracked" not in k and ("depth_model.encoder_modules." + k) not in sd]
    assert not missing, missing[:5]
    dec.load_state_dict({k[len("depth_model.decoder_modules."):]: v for k, v in sd.items() if k.startswith("depth_model.decoder_modules.")}, strict=False)
    missing = [k for k in dec.state_dict() if "num_batches_tracked" not in k and ("depth_model.decoder_modules." + k) not in sd]
    assert not missing, missing[:5]
    out_dir = os.path.join(ROOT, "tests", "golden")
    for (H, W, seed) in [(96, 128, 0), (160, 128, 1)]:
        img = smooth_image(H, W, seed=40 + seed)                           # BGR uint8
        x = img.astype(np.float32) / 255.0                                 # kenburns_effect.py:571
        rgb = x[:, :, ::-1].copy()                                         # depthmap.py:35
        t = torch.from_numpy(rgb.transpose(2, 0, 1))                       # ToTensor on float32: no rescale
        t = (t - torch.tensor(L.IMAGENET_MEAN).view(3, 1, 1)) / torch.tensor(L.IMAGENET_STD).view(3, 1, 1)
        with torch.no_grad():
            feats = enc(t[None])
            y = dec(feats)
        np.savez_compressed(os.path.join(out_dir, f"leres_ref_{H}x{W}.npz"), image=img, depth=y[0, 0].numpy(),
                            feat_rms=np.array([f.pow(2).mean().sqrt().item() for f in feats], np.float32))
        print(H, W, "depth mean/std", y.mean().item(), y.std().item(), "feat rms", [round(f.pow(2).mean().sqrt().item(), 3) for f in feats])


if __name__ == "__main__":
    main()
