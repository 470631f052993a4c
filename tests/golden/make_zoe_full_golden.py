"""Generates tests/golden/zoe_full_ref_*.npz: the COMPOSED ZoeDepth estimator the reference's Ken-Burns pipeline runs
(`load_zoe(..., img_size=[672, 672])`, anime_3dkenburns/kenburns_effect.py:543, then `zoe.infer(x, with_flip_aug=True, pad_input=True)`, :813)
on the CPU in fp32, from these pieces:

  * UNMODIFIED reference code, loaded by path from /root/reference: `DepthModel.infer` (reflect pad, flip twin, bicubic back-resize, crop, mean;
    depth_modules/zoedepth/models/depth_model.py:57-129), `ZoeDepth.forward` + attractor / log-binomial / seed-bin layers (zoedepth_v1.py:124-202,
    layers/*.py), `PrepForMidas` (keep-aspect resize to 672 with multiples of 32, align_corners=True, Normalize 0.5/0.5; midas.py:164-186);
  * the one piece that is NOT in the reference repo -- the torch.hub MiDaS `DPT_BEiT_L_384` network (midas.py:341) -- from its restatement
    oracle/zoe_dpt_oracle.py (pinned at 1e-6 to transformers' independent port, tests/test_oracle_nets_cpu.py), returning the tensors the
    reference's forward hooks capture (midas.py:258-276, :289-311).

Weights: cartoonsegmentation_b200.depth_modules.zoedepth.{dpt_synthetic_state_dict, synthetic_state_dict}(0) = what `ZoeDepth(None)` builds.
Net input is 672 x 672 (1765 tokens) whatever the image size, so a 384 x 384 image exercises the full-size encoder with a small fixture.
Run in the build container (about two minutes of CPU):   python tests/golden/make_zoe_full_golden.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_zoe_golden as mz                                                         # noqa: E402  (reference loader by path)


class OracleMidasCore(nn.Module):
    """What `MidasCore.forward(x, denorm=False, return_rel_depth=True)` returns (midas.py:258-276), with the hub network replaced by the oracle."""
    output_channels = [256] * 5

    def __init__(self, prep, sd_core):
        super().__init__()
        self.prep, self.sd = prep, sd_core

    def forward(self, x, denorm=False, return_rel_depth=True):
        from oracle import zoe_dpt_oracle as DO
        with torch.no_grad():
            o = DO.forward(self.sd, self.prep(x))
        return o['rel'], [o['outconv'], o['btl']] + list(o['fused'])


def build(img_size=(672, 672)):
    from cartoonsegmentation_b200.depth_modules import zoedepth as Z
    ref = mz.load_reference_zoe()
    midas = mz._load("ref_midas_file", os.path.join(mz.ZM, "base_models", "midas.py"))
    prep = midas.PrepForMidas(keep_aspect_ratio=True, img_size=list(img_size))          # config 'infer': force_keep_ar = true
    core = OracleMidasCore(prep, Z.dpt_synthetic_state_dict(0))
    model = ref.ZoeDepth(core, n_bins=64, bin_centers_type="softplus", bin_embedding_dim=128, n_attractors=[16, 8, 4, 1], attractor_alpha=1000,
                         attractor_gamma=2, attractor_kind='mean', attractor_type='inv', min_temp=0.0212, max_temp=50.0, train_midas=False,
                         inverse_midas=False).eval()
    missing, unexpected = model.load_state_dict(Z.synthetic_state_dict(0), strict=False)
    assert not unexpected and all(k.startswith(('conditional_log_binomial.log_binomial_transform', 'core')) for k in missing), (missing, unexpected)
    return model


def main():
    from cartoonsegmentation_b200.utils.synthetic import smooth_image
    torch.set_num_threads(os.cpu_count() or 1)
    model = build()
    for (H, W, seed) in [(384, 384, 5)]:
        img = smooth_image(H, W, seed=seed)                                          # BGR uint8; the reference feeds BGR/255 unswapped (kenburns_effect.py:813)
        x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1)[None]
        with torch.no_grad():
            depth = model.infer(x, with_flip_aug=True, pad_input=True)[0, 0].numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"zoe_full_ref_{H}x{W}.npz"), depth=depth.astype(np.float32), seed=seed)
        print(H, W, "metric depth mean/std/min/max", depth.mean(), depth.std(), depth.min(), depth.max())


if __name__ == "__main__":
    main()
