"""Generates tests/golden/inpaint_ref_*.npz with the UNMODIFIED reference Inpaint network (anime_3dkenburns/models/pointcloud_inpainting.py, loaded by path
from /root/reference; its `.utils` import is served by the CPU oracle's depth_to_points / spatial_filter / render_pointcloud) and the
seeded synthetic weights of cartoonsegmentation_b200.anime_3dkenburns.models.pointcloud_inpainting.  Run in the build container."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("CSB_REFERENCE", "/root/reference")


def load_reference_inpaint():
    from oracle import kb_oracle as orc
    pkg = types.ModuleType("refkbm")
    pkg.__path__ = [os.path.join(REF, "anime_3dkenburns", "models")]
    sys.modules["refkbm"] = pkg
    u = types.ModuleType("refkbm.utils")
    u.depth_to_points = lambda d, f: torch.from_numpy(orc.depth_to_points(d.numpy(), f))
    u.spatial_filter = lambda x, t: torch.from_numpy(orc.spatial_filter(x.numpy(), t))

    def render(p, d, W, H, f, b):
        r, e = orc.render_pointcloud(p.detach().numpy(), d.detach().numpy(), W, H, f, b)
        return torch.from_numpy(r), torch.from_numpy(e)
    u.render_pointcloud = render
    sys.modules["refkbm.utils"] = u
    spec = importlib.util.spec_from_file_location("refkbm.pointcloud_inpainting", os.path.join(pkg.__path__[0], "pointcloud_inpainting.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["refkbm.pointcloud_inpainting"] = m
    spec.loader.exec_module(m)
    return m


def main():
    from cartoonsegmentation_b200.anime_3dkenburns.models import pointcloud_inpainting as P
    from cartoonsegmentation_b200.utils.synthetic import smooth_disparity, smooth_image
    net = load_reference_inpaint().Inpaint().eval()
    net.load_state_dict(P.synthetic_state_dict(0), strict=True)
    for (H, W, seed) in [(96, 128, 0)]:
        img = smooth_image(H, W, seed=80 + seed)
        disp = smooth_disparity(H, W, seed=81 + seed)
        disp = disp / disp.max() * 40.0
        tenImage = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)[None].astype(np.float32) * (1.0 / 255.0)))
        shift = torch.tensor([30.0, -20.0, -60.0]).view(1, 3, 1)
        common = {'fltFocal': 512.0, 'fltBaseline': 40.0, 'intWidth': W, 'intHeight': H}
        with torch.no_grad():
            o = net(tenImage, torch.from_numpy(disp), shift, common, None)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"inpaint_ref_{H}x{W}.npz"), image=img, disparity=disp, shift=shift.numpy().reshape(3),
                            tenExisting=o['tenExisting'].numpy(), tenImage=o['tenImage'].numpy(), tenDisparity=o['tenDisparity'].numpy())
        print(H, W, "existing frac", o['tenExisting'].mean().item(), "image mean/std", o['tenImage'].mean().item(), o['tenImage'].std().item(),
              "disp mean/std", o['tenDisparity'].mean().item(), o['tenDisparity'].std().item())


if __name__ == "__main__":
    main()
