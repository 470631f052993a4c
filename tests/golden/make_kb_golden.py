"""Generates tests/golden/kb_ref_*.npz by running the UNMODIFIED reference kernels (oracle/_ref/libref_kernels.so, built
from /root/reference by oracle/build_ref_kernels.py) on a GPU.  Run on the B200 box:

    gpurun -- 'python tests/golden/make_kb_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  The CPU suite (tests/test_oracle_cpu.py) checks the oracle
against these files; they are what pins the oracle to the reference on machines without a GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import kb_oracle as orc          # noqa: E402
from tests import ref_kernels as ref         # noqa: E402
from tests.kb_scene import make_scene        # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    for (H, W, extra, seed, tag) in [(128, 160, 3000, 0, "128x160")]:
        s = make_scene(H, W, seed=seed, extra_points=extra)
        c = s['common']
        st = {'tenPoints': s['points'], 'fltShiftU': 10.0, 'fltShiftV': -6.0, 'fltDepthFrom': c['objDepthrange'][0], 'fltDepthTo': c['objDepthrange'][0] * 0.9}
        pts, _ = orc.process_shift(st, c)
        data = s['data']
        r, e, z0, z1 = ref.render_pointcloud(torch.from_numpy(pts).cuda(), torch.from_numpy(data).cuda(), W, H, stages=True)
        f = ref.fill_disocclusion(r.contiguous(), (r[:, 3:4] * (e > 0).float()).contiguous())
        np.savez_compressed(os.path.join(out, f"kb_ref_{tag}.npz"), H=H, W=W, points=pts, data=data,
                            render=r.cpu().numpy(), existing=e.cpu().numpy(), zee_pre=z0.cpu().numpy(), zee_post=z1.cpu().numpy(), filled=f.cpu().numpy())
    print("golden written to", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
