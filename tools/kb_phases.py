"""Wall-clock phases of ONE full Ken-Burns image (BASELINE configs[3]) with a device synchronize between phases (diagnostic: the syncs themselves
cost a little, so the sum is an upper bound of the unsynchronised call).  `python tools/kb_phases.py [out.json]`."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import _lib                                                       # noqa: E402
from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb                     # noqa: E402
from cartoonsegmentation_b200.anime_3dkenburns.common import process_autozoom                   # noqa: E402
from cartoonsegmentation_b200.utils.synthetic import smooth_image                               # noqa: E402

H = W = 1024
cfg = kb.KenBurnsConfig(det_size=H, max_size=H, depth_est='leres', depth_est_size=640, pred_score_thr=0.3)
pipe = kb.KenBurnsPipeline(cfg)
imgs = [smooth_image(H, W, seed=1234 + i) for i in range(3)]
for i in range(2):
    pipe.autozoom(pipe.generate_kenburns_config(imgs[i]))
torch.cuda.synchronize()
res = {}


class T:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.synchronize(); self.t = time.perf_counter(); self.l = _lib.launch_count()

    def __exit__(self, *a):
        torch.cuda.synchronize()
        res[self.name] = {"ms": round((time.perf_counter() - self.t) * 1e3, 3), "launches": _lib.launch_count() - self.l}


img = imgs[2]
with torch.no_grad():
    with T("total_unsynchronised"):
        frames = pipe.autozoom(pipe.generate_kenburns_config(img))
    with T("seg (AnimeInsSeg.infer)"):
        inst, _ = pipe.run_instance_segmentation(img, scale_down_to_maxsize=False)
    with T("instances.resize"):
        inst.resize(H, W)
    with T("depth + adjust (infer_disparity)"):
        disp = pipe.infer_disparity(img, inst, None, kcfg=pipe.cfg)
    with T("generate_kenburns_config (all of the above + cloud)"):
        kcfg = pipe.generate_kenburns_config(img)
    objFrom = {'fltCenterU': W / 2.0, 'fltCenterV': H / 2.0, 'intCropWidth': int(math.floor(0.97 * W)), 'intCropHeight': int(math.floor(0.97 * H))}
    with T("autozoom search (256 candidates)"):
        objTo = process_autozoom({'fltShift': 100.0, 'fltZoom': 1.25, 'objFrom': objFrom}, kcfg)
    st = {'fltSteps': np.linspace(0.0, 1.0, kcfg.num_frame).tolist(), 'objFrom': objFrom, 'objTo': objTo, 'boolInpaint': True}
    with T("process_kenburns (2 x inpaint + 75 frames + D2H)"):
        pipe.process_kenburns(st, kcfg, True)
    with T("process_kenburns without inpaint (75 frames + D2H)"):
        pipe.process_kenburns(st, kcfg, False)
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    open(sys.argv[1], 'w').write(json.dumps(res, indent=1) + "\n")
