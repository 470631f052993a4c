"""Host-side (cProfile) view of one full Ken-Burns image (BASELINE configs[3]) + wall/device split: where the Python/ctypes time goes.
`python tools/kb_hostprof.py`"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb                     # noqa: E402
from cartoonsegmentation_b200.utils.synthetic import smooth_image                               # noqa: E402

H = W = 1024
cfg = kb.KenBurnsConfig(det_size=H, max_size=H, depth_est='leres', depth_est_size=640, pred_score_thr=0.3, refine_crf=False)
pipe = kb.KenBurnsPipeline(cfg)
imgs = [smooth_image(H, W, seed=1234 + i) for i in range(4)]


def one(i):
    return pipe.autozoom(pipe.generate_kenburns_config(imgs[i % 4]))


for i in range(2):
    one(i)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter()
    fr = one(2 + i)
    torch.cuda.synchronize()
    print(f"image {i}: wall {1e3 * (time.perf_counter() - t0):.1f} ms, {len(fr)} frames", flush=True)
    del fr
pr = cProfile.Profile()
pr.enable()
fr = one(1)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
