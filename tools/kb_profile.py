"""Per-kernel device-time breakdown of ONE full Ken-Burns image (BASELINE configs[3]: generate_kenburns_config + autozoom @1024^2) with the library's
event profiler (csb_profile_begin/end) + wall clock of the same call without profiling.  `python tools/kb_profile.py [leres|zoe] [out.json]`."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import _lib                                                       # noqa: E402
from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb                     # noqa: E402
from cartoonsegmentation_b200.utils.synthetic import smooth_image                               # noqa: E402

depth = sys.argv[1] if len(sys.argv) > 1 else 'leres'
H = W = 1024
cfg = kb.KenBurnsConfig(det_size=H, max_size=H, depth_est=depth, depth_est_size=640, pred_score_thr=0.3)
pipe = kb.KenBurnsPipeline(cfg)
imgs = [smooth_image(H, W, seed=1234 + i) for i in range(3)]


def one(i):
    return pipe.autozoom(pipe.generate_kenburns_config(imgs[i % 3]))


one(0)
torch.cuda.synchronize()
t0 = time.perf_counter(); l0 = _lib.launch_count()
frames = one(1)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
launches = _lib.launch_count() - l0
lib = _lib.lib()
lib.csb_profile_begin(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
one(2)
buf = ctypes.create_string_buffer(1 << 16)
lib.csb_profile_end(buf, len(buf))
prof = json.loads(buf.value.decode())
tot = sum(v['ms'] for v in prof.values())
out = {"workload": f"kenburns_full ({depth}) @1024^2, 1 image -> {len(frames)} frames", "wall_ms": round(wall, 2), "launches": launches, "profiled_device_ms": round(tot, 2),
       "per_kernel": {k: {"ms": round(v['ms'], 3), "count": v['count']} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}}
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(s + "\n")
