"""ZoeDepth (DPT-BEiT-L + metric head, flip + pad augmentation) throughput at the BASELINE shape: B images of 1024 x 1024 -> 2B net inputs of 384 x 384.
Usage: python tools/zoe_bench.py [B] [out.json]   (CUDA events, 2 warm-up + 5 timed passes; per-kernel split from the library profiler)"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import _lib                                   # noqa: E402
from cartoonsegmentation_b200.depth_modules.zoedepth import ZoeDepth       # noqa: E402
from cartoonsegmentation_b200.utils.synthetic import smooth_image          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/zoe_bench.json"
net = ZoeDepth(None, 'cuda', img_size=[672, 672])          # the Ken-Burns pipeline's resolution (kenburns_effect.py:543)
imgs = torch.from_numpy(np.stack([smooth_image(1024, 1024, seed=50 + i) for i in range(4)])).cuda().repeat(B // 4, 1, 1, 1).contiguous()
for _ in range(2):
    d = net.infer_batch(imgs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d = net.infer_batch(imgs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
lib = _lib.lib()
lib.csb_profile_begin(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
net.infer_batch(imgs)
buf = ctypes.create_string_buffer(1 << 18)
lib.csb_profile_end(buf, len(buf))
prof = json.loads(buf.value.decode())
tot = sum(v['ms'] for v in prof.values())
# encoder + decoder FLOPs per 672 x 672 net input (2 MAC): 24 blocks x 1765 tokens x 12 D^2 + attention 4 T^2 D; DPT decoder ~ 291 GFLOP
T, D = 1765, 1024
gflop = (24 * (T * 12 * D * D * 2 + 4 * T * T * D) + 291e9) / 1e9
print(f"ZoeDepth.infer_batch: {B} images 1024x1024 (2 x {B} net inputs 672x672): {ms:.2f} ms  = {B / ms * 1e3:.1f} images/s, ~{gflop * 2 * B / ms:.0f} TFLOP/s on ~{gflop:.0f} GFLOP/net input")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:10]:
    print(f"  {k:22s} {v['ms']:9.3f} ms  {100 * v['ms'] / tot:5.1f}%  launches {v['count']}")
print("depth range", float(d.min()), float(d.max()))
json.dump(dict(batch=B, ms=ms, images_per_s=B / ms * 1e3, per_kernel=prof), open(out, "w"), indent=1)
