"""Per-kernel device-time breakdown of the detector forward + post-processing at the BASELINE shape (batch B x 1024x1024),
using the library's event-per-launch profiler.  Usage: python tools/det_profile.py [batch] [out.json]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import _lib                                   # noqa: E402
from cartoonsegmentation_b200.animeinsseg import AnimeInsSeg, rtmdet_postprocess          # noqa: E402
from cartoonsegmentation_b200.utils.synthetic import smooth_image          # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/det_profile.json"
    seg = AnimeInsSeg(None, default_det_size=1024, refine_kwargs={'refine_method': 'none'})
    imgs = torch.from_numpy(np.stack([smooth_image(1024, 1024, seed=100 + i) for i in range(min(B, 8))])).cuda()
    imgs = imgs.repeat((B + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1)[:B].contiguous()
    cfg = seg.model.bbox_head.test_cfg

    def run():
        cls, reg, ker, mf = seg.model.net.forward(imgs)
        return rtmdet_postprocess(cls, reg, ker, mf, (1024, 1024), cfg)
    for _ in range(2):
        o = run()
    torch.cuda.synchronize()
    print("instances per image:", o['num'].tolist()[:8])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"detector forward + post-process, batch {B}: {ms:.2f} ms  ({B / ms * 1e3:.1f} images/s, {1011.9 * B / ms:.1f} TFLOP/s on 1011.9 GFLOP/image)")
    lib = _lib.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.csb_profile_begin(st)
    run()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.csb_profile_end(buf, len(buf))
    prof = json.loads(buf.value.decode())
    tot = sum(v['ms'] for v in prof.values())
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
        print(f"  {k:20s} {v['ms']:9.3f} ms  {100 * v['ms'] / tot:5.1f}%  launches {v['count']}")
    json.dump(dict(batch=B, ms=ms, images_per_s=B / ms * 1e3, per_kernel=prof), open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
