"""Time the ConvNeXt block's depthwise 7x7 + LayerNorm statistics (csb_dwconv_stats_nhwc) at the detector's stage shapes (batch 32 @1024), halo vs tile."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

shapes = [(32, 256, 256, 128, 3), (32, 128, 128, 256, 3), (32, 64, 64, 512, 27), (32, 32, 32, 1024, 3)]
res = {}
for mode in ('1', '0'):
    os.environ['CSB_DW_HALO'] = mode
    tot = 0.0
    for N, H, W, C, reps in shapes:
        xs = [torch.randn(N, H, W, C, device='cuda').half() for _ in range(3)]         # rotate inputs: > L2
        w = torch.randn(7, 7, C, device='cuda'); b = torch.randn(C, device='cuda')
        for i in range(3):
            E.dwconv_stats_nhwc(xs[i], w, b)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize(); e0.record()
        for i in range(9):
            E.dwconv_stats_nhwc(xs[i % 3], w, b)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 9
        res[f"{'halo' if mode == '1' else 'tile'} {H}x{W}x{C}"] = round(ms, 4)
        tot += ms * reps
    res[f"{'halo' if mode == '1' else 'tile'} detector_total_ms"] = round(tot, 2)
print(json.dumps(res))
