"""Time csb_dwconv_nhwc (+LayerNorm) at the ConvNeXt-B detector's stage shapes (batch 32 @1024).  Run on the GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

shapes = [(32, 256, 256, 128, 3), (32, 128, 128, 256, 3), (32, 64, 64, 512, 27), (32, 32, 32, 1024, 3)]
dev = torch.device('cuda')
tot = 0.0
out = {}
for N, H, W, C, reps in shapes:
    x = torch.randn(N, H, W, C, device=dev).half()
    w = torch.randn(7, 7, C, device=dev)
    b = torch.randn(C, device=dev)
    g, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for ln in (False, True):
        for _ in range(3):
            E.dwconv_nhwc(x, w, b, ln=(g, be) if ln else None)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            E.dwconv_nhwc(x, w, b, ln=(g, be) if ln else None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"{H}x{W}x{C}{'+ln' if ln else ''}"] = round(ms, 3)
        if ln:
            tot += ms * reps
out['detector_total_ms'] = round(tot, 2)
print(json.dumps(out))
