"""A few depthwise 7x7 + LayerNorm-statistics launches (the ConvNeXt block's first kernel) for `ncu --set full -k regex:k_dwconv`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

shapes = ((16, 64, 64, 512), (4, 256, 256, 128)) if len(sys.argv) < 2 else (tuple(int(v) for v in sys.argv[1:5]),)
for N, H, W, C in shapes:
    x = torch.randn(N, H, W, C, device='cuda').half()
    w = torch.randn(7, 7, C, device='cuda')
    b = torch.randn(C, device='cuda')
    for _ in range(3):
        E.dwconv_stats_nhwc(x, w, b)
torch.cuda.synchronize()
print("done")
