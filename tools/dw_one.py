"""One depthwise 7x7 launch per ConvNeXt stage shape, for `ncu --set full -k regex:k_dwconv_tile`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

for N, H, W, C in ((32, 64, 64, 512), (32, 256, 256, 128)):
    x = torch.randn(N, H, W, C, device='cuda').half()
    w = torch.randn(7, 7, C, device='cuda')
    b = torch.randn(C, device='cuda')
    for _ in range(2):
        E.dwconv_nhwc(x, w, b)
torch.cuda.synchronize()
