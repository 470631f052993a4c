"""One conv shape for `ncu --set full -k regex:k_conv_tc`: python tools/conv_one.py N H W Cin Cout k act"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

N, H, W, Cin, Cout, k = map(int, sys.argv[1:7])
act = sys.argv[7] if len(sys.argv) > 7 and sys.argv[7] != 'none' else None
x = torch.randn(N, H, W, Cin, device='cuda').half()
w = E.pack_conv_weight(torch.randn(Cout, Cin, k, k, device='cuda') * 0.05)
b = torch.randn(Cout, device='cuda')
for _ in range(3):
    y = E.conv2d_nhwc(x, w, b, pad=k // 2, act=act)
torch.cuda.synchronize()
