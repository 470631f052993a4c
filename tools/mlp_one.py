"""One fused ConvNeXt-MLP shape for `ncu --set full -k regex:k_mlp_tc`: python tools/mlp_one.py [C] [N H W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else ((32, 256, 256) if C == 128 else (32, 128, 128))
Hd = 4 * C
g = torch.Generator(device='cuda').manual_seed(C)
x = torch.randn(N, H, W, C, device='cuda', generator=g).half()
t = torch.randn(N, H, W, C, device='cuda', generator=g).half()
w1 = torch.randn(Hd, C, device='cuda', generator=g) / C ** 0.5
b1 = torch.randn(Hd, device='cuda', generator=g) * 0.1
wf, bf, colsum = E.fold_layernorm(w1, b1, torch.rand(C, device='cuda', generator=g) + 0.5, torch.randn(C, device='cuda', generator=g) * 0.1)
w2p = E.pack_conv_weight((torch.randn(C, Hd, device='cuda', generator=g) / Hd ** 0.5).reshape(C, Hd, 1, 1))
b2 = torch.randn(C, device='cuda', generator=g) * 0.1
stats = torch.rand(N * H * W, C // 64, 2, device='cuda')
stats[..., 1] += 64.0
out = torch.empty_like(x)
for _ in range(3):
    E.convnext_mlp_nhwc(x, stats, wf, bf, colsum, w2p, b2, t, out=out)
torch.cuda.synchronize()
