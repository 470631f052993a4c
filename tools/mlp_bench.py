"""Fused ConvNeXt MLP (k_mlp_tc) vs the two-launch path (csb_conv2d_ln_nhwc + csb_conv2d_nhwc) at the two ConvNeXt-B stage shapes, batch 32 @1024^2.
python tools/mlp_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E                            # noqa: E402

for (N, H, W, C) in ((32, 256, 256, 128), (32, 128, 128, 256)):
    Hd = 4 * C
    g = torch.Generator(device='cuda').manual_seed(C)
    xs = [(torch.randn(N, H, W, C, device='cuda', generator=g)).half() for _ in range(2)]
    ts = [torch.randn(N, H, W, C, device='cuda', generator=g).half() for _ in range(2)]
    w1 = torch.randn(Hd, C, device='cuda', generator=g) / C ** 0.5
    b1 = torch.randn(Hd, device='cuda', generator=g) * 0.1
    gamma, beta = torch.rand(C, device='cuda', generator=g) + 0.5, torch.randn(C, device='cuda', generator=g) * 0.1
    w2 = torch.randn(C, Hd, device='cuda', generator=g) / Hd ** 0.5
    b2 = torch.randn(C, device='cuda', generator=g) * 0.1
    wf, bf, colsum = E.fold_layernorm(w1, b1, gamma, beta)
    w2p = E.pack_conv_weight(w2.reshape(C, Hd, 1, 1))
    stats = torch.rand(N * H * W, C // 64, 2, device='cuda')
    stats[..., 1] += 64.0
    h = torch.empty(N, H, W, Hd, device='cuda', dtype=torch.float16)
    outs = [torch.empty_like(xs[0]) for _ in range(2)]

    def two(i):
        E.conv2d_ln_nhwc(xs[i], stats, wf, bf, colsum, act='gelu', out=h)
        E.conv2d_nhwc(h, w2p, b2, residual=ts[i], res_mode=2, out=outs[i])

    def one(i):
        E.convnext_mlp_nhwc(xs[i], stats, wf, bf, colsum, w2p, b2, ts[i], out=outs[i])

    res = {}
    for name, fn in (("two-launch", two), ("fused", one)):
        for i in range(3):
            fn(i % 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            fn(i % 2)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
    fl = 2 * 2 * N * H * W * C * Hd / 1e9
    print(f"{N}x{H}x{W} C={C}: two-launch {res['two-launch'] * 1e3:.1f} us ({fl / res['two-launch']:.0f} TFLOP/s)   fused {res['fused'] * 1e3:.1f} us ({fl / res['fused']:.0f} TFLOP/s)"
          f"   x{res['two-launch'] / res['fused']:.2f}")
