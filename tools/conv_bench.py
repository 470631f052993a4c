"""Micro-benchmark of the tcgen05 conv engine on the detector's dominant shapes, batch 32 @1024x1024 input.
Prints achieved TFLOP/s per shape (CUDA events, 3 warm-up + 10 timed, distinct buffers > L2 cycled)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E   # noqa: E402

SHAPES = [  # name, N, H, W, Cin, Cout, k, stride, pad, act
    ("convnext s2 fc1 512->2048 @64x64", 32, 64, 64, 512, 2048, 1, 1, 0, 'gelu'),
    ("convnext s2 fc2 2048->512 @64x64", 32, 64, 64, 2048, 512, 1, 1, 0, None),
    ("convnext s0 fc1 128->512 @256x256", 32, 256, 256, 128, 512, 1, 1, 0, 'gelu'),
    ("convnext s0 fc2 512->128 @256x256", 32, 256, 256, 512, 128, 1, 1, 0, None),
    ("convnext s1 fc1 256->1024 @128x128", 32, 128, 128, 256, 1024, 1, 1, 0, 'gelu'),
    ("convnext s3 fc1 1024->4096 @32x32", 32, 32, 32, 1024, 4096, 1, 1, 0, 'gelu'),
    ("head 3x3 256->256 @128x128", 32, 128, 128, 256, 256, 3, 1, 1, 'silu'),
    ("head tower0 3x3 256->768 @128x128", 32, 128, 128, 256, 768, 3, 1, 1, 'silu'),
    ("neck 3x3 128->128 @128x128", 32, 128, 128, 128, 128, 3, 1, 1, 'silu'),
    ("neck 3x3 256->256 @64x64", 32, 64, 64, 256, 256, 3, 1, 1, 'silu'),
    ("neck 3x3 512->512 @32x32", 32, 32, 32, 512, 512, 3, 1, 1, 'silu'),
    ("neck 1x1 1024->512 @64x64", 32, 64, 64, 1024, 512, 1, 1, 0, 'silu'),
    ("stem 4x4s4 16->128 @256x256", 32, 1024, 1024, 16, 128, 4, 4, 0, None),
    ("down 2x2s2 128->256 @128x128", 32, 256, 256, 128, 256, 2, 2, 0, None),
    ("inpaint 3x3 32->32 @1024x1024", 1, 1024, 1024, 32, 32, 3, 1, 1, 'prelu'),
    ("inpaint 3x3 64->64 @512x512", 1, 512, 512, 64, 64, 3, 1, 1, 'prelu'),
]


def main():
    dev = torch.device('cuda')
    rows = []
    for name, N, H, W, Cin, Cout, k, stride, pad, act in SHAPES:
        nbuf = max(2, int(300e6 // max(1, N * H * W * Cin * 2)) + 1)
        nbuf = min(nbuf, 8)
        xs = [torch.randn(N, H, W, Cin, device=dev).half() for _ in range(nbuf)]
        w = E.pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5)
        b = torch.zeros(Cout, device=dev)
        slope = torch.full((Cout,), 0.25, device=dev) if act == 'prelu' else None
        Ho, Wo = E.out_hw(H, W, k, k, stride, pad, 1)
        out = torch.empty((N, Ho, Wo, Cout), device=dev, dtype=torch.float16)
        for i in range(3):
            E.conv2d_nhwc(xs[i % nbuf], w, b, stride=stride, pad=pad, act=act, act_param=slope, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for i in range(iters):
            E.conv2d_nhwc(xs[i % nbuf], w, b, stride=stride, pad=pad, act=act, act_param=slope, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * N * Ho * Wo * Cout * k * k * Cin
        byts = 2.0 * (N * H * W * Cin + N * Ho * Wo * Cout + Cout * k * k * Cin)
        rows.append(dict(shape=name, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 1), gflop=round(flops / 1e9, 2)))
        print(f"{name:42s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s (compulsory)", flush=True)
        del xs, out
    json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
