"""A/B timing of the conv engine's CTA-pair path (tcgen05 cta_group::2) on the hot shapes, batch 32.  `python tools/pair_bench.py [out.json]`"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E   # noqa: E402
from cartoonsegmentation_b200._lib import lib      # noqa: E402

SHAPES = [  # name, N, H, W, Cin, Cout, k, stride, pad, act, residual
    ("fc1 512->2048 gelu @64x64", 32, 64, 64, 512, 2048, 1, 1, 0, 'gelu', False),
    ("fc2 2048->512 +res @64x64", 32, 64, 64, 2048, 512, 1, 1, 0, None, True),
    ("fc1 128->512 gelu @256x256", 32, 256, 256, 128, 512, 1, 1, 0, 'gelu', False),
    ("fc2 512->128 +res @256x256", 32, 256, 256, 512, 128, 1, 1, 0, None, True),
    ("fc1 256->1024 gelu @128x128", 32, 128, 128, 256, 1024, 1, 1, 0, 'gelu', False),
    ("fc2 1024->256 +res @128x128", 32, 128, 128, 1024, 256, 1, 1, 0, None, True),
    ("fc1 1024->4096 gelu @32x32", 32, 32, 32, 1024, 4096, 1, 1, 0, 'gelu', False),
    ("3x3 256->256 silu @128x128", 32, 128, 128, 256, 256, 3, 1, 1, 'silu', False),
    ("3x3 256->768 silu @128x128", 32, 128, 128, 256, 768, 3, 1, 1, 'silu', False),
    ("3x3 128->128 silu @128x128", 32, 128, 128, 128, 128, 3, 1, 1, 'silu', False),
    ("leres 1x1 1024->1024 relu+res @40x40", 32, 40, 40, 1024, 1024, 1, 1, 0, 'relu', True),
    ("leres 3x3 256->256 relu @160x160", 32, 160, 160, 256, 256, 3, 1, 1, 'relu', False),
    ("leres 3x3 256->128 relu @320x320", 32, 320, 320, 256, 128, 3, 1, 1, 'relu', False),
    ("isnet 3x3 64->64 relu @360x360 x20", 20, 360, 360, 64, 64, 3, 1, 1, 'relu', False),
    ("isnet 3x3 128->64 relu @360x360 x20", 20, 360, 360, 128, 64, 3, 1, 1, 'relu', False),
]


def main():
    dev = torch.device('cuda')
    rows = []
    for name, N, H, W, Cin, Cout, k, stride, pad, act, has_res in SHAPES:
        nbuf = min(8, max(2, int(300e6 // max(1, N * H * W * Cin * 2)) + 1))
        xs = [torch.randn(N, H, W, Cin, device=dev).half() for _ in range(nbuf)]
        w = E.pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5)
        b = torch.zeros(Cout, device=dev)
        Ho, Wo = E.out_hw(H, W, k, k, stride, pad, 1)
        out = torch.empty((N, Ho, Wo, Cout), device=dev, dtype=torch.float16)
        res = torch.randn((N, Ho, Wo, Cout), device=dev).half() if has_res else None
        t = {}
        for mode in (0, 2):
            lib().csb_conv_set_pair_mode(mode)
            for i in range(3):
                E.conv2d_nhwc(xs[i % nbuf], w, b, stride=stride, pad=pad, act=act, out=out, residual=res, res_mode=2 if has_res else 0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for i in range(iters):
                E.conv2d_nhwc(xs[i % nbuf], w, b, stride=stride, pad=pad, act=act, out=out, residual=res, res_mode=2 if has_res else 0)
            e1.record()
            torch.cuda.synchronize()
            t[mode] = e0.elapsed_time(e1) / iters
        lib().csb_conv_set_pair_mode(1)
        flops = 2.0 * N * Ho * Wo * Cout * k * k * Cin
        rows.append(dict(shape=name, ms_single=round(t[0], 4), ms_pair=round(t[2], 4), tflops_single=round(flops / t[0] / 1e9, 1), tflops_pair=round(flops / t[2] / 1e9, 1)))
        print(f"{name:40s} single {t[0]:7.3f} ms {flops / t[0] / 1e9:7.1f} TF/s | pair {t[2]:7.3f} ms {flops / t[2] / 1e9:7.1f} TF/s | x{t[0] / t[2]:.2f}", flush=True)
        del xs, out
    json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pair_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
