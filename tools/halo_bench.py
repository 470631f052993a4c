"""A/B timing of the halo-tile conv kernel (csrc/tc_halo.cu) against the per-tap kernel / the FFMA depthwise kernel on the shapes it replaces.
`python tools/halo_bench.py [out.json]`"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E   # noqa: E402
from cartoonsegmentation_b200._lib import lib      # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


DENSE = [  # name, N, H, W, Cin, Cout, k, pad, dil, act, f32
    ("isnet 64->64 @360 x20", 20, 360, 360, 64, 64, 3, 1, 1, 'relu', False),
    ("isnet 128->64 @360 x20", 20, 360, 360, 128, 64, 3, 1, 1, 'relu', False),
    ("isnet 64->16 @360 x20", 20, 360, 360, 64, 16, 3, 1, 1, 'relu', False),
    ("isnet 64->32 @360 x20", 20, 360, 360, 64, 32, 3, 1, 1, 'relu', False),
    ("isnet 64->1 @360 x20 f32", 20, 360, 360, 64, 1, 3, 1, 1, None, True),
    ("isnet 256->64 @180 x20", 20, 180, 180, 256, 64, 3, 1, 1, 'relu', False),
    ("isnet 64->128 @180 x20", 20, 180, 180, 64, 128, 3, 1, 1, 'relu', False),
    ("isnet 512->128 @90 x20", 20, 90, 90, 512, 128, 3, 1, 1, 'relu', False),
    ("leres 128->1 @320 x32 f32", 32, 320, 320, 128, 1, 3, 1, 1, None, True),
    ("neck 128->128 silu @128 x32", 32, 128, 128, 128, 128, 3, 1, 1, 'silu', False),
    ("inpaint 64->64 prelu @512 x1", 1, 512, 512, 64, 64, 3, 1, 1, 'prelu', False),
]
GROUPED = [  # name, N, H, W, C, groups
    ("leres g32 1024 @40 x32", 32, 40, 40, 1024, 32),
    ("leres g32 256 @160 x32", 32, 160, 160, 256, 32),
    ("leres g32 512 @80 x32", 32, 80, 80, 512, 32),
    ("leres g32 2048 @20 x32", 32, 20, 20, 2048, 32),
]
DW = [  # name, N, H, W, C, k
    ("convnext dw7 128 @256 x32", 32, 256, 256, 128, 7),
    ("convnext dw7 256 @128 x32", 32, 128, 128, 256, 7),
    ("convnext dw7 512 @64 x32", 32, 64, 64, 512, 7),
    ("convnext dw7 1024 @32 x32", 32, 32, 32, 1024, 7),
    ("neck dw5 128 silu @128 x32", 32, 128, 128, 128, 5),
]


def main():
    dev = torch.device('cuda')
    rows = []
    for name, N, H, W, Cin, Cout, k, pad, dil, act, f32 in DENSE:
        x = torch.randn(N, H, W, Cin, device=dev).half()
        w = E.pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5)
        b = torch.zeros(Cout, device=dev)
        out = torch.empty((N, H, W, Cout), device=dev, dtype=torch.float32 if f32 else torch.float16)
        t = {}
        for mode in (0, 1):
            lib().csb_conv_set_halo_mode(mode)
            t[mode] = timeit(lambda: E.conv2d_nhwc(x, w, b, pad=pad, dil=dil, act=act, out=out))
        lib().csb_conv_set_halo_mode(1)
        fl = 2.0 * N * H * W * Cout * k * k * Cin
        rows.append(dict(shape=name, ms_old=t[0], ms_halo=t[1], tflops_old=fl / t[0] / 1e9, tflops_halo=fl / t[1] / 1e9))
        print(f"{name:32s} per-tap {t[0]:7.3f} ms {fl / t[0] / 1e9:7.1f} TF/s | halo {t[1]:7.3f} ms {fl / t[1] / 1e9:7.1f} TF/s | x{t[0] / t[1]:.2f}", flush=True)
    for name, N, H, W, Cc, groups in GROUPED:
        cpg = Cc // groups
        x = torch.randn(N, H, W, Cc, device=dev).half()
        wt = torch.randn(Cc, cpg, 3, 3, device=dev) / (cpg * 9) ** 0.5
        w64, wc = E.pack_grouped_weight(wt, groups), E.pack_grouped_weight_compact(wt, groups)
        b = torch.zeros(Cc, device=dev)
        out = torch.empty((N, H, W, Cc), device=dev, dtype=torch.float16)
        t0 = timeit(lambda: E.conv2d_nhwc(x, w64, b, pad=1, act='relu', groups=groups, out=out))
        t1 = timeit(lambda: E.conv2d_halo_nhwc(x, wc, b, pad=1, act='relu', groups=groups, out=out))
        fl = 2.0 * N * H * W * Cc * 9 * cpg
        gb = 4.0 * N * H * W * Cc
        rows.append(dict(shape=name, ms_old=t0, ms_halo=t1, tflops_old=fl / t0 / 1e9, tflops_halo=fl / t1 / 1e9, gbs_halo=gb / t1 / 1e6))
        print(f"{name:32s} 64-slice {t0:7.3f} ms {fl / t0 / 1e9:7.1f} TF/s | halo {t1:7.3f} ms {fl / t1 / 1e9:7.1f} TF/s {gb / t1 / 1e6:6.0f} GB/s | x{t0 / t1:.2f}", flush=True)
    for name, N, H, W, Cc, k in DW:
        x = torch.randn(N, H, W, Cc, device=dev).half()
        dw = (torch.randn(k, k, Cc, device=dev) / k).contiguous()
        wc = E.pack_dw_weight_compact(dw)
        b = torch.zeros(Cc, device=dev)
        out = torch.empty((N, H, W, Cc), device=dev, dtype=torch.float16)
        if k == 7:
            t0 = timeit(lambda: E.dwconv_stats_nhwc(x, dw, b, out=out))
            t1 = timeit(lambda: E.conv2d_halo_nhwc(x, wc, b, pad=3, groups=Cc, out=out, stats=True))
        else:
            t0 = timeit(lambda: E.dwconv_nhwc(x, dw, b, act='silu', out=out))
            t1 = timeit(lambda: E.conv2d_halo_nhwc(x, wc, b, pad=k // 2, act='silu', groups=Cc, out=out))
        gout = N * H * W * Cc / 1e9
        gb = 4.0 * N * H * W * Cc
        rows.append(dict(shape=name, ms_old=t0, ms_halo=t1, gout_s_old=gout / t0 * 1e3, gout_s_halo=gout / t1 * 1e3, gbs_halo=gb / t1 / 1e6))
        print(f"{name:32s} ffma {t0:7.3f} ms {gout / t0 * 1e3:6.1f} Gout/s | halo-tc {t1:7.3f} ms {gout / t1 * 1e3:6.1f} Gout/s {gb / t1 / 1e6:6.0f} GB/s | x{t0 / t1:.2f}", flush=True)
    json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/halo_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
