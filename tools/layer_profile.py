"""Per-layer-shape device time of every tcgen05 conv launch in the detector forward (batch B @1024^2) and the LeReS forward (batch B @640^2),
using the library profiler's detailed mode (CSB_PROFILE_DETAIL=1).  Usage: python tools/layer_profile.py [batch] [out.json]"""
import ctypes
import json
import os
import re
import sys

os.environ["CSB_PROFILE_DETAIL"] = "1"
import numpy as np          # noqa: E402
import torch                # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import _lib                                   # noqa: E402
from cartoonsegmentation_b200.animeinsseg import AnimeInsSeg               # noqa: E402
from cartoonsegmentation_b200.depth_modules.leres import LeReS             # noqa: E402
from cartoonsegmentation_b200.utils.synthetic import smooth_image          # noqa: E402

PAT = re.compile(r"(k_conv_tc|k_conv_halo)\[(\d+)x(\d+)x(\d+)x(\d+)->(\d+) k(\d+)x(\d+) s(\d+) d(\d+) g(\d+) act(\d+) res(\d+)\]")


def profile(fn):
    lib = _lib.lib()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.csb_profile_begin(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    fn()
    buf = ctypes.create_string_buffer(1 << 18)
    lib.csb_profile_end(buf, len(buf))
    return json.loads(buf.value.decode())


def table(name, prof):
    rows, tot, conv = [], sum(v['ms'] for v in prof.values()), 0.0
    for k, v in prof.items():
        m = PAT.match(k)
        if not m:
            continue
        N, H, W, Cin, Cout, R, S, st, dil, g, act, res = map(int, m.groups()[1:])
        Ho, Wo = (H + st - 1) // st, (W + st - 1) // st
        gflop = 2.0 * N * Ho * Wo * Cout * (Cin // g) * R * S / 1e9
        byts = 2.0 * N * (H * W * Cin + Ho * Wo * Cout)
        conv += v['ms']
        rows.append(dict(shape=('halo ' if k.startswith('k_conv_halo') else '') + k[k.index('['):], ms=v['ms'], count=v['count'], tflops=gflop * v['count'] / v['ms'], gbs=byts * v['count'] / v['ms'] / 1e6))
    rows.sort(key=lambda r: -r['ms'])
    print(f"== {name}: all kernels {tot:.2f} ms, conv {conv:.2f} ms")
    for r in rows[:28]:
        print(f"  {r['shape']:62s} x{r['count']:3d} {r['ms']:8.3f} ms  {r['tflops']:7.1f} TFLOP/s(useful)  {r['gbs']:7.0f} GB/s")
    return rows


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/layer_profile.json"
    seg = AnimeInsSeg(None, default_det_size=1024, refine_kwargs={'refine_method': 'none'})
    imgs = torch.from_numpy(np.stack([smooth_image(1024, 1024, seed=100 + i) for i in range(8)])).cuda().repeat(B // 8, 1, 1, 1).contiguous()
    det = table("detector", profile(lambda: seg.model.net.forward(imgs)))
    del seg
    net = LeReS(None)
    im2 = torch.from_numpy(np.stack([smooth_image(640, 640, seed=i) for i in range(8)])).cuda().repeat(B // 8, 1, 1, 1).contiguous()
    ler = table("leres", profile(lambda: net.forward(im2)))
    json.dump(dict(batch=B, detector=det, leres=ler), open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
