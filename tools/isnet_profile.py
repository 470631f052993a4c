"""Per-layer-shape device time of the ISNet refinement forward (A10) for a batch of B instances at SxS (reference: refine_size 720), with the library
profiler's detailed mode.  Usage: python tools/isnet_profile.py [B] [S] [out.json]"""
import ctypes
import json
import os
import re
import sys

os.environ["CSB_PROFILE_DETAIL"] = "1"
import torch                # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import _lib                                   # noqa: E402
from cartoonsegmentation_b200.animeinsseg import isnet as I                 # noqa: E402

PAT = re.compile(r"(k_conv_tc|k_conv_halo)\[(\d+)x(\d+)x(\d+)x(\d+)->(\d+) k(\d+)x(\d+) s(\d+) d(\d+) g(\d+) act(\d+) res(\d+)\]")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 720
    out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/isnet_profile.json"
    net = I.ISNetDIS(None)
    x = (torch.rand((B, S, S, 16), device='cuda') * 0.5).half()
    x[..., 4:] = 0
    lib = _lib.lib()
    for _ in range(2):
        net.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net.forward(x); e1.record(); torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    lib.csb_profile_begin(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    net.forward(x)
    buf = ctypes.create_string_buffer(1 << 18)
    lib.csb_profile_end(buf, len(buf))
    prof = json.loads(buf.value.decode())
    rows, other, conv = [], {}, 0.0
    for k, v in prof.items():
        m = PAT.match(k)
        if not m:
            other[k] = v
            continue
        N, H, W, Cin, Cout, R, S_, st, dil, g, act, res = map(int, m.groups()[1:])
        Ho, Wo = (H + st - 1) // st, (W + st - 1) // st
        gflop = 2.0 * N * Ho * Wo * Cout * (Cin // g) * R * S_ / 1e9
        byts = 2.0 * N * (H * W * Cin + Ho * Wo * Cout)
        conv += v['ms']
        rows.append(dict(shape=('halo ' if k.startswith('k_conv_halo') else '') + k[k.index('['):], ms=v['ms'], count=v['count'], gflop=gflop * v['count'], tflops=gflop * v['count'] / v['ms'], gbs=byts * v['count'] / v['ms'] / 1e6))
    rows.sort(key=lambda r: -r['ms'])
    tot = sum(v['ms'] for v in prof.values())
    print(f"== ISNet B={B} @{S}^2: forward {wall:.2f} ms unprofiled ({wall / B:.3f} ms/instance), profiled sum {tot:.2f} ms, conv {conv:.2f} ms, "
          f"total conv GFLOP {sum(r['gflop'] for r in rows):.0f}")
    for r in rows[:45]:
        print(f"  {r['shape']:58s} x{r['count']:3d} {r['ms']:8.3f} ms  {r['tflops']:7.1f} TFLOP/s  {r['gbs']:7.0f} GB/s")
    for k, v in sorted(other.items(), key=lambda kv: -kv[1]['ms']):
        print(f"  {k:58s} x{v['count']:3d} {v['ms']:8.3f} ms")
    json.dump(dict(batch=B, size=S, forward_ms=wall, conv=rows, other=other), open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
