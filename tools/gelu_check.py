"""GELU epilogue accuracy + speed for the C -> 4C GEMM shapes: python tools/gelu_check.py  (env CSB_GELU_MUFU selects the experimental split)"""
import os, sys
ACT = os.environ.get('ACT', 'gelu')
ACT = None if ACT == 'none' else ACT
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200 import engine as E

torch.manual_seed(0)
# accuracy: identity-like GEMM so the epilogue input is known: x in [-12, 12], W = I (64x64)
x = torch.linspace(-12, 12, 64 * 4096, device='cuda').reshape(1, 64, 64, 64).half()
w = E.pack_conv_weight(torch.eye(64, device='cuda').reshape(64, 64, 1, 1))
y = E.conv2d_nhwc(x, w, torch.zeros(64, device='cuda'), act='gelu').float()
ref = torch.nn.functional.gelu(x.float())
err = (y - ref.half().float()).abs()        # vs the correctly rounded fp16 result
print("gelu max abs err %.3e  (rel to |x| max %.3e)" % (err.max().item(), (err / x.float().abs().clamp_min(1)).max().item()))
ulp = torch.maximum(ref.half().float().abs(), torch.tensor(6.1e-5, device='cuda')).log2().floor().exp2() * 2.0 ** -10        # fp16 ulp at the reference value
print("gelu error in fp16 ulps of the result: max %.2f  rms %.3f  fraction != correctly rounded %.4f; rms abs err %.3e" %
      ((err / ulp).max().item(), (err / ulp).pow(2).mean().sqrt().item(), (err > 0).float().mean().item(), err.pow(2).mean().sqrt().item()))
for (N, H, W, Cin, Cout) in ((32, 64, 64, 512, 2048), (16, 256, 256, 128, 512), (32, 128, 128, 256, 1024)):
    xs = [torch.randn(N, H, W, Cin, device='cuda').half() for _ in range(3)]
    ww = E.pack_conv_weight(torch.randn(Cout, Cin, 1, 1, device='cuda') * Cin ** -0.5)
    b = torch.randn(Cout, device='cuda')
    outs = [torch.empty(N, H, W, Cout, device='cuda', dtype=torch.float16) for _ in range(3)]
    for i in range(3):
        E.conv2d_nhwc(xs[i], ww, b, act=ACT, out=outs[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        E.conv2d_nhwc(xs[i % 3], ww, b, act=ACT, out=outs[i % 3])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    print(f"{N}x{H}x{W}x{Cin}->{Cout}: {ms*1e3:.1f} us  {2*N*H*W*Cin*Cout/ms/1e9:.0f} TFLOP/s  {2*N*H*W*(Cin+Cout)/ms/1e6:.0f} GB/s")
