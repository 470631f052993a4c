"""One attention launch for `ncu --set full -k regex:k_attention_tc`: python tools/attn_one.py [B] [T]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartoonsegmentation_b200._lib import check, lib, ptr, stream          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1765
heads, d = 16, 64
qkv = (torch.randn(B, T, 3 * heads * d, device='cuda') * 1.5).half()
Tp = (T + 127) // 128 * 128
bias = torch.full((heads, Tp, Tp), -60000.0, device='cuda', dtype=torch.float16)
bias[:, :T, :T] = (torch.randn(heads, T, T, device='cuda') * 2).half()
out = torch.empty(B, T, heads * d, device='cuda', dtype=torch.float16)
lib().csb_attention_tc_scratch_bytes.restype = C.c_longlong
vt = torch.empty(int(lib().csb_attention_tc_scratch_bytes(B, T, heads)), device='cuda', dtype=torch.uint8)
for _ in range(3):
    check(lib().csb_attention_bias_tc(ptr(qkv), B, T, heads, d, ptr(bias), Tp, C.c_float(d ** -0.5), ptr(vt), ptr(out), stream()), "csb_attention_bias_tc")
torch.cuda.synchronize()
