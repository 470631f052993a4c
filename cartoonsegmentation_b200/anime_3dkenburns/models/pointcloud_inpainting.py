"""Point-cloud inpainting network on the B200 engine: `Inpaint` of
`anime_3dkenburns/models/pointcloud_inpainting.py:81-204` -- context extractor, 68-channel context render, 4x4 GridNet, colour / disparity heads.

Every conv is one tcgen05 launch.  The GridNet blocks are pre-activation (`PReLU -> conv -> PReLU -> conv`, :10-15): the inner PReLU is the first
conv's epilogue activation, the leading one is an elementwise kernel on the block input (which is also the skip), and every `+=` of the grid
(block skip, row merge) is the second conv's epilogue residual, written in place.  The context render goes through
`csb_inpaint_context_render` (fp16 interleaved payload straight from the conv engine; coverage median-5 and masking fused with the normalise).
Parameters: a state_dict with the reference's names (e.g. `'0x0 - 0x1.netMain.1.weight'`), so `kenburns_inpaintnet.ckpt` drops in.
"""
import ctypes as C
import math

import torch

from ... import engine as E
from ..._lib import check, f3, lib, ptr, stream

ROWS = [(0, 32), (1, 64), (2, 128), (3, 256)]


def _basic_specs(s, name, chans, kind):
    c0, c1, c2 = chans
    if kind == 'relu-conv-relu-conv':
        s += [(f"{name}.netMain.0.weight", (c0,), 'prelu'), (f"{name}.netMain.1.weight", (c1, c0, 3, 3), 'conv_act'), (f"{name}.netMain.1.bias", (c1,), 'bias'),
              (f"{name}.netMain.2.weight", (c1,), 'prelu'), (f"{name}.netMain.3.weight", (c2, c1, 3, 3), 'conv_res'), (f"{name}.netMain.3.bias", (c2,), 'bias')]
    else:
        s += [(f"{name}.netMain.0.weight", (c1, c0, 3, 3), 'conv_act'), (f"{name}.netMain.0.bias", (c1,), 'bias'), (f"{name}.netMain.1.weight", (c1,), 'prelu'),
              (f"{name}.netMain.2.weight", (c2, c1, 3, 3), 'conv_res'), (f"{name}.netMain.2.bias", (c2,), 'bias')]
    if c0 != c2:
        s += [(f"{name}.netShortcut.weight", (c2, c0, 1, 1), 'conv_lin'), (f"{name}.netShortcut.bias", (c2,), 'bias')]


def param_specs():
    s = [("netContext.0.weight", (64, 4, 3, 3), 'conv_act'), ("netContext.0.bias", (64,), 'bias'), ("netContext.1.weight", (64,), 'prelu'),
         ("netContext.2.weight", (64, 64, 3, 3), 'conv_act'), ("netContext.2.bias", (64,), 'bias'), ("netContext.3.weight", (64,), 'prelu')]
    _basic_specs(s, "netInput", (69, 32, 32), 'conv-relu-conv')
    for r, f in ROWS:
        for c in range(3):
            _basic_specs(s, f"{r}x{c} - {r}x{c + 1}", (f, f, f), 'relu-conv-relu-conv')
    for col in (0, 1):
        for (r, f), (_, f2) in zip(ROWS[:-1], ROWS[1:]):
            n = f"{r}x{col} - {r + 1}x{col}"
            s += [(f"{n}.netMain.0.weight", (f,), 'prelu'), (f"{n}.netMain.1.weight", (f2, f, 3, 3), 'conv_act'), (f"{n}.netMain.1.bias", (f2,), 'bias'),
                  (f"{n}.netMain.2.weight", (f2,), 'prelu'), (f"{n}.netMain.3.weight", (f2, f2, 3, 3), 'conv_res'), (f"{n}.netMain.3.bias", (f2,), 'bias')]
    for col in (2, 3):
        for (r, f), (_, f2) in zip(ROWS[:-1], ROWS[1:]):
            n = f"{r + 1}x{col} - {r}x{col}"
            s += [(f"{n}.netMain.1.weight", (f2,), 'prelu'), (f"{n}.netMain.2.weight", (f, f2, 3, 3), 'conv_act'), (f"{n}.netMain.2.bias", (f,), 'bias'),
                  (f"{n}.netMain.3.weight", (f,), 'prelu'), (f"{n}.netMain.4.weight", (f, f, 3, 3), 'conv_res'), (f"{n}.netMain.4.bias", (f,), 'bias')]
    _basic_specs(s, "netImage", (32, 32, 3), 'conv-relu-conv')
    _basic_specs(s, "netDisparity", (32, 32, 1), 'conv-relu-conv')
    return s


def synthetic_state_dict(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in param_specs():
        if kind.startswith('conv'):
            fan_in = shape[1] * shape[2] * shape[3]
            std = math.sqrt(2.0 / fan_in) if kind == 'conv_act' else (0.5 / math.sqrt(fan_in) if kind == 'conv_res' else 1.0 / math.sqrt(fan_in))
            sd[name] = torch.randn(shape, generator=g) * std
        elif kind == 'bias':
            sd[name] = torch.rand(shape, generator=g) * 0.2 - 0.1
        elif kind == 'prelu':
            sd[name] = torch.rand(shape, generator=g) * 0.3 + 0.1          # PReLU slopes in [0.1, 0.4] (reference init: 0.25)
    return sd


class _Conv:
    def __init__(self, sd, name, dev, cin_pad=None):
        self.w = E.pack_conv_weight(sd[f"{name}.weight"].to(dev), torch.float16, cin_pad)
        self.b = sd[f"{name}.bias"].float().contiguous().to(dev)

    def __call__(self, x, **kw):
        return E.conv2d_nhwc(x, self.w, self.b, **kw)


def _prelu(x, slope, out=None):
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib().csb_prelu_nhwc(ptr(x), Cc, 0, ptr(slope), C.c_longlong(N * H * W), Cc, ptr(out), Cc, 0, stream()), "csb_prelu_nhwc")
    return out


class Inpaint:
    """forward(tenImage [1,3,H,W], tenDisparity [1,1,H,W], tenShift (3 floats), objCommon) -> dict like the reference (:197-202)."""

    def __init__(self, state_dict=None, device='cuda'):
        sd = synthetic_state_dict(0) if state_dict is None else state_dict
        dev = self.dev = torch.device(device)
        sl = lambda n: sd[n].float().contiguous().to(dev)
        self.ctx = (_Conv(sd, "netContext.0", dev, 16), sl("netContext.1.weight"), _Conv(sd, "netContext.2", dev), sl("netContext.3.weight"))
        self.inp = (_Conv(sd, "netInput.netMain.0", dev, 80), sl("netInput.netMain.1.weight"), _Conv(sd, "netInput.netMain.2", dev),
                    _Conv(sd, "netInput.netShortcut", dev, 80))
        self.basic, self.down, self.up = {}, {}, {}
        for r, f in ROWS:
            for c in range(3):
                n = f"{r}x{c} - {r}x{c + 1}"
                self.basic[(r, c + 1)] = (sl(f"{n}.netMain.0.weight"), _Conv(sd, f"{n}.netMain.1", dev), sl(f"{n}.netMain.2.weight"), _Conv(sd, f"{n}.netMain.3", dev))
        for col in (0, 1):
            for r in range(3):
                n = f"{r}x{col} - {r + 1}x{col}"
                self.down[(r + 1, col)] = (sl(f"{n}.netMain.0.weight"), _Conv(sd, f"{n}.netMain.1", dev), sl(f"{n}.netMain.2.weight"), _Conv(sd, f"{n}.netMain.3", dev))
        for col in (2, 3):
            for r in range(3):
                n = f"{r + 1}x{col} - {r}x{col}"
                self.up[(r, col)] = (sl(f"{n}.netMain.1.weight"), _Conv(sd, f"{n}.netMain.2", dev), sl(f"{n}.netMain.3.weight"), _Conv(sd, f"{n}.netMain.4", dev))
        self.head = {k: (_Conv(sd, f"{k}.netMain.0", dev), sl(f"{k}.netMain.1.weight"), _Conv(sd, f"{k}.netMain.2", dev), _Conv(sd, f"{k}.netShortcut", dev))
                     for k in ("netImage", "netDisparity")}

    # ---- grid pieces
    def _basic(self, key, x):
        s1, c1, s2, c2 = self.basic[key]
        b = c1(_prelu(x, s1), pad=1, act='prelu', act_param=s2)
        return c2(b, pad=1, residual=x, res_mode=2)                      # netMain(x) + x

    def _down_into(self, key, x, target):
        s1, c1, s2, c2 = self.down[key]
        b = c1(_prelu(x, s1), stride=2, pad=1, act='prelu', act_param=s2)
        c2(b, pad=1, residual=target, res_mode=2, out=target)            # target += Downsample(x)

    def _down_new(self, key, x):
        s1, c1, s2, c2 = self.down[key]
        b = c1(_prelu(x, s1), stride=2, pad=1, act='prelu', act_param=s2)
        return c2(b, pad=1)

    def _up_into(self, key, x, target):
        s1, c1, s2, c2 = self.up[key]
        u = E.resample_nhwc(x, x.shape[1] * 2, x.shape[2] * 2, 'bilinear')
        b = c1(_prelu(u, s1, out=u), pad=1, act='prelu', act_param=s2)
        assert b.shape[1:3] == target.shape[1:3], "Inpaint: H and W must be multiples of 8 (the reference crops odd sizes by negative padding, :165-166)"
        c2(b, pad=1, residual=target, res_mode=2, out=target)            # target += Upsample(x)

    def forward(self, tenImage, tenDisparity, tenShift, objCommon, segmasks=None):
        from .utils import _f32
        dev = self.dev
        img, disp = _f32(tenImage).to(dev), _f32(tenDisparity).to(dev)
        _, _, H, W = img.shape
        assert H % 8 == 0 and W % 8 == 0
        focal, baseline = objCommon['fltFocal'], objCommon['fltBaseline']
        # :117-120 -- geometry of the raw frame: depth = fb / (disp + 1e-7), valid = |laplacian(disp / max)| < 0.03, points of depth * valid (one kernel)
        from .utils import net_output, pack_norm16, tensor_stats
        st_i, st_d = tensor_stats(img), tensor_stats(disp)                                               # {mean, std, max} on the device
        tenPoints = torch.empty((1, 3, H * W), device=dev, dtype=torch.float32)
        check(lib().csb_inpaint_points(ptr(disp), H, W, C.c_double(focal), C.c_double(baseline), ptr(st_d), ptr(tenPoints), stream()), "csb_inpaint_points")
        # :122-131 -- per-tensor normalisation, packed as the NHWC fp16 input [img_n(3) | disp_n(1) | zeros]
        x16 = pack_norm16(img, st_i, disp, st_d)
        c0, sl0, c1, sl1 = self.ctx
        ctx = c1(c0(x16, pad=1, act='prelu', act_param=sl0), pad=1, act='prelu', act_param=sl1)          # [1,H,W,64]
        payload = torch.empty((H * W, 72), device=dev, dtype=torch.float16)                               # [img(3) | disp(1) | context(64) | pad]
        check(lib().csb_inpaint_payload(ptr(x16), ptr(ctx), C.c_longlong(H * W), ptr(payload), stream()), "csb_inpaint_payload")
        # :135-142 -- context render + coverage median + masking, written as netInput's NHWC input
        sh = [float(v) for v in torch.as_tensor(tenShift).flatten().tolist()]
        CP = lib().csb_render_acc_channels(68)
        zkey = torch.empty((H, W), device=dev, dtype=torch.int32); zee = torch.empty((H, W), device=dev)
        acc = torch.empty((H, W, CP), device=dev); flags = torch.empty((H * W,), device=dev, dtype=torch.uint8)
        x80 = torch.empty((1, H, W, 80), device=dev, dtype=torch.float16); existing = torch.empty((1, 1, H, W), device=dev)
        check(lib().csb_inpaint_context_render(ptr(tenPoints.contiguous()), ptr(payload), 72, H * W, 68, H, W, C.c_double(focal), C.c_double(baseline), f3(sh),
                                               ptr(zkey), ptr(zee), ptr(acc), ptr(flags), ptr(x80), 80, ptr(existing), stream()), "csb_inpaint_context_render")
        # :146-149 -- column 0
        ci, sli, ci2, csc = self.inp
        col = [None] * 4
        col[0] = ci2(ci(x80, pad=1, act='prelu', act_param=sli), pad=1, residual=csc(x80), res_mode=2)
        for r in (1, 2, 3):
            col[r] = self._down_new((r, 0), col[r - 1])
        # :151-157 -- column 1
        for r in range(4):
            col[r] = self._basic((r, 1), col[r])
            if r != 0:
                self._down_into((r, 1), col[r - 1], col[r])
        # :159-187 -- columns 2 and 3
        for c in (2, 3):
            for r in (3, 2, 1, 0):
                col[r] = self._basic((r, c), col[r])
                if r != 3:
                    self._up_into((r, c), col[r + 1], col[r])
        out = {}
        for k, st_k, post in (("netImage", st_i, 1), ("netDisparity", st_d, 2)):
            h0, slh, h1, hsc = self.head[k]
            # :190-200 (eval mode): (head + shortcut) * (std + 1e-7) + mean, image clipped to [0, 1], disparity thresholded at 0 -> [1,c,H,W] fp32
            out[k] = net_output(h1(h0(col[0], pad=1, act='prelu', act_param=slh), pad=1, out_f32=True), hsc(col[0], out_f32=True), st_k, post)
        tenImage, tenDisp = out["netImage"], out["netDisparity"]
        return {'tenExisting': existing, 'tenImage': tenImage, 'tenDisparity': tenDisp, 'segmasks': None}
