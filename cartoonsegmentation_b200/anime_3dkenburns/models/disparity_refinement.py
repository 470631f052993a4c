"""Disparity refinement network on the B200 engine: `Refine` of `anime_3dkenburns/models/disparity_refinement.py:84-135`,
called through `disparity_refinement(net, tenImage, tenDisparity)` (`models/__init__.py:13-14`, `kenburns_effect.py:619-620,821-829`).

Image pyramid (3 -> 24 -> 48 -> 96, two stride-2 `Downsample`s), disparity branch (1 -> 96) merged coarse-to-fine through two `Upsample`s and a
`Basic`, 1-channel head; per-tensor mean / std normalisation in, de-normalisation + threshold(0) out.  Every conv is one tcgen05 launch; the
24- and 72-channel tensors live in 32- / 80-channel zero-padded NHWC buffers (the engine wants Cin % 16 == 0), concats are channel-slice writes.
Parameters: a state_dict with the reference's names (`netImageOne.netMain.0.weight`, ...), so `kenburns_refinenet.ckpt` drops in.
"""
import math

import torch

from ... import engine as E
from .pointcloud_inpainting import _prelu

SPEC = [("netImageOne", 'basic', (3, 24, 24)), ("netImageTwo", 'down', (24, 48, 48)), ("netImageThr", 'down', (48, 96, 96)),
        ("netDisparityOne", 'basic', (1, 96, 96)), ("netDisparityTwo", 'up', (192, 96, 96)), ("netDisparityThr", 'up', (144, 48, 48)),
        ("netDisparityFou", 'basic', (72, 24, 24)), ("netRefine", 'basic', (24, 24, 1))]


def param_specs():
    s = []
    for name, kind, (c0, c1, c2) in SPEC:
        if kind == 'basic':                     # 'conv-relu-conv' (+ 1x1 shortcut when c0 != c2)
            s += [(f"{name}.netMain.0.weight", (c1, c0, 3, 3), 'conv_act'), (f"{name}.netMain.0.bias", (c1,), 'bias'), (f"{name}.netMain.1.weight", (c1,), 'prelu'),
                  (f"{name}.netMain.2.weight", (c2, c1, 3, 3), 'conv_res'), (f"{name}.netMain.2.bias", (c2,), 'bias')]
            if c0 != c2:
                s += [(f"{name}.netShortcut.weight", (c2, c0, 1, 1), 'conv_lin'), (f"{name}.netShortcut.bias", (c2,), 'bias')]
        else:                                   # Downsample: PReLU, conv s2, PReLU, conv; Upsample: bilinear x2, PReLU, conv, PReLU, conv
            o = 0 if kind == 'down' else 1
            s += [(f"{name}.netMain.{o}.weight", (c0,), 'prelu'), (f"{name}.netMain.{o + 1}.weight", (c1, c0, 3, 3), 'conv_act'), (f"{name}.netMain.{o + 1}.bias", (c1,), 'bias'),
                  (f"{name}.netMain.{o + 2}.weight", (c1,), 'prelu'), (f"{name}.netMain.{o + 3}.weight", (c2, c1, 3, 3), 'conv_res'), (f"{name}.netMain.{o + 3}.bias", (c2,), 'bias')]
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
        else:
            sd[name] = torch.rand(shape, generator=g) * 0.3 + 0.1
    return sd


def _pad16(c):
    return (c + 15) // 16 * 16


class _C:
    """conv weight packed with Cin padded to a multiple of 16 (zero taps on the padding channels)."""

    def __init__(self, sd, name, dev):
        w = sd[f"{name}.weight"]
        self.w = E.pack_conv_weight(w.to(dev), torch.float16, _pad16(w.shape[1]))
        self.b = sd[f"{name}.bias"].float().contiguous().to(dev)

    def __call__(self, x, **kw):
        return E.conv2d_nhwc(x, self.w, self.b, **kw)


def _slope(sd, name, dev):
    """PReLU slopes padded to the buffer's channel count (the padding channels hold zeros, any slope does)."""
    s = sd[name].float()
    return torch.cat([s, torch.zeros(_pad16(s.numel()) - s.numel())]).contiguous().to(dev)


class Refine:
    """forward(tenImage [1,3,H,W] fp32, tenDisparity [1,1,h,w] fp32) -> refined disparity [1,1,H,W] fp32 (reference :98-132)."""

    def __init__(self, state_dict=None, device='cuda'):
        sd = synthetic_state_dict(0) if state_dict is None else state_dict
        dev = self.dev = torch.device(device)
        self.m = {}
        for name, kind, (c0, c1, c2) in SPEC:
            if kind == 'basic':
                self.m[name] = dict(c1=_C(sd, f"{name}.netMain.0", dev), s=_slope(sd, f"{name}.netMain.1.weight", dev), c2=_C(sd, f"{name}.netMain.2", dev),
                                    sc=_C(sd, f"{name}.netShortcut", dev) if c0 != c2 else None, out=c2)
            else:
                o = 0 if kind == 'down' else 1
                self.m[name] = dict(s0=_slope(sd, f"{name}.netMain.{o}.weight", dev), c1=_C(sd, f"{name}.netMain.{o + 1}", dev),
                                    s1=_slope(sd, f"{name}.netMain.{o + 2}.weight", dev), c2=_C(sd, f"{name}.netMain.{o + 3}", dev), out=c2)

    def _buf(self, N, H, W, Cc):
        return torch.zeros((N, H, W, _pad16(Cc)), device=self.dev, dtype=torch.float16)

    def _basic(self, name, x, out=None, out_coff=0, out_f32=False):
        m = self.m[name]
        N, H, W, _ = x.shape
        h = self._buf(N, H, W, m['c1'].b.numel())
        m['c1'](x, pad=1, act='prelu', act_param=m['s'], out=h)
        sc = x if m['sc'] is None else m['sc'](x, out=self._buf(N, H, W, m['out']))                   # netShortcut(x), or x itself
        if out is None and not out_f32:
            out = self._buf(N, H, W, m['out'])
        return m['c2'](h, pad=1, residual=sc, res_mode=2, out=out, out_coff=out_coff, out_f32=out_f32 and out is None)

    def _down(self, name, x):
        m = self.m[name]
        N, H, W, _ = x.shape
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        h = self._buf(N, Ho, Wo, m['c1'].b.numel())
        m['c1'](_prelu(x, m['s0']), stride=2, pad=1, act='prelu', act_param=m['s1'], out=h)
        return m['c2'](h, pad=1, out=self._buf(N, Ho, Wo, m['out']))

    def _up(self, name, x, out, out_coff):
        m = self.m[name]
        N, H, W, _ = x.shape
        u = E.resample_nhwc(x, H * 2, W * 2, 'bilinear')
        h = self._buf(N, H * 2, W * 2, m['c1'].b.numel())
        m['c1'](_prelu(u, m['s0'], out=u), pad=1, act='prelu', act_param=m['s1'], out=h)
        if out.shape[1:3] == (H * 2, W * 2):
            return m['c2'](h, pad=1, out=out, out_coff=out_coff)
        t = m['c2'](h, pad=1)                                                                          # "not ideal" resize of the reference (:116,118)
        return E.resample_nhwc(t, out.shape[1], out.shape[2], 'bilinear', out=out, yoff=out_coff)

    def forward(self, tenImage, tenDisparity):
        dev = self.dev
        img, disp = tenImage.float().to(dev), tenDisparity.float().to(dev)
        assert img.shape[0] == 1 and disp.shape[0] == 1, "the reference normalises over the whole batch tensor; batch 1 only"
        from .utils import net_output, pack_norm16, tensor_stats
        st_i, st_d = tensor_stats(img), tensor_stats(disp)                                             # :99-100: {mean, std} on the device
        _, _, H, W = img.shape
        h, w = disp.shape[2:]
        xi = pack_norm16(img, st_i)                                                                    # [(img - m) / (s + 1e-7) (3) | zeros]
        xd = pack_norm16(disp, st_d)
        cat1 = self._buf(1, H, W, 72)                                                                  # [imageOne(24) | upsample(48)]
        self._basic("netImageOne", xi, out=cat1, out_coff=0)
        one = cat1                                                                                     # channels 0..23 (+ the rest read as zero taps... see _down)
        two_in = self._buf(1, H, W, 24)
        two_in[..., :24] = cat1[..., :24]
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        cat2 = self._buf(1, H2, W2, 144)                                                               # [imageTwo(48) | upsample(96)]
        two = self._down("netImageTwo", two_in)
        cat2[..., :48] = two[..., :48]
        thr = self._down("netImageThr", two)
        H4, W4 = thr.shape[1:3]
        cat3 = self._buf(1, H4, W4, 192)                                                               # [imageThr(96) | disparityOne(96)]
        cat3[..., :96] = thr
        d1 = self._basic("netDisparityOne", xd)
        if d1.shape[1:3] == (H4, W4):
            cat3[..., 96:] = d1
        else:
            E.resample_nhwc(d1, H4, W4, 'bilinear', out=cat3, yoff=96)                                 # :114
        self._up("netDisparityTwo", cat3, cat2, 48)
        self._up("netDisparityThr", cat2, cat1, 24)
        fou = self._basic("netDisparityFou", cat1)
        ref = self._basic("netRefine", fou, out_f32=True)                                              # [1,H,W,1] fp32
        return net_output(ref, None, st_d, 2)                                                          # :128-135: * (s + 1e-7) + m, threshold at 0 -> [1,1,H,W]
