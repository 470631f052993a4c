"""ISNet-DIS mask refinement on the B200 engine.

Reference: `animeinsseg/models/animeseg_refine/isnet.py` -- REBNCONV :95-108 (3x3 conv, dilation d, BN, ReLU), `_upsample_like` :111-115
(bilinear, align_corners=False), RSU7/6/5/4 :118-372, RSU4F :375-407, ISNetDIS(in_ch=4) :524-645; loaded by `load_refinenet('refinenet_isnet')`
(`animeseg_refine/__init__.py:159-167`) and used by `AnimeInsSeg._postprocess_refine` (`animeinsseg/__init__.py:638-665`), which only reads
side output d1 -- so side2..side6 are never computed here.

Every REBNCONV is one tcgen05 conv launch (BN folded, ReLU fused); `torch.cat((up, skip))` never copies: the encoder conv writes its output
straight into the upper half of the decoder's concat buffer and the upsampler writes the lower half; the RSU residual `hx1d + hxin` is the
conv epilogue's post-activation residual.  Parameters: a state_dict with the reference's names (`stage1.rebnconvin.conv_s1.weight`, ...).
"""
import math

import torch

from .. import engine as E

# name, RSU height L (0 = RSU4F), in, mid, out
STAGES = [("stage1", 7, 64, 32, 64), ("stage2", 6, 64, 32, 128), ("stage3", 5, 128, 64, 256), ("stage4", 4, 256, 128, 512), ("stage5", 0, 512, 256, 512),
          ("stage6", 0, 512, 256, 512), ("stage5d", 0, 1024, 256, 512), ("stage4d", 4, 1024, 128, 256), ("stage3d", 5, 512, 64, 128),
          ("stage2d", 6, 256, 32, 64), ("stage1d", 7, 128, 16, 64)]


def _rebn(specs, name, cin, cout):
    specs.append((f"{name}.conv_s1.weight", (cout, cin, 3, 3), 'conv_act'))
    specs.append((f"{name}.conv_s1.bias", (cout,), 'bias'))
    for p, kind in (('weight', 'bn_w'), ('bias', 'bn_b'), ('running_mean', 'bn_m'), ('running_var', 'bn_v')):
        specs.append((f"{name}.bn_s1.{p}", (cout,), kind))


def param_specs(in_ch=4):
    s = [("conv_in.weight", (64, in_ch, 3, 3), 'conv_lin'), ("conv_in.bias", (64,), 'bias')]
    for name, L, cin, mid, cout in STAGES:
        n = 4 if L == 0 else L
        _rebn(s, f"{name}.rebnconvin", cin, cout)
        _rebn(s, f"{name}.rebnconv1", cout, mid)
        for k in range(2, n + 1):
            _rebn(s, f"{name}.rebnconv{k}", mid, mid)
        for k in range(n - 1, 1, -1):
            _rebn(s, f"{name}.rebnconv{k}d", 2 * mid, mid)
        _rebn(s, f"{name}.rebnconv1d", 2 * mid, cout)
    for i, c in zip(range(1, 7), (64, 64, 128, 256, 512, 512)):
        s += [(f"side{i}.weight", (1, c, 3, 3), 'conv_zm:2'), (f"side{i}.bias", (1,), 'bias')]
    return s


def synthetic_state_dict(seed=0, in_ch=4):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    u = lambda shape, lo, hi: torch.rand(shape, generator=g) * (hi - lo) + lo
    for name, shape, kind in param_specs(in_ch):
        gain = 1.0
        if ':' in kind:
            kind, gain = kind.split(':')[0], float(kind.split(':')[1])
        if kind.startswith('conv'):
            fan_in = shape[1] * shape[2] * shape[3]
            std = math.sqrt(2.0 / fan_in) if kind == 'conv_act' else 1.0 / math.sqrt(fan_in)
            sd[name] = torch.randn(shape, generator=g) * (std * gain)
            if kind == 'conv_zm':            # zero-mean filter: the side output reacts to feature variation, not to the (positive) ReLU DC level,
                sd[name] = sd[name] - sd[name].mean()      # so the refined mask is not trivially all-ones under random weights
        elif kind in ('bias', 'bn_b'):
            sd[name] = u(shape, -0.1, 0.1)
        elif kind in ('bn_w', 'bn_v'):
            sd[name] = u(shape, 0.8, 1.2)
        elif kind == 'bn_m':
            sd[name] = torch.randn(shape, generator=g) * 0.1
    return sd


class _REBN:
    def __init__(self, sd, name, dev, dil=1, eps=1e-5, cin_pad=None):
        w, b = sd[f"{name}.conv_s1.weight"].float(), sd[f"{name}.conv_s1.bias"].float()
        g, beta, m, v = (sd[f"{name}.bn_s1.{k}"].float() for k in ("weight", "bias", "running_mean", "running_var"))
        s = g / torch.sqrt(v + eps)
        self.w = E.pack_conv_weight((w * s.view(-1, 1, 1, 1)).to(dev), torch.float16, cin_pad)
        self.b = ((b - m) * s + beta).contiguous().to(dev)
        self.dil = dil

    def __call__(self, x, **kw):
        return E.conv2d_nhwc(x, self.w, self.b, pad=self.dil, dil=self.dil, act='relu', **kw)


class _RSU:
    """RSU-L (L = 4..7) or RSU-4F (L = 0).  forward(x [N,h,w,Cx] slice in_coff..in_coff+cin, out, out_coff)."""

    def __init__(self, sd, name, L, cin, mid, cout, dev):
        self.L, self.cin, self.mid, self.cout = L, cin, mid, cout
        n = 4 if L == 0 else L
        dil_enc = [1, 2, 4, 8] if L == 0 else [1] * (n - 1) + [2]
        self.cin_conv = _REBN(sd, f"{name}.rebnconvin", dev)
        self.enc = [_REBN(sd, f"{name}.rebnconv{k}", dev, dil_enc[k - 1]) for k in range(1, n + 1)]
        dil_dec = {3: 4, 2: 2, 1: 1} if L == 0 else {k: 1 for k in range(1, n)}
        self.dec = {k: _REBN(sd, f"{name}.rebnconv{k}d", dev, dil_dec[k]) for k in range(n - 1, 0, -1)}

    def __call__(self, x, out=None, out_coff=0, in_coff=0):
        N, h, w, _ = x.shape
        dev, f16, mid, n = x.device, torch.float16, self.mid, (4 if self.L == 0 else self.L)
        hxin = self.cin_conv(x, in_coff=in_coff)
        if self.L == 0:         # RSU-4F: no pooling, dilations 1,2,4,8
            cats = {k: torch.empty((N, h, w, 2 * mid), device=dev, dtype=f16) for k in (1, 2, 3)}      # cat_k = [deeper | hx_k]
            self.enc[0](hxin, out=cats[1], out_coff=mid)
            self.enc[1](cats[1], in_coff=mid, out=cats[2], out_coff=mid)
            self.enc[2](cats[2], in_coff=mid, out=cats[3], out_coff=mid)
            self.enc[3](cats[3], in_coff=mid, out=cats[3], out_coff=0)                                 # hx4 -> lower half of cat_3
            self.dec[3](cats[3], out=cats[2], out_coff=0)
            self.dec[2](cats[2], out=cats[1], out_coff=0)
            return E.conv2d_nhwc(cats[1], self.dec[1].w, self.dec[1].b, pad=1, act='relu', residual=hxin, res_mode=2, out=out, out_coff=out_coff)
        sizes = [(h, w)]
        for _ in range(n - 2):
            sizes.append((E.pool_out(sizes[-1][0], 2, 2, 0, True), E.pool_out(sizes[-1][1], 2, 2, 0, True)))
        cats = {k: torch.empty((N, sizes[k - 1][0], sizes[k - 1][1], 2 * mid), device=dev, dtype=f16) for k in range(1, n)}
        cur, coff = hxin, 0
        for k in range(1, n):                       # hx_k -> upper half of cat_k, pooled for the next level (ceil-mode 2x2)
            self.enc[k - 1](cur, in_coff=coff, out=cats[k], out_coff=mid)
            if k < n - 1:
                cur, coff = E.maxpool2d_nhwc(cats[k], 2, 2, 0, True, xoff=mid, channels=mid), 0
        self.enc[n - 1](cats[n - 1], in_coff=mid, out=cats[n - 1], out_coff=0)                         # dilated hx_n -> lower half of cat_{n-1}
        for k in range(n - 1, 1, -1):
            d = self.dec[k](cats[k])
            E.resample_nhwc(d, sizes[k - 2][0], sizes[k - 2][1], 'bilinear', out=cats[k - 1], yoff=0)  # _upsample_like -> lower half of cat_{k-1}
        return E.conv2d_nhwc(cats[1], self.dec[1].w, self.dec[1].b, pad=1, act='relu', residual=hxin, res_mode=2, out=out, out_coff=out_coff)


class ISNetDIS:
    """B200 forward of ISNetDIS(in_ch=4): x [N,H,W,16] fp16 (4 real channels) -> d1 logits [N,H,W] fp32 (the only output the refine path reads)."""

    def __init__(self, state_dict=None, device='cuda', in_ch=4):
        sd = synthetic_state_dict(0, in_ch) if state_dict is None else state_dict
        dev = self.dev = torch.device(device)
        self.conv_in = (E.pack_conv_weight(sd["conv_in.weight"].to(dev), torch.float16, 16), sd["conv_in.bias"].float().contiguous().to(dev))
        self.rsu = {name: _RSU(sd, name, L, cin, mid, cout, dev) for name, L, cin, mid, cout in STAGES}
        self.side1 = (E.pack_conv_weight(sd["side1.weight"].to(dev)), sd["side1.bias"].float().contiguous().to(dev))

    def forward(self, x16):
        N, H, W, _ = x16.shape
        dev, f16 = x16.device, torch.float16
        hxin = E.conv2d_nhwc(x16, self.conv_in[0], self.conv_in[1], stride=2, pad=1)
        szs = [hxin.shape[1:3]]
        for _ in range(5):
            szs.append((E.pool_out(szs[-1][0], 2, 2, 0, True), E.pool_out(szs[-1][1], 2, 2, 0, True)))
        chans = (64, 128, 256, 512, 512)
        cat = [torch.empty((N, szs[i][0], szs[i][1], 2 * chans[i]), device=dev, dtype=f16) for i in range(5)]     # cat_i = [up(deeper) | hx_{i+1}]
        names = ("stage1", "stage2", "stage3", "stage4", "stage5")
        cur, coff = hxin, 0
        for i in range(5):
            self.rsu[names[i]](cur, out=cat[i], out_coff=chans[i], in_coff=coff)
            cur, coff = E.maxpool2d_nhwc(cat[i], 2, 2, 0, True, xoff=chans[i], channels=chans[i]), 0
        hx6 = self.rsu["stage6"](cur)
        E.resample_nhwc(hx6, szs[4][0], szs[4][1], 'bilinear', out=cat[4], yoff=0)
        d = None
        for i, name in zip((4, 3, 2, 1, 0), ("stage5d", "stage4d", "stage3d", "stage2d", "stage1d")):
            d = self.rsu[name](cat[i])
            if i > 0:
                E.resample_nhwc(d, szs[i - 1][0], szs[i - 1][1], 'bilinear', out=cat[i - 1], yoff=0)
        d1 = E.conv2d_nhwc(d, self.side1[0], self.side1[1], pad=1, out_f32=True)          # [N,h/2,w/2,1]
        return E.resample_f32(d1.view(N, d1.shape[1], d1.shape[2]), H, W, False)          # _upsample_like(d1, x)
