"""LeReS monocular depth on the B200 engine: ResNeXt-101 32x8d encoder + FTB/FFM/AO decoder.

Reference: `depth_modules/leres/leres/Resnext_torch.py:70-236` (encoder), `network_auxi.py:15-62,100-124,191-213,238-259` (Decoder, FTB, FFM, AO),
`multi_depth_model_woauxi.py:6-32` (RelDepthModel / DepthModel), `depthmap.py:15-47` (estimateleres), `depth_modules/leres/__init__.py:69-147`
(apply_leres: 16 -> 8 bit quantisation + inversion), `anime_3dkenburns/kenburns_effect.py:563-581` (_depth_est_leres).

All convs run on the tcgen05 engine (NHWC fp16, fp32 accumulate): BatchNorm folded, ReLU / residual fused into the epilogue; the grouped 3x3
convs (groups=32, 8..64 channels per group) run as block-diagonal 64-channel slices (`csb_conv2d_nhwc` groups mode).  Parameters are a
state_dict with the reference's names (`depth_model.encoder_modules.encoder.*`, `depth_model.decoder_modules.*`), so `res101.pth` drops in.

Quirk reproduced (network_auxi.py:113-118): FTB's branch starts with `nn.ReLU(inplace=True)`, which also rectifies the tensor used by the skip:
    y = relu(conv1(x));  out = relu(y + conv_b(relu(bn(conv_a(y)))))
"""
import math

import numpy as np
import torch

from .. import engine as E

LAYERS = (3, 4, 23, 3)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
ENC, DEC = "depth_model.encoder_modules.encoder", "depth_model.decoder_modules"


# ------------------------------------------------------------------------------------------------ parameter inventory
def _bn(specs, name, c):
    for p, kind in (('weight', 'bn_w'), ('bias', 'bn_b'), ('running_mean', 'bn_m'), ('running_var', 'bn_v')):
        specs.append((f"{name}.{p}", (c,), kind))


def _ftb(specs, name, cin, mid):
    specs += [(f"{name}.conv1.weight", (mid, cin, 3, 3), 'conv_act'), (f"{name}.conv1.bias", (mid,), 'bias'),
              (f"{name}.conv_branch.1.weight", (mid, mid, 3, 3), 'conv_act'), (f"{name}.conv_branch.1.bias", (mid,), 'bias')]
    _bn(specs, f"{name}.conv_branch.2", mid)
    specs += [(f"{name}.conv_branch.4.weight", (mid, mid, 3, 3), 'conv_res'), (f"{name}.conv_branch.4.bias", (mid,), 'bias')]


def param_specs():
    s = [(f"{ENC}.conv1.weight", (64, 3, 7, 7), 'conv_act')]
    _bn(s, f"{ENC}.bn1", 64)
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), LAYERS), start=1):
        width = planes * 4                       # int(planes * 8/64) * 32
        for b in range(nblk):
            p = f"{ENC}.layer{li}.{b}"
            s.append((f"{p}.conv1.weight", (width, inplanes, 1, 1), 'conv_act')); _bn(s, f"{p}.bn1", width)
            s.append((f"{p}.conv2.weight", (width, width // 32, 3, 3), 'conv_act')); _bn(s, f"{p}.bn2", width)
            s.append((f"{p}.conv3.weight", (planes * 4, width, 1, 1), 'conv_res')); _bn(s, f"{p}.bn3", planes * 4)
            if b == 0:
                s.append((f"{p}.downsample.0.weight", (planes * 4, inplanes, 1, 1), 'conv_lin')); _bn(s, f"{p}.downsample.1", planes * 4)
            inplanes = planes * 4
    _ftb(s, f"{DEC}.conv", 2048, 512)
    s += [(f"{DEC}.conv1.weight", (256, 512, 3, 3), 'conv_lin'), (f"{DEC}.conv1.bias", (256,), 'bias')]
    for name, cin in (("ffm2", 1024), ("ffm1", 512), ("ffm0", 256)):
        _ftb(s, f"{DEC}.{name}.ftb1", cin, 256)
        _ftb(s, f"{DEC}.{name}.ftb2", 256, 256)
    s += [(f"{DEC}.outconv.adapt_conv.0.weight", (128, 256, 3, 3), 'conv_act'), (f"{DEC}.outconv.adapt_conv.0.bias", (128,), 'bias')]
    _bn(s, f"{DEC}.outconv.adapt_conv.1", 128)
    s += [(f"{DEC}.outconv.adapt_conv.3.weight", (1, 128, 3, 3), 'conv_lin'), (f"{DEC}.outconv.adapt_conv.3.bias", (1,), 'bias')]
    return s


def synthetic_state_dict(seed=0):
    """Seeded variance-preserving weights with the reference's parameter names (see animeinsseg/rtmdet.py:synthetic_state_dict)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    u = lambda shape, lo, hi: torch.rand(shape, generator=g) * (hi - lo) + lo
    for name, shape, kind in param_specs():
        if kind.startswith('conv'):
            fan_in = shape[1] * shape[2] * shape[3]
            std = math.sqrt(2.0 / fan_in) if kind == 'conv_act' else (0.5 / math.sqrt(fan_in) if kind == 'conv_res' else 1.0 / math.sqrt(fan_in))
            sd[name] = torch.randn(shape, generator=g) * std
        elif kind in ('bias', 'bn_b'):
            sd[name] = u(shape, -0.1, 0.1)
        elif kind in ('bn_w', 'bn_v'):
            sd[name] = u(shape, 0.8, 1.2)
        elif kind == 'bn_m':
            sd[name] = torch.randn(shape, generator=g) * 0.1
    return sd


# ------------------------------------------------------------------------------------------------ folding
def _fold(w, bias, sd, bn, eps=1e-5):
    """conv (+bias) followed by BatchNorm -> (w', b')"""
    g, b, m, v = (sd[f"{bn}.{k}"].float() for k in ("weight", "bias", "running_mean", "running_var"))
    s = g / torch.sqrt(v + eps)
    b0 = bias.float() if bias is not None else torch.zeros_like(m)
    return w.float() * s.view(-1, 1, 1, 1), (b0 - m) * s + b


class _C:
    def __init__(self, w, b, dev, groups=1, cin_pad=None):
        self.groups = groups
        self.w = E.pack_grouped_weight(w.to(dev), groups) if groups > 1 else E.pack_conv_weight(w.to(dev), torch.float16, cin_pad)
        # stride-1 grouped 3x3: compact diagonal blocks for the halo-tile kernel (N = K = channels per group, no zero multiplies)
        self.w_halo = E.pack_grouped_weight_compact(w.to(dev), groups) if groups > 1 else None
        self.b = None if b is None else b.float().contiguous().to(dev)

    def __call__(self, x, **kw):
        if self.w_halo is not None and kw.get('stride', 1) == 1:
            N, H, W, Cx = x.shape
            Cout, R, S, _ = self.w_halo.shape
            if E.halo_supported(N, H, W, Cout, Cx, kw.get('in_coff', 0), Cout, R, S, 1, kw.get('pad', 0), kw.get('dil', 1), self.groups):
                kw = {k: v for k, v in kw.items() if k != 'stride'}
                return E.conv2d_halo_nhwc(x, self.w_halo, self.b, groups=self.groups, **kw)
        return E.conv2d_nhwc(x, self.w, self.b, groups=self.groups, **kw)


class _FTB:
    def __init__(self, sd, name, dev):
        self.c1 = _C(sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], dev)
        self.ca = _C(*_fold(sd[f"{name}.conv_branch.1.weight"], sd[f"{name}.conv_branch.1.bias"], sd, f"{name}.conv_branch.2"), dev)
        self.cb = _C(sd[f"{name}.conv_branch.4.weight"], sd[f"{name}.conv_branch.4.bias"], dev)

    def __call__(self, x):
        y = self.c1(x, pad=1, act='relu')                        # relu(conv1(x)) -- the inplace-ReLU quirk
        t = self.ca(y, pad=1, act='relu')
        return self.cb(t, pad=1, act='relu', residual=y, res_mode=1)


STEM_S2D = __import__("os").environ.get("CSB_LERES_S2D", "1") != "0"       # space-to-depth form of the 7x7 stride-2 stem (see LeReS.__init__)
# channels of the space-to-depth stem input: 12 real ones padded to 16 (per-tap kernel, 25 TMA boxes of 32 B rows per tile: 1.25 ms per 32 frames) or to
# 64 (halo-tile kernel: one halo box per tile, the 25 taps are shifted descriptors on it: conv -0.45 ms, but the prep kernel writes 4x the bytes:
# +0.6 ms).  Measured on B200 (gpurun r2c32 / r2c33): 16 wins by 0.2-0.6 ms per step.
STEM_CP = int(__import__("os").environ.get("CSB_LERES_STEM_CP", "16"))


class LeReS:
    """B200 forward of RelDepthModel(backbone='resnext101').depth_model: [N,H,W,3] uint8 BGR -> [N,H,W] fp32 depth logits."""

    def __init__(self, state_dict=None, device='cuda'):
        sd = synthetic_state_dict(0) if state_dict is None else {k[7:] if k.startswith('module.') else k: v for k, v in state_dict.items()}
        dev = self.dev = torch.device(device)
        w7, b7 = _fold(sd[f"{ENC}.conv1.weight"], None, sd, f"{ENC}.bn1")
        self.stem = _C(w7, b7, dev, cin_pad=16)
        # The 7x7 stride-2 pad-3 stem (Resnext_torch.py:156) as a 5x5 stride-1 pad-2 conv over the 2x2 space-to-depth image (12 channels, padded to
        # 16): in[2y + r - 3] = s2d[y + a][dy] with r - 3 = 2a + dy, a in [-2, 1] (tap a = 2 is zero).  K shrinks from 49 x 16 to 25 x 16 and the
        # prep kernel writes a quarter of the bytes; same products, same zero padding.
        w5 = torch.zeros((w7.shape[0], 12, 5, 5), dtype=w7.dtype)
        for a in range(-2, 2):
            for dy in range(2):
                r = 2 * a + dy + 3
                if not 0 <= r <= 6:
                    continue
                for bb in range(-2, 2):
                    for dx in range(2):
                        s_ = 2 * bb + dx + 3
                        if 0 <= s_ <= 6:
                            w5[:, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3, a + 2, bb + 2] = w7[:, :, r, s_]
        self.stem_s2d = _C(w5, b7, dev, cin_pad=STEM_CP)
        self.layers = []
        for li, nblk in enumerate(LAYERS, start=1):
            blocks = []
            for b in range(nblk):
                p = f"{ENC}.layer{li}.{b}"
                blk = dict(c1=_C(*_fold(sd[f"{p}.conv1.weight"], None, sd, f"{p}.bn1"), dev),
                           c2=_C(*_fold(sd[f"{p}.conv2.weight"], None, sd, f"{p}.bn2"), dev, groups=32),
                           c3=_C(*_fold(sd[f"{p}.conv3.weight"], None, sd, f"{p}.bn3"), dev),
                           stride=2 if (b == 0 and li > 1) else 1, down=None)
                if f"{p}.downsample.0.weight" in sd:
                    blk['down'] = _C(*_fold(sd[f"{p}.downsample.0.weight"], None, sd, f"{p}.downsample.1"), dev)
                blocks.append(blk)
            self.layers.append(blocks)
        self.conv = _FTB(sd, f"{DEC}.conv", dev)
        self.conv1 = _C(sd[f"{DEC}.conv1.weight"], sd[f"{DEC}.conv1.bias"], dev)
        self.ffm = {n: (_FTB(sd, f"{DEC}.{n}.ftb1", dev), _FTB(sd, f"{DEC}.{n}.ftb2", dev)) for n in ("ffm2", "ffm1", "ffm0")}
        self.ao0 = _C(*_fold(sd[f"{DEC}.outconv.adapt_conv.0.weight"], sd[f"{DEC}.outconv.adapt_conv.0.bias"], sd, f"{DEC}.outconv.adapt_conv.1"), dev)
        self.ao1 = _C(sd[f"{DEC}.outconv.adapt_conv.3.weight"], sd[f"{DEC}.outconv.adapt_conv.3.bias"], dev)

    def encoder(self, x16, s2d=False):
        x = self.stem_s2d(x16, stride=1, pad=2, act='relu') if s2d else self.stem(x16, stride=2, pad=3, act='relu')
        x = E.maxpool3s2_nhwc(x)
        feats = []
        for blocks in self.layers:
            for blk in blocks:
                a = blk['c1'](x, act='relu')
                b = blk['c2'](a, stride=blk['stride'], pad=1, act='relu')
                identity = x if blk['down'] is None else blk['down'](x, stride=blk['stride'])
                x = blk['c3'](b, act='relu', residual=identity, res_mode=1)              # relu(bn3(conv3) + identity)
            feats.append(x)
        return feats

    def _ffm(self, name, low, high):
        f1, f2 = self.ffm[name]
        x = E.add_nhwc(f1(low), high)
        x = f2(x)
        return E.resample_nhwc(x, x.shape[1] * 2, x.shape[2] * 2, 'bilinear_ac')

    def forward(self, img_u8, rgb_input=False):
        """img_u8 [N,H,W,3] uint8 (BGR unless rgb_input), H and W multiples of 32 -> depth logits [N,H,W] fp32."""
        if img_u8.dim() == 3:
            img_u8 = img_u8[None]
        if rgb_input:
            return self._forward_impl(img_u8, True)
        if getattr(self, '_graphed', None) is None:                # batches <= 4 replay a CUDA graph per input shape (utils/graphs.py)
            from ..utils.graphs import GraphedForward
            self._graphed = GraphedForward(self._forward_impl)
        return self._graphed(img_u8.contiguous())

    def _forward_impl(self, img_u8, rgb_input=False):
        N, H, W, _ = img_u8.shape
        assert H % 32 == 0 and W % 32 == 0
        # estimateleres: BGR -> RGB (depthmap.py:35), ToTensor on float (no /255 again), Normalize(ImageNet) (:26) on img/255
        mean, std = [255.0 * m for m in IMAGENET_MEAN], [255.0 * s for s in IMAGENET_STD]
        if STEM_S2D:
            f = self.encoder(E.image_prep_s2d_nhwc(img_u8, mean, std, 2, STEM_CP, swap_rb=not rgb_input), s2d=True)
        else:
            f = self.encoder(E.image_prep_nhwc(img_u8, mean, std, swap_rb=not rgb_input, CP=16))
        x32 = self.conv1(self.conv(f[3]), pad=1)
        x16_ = E.resample_nhwc(x32, x32.shape[1] * 2, x32.shape[2] * 2, 'bilinear_ac')
        x8 = self._ffm("ffm2", f[2], x16_)
        x4 = self._ffm("ffm1", f[1], x8)
        x2 = self._ffm("ffm0", f[0], x4)
        a = self.ao0(x2, pad=1, act='relu')
        o = self.ao1(a, pad=1, out_f32=True)                                             # [N,H/2,W/2,1] fp32
        return E.resample_f32(o.view(N, H // 2, W // 2), H, W, True)


# ------------------------------------------------------------------------------------------------ reference-shaped host wrappers
_model = None


def apply_leres(input_image, thr_a: int = 0, thr_b: int = 0, boost: bool = False, device: str = 'cuda', model: LeReS = None):
    """depth_modules/leres/__init__.py:69-147.  input_image: HxWx3 float32 BGR in [0,1] (as _depth_est_leres passes it) or uint8 BGR.
    -> uint8 HxW 'depth image' (min-max normalised, 16->8 bit, inverted).  H, W multiples of 32."""
    import cv2
    global _model
    if boost:
        raise NotImplementedError("estimateboost / pix2pix merge is out of scope (boost=False at kenburns_effect.py:572)")
    if model is None:
        if _model is None:
            _model = LeReS(None, device)
        model = _model
    assert input_image.ndim == 3
    u8 = input_image if input_image.dtype == np.uint8 else np.clip(np.rint(input_image * 255.0), 0, 255).astype(np.uint8)
    depth = model.forward(torch.from_numpy(np.ascontiguousarray(u8)).to(model.dev))[0].cpu().numpy()        # estimateleres; same-size INTER_CUBIC is a copy
    return quantise_depth(depth, thr_a, thr_b)


def quantise_depth(depth, thr_a=0, thr_b=0):
    """apply_leres tail, :117-147 (numpy/OpenCV on the host exactly as the reference)."""
    import cv2
    depth_min, depth_max = depth.min(), depth.max()
    max_val = (2 ** 16) - 1
    if depth_max - depth_min > np.finfo("float").eps:
        out = max_val * (depth - depth_min) / (depth_max - depth_min)
    else:
        out = np.zeros(depth.shape)
    depth_image = cv2.convertScaleAbs(out.astype("uint16"), alpha=(255.0 / 65535.0))
    if thr_a != 0:
        depth_image = cv2.threshold(depth_image, ((thr_a / 100) * 255), 255, cv2.THRESH_TOZERO)[1]
    depth_image = cv2.bitwise_not(depth_image)
    if thr_b != 0:
        depth_image = cv2.threshold(depth_image, ((thr_b / 100) * 255), 255, cv2.THRESH_TOZERO)[1]
    return depth_image
