"""TEST INFRASTRUCTURE ONLY -- CPU oracle (plain PyTorch fp32, NCHW) of the RTMDet-Ins detector that
`AnimeInsSeg.infer` runs.

Why a restatement: the detector's arithmetic is NOT in the reference repo.  `AnimeInsSeg.__init__` builds it from a config
string inside the checkpoint through mmdet/mmcv/mmengine (animeinsseg/__init__.py:196-209), none of which is installed
here or vendored.  Parity status, stated per piece:

  * mask head `_mask_predict_by_feat_single`: follows the reference's own code line by line
    (animeinsseg/models/rtmdet_inshead_custom.py:253-303) -- pinned by the reference source.
  * mask tail (x8 bilinear, resize, crop, sigmoid, threshold): follows the reference's own restatement
    (animeinsseg/__init__.py:361-370) -- pinned by the reference source.
  * `_det_forward` tail (score filter, int32 truncation, xyxy->xywh): animeinsseg/__init__.py:451-462 -- pinned.
  * ConvNeXt-B backbone: mmpretrain `ConvNeXt(arch='base', out_indices=[1,2,3])` restated from SURVEY.md Appendix A.4;
    cross-checked in tests against torchvision.models.convnext_base (independent implementation of the same block math).
  * CSPNeXtPAFPN neck, RTMDetInsSepBNHead, predict_by_feat / batched NMS (mmdet 3.3.0, mmcv 2.1.0): restated from
    SURVEY.md Appendix A.5-A.7; NMS cross-checked against torchvision.ops.batched_nms.  PARITY UNPINNED against the real
    mmdet package (absent); every recalled choice is one named function / constructor argument here.

Module attribute names follow mmdet's so that a real checkpoint's state_dict keys line up (`backbone.*`, `neck.*`, `bbox_head.*`).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- building blocks
class LayerNorm2d(nn.LayerNorm):
    """mmpretrain LayerNorm2d: LayerNorm over the channel dim of NCHW."""

    def forward(self, x):
        return F.layer_norm(x.permute(0, 2, 3, 1), self.normalized_shape, self.weight, self.bias, self.eps).permute(0, 3, 1, 2)


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule: conv (bias only if no norm) -> BN -> act."""

    def __init__(self, cin, cout, k, stride=1, padding=0, groups=1, act='silu', norm=True, bn_eps=1e-5):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, padding, groups=groups, bias=not norm)
        self.bn = nn.BatchNorm2d(cout, eps=bn_eps) if norm else None
        self.act = act

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        if self.act == 'silu':
            x = F.silu(x)
        elif self.act == 'relu':
            x = F.relu(x)
        return x


class DepthwiseSeparableConvModule(nn.Module):
    def __init__(self, cin, cout, k, padding, act='silu'):
        super().__init__()
        self.depthwise_conv = ConvModule(cin, cin, k, 1, padding, groups=cin, act=act)
        self.pointwise_conv = ConvModule(cin, cout, 1, act=act)

    def forward(self, x):
        return self.pointwise_conv(self.depthwise_conv(x))


class CSPNeXtBlock(nn.Module):
    """3x3 ConvModule -> depthwise 5x5 ConvModule -> pointwise 1x1 ConvModule (+x if add_identity)."""

    def __init__(self, cin, cout, add_identity, act='silu'):
        super().__init__()
        self.conv1 = ConvModule(cin, cout, 3, 1, 1, act=act)
        self.conv2 = DepthwiseSeparableConvModule(cout, cout, 5, 2, act=act)
        self.add_identity = add_identity and cin == cout

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        return out + x if self.add_identity else out


class ChannelAttention(nn.Module):
    """mmdet `ChannelAttention` ‡ (layers/se_layer.py): x * Hardsigmoid(Conv2d(C, C, 1, bias=True)(AdaptiveAvgPool2d(1)(x)))."""

    def __init__(self, channels):
        super().__init__()
        self.fc = nn.Conv2d(channels, channels, 1, 1, 0, bias=True)

    def forward(self, x):
        return x * F.hardsigmoid(self.fc(x.mean((2, 3), keepdim=True)))


class CSPLayer(nn.Module):
    def __init__(self, cin, cout, num_blocks, add_identity, expand_ratio=0.5, act='silu', channel_attention=False):
        super().__init__()
        mid = int(cout * expand_ratio)
        self.main_conv = ConvModule(cin, mid, 1, act=act)
        self.short_conv = ConvModule(cin, mid, 1, act=act)
        self.final_conv = ConvModule(2 * mid, cout, 1, act=act)
        self.blocks = nn.Sequential(*[CSPNeXtBlock(mid, mid, add_identity, act) for _ in range(num_blocks)])
        if channel_attention:
            self.attention = ChannelAttention(2 * mid)

    def forward(self, x):
        x = torch.cat((self.blocks(self.main_conv(x)), self.short_conv(x)), dim=1)
        if hasattr(self, 'attention'):
            x = self.attention(x)
        return self.final_conv(x)


class SPPBottleneck(nn.Module):
    """mmdet `SPPBottleneck` ‡ (backbones/csp_darknet.py): 1x1 (C -> C/2), cat(x, MaxPool 5/9/13, stride 1, pad k//2), 1x1 (2C -> C)."""

    def __init__(self, cin, cout, kernel_sizes=(5, 9, 13), act='silu'):
        super().__init__()
        mid = cin // 2
        self.conv1 = ConvModule(cin, mid, 1, act=act)
        self.poolings = nn.ModuleList([nn.MaxPool2d(k, 1, k // 2) for k in kernel_sizes])
        self.conv2 = ConvModule(mid * (len(kernel_sizes) + 1), cout, 1, act=act)

    def forward(self, x):
        x = self.conv1(x)
        return self.conv2(torch.cat([x] + [p(x) for p in self.poolings], dim=1))


class CSPNeXt(nn.Module):
    """mmdet `CSPNeXt(arch='P5', deepen_factor=1, widen_factor=1, expand_ratio=.5, channel_attention=True, out_indices=(2,3,4))` ‡
    -- the backbone of the shipped rtmdetl_e60.ckpt."""
    ARCH = ((64, 128, 3, True, False), (128, 256, 6, True, False), (256, 512, 6, True, False), (512, 1024, 3, False, True))

    def __init__(self, act='silu'):
        super().__init__()
        self.stem = nn.Sequential(ConvModule(3, 32, 3, 2, 1, act=act), ConvModule(32, 32, 3, 1, 1, act=act), ConvModule(32, 64, 3, 1, 1, act=act))
        for i, (cin, cout, n, identity, spp) in enumerate(self.ARCH, 1):
            stage = [ConvModule(cin, cout, 3, 2, 1, act=act)]
            if spp:
                stage.append(SPPBottleneck(cout, cout, act=act))
            stage.append(CSPLayer(cout, cout, n, identity, 0.5, act, channel_attention=True))
            setattr(self, f'stage{i}', nn.Sequential(*stage))

    def forward(self, x):
        outs = []
        x = self.stem(x)
        for i in range(1, 5):
            x = getattr(self, f'stage{i}')(x)
            if i >= 2:
                outs.append(x)
        return outs


# ----------------------------------------------------------------------------------------------- backbone (A.4)
class ConvNeXtBlock(nn.Module):
    def __init__(self, dim, layer_scale_init_value=1.0):
        super().__init__()
        self.depthwise_conv = nn.Conv2d(dim, dim, 7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pointwise_conv1 = nn.Linear(dim, 4 * dim)
        self.pointwise_conv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim))

    def forward(self, x):
        s = x
        x = self.depthwise_conv(x).permute(0, 2, 3, 1)
        x = self.pointwise_conv2(F.gelu(self.pointwise_conv1(self.norm(x))))
        return s + (x * self.gamma).permute(0, 3, 1, 2)


class ConvNeXt(nn.Module):
    def __init__(self, depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024), out_indices=(1, 2, 3)):
        super().__init__()
        self.out_indices = out_indices
        self.downsample_layers = nn.ModuleList([nn.Sequential(nn.Conv2d(3, dims[0], 4, 4), LayerNorm2d(dims[0], eps=1e-6))])
        for i in range(1, 4):
            self.downsample_layers.append(nn.Sequential(LayerNorm2d(dims[i - 1], eps=1e-6), nn.Conv2d(dims[i - 1], dims[i], 2, 2)))
        self.stages = nn.ModuleList([nn.Sequential(*[ConvNeXtBlock(dims[i]) for _ in range(depths[i])]) for i in range(4)])
        for i in out_indices:
            setattr(self, f'norm{i}', LayerNorm2d(dims[i], eps=1e-6))

    def forward(self, x):
        outs = []
        for i in range(4):
            x = self.stages[i](self.downsample_layers[i](x))
            if i in self.out_indices:
                outs.append(getattr(self, f'norm{i}')(x))
        return outs


# ----------------------------------------------------------------------------------------------- neck (A.5)
class CSPNeXtPAFPN(nn.Module):
    def __init__(self, in_channels=(256, 512, 1024), out_channels=256, num_csp_blocks=3, act='silu'):
        super().__init__()
        c = in_channels
        self.reduce_layers = nn.ModuleList([ConvModule(c[2], c[1], 1, act=act), ConvModule(c[1], c[0], 1, act=act)])
        self.top_down_blocks = nn.ModuleList([CSPLayer(c[1] * 2, c[1], num_csp_blocks, False, act=act), CSPLayer(c[0] * 2, c[0], num_csp_blocks, False, act=act)])
        self.downsamples = nn.ModuleList([ConvModule(c[0], c[0], 3, 2, 1, act=act), ConvModule(c[1], c[1], 3, 2, 1, act=act)])
        self.bottom_up_blocks = nn.ModuleList([CSPLayer(c[0] * 2, c[1], num_csp_blocks, False, act=act), CSPLayer(c[1] * 2, c[2], num_csp_blocks, False, act=act)])
        self.out_convs = nn.ModuleList([ConvModule(ch, out_channels, 3, 1, 1, act=act) for ch in c])

    def forward(self, inputs):
        inner_outs = [inputs[2]]
        for idx in (2, 1):
            feat_high = self.reduce_layers[2 - idx](inner_outs[0])
            inner_outs[0] = feat_high
            up = F.interpolate(feat_high, scale_factor=2, mode='nearest')
            inner_outs.insert(0, self.top_down_blocks[2 - idx](torch.cat([up, inputs[idx - 1]], 1)))
        outs = [inner_outs[0]]
        for idx in (0, 1):
            down = self.downsamples[idx](outs[-1])
            outs.append(self.bottom_up_blocks[idx](torch.cat([down, inner_outs[idx + 1]], 1)))
        return [conv(o) for conv, o in zip(self.out_convs, outs)]


# ----------------------------------------------------------------------------------------------- head (A.6)
class MaskFeatModule(nn.Module):
    def __init__(self, in_channels=256, feat_channels=256, stacked_convs=4, num_levels=3, num_prototypes=8, act='silu'):
        super().__init__()
        self.fusion_conv = nn.Conv2d(num_levels * in_channels, in_channels, 1)
        self.stacked_convs = nn.Sequential(*[ConvModule(in_channels if i == 0 else feat_channels, feat_channels, 3, 1, 1, act=act) for i in range(stacked_convs)])
        self.projection = nn.Conv2d(feat_channels, num_prototypes, 1)

    def forward(self, feats):
        size = feats[0].shape[-2:]
        fused = [feats[0]] + [F.interpolate(f, size=size, mode='bilinear') for f in feats[1:]]
        return self.projection(self.stacked_convs(self.fusion_conv(torch.cat(fused, 1))))


class RTMDetInsSepBNHead(nn.Module):
    def __init__(self, num_classes=1, in_channels=256, feat_channels=256, stacked_convs=2, strides=(8, 16, 32), num_prototypes=8, dyconv_channels=8,
                 num_dyconvs=3, act='silu'):
        super().__init__()
        self.strides, self.num_prototypes, self.dyconv_channels, self.num_dyconvs = strides, num_prototypes, dyconv_channels, num_dyconvs
        self.num_gen_params = (num_prototypes + 2) * dyconv_channels + dyconv_channels * dyconv_channels * (num_dyconvs - 2) + dyconv_channels \
            + dyconv_channels * (num_dyconvs - 1) + 1     # 169
        L = len(strides)

        def tower():
            return nn.ModuleList([nn.ModuleList([ConvModule(in_channels if i == 0 else feat_channels, feat_channels, 3, 1, 1, act=act) for i in range(stacked_convs)])
                                  for _ in range(L)])
        self.cls_convs, self.reg_convs, self.kernel_convs = tower(), tower(), tower()
        self.rtm_cls = nn.ModuleList([nn.Conv2d(feat_channels, num_classes, 1) for _ in range(L)])
        self.rtm_reg = nn.ModuleList([nn.Conv2d(feat_channels, 4, 1) for _ in range(L)])
        self.rtm_kernel = nn.ModuleList([nn.Conv2d(feat_channels, self.num_gen_params, 1) for _ in range(L)])
        for towers in (self.cls_convs, self.reg_convs, self.kernel_convs):     # share_conv=True: conv weights shared across levels, BN per level
            for n in range(1, L):
                for i in range(stacked_convs):
                    towers[n][i].conv = towers[0][i].conv
        self.mask_head = MaskFeatModule(in_channels, feat_channels, 4, L, num_prototypes, act)

    def forward(self, feats):
        mask_feat = self.mask_head(feats)
        cls, reg, ker = [], [], []
        for idx, (x, stride) in enumerate(zip(feats, self.strides)):
            c, r, k = x, x, x
            for m in self.cls_convs[idx]:
                c = m(c)
            for m in self.kernel_convs[idx]:
                k = m(k)
            for m in self.reg_convs[idx]:
                r = m(r)
            cls.append(self.rtm_cls[idx](c))
            ker.append(self.rtm_kernel[idx](k))
            reg.append(F.relu(self.rtm_reg[idx](r)) * stride)
        return cls, reg, ker, mask_feat


class RTMDetIns(nn.Module):
    def __init__(self, backbone='convnext_b'):
        super().__init__()
        assert backbone in ('convnext_b', 'cspnext_l')
        self.backbone = ConvNeXt() if backbone == 'convnext_b' else CSPNeXt()
        self.neck = CSPNeXtPAFPN()
        self.bbox_head = RTMDetInsSepBNHead()

    def forward(self, x):
        return self.bbox_head(self.neck(self.backbone(x)))


# ----------------------------------------------------------------------------------------------- pre / post-processing
DEFAULT_TEST_CFG = dict(nms_pre=1000, score_thr=0.05, iou_threshold=0.6, max_per_img=100, min_bbox_size=0, mask_thr_binary=0.5)
MEAN_BGR, STD_BGR = (103.53, 116.28, 123.675), (57.375, 57.12, 58.395)


def preprocess(img_bgr_u8):
    """DetDataPreprocessor for det_size == image size (Resize/Pad are identities, SURVEY §8 preamble): [H,W,3] u8 -> [1,3,H,W] f32."""
    x = torch.from_numpy(img_bgr_u8).float().permute(2, 0, 1)[None]
    mean = torch.tensor(MEAN_BGR).view(1, 3, 1, 1)
    std = torch.tensor(STD_BGR).view(1, 3, 1, 1)
    return (x - mean) / std


def grid_priors(h, w, stride):
    """MlvlPointGenerator(offset=0).single_level_grid_priors(with_stride=True): [x, y, stride, stride], row-major."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32) * stride, torch.arange(w, dtype=torch.float32) * stride, indexing='ij')
    s = torch.full_like(xs, float(stride))
    return torch.stack([xs, ys, s, s], -1).view(-1, 4)


def nms_greedy(boxes, scores, iou_thr):
    """mmcv.ops.nms semantics: sort by score desc (stable), greedy, suppress IoU > thr, IoU without +1 offset."""
    order = torch.sort(scores, descending=True, stable=True)[1]
    b = boxes[order]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep, dead = [], torch.zeros(len(b), dtype=torch.bool)
    for i in range(len(b)):
        if dead[i]:
            continue
        keep.append(i)
        xx1 = torch.maximum(b[i, 0], b[i + 1:, 0]); yy1 = torch.maximum(b[i, 1], b[i + 1:, 1])
        xx2 = torch.minimum(b[i, 2], b[i + 1:, 2]); yy2 = torch.minimum(b[i, 3], b[i + 1:, 3])
        inter = (xx2 - xx1).clamp(min=0) * (yy2 - yy1).clamp(min=0)
        iou = inter / (area[i] + area[i + 1:] - inter)
        dead[i + 1:] |= iou > iou_thr
    return order[torch.tensor(keep, dtype=torch.long)]


def decode_and_select(cls, reg, ker, strides, img_hw, cfg):
    """predict_by_feat up to (and including) NMS for one image.  Inputs: per-level [1,C,h,w] tensors.
    Returns boxes [K,4], scores [K], labels [K], kernels [K,169], priors [K,4] in descending score order."""
    sel = []
    for c, r, k, stride in zip(cls, reg, ker, strides):
        h, w = c.shape[-2:]
        scores = c[0].permute(1, 2, 0).reshape(-1, c.shape[1]).sigmoid()
        dist = r[0].permute(1, 2, 0).reshape(-1, 4)
        kern = k[0].permute(1, 2, 0).reshape(-1, k.shape[1])
        pri = grid_priors(h, w, stride)
        valid = scores > cfg['score_thr']                                  # filter_scores_and_topk
        idx = valid.nonzero()
        s = scores[valid]
        s, order = torch.sort(s, descending=True, stable=True)
        order = order[:cfg['nms_pre']]
        s = s[:cfg['nms_pre']]
        keep_idx, labels = idx[order, 0], idx[order, 1]
        sel.append((s, labels, dist[keep_idx], pri[keep_idx], kern[keep_idx]))
    s = torch.cat([t[0] for t in sel]); labels = torch.cat([t[1] for t in sel]); dist = torch.cat([t[2] for t in sel])
    pri = torch.cat([t[3] for t in sel]); kern = torch.cat([t[4] for t in sel])
    H, W = img_hw
    boxes = torch.stack([pri[:, 0] - dist[:, 0], pri[:, 1] - dist[:, 1], pri[:, 0] + dist[:, 2], pri[:, 1] + dist[:, 3]], -1)   # distance2bbox
    boxes[:, 0::2] = boxes[:, 0::2].clamp(0, W)
    boxes[:, 1::2] = boxes[:, 1::2].clamp(0, H)
    if cfg['min_bbox_size'] >= 0:
        wh = boxes[:, 2:] - boxes[:, :2]
        ok = (wh[:, 0] > cfg['min_bbox_size']) & (wh[:, 1] > cfg['min_bbox_size'])
        boxes, s, labels, kern, pri = boxes[ok], s[ok], labels[ok], kern[ok], pri[ok]
    if len(boxes) == 0:
        return boxes, s, labels, kern, pri
    offs = labels.to(boxes) * (boxes.max() + 1)                            # batched_nms class-offset trick
    keep = nms_greedy(boxes + offs[:, None], s, cfg['iou_threshold'])[:cfg['max_per_img']]
    return boxes[keep], s[keep], labels[keep], kern[keep], pri[keep]


def parse_dynamic_params(flatten_kernels, num_prototypes=8, dyconv_channels=8, num_dyconvs=3):
    """mmdet RTMDetInsHead.parse_dynamic_params: split order [w0(80), w1(64), w2(8), b0(8), b1(8), b2(1)]."""
    n_inst = flatten_kernels.size(0)
    weight_nums = [(num_prototypes + 2) * dyconv_channels] + [dyconv_channels * dyconv_channels] * (num_dyconvs - 2) + [dyconv_channels]
    bias_nums = [dyconv_channels] * (num_dyconvs - 1) + [1]
    splits = list(torch.split_with_sizes(flatten_kernels, weight_nums + bias_nums, dim=1))
    weights, biases = splits[:num_dyconvs], splits[num_dyconvs:]
    for i in range(num_dyconvs):
        if i < num_dyconvs - 1:
            weights[i] = weights[i].reshape(n_inst * dyconv_channels, -1, 1, 1)
            biases[i] = biases[i].reshape(n_inst * dyconv_channels)
        else:
            weights[i] = weights[i].reshape(n_inst, -1, 1, 1)
            biases[i] = biases[i].reshape(n_inst)
    return weights, biases


def mask_predict_by_feat_single(mask_feat, kernels, priors, stride0=8):
    """RTMDetInsSepBNHeadCustom._mask_predict_by_feat_single, animeinsseg/models/rtmdet_inshead_custom.py:253-303.
    mask_feat [8,h,w], kernels [K,169], priors [K,4] -> logits [K,h,w]."""
    num_inst = kernels.shape[0]
    h, w = mask_feat.size()[-2:]
    if num_inst < 1:
        return torch.empty(size=(num_inst, h, w), dtype=mask_feat.dtype)
    coord = grid_priors(h, w, stride0)[:, :2].reshape(1, -1, 2)                          # :268
    points = priors[:, :2].reshape(-1, 1, 2)
    strides = priors[:, 2].reshape(-1, 1, 1)
    relative_coord = (points - coord).permute(0, 2, 1) / (strides * 8)                   # :273-275
    relative_coord = relative_coord.reshape(num_inst, 2, h, w)
    x = torch.cat([relative_coord, mask_feat.unsqueeze(0).repeat(num_inst, 1, 1, 1)], dim=1)   # :277-278
    weights, biases = parse_dynamic_params(kernels)
    n_layers = len(weights)
    x = x.reshape(1, -1, h, w)
    for i, (weight, bias) in enumerate(zip(weights, biases)):                            # :286-294
        x = F.conv2d(x, weight, bias=bias, stride=1, padding=0, groups=num_inst)
        if i < n_layers - 1:
            x = F.relu(x)
    return x.reshape(num_inst, h, w)


def mask_tail(logits, ori_hw, scale_factor=(1.0, 1.0), stride0=8, mask_thr_binary=0.5):
    """animeinsseg/__init__.py:361-370: x8 bilinear -> resize to ceil(size / scale) -> crop -> sigmoid -> > thr.  [K,h,w] -> bool [K,H,W]."""
    if logits.shape[0] == 0:
        return torch.zeros((0, ori_hw[0], ori_hw[1]), dtype=torch.bool)
    m = F.interpolate(logits.unsqueeze(0), scale_factor=stride0, mode='bilinear')
    sf = [1.0 / s for s in scale_factor]
    m = F.interpolate(m, size=[math.ceil(m.shape[-2] * sf[0]), math.ceil(m.shape[-1] * sf[1])], mode='bilinear', align_corners=False)[..., :ori_hw[0], :ori_hw[1]]
    return (m.sigmoid().squeeze(0) > mask_thr_binary)


def det_forward_tail(boxes, scores, masks, pred_score_thr=0.3):
    """AnimeInsSeg._det_forward tail, animeinsseg/__init__.py:451-462: score filter, int32 truncation, xyxy -> xywh."""
    keep = scores > pred_score_thr
    boxes, scores, masks = boxes[keep], scores[keep], masks[keep]
    b = boxes.to(torch.int32)
    b[:, 2:] -= b[:, :2]
    return masks, b, scores


@torch.no_grad()
def infer(model, img_bgr_u8, cfg=None, pred_score_thr=0.3):
    """AnimeInsSeg.infer (refine off) for one image with det_size == image size -> dict(masks bool [K,H,W], bboxes int32 xywh, scores, + intermediates)."""
    cfg = dict(DEFAULT_TEST_CFG, **(cfg or {}))
    H, W = img_bgr_u8.shape[:2]
    cls, reg, ker, mask_feat = model(preprocess(img_bgr_u8))
    boxes, scores, labels, kern, pri = decode_and_select(cls, reg, ker, model.bbox_head.strides, (H, W), cfg)
    logits = mask_predict_by_feat_single(mask_feat[0], kern, pri)
    masks = mask_tail(logits, (H, W), mask_thr_binary=cfg['mask_thr_binary'])
    m, b, s = det_forward_tail(boxes, scores, masks, pred_score_thr)
    return dict(masks=m, bboxes=b, scores=s, raw=dict(cls=cls, reg=reg, ker=ker, mask_feat=mask_feat, boxes=boxes, all_scores=scores, kernels=kern,
                                                       priors=pri, logits=logits))
