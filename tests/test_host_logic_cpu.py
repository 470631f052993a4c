"""CPU suite: host-side logic added around the CUDA path (no GPU, no compute calls into the library)."""
import numpy as np
import torch
import torch.nn.functional as F


def test_fold_layernorm_is_layernorm_then_linear():
    """engine.fold_layernorm: LayerNorm(gamma, beta) -> Linear(W, b) == rstd * (x W'^T - mean * colsum) + b' on the un-normalised x -- the identity the
    fc1 GEMM epilogue of the ConvNeXt block (csb_conv2d_ln_nhwc) relies on (mmpretrain ConvNeXtBlock: norm -> pointwise_conv1)."""
    from cartoonsegmentation_b200 import engine as E
    g = torch.Generator().manual_seed(0)
    C, Co, P = 128, 512, 37
    x = torch.randn(P, C, generator=g) * 2.0 + 1.5
    w, b = torch.randn(Co, C, generator=g) / C ** 0.5, torch.randn(Co, generator=g) * 0.1
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    ref = F.linear(F.layer_norm(x, (C,), gamma, beta, 1e-6), w, b)
    wp, bp, cs = E.fold_layernorm(w, b, gamma, beta, dtype=torch.float32)
    assert wp.shape == (Co, 1, 1, C) and bp.shape == (Co,) and cs.shape == (Co,)
    mean = x.mean(1, keepdim=True)
    rstd = torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-6)
    out = rstd * (x @ wp.reshape(Co, C).t() - mean * cs[None]) + bp[None]
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)
    # the column sums are those of the ROUNDED fp16 weights, so that the mean term cancels exactly against the fp16 GEMM
    wp16, _, cs16 = E.fold_layernorm(w, b, gamma, beta)
    assert wp16.dtype == torch.float16 and torch.equal(cs16, wp16.float().reshape(Co, C).sum(1))


def test_midas_net_size_matches_reference_resize():
    """`Resize(width, height, keep_aspect_ratio=True, ensure_multiple_of=32, resize_method='minimal').get_size` (midas.py:104-148) for the sizes on
    the path: the Ken-Burns pipeline's img_size [672, 672] on a reflect-padded 1024^2 frame, MidasCore's default 384, load_zoe's [512, 672]."""
    from cartoonsegmentation_b200.depth_modules.zoedepth import midas_net_size

    def ref(h, w, net_h, net_w, m=32):
        sh, sw = net_h / h, net_w / w
        if abs(1 - sw) < abs(1 - sh):
            sh = sw
        else:
            sw = sh
        c = lambda v: int(np.round(v / m) * m)
        return c(sh * h), c(sw * w)

    pad = int(np.sqrt(1024 / 2) * 3)
    assert midas_net_size(1024 + 2 * pad, 1024 + 2 * pad, [672, 672]) == (672, 672)
    assert midas_net_size(1158, 1158) == (384, 384) == midas_net_size(1158, 1158, 384)
    for (h, w, net) in [(720, 960, (512, 672)), (480, 640, (384, 384)), (1000, 700, (672, 672)), (333, 517, (512, 672))]:
        assert midas_net_size(h, w, net) == ref(h, w, *net)


def test_bench_workload_config_names_the_stages():
    import bench
    c = bench.workload_config(32, 8, ("seg", "depth", "warp"), "leres")
    assert c["stages"] == ["seg", "depth", "warp"] and c["depth"] == "leres" and "LeReS" in c["workload"] and c["batch_frames_per_step"] == 32
    z = bench.workload_config(32, 8, ("seg", "depth", "warp"), "zoe")
    assert "ZoeDepth" in z["workload"] and "672" in z["workload"]
    w = bench.workload_config(4, 2, ("warp",))
    assert w["depth"] is None and any("seg" in m for m in w["stages_missing"])


def test_zoe_cond_input_channel_permutation_is_a_permutation():
    """ZoeHead feeds ConditionalLogBinomial.mlp.0 the channel order [feat 0..31 | embedding 32..159 | rel 160] (csb_zoe_cond_input); the reference
    concatenates [feat, rel, embedding] (zoedepth_v1.py:176-184).  The weight columns must be the matching permutation."""
    w0 = torch.arange(161, dtype=torch.float32)[None].repeat(3, 1)
    perm = torch.cat([w0[:, :32], w0[:, 33:161], w0[:, 32:33]], 1)
    assert perm.shape == w0.shape and sorted(perm[0].tolist()) == list(range(161))
    assert perm[0, 160] == 32 and perm[0, 32] == 33 and perm[0, 159] == 160


def test_leres_stem_space_to_depth_rearrangement():
    """LeReS.__init__: the 7x7 stride-2 pad-3 stem (Resnext_torch.py:156) == a 5x5 stride-1 pad-2 conv over the 2x2 space-to-depth image with the
    weights rearranged as in depth_modules/leres.py (channel (dy*2+dx)*3 + c = pixel (2Y+dy, 2X+dx), what csb_image_prep_s2d_nhwc writes)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 32, 48, generator=g)
    w7 = torch.randn(8, 3, 7, 7, generator=g)
    w5 = torch.zeros(8, 12, 5, 5)
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
    s2d = torch.stack([x[:, :, dy::2, dx::2] for dy in range(2) for dx in range(2)], 1).reshape(2, 12, 16, 24)
    assert torch.allclose(F.conv2d(s2d, w5, stride=1, padding=2), F.conv2d(x, w7, stride=2, padding=3), atol=1e-4)


def _constants_from(path, pattern):
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), path)).read()
    return [float(x) for x in re.findall(pattern, src)]


def test_gelu_sigmoid_form_constants_reproduce_exact_gelu():
    """csrc/tc_conv.cu gelu2_mufu: gelu(x) = x / (1 + 2^(x * q(min(x^2, 64)))) with the three constants in the source (which carry -2 log2 e): the
    float32 evaluation stays within 3e-5 of the exact erf GELU (nn.GELU(), mmpretrain ConvNeXtBlock) over [-12, 12], saturates correctly beyond."""
    from scipy.special import erf
    body = open(__file__.replace("tests/test_host_logic_cpu.py", "cartoonsegmentation_b200/csrc/tc_conv.cu")).read()
    body = body[body.index("void gelu2_mufu"):body.index("#ifndef CSB_GELU_MUFU_PAIRS")]
    import re
    c5, c3, c1 = [np.float32(v) for v in re.findall(r"CSB_C2\((-?[0-9.eE+-]+)f\)", body)[:3]]
    assert c1 < 0 and abs(c1 / (-2 * np.log2(np.e)) - 0.7975) < 1e-3          # leading coefficient ~ sqrt(2/pi)
    x = np.linspace(-12, 12, 200001).astype(np.float32)
    x2 = np.minimum(x * x, np.float32(64))
    q = (c5 * x2 + c3) * x2 + c1
    with np.errstate(over='ignore'):
        y = x * (np.float32(1) / (np.float32(1) + np.exp2((q * x).astype(np.float64)).astype(np.float32)))
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    assert np.abs(y - ref).max() < 3e-5
    for big in (-1e4, -50.0, 50.0, 1e4):                                      # exp overflow -> 0, underflow -> x
        xb = np.float32(big)
        qb = (c5 * np.float32(64) + c3) * np.float32(64) + c1
        with np.errstate(over='ignore'):
            eb = np.float32(np.exp2(np.float64(qb * xb)))                      # ex2.approx.ftz.f32: overflows to +inf, underflows to 0
            yb = xb * (np.float32(1) / (np.float32(1) + eb))
        assert yb == (0.0 if big < 0 else big)


def test_lazy_rescaling_online_softmax_matches_attention():
    """The algorithm of csrc/zoe_attn_tc.cu restated in numpy: 64-key blocks, scores in log2 units, a reference maximum that only moves (and
    rescales O and l) when a row's block maximum exceeds it by more than 8, P rounded to fp16 before P V, fp32 accumulation -- against plain
    softmax attention with an additive bias (timm BEiT Attention as MiDaS patches it)."""
    rng = np.random.default_rng(0)
    T, d = 300, 64
    q, k, v = (rng.standard_normal((T, d)).astype(np.float32) * 1.5 for _ in range(3))
    bias = (rng.standard_normal((T, T)) * 4).astype(np.float32)
    bias[:, 5] += 40.0                                                        # a late, dominant key forces rescaling steps
    scale = d ** -0.5
    s = q @ k.T * scale + bias
    ref = np.exp(s - s.max(1, keepdims=True))
    ref = (ref / ref.sum(1, keepdims=True)) @ v
    log2e = np.float32(1.4426950408889634)
    Tp = (T + 63) // 64 * 64
    sp = np.full((T, Tp), -60000.0, np.float32)
    sp[:, :T] = bias
    kp = np.zeros((Tp, d), np.float32); kp[:T] = k
    vp = np.zeros((Tp, d), np.float32); vp[:T] = v
    m_ref = np.full(T, -np.inf, np.float32)
    l = np.zeros(T, np.float32)
    o = np.zeros((T, d), np.float32)
    rescales = 0
    for j in range(Tp // 64):
        t = (q @ kp[j * 64:(j + 1) * 64].T) * np.float32(scale) * log2e + sp[:, j * 64:(j + 1) * 64] * log2e
        mx = t.max(1)
        for w in range(0, T, 32):                                             # the decision is per warp (32 rows)
            rows = slice(w, min(T, w + 32))
            if np.any(mx[rows] - m_ref[rows] > 8.0):
                m_new = np.maximum(m_ref[rows], mx[rows])
                with np.errstate(invalid='ignore'):
                    alpha = np.where(np.isinf(m_ref[rows]), 0.0, np.exp2(m_ref[rows] - m_new)).astype(np.float32)
                m_ref[rows] = m_new
                l[rows] *= alpha
                o[rows] *= alpha[:, None]
                rescales += 1
        p = np.exp2(t - m_ref[:, None]).astype(np.float32)
        assert p.max() <= 256.0 * 1.0001                                       # what keeps fp16 P exact enough
        l += p.sum(1)
        o += p.astype(np.float16).astype(np.float32) @ vp[j * 64:(j + 1) * 64]
    out = o / l[:, None]
    rr = np.sqrt(((out - ref) ** 2).mean() / (ref ** 2).mean())
    assert rr < 1e-3 and rescales < (Tp // 64) * ((T + 31) // 32)              # exact enough, and genuinely lazy


def test_bench_roofline_accounting_from_launch_labels():
    """bench.aggregate_detail: FLOPs / compulsory bytes of the tensor-core launches from their detailed profile labels (what roofline.achieved,
    algorithmic_bytes_per_launch and conv_engine are built from)."""
    import bench
    prof = {
        "k_conv_tc[32x64x64x512->2048 k1x1 s1 d1 g1 act3 res0]": {"ms": 10.0, "count": 27},
        "k_conv_tc[32x128x128x256->256 k3x3 s2 d1 g1 act2 res2]": {"ms": 1.0, "count": 1},
        "k_conv_halo[32x40x40x1024->1024 k3x3 s1 d1 g32 act1 res0]": {"ms": 2.0, "count": 22},
        "k_mlp_tc[2097152x128->512->128]": {"ms": 3.6, "count": 3},
        "k_layernorm": {"ms": 1.2, "count": 7},
    }
    agg = bench.aggregate_detail(prof)
    px = 32 * 64 * 64
    assert agg["k_conv_tc"]["count"] == 28 and abs(agg["k_conv_tc"]["ms"] - 11.0) < 1e-9
    g1 = 27 * 2.0 * px * 2048 * 512 / 1e9
    g2 = 2.0 * 32 * 64 * 64 * 256 * 9 * 256 / 1e9                               # stride 2: 64 x 64 output pixels
    assert abs(agg["k_conv_tc"]["gflop"] - (g1 + g2)) < 1e-6 * (g1 + g2)
    b1 = 27 * (2.0 * px * (512 + 2048) + 2.0 * 2048 * 512) / 1e9
    b2 = (2.0 * 32 * (128 * 128 * 256 + 64 * 64 * 256 * 2) + 2.0 * 256 * 9 * 256) / 1e9      # residual read counted
    assert abs(agg["k_conv_tc"]["gbytes"] - (b1 + b2)) < 1e-6 * (b1 + b2)
    gh = 22 * 2.0 * 32 * 40 * 40 * 1024 * 9 * (1024 // 32) / 1e9                # grouped: the group-sparse count
    assert abs(agg["k_conv_halo"]["gflop"] - gh) < 1e-6 * gh
    gm = 3 * 2.0 * 2.0 * 2097152 * 128 * 512 / 1e9                              # two GEMMs per fused MLP launch
    assert abs(agg["k_mlp_tc"]["gflop"] - gm) < 1e-6 * gm
    bm = 3 * (3.0 * 2097152 * 128 * 2 + 2.0 * 128 * 512 * 2 + 2097152 * 2 * 8) / 1e9
    assert abs(agg["k_mlp_tc"]["gbytes"] - bm) < 1e-6 * bm
    assert agg["k_layernorm"] == {"ms": 1.2, "count": 7, "gflop": 0.0}
