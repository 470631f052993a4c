#!/usr/bin/env python
"""bench.py -- hot-path throughput on N B200s (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # CPU arm: the oracle port of the same path on the host cores

A "step" is one pass of the hot path over one batch of `--batch` synthetic 1024x1024 input frames.  Workload of the headline line
(`config.workload`, BASELINE.json metric "seg+depth+warp"): per input frame the AnimeInsSeg.infer body (ConvNeXt-B RTMDet-Ins forward,
decode / NMS / dynamic mask head / x8 mask tail), the depth forward (`--depth leres` default: ResNeXt-101 32x8d + decoder and the
reference's 16->8 bit tail; `--depth zoe`: DPT-BEiT-L + metric bins head with pad + flip augmentation), instance-guided depth
flattening, disparity -> point cloud (kenburns_effect.py:928-937), camera shift, z-buffered splat render, disocclusion fill, uint8
pack, centre crop + bilinear resize (kenburns_effect.py:1028-1040,1069-1070).

The same JSON line carries `other_workloads` (skipped with --no-other), each measured through the reference-facing Python API with
host buffers on both sides:
  infer_batch32    BASELINE configs[1]: AnimeInsSeg.infer(list of 32 numpy images), refine off; and with the default-on ISNet refine
  kenburns_full    BASELINE configs[3]: KenBurnsPipeline.generate_kenburns_config(img) + .autozoom(cfg): seg + depth + 256-candidate
                   autozoom + 2 inpaint passes + 75 output frames per input image

Timing: W >= 3 warm-up steps, then exactly K steps between barrier + cuda.synchronize; device time from CUDA events on
the launching stream, max over ranks.  Inputs cycle over `--scenes` distinct scenes whose footprint exceeds the 126 MB L2.
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 1024
FOCAL, BASELINE = 512.0, 40.0
SHIFT_U, SHIFT_V, DEPTH_RATIO = 40.0, -25.0, 0.8          # a mid-path Ken-Burns camera position
CROP = int(np.floor(0.97 * W))                              # objFrom crop, kenburns_effect.py:958-959
METRIC = "end-to-end frames/sec @1024x1024 (seg+depth+warp)"


def make_inputs(n_scenes):
    from cartoonsegmentation_b200.utils.synthetic import smooth_disparity, smooth_image
    imgs = np.stack([smooth_image(H, W, seed=1234 + i) for i in range(n_scenes)])
    disp = np.stack([smooth_disparity(H, W, seed=4321 + i)[0, 0] for i in range(n_scenes)])
    return imgs, disp


# ------------------------------------------------------------------------------------------------ exactly ONE line on stdout
# Libraries write to file descriptor 1 behind Python's back (NCCL prints "NCCL version ..." at communicator creation).  Everything written to
# fd 1 during the run is sent to stderr; the JSON line goes to the saved, real stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_frame(orc, img, raw):
    """One input frame through the oracle port of the warp stage -> uint8 frame."""
    cloud = orc.disparity_to_cloud(raw[None, None], FOCAL, BASELINE)
    common = {'objDepthrange': cloud['depthrange'], 'intWidth': W, 'intHeight': H, 'fltFocal': FOCAL, 'fltBaseline': BASELINE}
    dmin = cloud['depthrange'][0]
    pts, _ = orc.process_shift({'tenPoints': cloud['points'].reshape(1, 3, -1), 'fltShiftU': SHIFT_U, 'fltShiftV': SHIFT_V,
                                'fltDepthFrom': dmin, 'fltDepthTo': dmin * DEPTH_RATIO}, common)
    img_t = np.ascontiguousarray(img.transpose(2, 0, 1)[None].astype(np.float32) * np.float32(1.0 / 255.0))
    data = np.concatenate([img_t.reshape(1, 3, -1), cloud['depth'].reshape(1, 1, -1)], 1)
    r, e = orc.render_pointcloud(pts, data, W, H, FOCAL, BASELINE)
    f = orc.fill_disocclusion(r, r[:, 3:4] * (e > 0.0))
    u8 = orc.frame_pack_u8(f[0])
    return orc.resize_linear(orc.get_rect_sub_pix(u8, (CROP, CROP), (W / 2.0, H / 2.0)), (W, H))


def cpu_arm(imgs, disp, n_frames, threads):
    """Warp stage only, `threads` frames in flight (ctypes releases the GIL); returns frames/s."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import kb_oracle as orc
    orc.lib()
    cpu_frame(orc, imgs[0], disp[0])                          # warm-up (page-in, build)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: cpu_frame(orc, imgs[i % len(imgs)], disp[i % len(disp)]), range(n_frames)))
    return n_frames / (time.perf_counter() - t0)


class CpuPath:
    """The CPU arm: the oracle port of every stage of the workload for ONE input frame (PyTorch fp32 on all host threads for the networks,
    oracle/kb_oracle.c -- scalar C, ONE thread -- for the Ken-Burns kernels).  The real reference package cannot run on a CPU at all (mmdet / mmcv /
    cupy absent in the image; its Ken-Burns kernels are GPU-only cupy strings and anime_3dkenburns/common.py:74 hard-codes .cuda()), so its CPU
    implementation of this path *is* the port."""

    def __init__(self, stages, depth='leres'):
        import torch
        from cartoonsegmentation_b200.animeinsseg import rtmdet
        from cartoonsegmentation_b200.depth_modules import leres as L
        from oracle import det_oracle as D, kb_adjust_oracle as AO, kb_oracle as orc, leres_oracle as LO
        self.torch, self.D, self.orc, self.LO, self.AO = torch, D, orc, LO, AO
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.stages = stages
        self.depth = depth if 'depth' in stages else None
        self.det = self.ler = None
        if 'seg' in stages:
            self.det = D.RTMDetIns().eval(); self.det.load_state_dict(rtmdet.synthetic_state_dict(0))
        if self.depth == 'leres':
            self.ler = LO.RelDepthModel().eval(); self.ler.load_state_dict(L.synthetic_state_dict(0), strict=False)
        elif self.depth == 'zoe':
            raise SystemExit("--impl reference --depth zoe: the CPU oracle of DPT-BEiT-L @672^2 x 2 flips needs minutes per frame; time the LeReS workload")
        orc.lib()
        self.stage_s = {"seg": 0.0, "depth": 0.0, "adjust": 0.0, "warp_single_thread_c": 0.0}

    def frame(self, img, raw_synth, keep=False):
        torch = self.torch
        masks = None
        t = time.perf_counter()
        with torch.no_grad():
            if self.det is not None:
                masks = self.D.infer(self.det, img)['masks']
            t1 = time.perf_counter(); self.stage_s["seg"] += t1 - t
            raw = self.LO.depth_est_leres(self.ler, img) if self.ler is not None else raw_synth
            t2 = time.perf_counter(); self.stage_s["depth"] += t2 - t1
            if masks is not None and len(masks):                       # depth_adjustment_animesseg (kenburns_effect.py:39-91)
                raw = self.AO.depth_adjustment_animesseg(masks, torch.from_numpy(np.ascontiguousarray(raw))[None, None], torch.zeros(1, 3, H, W))[0, 0].numpy()
            t3 = time.perf_counter(); self.stage_s["adjust"] += t3 - t2
        out = cpu_frame(self.orc, img, np.ascontiguousarray(raw)) if 'warp' in self.stages else None
        self.stage_s["warp_single_thread_c"] += time.perf_counter() - t3
        return (out, raw, 0 if masks is None else len(masks)) if keep else out

    def time(self, imgs, disp, steps, warm, keep=False):
        for i in range(warm):
            self.frame(imgs[i % len(imgs)], disp[i % len(disp)])
        self.stage_s = {k: 0.0 for k in self.stage_s}
        kept = []
        t0 = time.perf_counter()
        for i in range(steps):
            r = self.frame(imgs[i % len(imgs)], disp[i % len(disp)], keep)
            if keep:
                kept.append(r)
        sec = (time.perf_counter() - t0) / steps
        self.kept = kept
        return sec

    def sample_text(self, steps):
        st = {k: round(v / max(1, steps), 3) for k, v in self.stage_s.items()}
        return (f"{steps} frame(s) of 1024x1024 (after 1 warm-up) through the oracle port of every stage of the workload: oracle/det_oracle.py + oracle/leres_oracle.py + "
                f"oracle/kb_adjust_oracle.py (PyTorch fp32, {self.cores} threads) + oracle/kb_oracle.c (scalar C on ONE thread); seconds per frame by stage: {st}")


def run_reference(args):
    """--impl reference: CpuPath, one image per step, on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    stages = [s_ for s_ in args.stages.split(",") if s_]
    imgs, disp = make_inputs(2)
    cpu = CpuPath(stages, args.depth)
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    sec = cpu.time(imgs, disp, steps, warm)
    fps = 1.0 / sec
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1000.0 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, 2, stages, args.depth),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cpu.cores, "kind": "port", "sample": cpu.sample_text(steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def workload_config(batch, scenes, stages=("seg", "depth", "warp"), depth="leres"):
    names = {"seg": "AnimeInsSeg.infer body (ConvNeXt-B RTMDet-Ins forward @1024^2, det_size=1024, random-init seeded weights, decode/NMS/mask head/mask tail, "
                    "max_instances=100, refine off)",
             "depth": ("LeReS depth (ResNeXt-101 32x8d + decoder @640^2) + the reference's 16->8 bit quantisation tail" if depth == "leres" else
                       "ZoeDepth (DPT-BEiT-L @672^2, reflect pad + flip augmentation = 2 net inputs per frame, metric bins head, depth->disparity)")
                      + " + instance-guided depth flattening",
             "warp": "disparity->cloud + camera shift + z-buffered render (C=4) + disocclusion fill + u8 pack + centre crop/resize: one Ken-Burns frame"}
    missing = [v for k, v in {"seg": "seg", "depth": "depth (raw disparity is then a synthetic input)", "warp": "warp"}.items() if k not in stages]
    missing += ["ISNet mask refine (A10) and Inpaint net + autozoom (C4-C6) are per-image, not per-frame: measured in other_workloads"]
    return {"workload": "per input frame @1024x1024: " + " -> ".join(names[s_] for s_ in ("seg", "depth", "warp") if s_ in stages),
            "api": "KenBurnsPipeline.render_frame_batch(frames uint8 [B,1024,1024,3]) -> one warped uint8 frame per input frame",
            "stages": list(stages), "depth": depth if "depth" in stages else None, "stages_missing": missing, "frame": [H, W], "batch_frames_per_step": batch, "focal": FOCAL, "baseline": BASELINE,
            "partitioning": "one frame stream, frame i -> rank i mod G (utils/dist.py:shard_indices); scene of frame i = i mod scenes",
            "l2_policy": f"inputs cycle over {scenes} distinct scenes in two alternating batch buffers; every frame streams > 126 MB of intermediates, so no input survives in L2 between uses"}


def timed_call(fn):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), r


MLP_LABEL = re.compile(r"(k_mlp_tc)\[(\d+)x(\d+)->(\d+)->(\d+)\]")
CONV_LABEL = re.compile(r"(k_conv_tc|k_conv_halo)\[(\d+)x(\d+)x(\d+)x(\d+)->(\d+) k(\d+)x(\d+) s(\d+) d(\d+) g(\d+) act(\d+) res(\d+)\]")


def profile_detail(lib, fn):
    """profile_call in the library's detailed mode (conv launches keyed by layer shape) -> per kernel name {ms, count, gflop}: the algorithmic
    FLOPs of the tensor-core kernels are computed from the launches themselves, 2 * pixels_out * Cout * R * S * Cin / groups per launch."""
    old = os.environ.get("CSB_PROFILE_DETAIL")
    os.environ["CSB_PROFILE_DETAIL"] = "1"
    try:
        prof = profile_call(lib, fn)
    finally:
        if old is None:
            os.environ.pop("CSB_PROFILE_DETAIL", None)
        else:
            os.environ["CSB_PROFILE_DETAIL"] = old
    return aggregate_detail(prof)


def aggregate_detail(prof):
    """{detailed launch label: {ms, count}} -> {kernel: {ms, count, gflop, gbytes}} (pure: unit-tested in tests/test_host_logic_cpu.py)"""
    agg = {}
    for label, v in prof.items():
        m = CONV_LABEL.match(label)
        mm = MLP_LABEL.match(label)
        name = m.group(1) if m else (mm.group(1) if mm else label)
        a = agg.setdefault(name, {"ms": 0.0, "count": 0, "gflop": 0.0})
        a["ms"] += v["ms"]
        a["count"] += v["count"]
        if mm:                                                        # fused ConvNeXt MLP: two GEMMs; compulsory bytes = x + residual + out + both filters + LN partials
            px, C_, Hd = int(mm.group(2)), int(mm.group(3)), int(mm.group(4))
            a["gflop"] += v["count"] * 2.0 * 2.0 * px * C_ * Hd / 1e9
            a["gbytes"] = a.get("gbytes", 0.0) + v["count"] * (3.0 * px * C_ * 2 + 2.0 * C_ * Hd * 2 + px * (C_ // 64) * 8) / 1e9
        if m:
            N, Hi, Wi, Cin, Cout, R, S_, st, dil, g = map(int, m.groups()[1:11])
            Ho, Wo = (Hi + st - 1) // st, (Wi + st - 1) // st        # the engine's layers are 'same'-padded (stride 1) or strided: ceil(n / stride)
            a["gflop"] += v["count"] * 2.0 * N * Ho * Wo * Cout * R * S_ * (Cin // max(1, g)) / 1e9
            res = int(m.group(13))                                    # compulsory bytes: fp16 input + output (+ residual read) + the packed filter
            a["gbytes"] = a.get("gbytes", 0.0) + v["count"] * (2.0 * N * (Hi * Wi * Cin + Ho * Wo * Cout * (2 if res else 1)) + 2.0 * Cout * R * S_ * (Cin // max(1, g))) / 1e9
    return agg


def profile_call(lib, fn):
    """per-kernel device time of fn(): a CUDA event after every library launch (csb_profile_begin/end)"""
    import torch
    torch.cuda.synchronize()
    lib.csb_profile_begin(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    fn()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.csb_profile_end(buf, len(buf))
    return json.loads(buf.value.decode())


# ------------------------------------------------------------------------------------------------ API-level workloads (BASELINE configs[0..3])
def run_other(pipe, imgs_np, args, lib):
    """Workloads measured through the reference-facing Python API, host numpy in / host numpy (or AnimeInstances) out, CUDA events around the
    calls (the host gaps between launches are inside the timed region)."""
    import torch
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb
    seg = pipe.animeinsseg
    S = len(imgs_np)
    res = {}
    off, on = {'refine_method': 'none'}, {'refine_method': 'refinenet_isnet'}
    # ---- BASELINE configs[0]: AnimeInsSeg.infer on ONE 512 x 512 example image (examples/1562990.jpg, fixture made by tests/golden/make_example_fixture.py)
    ex = os.path.join(ROOT, "tests", "golden", "example_1562990_512.jpg")
    if os.path.exists(ex):
        import cv2
        img512 = cv2.imread(ex)
        for kw, name in ((off, "refine_off"), (on, "refine_isnet")):
            for _ in range(2):
                seg.infer(img512, 0.3, kw, 'tensor', det_size=512)
            ms, inst = timed_call(lambda: [seg.infer(img512, 0.3, kw, 'tensor', det_size=512) for _ in range(5)])
            res.setdefault("infer_512_example", {"api": "AnimeInsSeg.infer(ndarray 512x512x3 = examples/1562990.jpg resized + centre-cropped, pred_score_thr=0.3, det_size=512), "
                                                       "one image per call (BASELINE configs[0]; the CPU arm of the same call is cpu_baseline_512)"})[name] = \
                {"ms_per_image": ms / 5, "images_per_s": 5e3 / ms, "instances": len(inst[-1])}
        seg.set_refine_method('none')
    # ---- BASELINE configs[1]: AnimeInsSeg.infer on a batch (python list) of 32 images of 1024x1024, det_size 1024
    lst = [imgs_np[i % S] for i in range(32)]
    seg.infer(lst, 0.3, off, 'tensor', det_size=H)                      # steady state: the same call once untimed (first-use allocations of the batch-32 buffers)
    ms, inst = timed_call(lambda: seg.infer(lst, 0.3, off, 'tensor', det_size=H))
    K = float(np.mean([len(i) for i in inst]))
    res["infer_batch32"] = {"api": "AnimeInsSeg.infer(list of 32 ndarray 1024x1024x3, pred_score_thr=0.3, refine off, output_type='tensor', det_size=1024)",
                            "frames_per_s": 32 / (ms * 1e-3), "ms": ms, "instances_per_image": K, "h2d_bytes": 32 * H * W * 3}
    del inst
    # ---- the same with the reference's default-on ISNet refinement (A10): 159.5 GFLOP per instance at 720^2
    n_ref = 4
    for _ in range(2):                                            # steady state: the second call of a shape captures its CUDA graph, later ones replay it
        seg.infer(lst[:n_ref], 0.3, on, 'tensor', det_size=H)
    ms, inst = timed_call(lambda: seg.infer(lst[:n_ref], 0.3, on, 'tensor', det_size=H))
    K = float(np.mean([len(i) for i in inst]))
    tf = 159.5e-3 * K * n_ref / (ms * 1e-3)
    res["infer_refine_isnet"] = {"api": f"AnimeInsSeg.infer(list of {n_ref} ndarray 1024x1024x3, refine_method='refinenet_isnet' (reference default), det_size=1024)",
                                 "frames_per_s": n_ref / (ms * 1e-3), "ms": ms, "instances_per_image": K, "isnet_tflops": tf,
                                 "roofline": {"bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "note": "159.5 GFLOP per instance x K x images over the whole call (detector included in the time)"}}
    del inst
    seg.set_refine_method('none')
    # ---- BASELINE configs[2]: seg + ZoeDepth forward over 32 frames (DPT-BEiT-L @672^2, 2 net inputs per frame), no warp
    if not args.no_zoe:
        zcfg = kb.KenBurnsConfig(det_size=H, max_size=H, depth_est='zoe', pred_score_thr=0.3, refine_crf=False)
        zp = kb.KenBurnsPipeline.__new__(kb.KenBurnsPipeline)
        zp.__dict__.update(pipe.__dict__)                      # shares the detector; adds the ZoeDepth estimator
        zp.cfg, zp.depth_zoe = zcfg, None
        zp.set_depth_estimation('zoe')
        zb = torch.from_numpy(np.stack(lst)).to(pipe.device)
        zp.render_frame_batch(zb, warp=False)
        ms, _ = timed_call(lambda: zp.render_frame_batch(zb, warp=False))
        prof = profile_call(lib, lambda: zp.render_frame_batch(zb, warp=False))
        zoe_lin = 2 * (24 * 1765 * 12 * 1024 * 1024 * 2 / 1e9 + 291.0) * 32 + 1011.9 * 32          # GFLOP on the conv engine (k_conv_tc + k_conv_halo): BEiT linears + DPT convs (2 net inputs) + detector
        att = 2 * 24 * 16 * 4 * 1765 * 1765 * 64 / 1e9 * 32                                          # GFLOP on k_attention_tc
        res["seg_zoedepth_batch32"] = {"api": "KenBurnsPipeline(depth_est='zoe').render_frame_batch(32 frames 1024x1024x3 resident in HBM, warp=False): detector + "
                                              "ZoeDepth.infer(pad_input, with_flip_aug) + depth->disparity + instance flattening (BASELINE configs[2])",
                                       "frames_per_s": 32e3 / ms, "ms": ms,
                                       "conv_engine_tflops": zoe_lin / sum(prof[k]["ms"] for k in ("k_conv_tc", "k_conv_halo") if k in prof) if "k_conv_tc" in prof else None,
                                       "k_attention_tc_tflops": att / prof["k_attention_tc"]["ms"] if "k_attention_tc" in prof else None,
                                       "per_kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:10]}}
        del zp, zb
        torch.cuda.empty_cache()
    # ---- BASELINE configs[3]: full 3D Ken Burns per input image: seg + depth + adjust + cloud + autozoom (256 candidate renders) + 2 inpaint passes
    #      + num_frame (75) output frames, uint8 frames returned on the host (run_kenburns.py:19-33)
    n_img = max(1, args.other_images)

    def kb_image(i):
        kcfg = pipe.generate_kenburns_config(imgs_np[i % S])
        return pipe.autozoom(kcfg)
    for i in range(n_img + 1):                                # steady state: every image once untimed (per-image cloud sizes differ -> first-use allocations)
        kb_image(i)
    ms, frames = timed_call(lambda: [len(kb_image(1 + i)) for i in range(n_img)])
    l0 = lib.csb_launch_count()
    prof = profile_call(lib, lambda: kb_image(0))
    n_launch = lib.csb_launch_count() - l0
    res["kenburns_full"] = {"api": "KenBurnsPipeline.generate_kenburns_config(img) + .autozoom(cfg): seg + depth(%s) + autozoom + 2 x inpaint + %d frames, "
                                   "host ndarray in, list of host uint8 frames out" % (pipe.cfg.depth_est, pipe.cfg.num_frame),
                            "input_images_per_s": n_img / (ms * 1e-3), "output_frames_per_s": sum(frames) / (ms * 1e-3), "ms_per_image": ms / n_img,
                            "images": n_img, "h2d_bytes_per_image": H * W * 3, "d2h_bytes_per_image": pipe.cfg.num_frame * H * W * 3, "launches_per_image": int(n_launch),
                            "per_kernel_ms_profiled": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if k != "memset"},
                            "profile_note": "a CUDA event after every library launch; 'memset' (dropped) also absorbs non-library work between launches (copies, host gaps)"}
    # ---- the same with the options the reference ships in configs/3dkenburns.yaml: ISNet mask refinement (refine_size 720) + depth_field (bokeh)
    pipe.cfg.mask_refine_kwargs = {'refine_method': 'refinenet_isnet', 'refine_size': 720}
    pipe.cfg.depth_field = True
    n_y = min(2, n_img)
    for i in range(n_y + 1):
        kb_image(i)
    ms, frames = timed_call(lambda: [len(kb_image(1 + i)) for i in range(n_y)])
    res["kenburns_full_shipped_yaml"] = {"api": "as kenburns_full with mask_refine_kwargs={refinenet_isnet, 720} and depth_field=True (configs/3dkenburns.yaml:16,36-38)",
                                         "input_images_per_s": n_y / (ms * 1e-3), "output_frames_per_s": sum(frames) / (ms * 1e-3), "ms_per_image": ms / n_y,
                                         "images": n_y}
    pipe.cfg.mask_refine_kwargs = {}
    pipe.cfg.depth_field = False
    seg.set_refine_method('none')
    return res


def parity_block(pipe, cpu, imgs_np, disp_np, stages):
    """The oracle frames of the cpu_baseline leg double as the checker: the same frames through the CUDA path (same API call as the timed one),
    compared stage by stage.  Detailed full-size numbers: profiles/r2_parity_full.json (tests/parity_full.py), asserted in tests/test_parity_full_gpu.py."""
    import torch
    n = len(cpu.kept)
    batch = torch.from_numpy(imgs_np[:n]).to(pipe.device)
    raw = None if 'depth' in stages else torch.from_numpy(disp_np[:n]).to(pipe.device)
    r = pipe.render_frame_batch(batch, SHIFT_U, SHIFT_V, DEPTH_RATIO, raw_disparity=raw, segment='seg' in stages, warp='warp' in stages)
    out = {"frames_compared": n, "oracle": "oracle/det_oracle.py + leres_oracle.py + kb_adjust_oracle.py + kb_oracle.c (fp32 CPU) on the same seeded frames",
           "parity_pinning": "K1-K4/C8 pinned to the unmodified reference kernels, A10/B4/B6/C5 to reference-module goldens; A1-A6 (mmdet) and B3 (MiDaS) are "
                             "restatements of un-vendored third-party code: parity UNPINNED for those rows",
           "full_size_measurements": "profiles/r2_parity_full.json"}
    disp = r['disparity'].cpu().numpy()
    dd, fd, inst = [], [], []
    for i, (frame_o, raw_o, k_o) in enumerate(cpu.kept):
        d = np.abs(disp[i] - raw_o)
        dd.append({"max_abs_levels": float(d.max()), "frac_pixels_differ": float((d > 0).mean()), "frac_gt_1_level": float((d > 1.0).mean())})
        if r['frames'] is not None and frame_o is not None:
            f = np.abs(r['frames'][i].cpu().numpy().astype(int) - frame_o.astype(int))
            fd.append({"max_abs_lsb": int(f.max()), "frac_bytes_differ": float((f > 0).mean()), "frac_gt_2_lsb": float((f > 2).mean())})
        inst.append({"ours": None if r['num_instances'] is None else r['num_instances'][i], "oracle": k_o})
    out["adjusted_disparity_8bit_levels"] = dd
    out["warped_frame_u8"] = fd
    out["instances"] = inst
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stages", default="seg,depth,warp", help="comma list out of seg,depth,warp (default: the full metric)")
    ap.add_argument("--depth", default="leres", choices=["leres", "zoe"], help="depth estimator of the depth stage (reference: KenBurnsConfig.depth_est)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip other_workloads (infer_512_example, infer_batch32, seg_zoedepth_batch32, kenburns_full)")
    ap.add_argument("--no-zoe", action="store_true", help="skip the seg + ZoeDepth workload (BASELINE configs[2]) inside other_workloads")
    ap.add_argument("--other-images", type=int, default=3, help="input images of the kenburns_full workload")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    stages = [s_ for s_ in args.stages.split(",") if s_]

    import torch
    import torch.distributed as dist
    from cartoonsegmentation_b200 import _lib
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb
    from cartoonsegmentation_b200.utils.dist import gather_counters, max_over_ranks, shard_indices

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    B, S = args.batch, args.scenes

    # ---- the reference call surface: one pipeline object per rank, weights replicated from the same seed
    cfg = kb.KenBurnsConfig(det_size=H, max_size=H, depth_est=args.depth if 'depth' in stages else 'external', depth_est_size=640, pred_score_thr=0.3,
                            refine_crf=False)
    pipe = kb.KenBurnsPipeline(cfg, device=dev)
    pipe.animeinsseg.set_detect_size(H)

    # ---- inputs: ONE frame stream, frame i -> rank i mod G (shard_indices); frame i shows scene i mod S.  Two alternating batch buffers per rank.
    imgs_np, disp_np = make_inputs(S)
    NB = 2
    mine = shard_indices(world * B * NB, rank, world)                  # this rank's frames of the first NB steps (the pattern repeats)
    scene = [[mine[k * B + b] % S for b in range(B)] for k in range(NB)]
    batch_pin = [torch.from_numpy(np.stack([imgs_np[i] for i in scene[k]])).pin_memory() for k in range(NB)]
    batch_dev = [t.to(dev) for t in batch_pin]
    disp_dev = [torch.from_numpy(np.stack([disp_np[i] for i in scene[k]])).to(dev) for k in range(NB)] if 'depth' not in stages else None
    outs_pin = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
    stats = {"instances": 0, "calls": 0}

    def step(k, e2e):
        j = k % NB
        r = pipe.render_frame_batch(batch_pin[j] if e2e else batch_dev[j], SHIFT_U, SHIFT_V, DEPTH_RATIO,
                                    raw_disparity=None if disp_dev is None else disp_dev[j].clone() if 'seg' in stages else disp_dev[j],
                                    segment='seg' in stages, warp='warp' in stages, out_host=outs_pin if (e2e and 'warp' in stages) else None)
        if r['num_instances'] is not None:
            stats["instances"] += sum(r['num_instances']); stats["calls"] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for k in range(steps):
            step(k, e2e)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev), _lib.launch_count() - l0

    for k in range(args.warmup):
        step(k, False)
    step(0, True)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed(False, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(True, args.steps)

    # ---- per-kernel device time of the same step (second pass, a CUDA event after every launch) -> roofline
    # three profile passes, per-kernel median: the step runs at the board power cap and the SM clock of a single pass moves the per-kernel times by +-5 %
    passes = [profile_call(lib, lambda: step(0, False)) for _ in range(3)]
    prof = {k: {"ms": sorted(p_[k]["ms"] for p_ in passes if k in p_)[len([1 for p_ in passes if k in p_]) // 2], "count": passes[0][k]["count"]} for k in passes[0]}
    prof_d = profile_detail(lib, lambda: step(0, False))       # the same step with conv launches keyed by layer shape: FLOPs per tensor-core kernel

    other = run_other(pipe, imgs_np, args, lib) if (world == 1 and not args.no_other) else None

    # throughput counters: ONE all-gather of a per-rank struct over NCCL/NVLink
    frames = args.steps * B
    gathered = gather_counters(torch.tensor([frames, ms, ms_e2e], device=dev, dtype=torch.float64))
    total_frames = float(gathered[:, 0].sum())

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")        # ncu dram__bytes per kernel of this same step (tools/ncu_traffic.py)
        if os.path.exists(tp):
            traffic = json.load(open(tp))
        P = H * W
        dom = max((k for k in prof if k != "memset"), key=lambda k: prof[k]["ms"]) if prof else None
        roof = None
        tr = traffic.get("kernels", {}).get(dom) if dom else None
        tr_same = bool(tr) and traffic.get("batch") == B and traffic.get("stages") == ",".join(stages) and traffic.get("depth") == args.depth
        if dom in ("k_conv_tc", "k_conv_halo"):
            # Algorithmic FLOPs of the tensor-core launches of this step, computed from the launches themselves (profile_detail: 2 * output pixels *
            # Cout * R * S * Cin / groups per launch).  Cross-check: detector 1011.9 GFLOP / image @1024^2 (ConvNeXt-B 641.7 + neck 132.2
            # + head 237.9), LeReS 591.9 GFLOP / image @640^2; ZoeDepth per 672^2 net input 24 x 12 T D^2 MAC of Linear layers + ~291 GFLOP of DPT convs.
            d = prof_d.get(dom, {"ms": prof[dom]["ms"], "count": prof[dom]["count"], "gflop": 0.0})
            gflop, ach = d["gflop"], d["gflop"] / prof[dom]["ms"]      # GFLOP / ms = TFLOP/s, on the time of the undetailed pass
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak,
                    "traffic": (tr["dram_bytes_per_launch"] if tr_same else None),
                    "peak_source": "measured sustained bf16 cuBLAS (MEASURED_PEAKS.json)" if "bf16_tflops_sustained" in peaks else "fallback",
                    "launches_per_step": prof[dom]["count"], "algorithmic_gflop_per_step": gflop, "avg_launch_us": 1e3 * prof[dom]["ms"] / prof[dom]["count"],
                    "flop_note": "FLOPs summed over this kernel's launches of the step from their layer shapes (grouped convs: the group-sparse count)",
                    "algorithmic_bytes_per_launch": (d.get("gbytes", 0.0) * 1e9 / d["count"]) if d.get("count") else None}
            if roof["traffic"] and roof["algorithmic_bytes_per_launch"]:
                roof["traffic_over_algorithmic"] = roof["traffic"] / roof["algorithmic_bytes_per_launch"]
            # the conv engine as a whole: both tcgen05 kernels (k_conv_tc: per-tap pipeline, k_conv_halo: halo tiles for thin / grouped layers)
            TCK = ("k_conv_tc", "k_conv_halo", "k_mlp_tc")             # per-tap pipeline, halo tiles (thin / grouped layers), fused ConvNeXt MLP (stages 1-2)
            eng_ms = sum(prof[k]["ms"] for k in TCK if k in prof)
            eng_gf = sum(prof_d[k]["gflop"] for k in TCK if k in prof_d)
            roof["conv_engine"] = {"kernels": [k for k in TCK if k in prof], "ms_per_step": round(eng_ms, 3), "algorithmic_gflop_per_step": round(eng_gf, 1),
                                   "achieved": eng_gf / eng_ms if eng_ms else None, "frac": (eng_gf / eng_ms / tc_peak) if eng_ms else None,
                                   "per_kernel": {k: {"ms": round(prof[k]["ms"], 3), "gflop": round(prof_d[k]["gflop"], 1), "tflops": round(prof_d[k]["gflop"] / prof[k]["ms"], 1)}
                                                  for k in TCK if k in prof and k in prof_d},
                                   "note": "all tcgen05 kernels of the step; the dominant kernel's own figure is `frac` above (the GELU-epilogue-bound C -> 4C layers of "
                                           "ConvNeXt stages 1-2 run in k_mlp_tc, not in k_conv_tc)"}
            if tr:
                roof["traffic_source"] = {"file": "profiles/r2_traffic.json", "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the "
                                          "kernel's launches of one step)", "same_config_as_this_run": tr_same, **{k: tr[k] for k in ("launches", "dram_read_bytes", "dram_write_bytes")}}
        elif dom:
            alg_bytes = {"k_splat": 52 * P, "k_zpass": 16 * P, "k_degrid": 8 * P, "k_norm_pack_mark": 24 * P, "k_crop_resize": 6 * P, "k_d2c_max": 4 * P,
                         "k_d2c_scale": 12 * P, "k_d2c_points": 55 * P}
            dur_ms = prof[dom]["ms"] / prof[dom]["count"]
            ab = alg_bytes.get(dom)
            ach = ab / (dur_ms * 1e-3) / 1e9 if ab else None
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": (ach / hbm_peak) if ach else None,
                    "traffic": (tr["dram_bytes_per_launch"] if tr_same else None),
                    "peak_source": peak_src, "avg_launch_us": dur_ms * 1e3, "algorithmic_bytes_per_launch": ab}
        if roof is not None:
            roof["per_kernel_ms_per_step"] = {k: round(v["ms"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
            if "k_dwconv_halo" in prof:                               # second kernel: fp32 FMA issue bound (49 FMA per output element), HBM floor 4 C B/px
                dw_gflop = 2 * 49 * 97.5e6 * B / 1e9                  # ConvNeXt-B @1024^2: 97.5 M (pixel, channel) outputs per image
                roof["second_kernel"] = {"kernel": "k_dwconv_halo", "ms_per_step": round(prof["k_dwconv_halo"]["ms"], 3), "bound": "fp32 FMA issue",
                                         "achieved_tflops": dw_gflop / prof["k_dwconv_halo"]["ms"], "peak_tflops": 72.0,
                                         "peak_source": "148 SMs x 128 FMA/clk x 2 x 1.9 GHz (packed FFMA2)", "hbm_floor_ms": 4 * 97.5e6 * B / (hbm_peak * 1e6)}
        h2d = B * H * W * 3                                           # the uint8 BGR frames; every later stage stays on the device
        d2h = (B * H * W * 3 if 'warp' in stages else 0) + (4 * B if 'seg' in stages else 0)     # rendered uint8 frames + instance counts
        out = {"metric": METRIC, "value": total_frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate) / f32 render",
               "data": "synthetic", "config": workload_config(B, S, stages, args.depth), "clocks": clocks,
               "e2e": {"value": total_frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "KenBurnsPipeline.render_frame_batch(pinned host uint8 frames, out_host=pinned uint8 frames)"},
               "gpu_launches": launches, "roofline": roof,
               "instances_per_image": stats["instances"] / max(1, stats["calls"] * B) if 'seg' in stages else None}
        if world == 1 and not args.no_other:
            out["other_workloads"] = other
        if world == 1 and not args.no_cpu_baseline:
            # the oracle port of the SAME workload (all stages) on the box's host cores, a bounded sample: 1 warm-up + 2 timed frames; its frames are also
            # the checker of the `parity` block
            if args.depth == 'leres':
                cpu = CpuPath(stages, 'leres')
                sec = cpu.time(imgs_np[:2], disp_np[:2], 2, 1, keep=True)
                out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "frames/s", "cores": cpu.cores, "kind": "port", "sample": cpu.sample_text(2)}
                out["parity"] = parity_block(pipe, cpu, imgs_np, disp_np, stages)
                ex = os.path.join(ROOT, "tests", "golden", "example_1562990_512.jpg")
                if 'seg' in stages and os.path.exists(ex):            # BASELINE configs[0]: the CPU arm of AnimeInsSeg.infer on the 512 x 512 example image
                    import cv2
                    img512 = cv2.imread(ex)
                    cpu.D.infer(cpu.det, img512)
                    t0 = time.perf_counter()
                    for _ in range(5):
                        n512 = len(cpu.D.infer(cpu.det, img512)['scores'])
                    s512 = (time.perf_counter() - t0) / 5
                    out["cpu_baseline_512"] = {"value": 1.0 / s512, "unit": "images/s", "cores": cpu.cores, "kind": "port", "instances": n512,
                                               "sample": "5 calls of oracle.det_oracle.infer (refine off) on tests/golden/example_1562990_512.jpg, det_size=512 (BASELINE configs[0])"}
            else:
                out["cpu_baseline"] = {"value": cpu_arm(imgs_np[:2], disp_np[:2], 4, 1), "unit": "frames/s", "cores": 1, "kind": "port",
                                       "sample": "4 frames of 1024x1024 on 1 thread through the WARP stage only (oracle/kb_oracle.c); the DPT-BEiT-L CPU oracle @672^2 "
                                                 "needs minutes per frame -- run the default --depth leres workload for an all-stage CPU baseline"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
