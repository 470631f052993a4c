#!/usr/bin/env python
"""bench.py -- hot-path throughput on N B200s (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # CPU arm: the oracle port of the same path on the host cores

A "step" is one pass of the hot path over one batch of `--batch` synthetic 1024x1024 input frames.
Round-1 workload (`config.workload`): the Ken-Burns *warp* leg of BASELINE.json's metric -- per input frame:
disparity -> point cloud (kenburns_effect.py:928-937), camera shift, z-buffered splat render, disocclusion fill,
uint8 pack, centre crop + bilinear resize (kenburns_effect.py:1028-1040,1069-1070).  The seg and depth forwards are not on
the path yet (`config.stages_missing`); raw disparity is a synthetic input until the depth net lands.

Timing: W >= 3 warm-up steps, then exactly K steps between barrier + cuda.synchronize; device time from CUDA events on
the launching stream, max over ranks.  Inputs cycle over `--scenes` distinct scenes whose footprint exceeds the 126 MB L2.
"""
import argparse
import ctypes
import json
import os
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
    """One input frame through the oracle port of the same path -> uint8 frame."""
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
    """Oracle port on `threads` host threads (ctypes releases the GIL); returns frames/s."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import kb_oracle as orc
    orc.lib()
    cpu_frame(orc, imgs[0], disp[0])                          # warm-up (page-in, build)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: cpu_frame(orc, imgs[i % len(imgs)], disp[i % len(disp)]), range(n_frames)))
    return n_frames / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference Ken-Burns kernels are GPU-only cupy strings (anime_3dkenburns/common.py:74 hard-codes
    .cuda()); its CPU implementation of this path is therefore the oracle port, run on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    imgs, disp = make_inputs(2)
    per_step = cores                                          # one frame per host thread per step
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_arm(imgs, disp, per_step, cores)
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        cpu_arm(imgs, disp, per_step, cores)
    dt = time.perf_counter() - t0
    fps = steps * per_step / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(per_step, 2),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {per_step} frames of 1024x1024 (one per host thread), oracle/kb_oracle.c"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def workload_config(batch, scenes):
    return {"workload": "kenburns_warp_1024: disparity->cloud + shift + render(C=4) + disocclusion fill + u8 pack + crop/resize per input frame",
            "stages_missing": ["seg (AnimeInsSeg.infer)", "depth forward (raw disparity is a synthetic input)"],
            "frame": [H, W], "batch_frames_per_step": batch, "focal": FOCAL, "baseline": BASELINE,
            "l2_policy": f"inputs cycle over {scenes} scenes (> 126 MB L2 footprint with per-scene clouds)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from cartoonsegmentation_b200 import _lib
    from cartoonsegmentation_b200.anime_3dkenburns import kenburns_effect as kb

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    # ---- inputs: each rank owns `scenes` distinct scenes (weak scaling: per-GPU work fixed); seeds differ per rank
    imgs_np, disp_np = make_inputs(args.scenes)
    imgs_np = np.roll(imgs_np, rank, axis=0)
    imgs_pin = torch.from_numpy(imgs_np).pin_memory(); disp_pin = torch.from_numpy(disp_np).pin_memory()
    imgs_dev = imgs_pin.to(dev); disp_dev = disp_pin.to(dev)
    S = args.scenes
    scratch = kb.FrameScratch(H, W, dev)
    clouds = [{k: torch.empty((1, c, H, W), device=dev) for k, c in (('disparity', 1), ('depth', 1), ('valid', 1), ('points', 3), ('unaltered', 3))} for _ in range(2)]
    data = [torch.empty((1, 4, H * W), device=dev) for _ in range(2)]
    scalars = torch.empty(8, device=dev); d2c_scratch = torch.empty(64, device=dev, dtype=torch.int64); shift_dev = torch.empty(3, device=dev)
    outs = torch.empty((args.batch, H, W, 3), device=dev, dtype=torch.uint8)
    outs_pin = torch.empty((args.batch, H, W, 3), dtype=torch.uint8).pin_memory()
    stage_img = torch.empty((args.batch, H, W, 3), device=dev, dtype=torch.uint8); stage_disp = torch.empty((args.batch, H, W), device=dev)
    cd = ctypes.c_double
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def frame(img_u8, raw, out, slot):
        """One input frame, all on the current stream, no host sync."""
        c = clouds[slot]
        _lib.check(lib.csb_disparity_to_cloud(_lib.ptr(raw), H, W, cd(FOCAL), cd(BASELINE), _lib.ptr(c['disparity']), _lib.ptr(c['depth']), _lib.ptr(c['valid']),
                                              _lib.ptr(c['points']), _lib.ptr(c['unaltered']), _lib.ptr(scalars), _lib.ptr(d2c_scratch), _lib.ptr(img_u8), _lib.ptr(data[slot]), st()))
        _lib.check(lib.csb_shift_from_scalars(_lib.ptr(scalars), W, H, cd(FOCAL), cd(SHIFT_U), cd(SHIFT_V), cd(DEPTH_RATIO), _lib.ptr(shift_dev), st()))
        d = data[slot]
        _lib.check(lib.csb_kenburns_frame(_lib.ptr(c['points']), _lib.ptr(d), H * W, H, W, cd(FOCAL), cd(BASELINE), None, _lib.ptr(shift_dev), CROP, CROP,
                                          cd(W / 2.0), cd(H / 2.0), _lib.ptr(scratch.zkey), _lib.ptr(scratch.zee), _lib.ptr(scratch.acc), _lib.ptr(scratch.packed),
                                          _lib.ptr(out), None, st()))

    def step_resident(k):
        for b in range(args.batch):
            i = (k * args.batch + b) % S
            frame(imgs_dev[i], disp_dev[i], outs[b], b & 1)

    def step_e2e(k):
        for b in range(args.batch):
            i = (k * args.batch + b) % S
            stage_img[b].copy_(imgs_pin[i], non_blocking=True)
            stage_disp[b].copy_(disp_pin[i], non_blocking=True)
            frame(stage_img[b], stage_disp[b], outs[b], b & 1)
            outs_pin[b].copy_(outs[b], non_blocking=True)
        torch.cuda.current_stream().synchronize()               # the caller receives the frames of this step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for k in range(steps):
            fn(k)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count() - l0

    for k in range(args.warmup):
        step_resident(k); step_e2e(k)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, args.steps)

    # ---- per-kernel device time of the same steps (second pass with an event after every launch) -> roofline
    torch.cuda.synchronize()
    lib.csb_profile_begin(st())
    for k in range(min(args.steps, 5)):
        step_resident(k)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.csb_profile_end(buf, len(buf))
    prof = json.loads(buf.value.decode())
    prof_steps = min(args.steps, 5)

    # throughput counters: ONE all-gather of a per-rank struct over NCCL/NVLink
    frames = args.steps * args.batch
    counters = torch.tensor([frames, ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        gathered = [torch.empty_like(counters) for _ in range(world)]
        dist.all_gather(gathered, counters)
        total_frames = sum(float(g[0]) for g in gathered)
    else:
        total_frames = frames

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        dom = max((k for k in prof if k not in ("memset",)), key=lambda k: prof[k]["ms"]) if prof else None
        P = H * W
        alg_bytes = {"k_splat": (12 + 16) * P + 4 * P + 20 * P,      # read points+data (28 B/pt) + z-buffer (4 B/px) ; write (C+1)*4 B/px
                     "k_zpass": 12 * P + 4 * P, "k_degrid": 8 * P, "k_norm_fill_pack": 20 * P + 3 * P, "k_crop_resize": 6 * P,
                     "k_d2c_max": 4 * P, "k_d2c_scale": 12 * P, "k_d2c_points": 55 * P}
        roof = None
        if dom:
            dur_ms = prof[dom]["ms"] / prof[dom]["count"]
            ab = alg_bytes.get(dom)
            ach = ab / (dur_ms * 1e-3) / 1e9 if ab else None
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": (ach / hbm_peak) if ach else None,
                    "traffic": None, "peak_source": peak_src, "avg_launch_us": dur_ms * 1e3, "algorithmic_bytes_per_launch": ab,
                    "per_kernel_ms_per_step": {k: v["ms"] / prof_steps for k, v in prof.items()}}
        out = {"metric": METRIC, "value": total_frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": workload_config(args.batch, S), "clocks": clocks,
               "e2e": {"value": total_frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": args.batch * (H * W * 3 + H * W * 4),
                       "d2h_bytes_per_step": args.batch * H * W * 3},
               "gpu_launches": launches, "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            imgs2, disp2 = imgs_np[:2], disp_np[:2]
            n = 4
            out["cpu_baseline"] = {"value": cpu_arm(imgs2, disp2, n, 1), "unit": "frames/s", "cores": 1, "kind": "port",
                                   "sample": f"{n} frames of 1024x1024 on 1 thread, oracle/kb_oracle.c (scalar C port)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
