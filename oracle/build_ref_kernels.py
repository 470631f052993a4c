"""TEST INFRASTRUCTURE ONLY -- builds a GPU oracle from the UNMODIFIED reference kernels.

The reference ships its five CUDA kernels as Python strings that are JIT-compiled by
cupy/NVRTC (`utils/cupy_utils.py:7-13`).  cupy is absent here, but nvcc is present, so
this script (run in the build container, where /root/reference exists):

  1. imports the reference's own `utils/cupy_utils.py` and
     `anime_3dkenburns/{models/utils.py,common.py}` BY PATH with a stub `cupy` module,
  2. replaces `launch_kernel` by a recorder and calls the reference's own
     `render_pointcloud` / `fill_disocclusion` on CPU tensors of the wanted shape, so the
     reference's own `preprocess_kernel` (`cupy_utils.py:16-122`) performs the textual
     SIZE_/STRIDE_/OFFSET_/VALUE_/{{var}} expansion,
  3. writes each expanded kernel to `oracle/_ref/gen/*.cu` (git-ignored; never committed;
     the only edit is a shape suffix appended to the entry-point name plus a host launcher
     of ours appended after it), and compiles them with nvcc for sm_100a into
     `oracle/_ref/libref_kernels.so`.

Nothing from the reference is copied into tracked files.  The .so travels to the GPU box
with the snapshot; `tests/test_ref_kernels_gpu.py` runs the reference kernels there and
compares them with the product kernels.  Only tests may load the result.
"""
import importlib.util
import os
import subprocess
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CSB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")

# (H, W, N, C) render shapes and (H, W, C) fill shapes the parity tests use
RENDER_SHAPES = [(64, 96, 64 * 96, 3), (128, 160, 128 * 160 + 3000, 4), (256, 256, 256 * 256, 4), (256, 256, 256 * 256 + 5000, 4),
                 (1024, 1024, 1024 * 1024, 4), (1024, 1024, 1024 * 1024, 3)]
FILL_SHAPES = [(64, 96, 4), (128, 160, 4), (256, 256, 4), (1024, 1024, 4)]
FOCAL, BASELINE = 512.0, 40.0


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class _Recorder:
    def __init__(self):
        self.records = []

    def launch_kernel(self, name, src):
        self.records.append((name, src))
        return lambda **kw: None


def _import_reference(rec):
    cupy = types.ModuleType("cupy")
    cupy.memoize = lambda **kw: (lambda f: f)
    cupy.int32 = int
    cupy.RawKernel = None
    sys.modules["cupy"] = cupy
    upkg = types.ModuleType("utils")
    upkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = upkg
    cu = _load("utils.cupy_utils", os.path.join(REF, "utils", "cupy_utils.py"))
    cu.launch_kernel = rec.launch_kernel
    pkg = types.ModuleType("refkb")
    pkg.__path__ = [os.path.join(REF, "anime_3dkenburns")]
    sys.modules["refkb"] = pkg
    mpkg = types.ModuleType("refkb.models")
    mpkg.__path__ = [os.path.join(REF, "anime_3dkenburns", "models")]
    sys.modules["refkb.models"] = mpkg
    mu = _load("refkb.models.utils", os.path.join(REF, "anime_3dkenburns", "models", "utils.py"))
    mu.launch_kernel = rec.launch_kernel
    cm = _load("refkb.common", os.path.join(REF, "anime_3dkenburns", "common.py"))
    cm.launch_kernel = rec.launch_kernel
    numba = types.ModuleType("numba")                      # utils/effects.py decorates its CPU variant with numba.njit; not needed here
    numba.jit = numba.njit = lambda f=None, **kw: (f if callable(f) else (lambda g: g))
    sys.modules.setdefault("numba", numba)
    cupy.float32 = float
    ef = _load("utils.effects", os.path.join(REF, "utils", "effects.py"))
    ef.launch_kernel = rec.launch_kernel
    return mu, cm, ef


LAUNCHER = '''
extern "C" void launch_{entry}(int n, void** p, void* stream) {{
    dim3 grid((n + 512 - 1) / 512, 1, 1), block(512, 1, 1);   // reference launch shape
    {entry}<<<grid, block, 0, (cudaStream_t) stream>>>(n, {args});
}}
'''


BOKEH_LAUNCHER = '''
extern "C" void launch_{entry}(int n, int h, int w, int nsamples, float dx, float dy, void** p, void* stream) {{
    dim3 grid((n + 512 - 1) / 512, 1, 1), block(512, 1, 1);   // reference launch shape (utils/effects.py:74-76)
    {entry}<<<grid, block, 0, (cudaStream_t) stream>>>(n, h, w, nsamples, dx, dy, (const float*) p[0], (const float*) p[1], (float*) p[2]);
}}
'''


def main():
    if not os.path.isdir(REF):
        print("reference not present; keeping prebuilt oracle/_ref", file=sys.stderr)
        return 0
    os.makedirs(GEN, exist_ok=True)
    rec = _Recorder()
    mu, cm, ef = _import_reference(rec)
    units = []
    for (H, W, N, C) in RENDER_SHAPES:
        rec.records.clear()
        mu.render_pointcloud(torch.zeros(1, 3, N), torch.zeros(1, C, N), W, H, FOCAL, BASELINE)
        for name, src in rec.records:
            units.append((name, f"H{H}_W{W}_N{N}_C{C}", src))
    for (H, W, C) in FILL_SHAPES:
        rec.records.clear()
        cm.fill_disocclusion(torch.zeros(1, C, H, W), torch.zeros(1, 1, H, W))
        for name, src in rec.records:
            units.append((name, f"H{H}_W{W}_C{C}", src))
    rec.records.clear()
    ef.bokeh_filter_cupy(torch.zeros(1, 3, 16), torch.zeros(1, 1, 16), 0.0, 1.0, 4, 4, 32)      # utils/effects.py:12-84 (shape-independent source)
    bokeh_units = [(name, src) for name, src in rec.records]
    nargs = {"kernel_pointrender_updateZee": 3, "kernel_pointrender_updateDegrid": 3,
             "kernel_pointrender_updateOutput": 4, "kernel_discfill_updateOutput": 3}
    objs = []
    for name, suffix, src in units:
        entry = f"{name}_{suffix}"
        body = src.replace(name, entry)
        args = ", ".join(f"(float*) p[{i}]" for i in range(nargs[name]))
        cu = os.path.join(GEN, entry + ".cu")
        with open(cu, "w") as f:
            f.write("#include <assert.h>\n" + body + LAUNCHER.format(entry=entry, args=args))
        obj = cu[:-3] + ".o"
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                               "-Xcompiler", "-fPIC", "-c", cu, "-o", obj])
        objs.append(obj)
    for name, src in bokeh_units:                          # kernel_bokeh(n, h, w, nsamples, dx, dy, img, depth, blurred)
        cu = os.path.join(GEN, name + ".cu")
        with open(cu, "w") as f:
            f.write(src + BOKEH_LAUNCHER.format(entry=name))
        obj = cu[:-3] + ".o"
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                               "-Xcompiler", "-fPIC", "-c", cu, "-o", obj])
        objs.append(obj)
    so = os.path.join(OUT, "libref_kernels.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs)
    print("built", so, "with", len(objs), "reference kernels")
    return 0


if __name__ == "__main__":
    sys.exit(main())
