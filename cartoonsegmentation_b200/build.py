"""Build libcsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

`python -m cartoonsegmentation_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A/B builds for experiments: CSB_BUILD_SUFFIX=_x CSB_EXTRA_NVCC="-DCSB_MBAR_HINT_NS=0" python -m cartoonsegmentation_b200.build -> libcsb200_x.so,
# loaded instead of the default library when CSB200_LIB points at it (_lib.py)
_SUFFIX = os.environ.get("CSB_BUILD_SUFFIX", "")
OBJ = os.path.join(HERE, "_build" + _SUFFIX)
SO = os.path.join(HERE, "libcsb200%s.so" % _SUFFIX)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-extended-lambda",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"] + os.environ.get("CSB_EXTRA_NVCC", "").split()
EXPORT = ["-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src):
    h = hashlib.sha1()
    for f in [src] + sorted(os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".cuh", ".h"))) + \
            [os.path.join(HERE, "..", "include", "csb200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ""
    r = subprocess.run(["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, True, r.stderr


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(_compile, _sources()))
    objs = [o for o, _, _ in res]
    if verbose:
        for _, rebuilt, log in res:
            if rebuilt:
                sys.stderr.write(log)
    if any(r for _, r, _ in res) or not os.path.exists(SO):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs + ["-lcuda"])
    return SO


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
