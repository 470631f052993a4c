"""CUDA-graph replay of launch-bound network forwards (small batches).

A batch-1 detector / LeReS forward is 150-200 kernel launches of 5-30 us each; issued from Python through ctypes (a ConvDesc, three
cuTensorMapEncodeTiled calls and a launch per conv, ~40 us of host time each) the GPU idles between them -- the reference has the same problem with
its eager PyTorch launches.  `GraphedForward` captures one forward per input shape into a CUDA graph (all library launches go to torch's current
stream, i.e. the capture stream; TMA descriptors travel as __grid_constant__ kernel parameters, scratch comes from the graph's private pool) and
replays it afterwards: one host call per forward.  First call with a shape: eager (warm-up); second: capture + replay; later: copy-in + replay.

The returned tensors are the graph's static outputs: they are overwritten by the next replay of the same shape, so callers consume them on the same
stream before calling again (the detector's post-process and the LeReS tail do).  CSB_CUDA_GRAPHS=0 disables capture.
"""
import os

import torch


class GraphedForward:
    def __init__(self, fn, max_batch=4, max_entries=6):
        self.fn, self.max_batch, self.max_entries = fn, max_batch, max_entries
        self.entries = {}
        self.enabled = os.environ.get("CSB_CUDA_GRAPHS", "1") != "0"

    def __call__(self, x):
        if not self.enabled or not x.is_cuda or x.shape[0] > self.max_batch or torch.cuda.is_current_stream_capturing():
            return self.fn(x)
        key = (tuple(x.shape), x.dtype, x.device.index)
        ent = self.entries.get(key)
        if ent is None:                                   # first sight of this shape: eager (also the warm-up every capture needs)
            if len(self.entries) >= self.max_entries:
                self.entries.pop(next(iter(self.entries)))
            self.entries[key] = {"graph": None}
            return self.fn(x)
        if ent["graph"] is None:
            try:
                static_in = x.clone()
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self.fn(static_in)
                ent.update(graph=g, static_in=static_in, out=out)
            except Exception as e:                        # a forward that cannot be captured (host read inside) keeps running eagerly
                self.enabled = False
                import warnings
                warnings.warn(f"CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly")
                return self.fn(x)
        else:
            ent["static_in"].copy_(x, non_blocking=True)
        ent["graph"].replay()
        return ent["out"]
