"""Readers for the reference's on-disk checkpoint formats -> plain state_dicts with the reference's parameter names.

| file (utils/constants.py)                         | format                                                        | reference loader |
|---|---|---|
| models/AnimeInstanceSegmentation/rtmdetl_e60.ckpt | mmengine checkpoint {'state_dict', 'meta': {'cfg': str, ...}} | animeinsseg/__init__.py:196-208 |
| models/AnimeInstanceSegmentation/refine_last.ckpt | plain state_dict of ISNetDIS(in_ch=4)                          | animeseg_refine/__init__.py:159-163 |
| models/leres/res101.pth                           | {'depth_model': state_dict} (keys may carry 'module.')         | depth_modules/leres/__init__.py:84-89 |
| models/zoedepth/ZoeD_M12_N.pt                     | {'model': state_dict} or a plain one ('module.' stripped)      | zoedepth/models/model_io.py:27-52 |
| kenburns_inpaintnet / depth refinenet checkpoints | plain state_dict                                               | anime_3dkenburns/models/__init__.py:7-20 |

mmengine checkpoints pickle objects of packages that are not installed here (mmengine `ConfigDict`, `HistoryBuffer` in 'message_hub', numpy scalars);
`torch_load` first tries the safe `weights_only` reader and then falls back to an unpickler that replaces every class it cannot import by an inert
stub, so only tensors and builtin containers are ever materialised.
"""
import pickle
from collections import OrderedDict

import torch


class _Stub:
    """Stand-in for a pickled object whose class is not importable: swallows construction and state."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        pass

    def __reduce_ex__(self, protocol):
        return (_Stub, ())


class _TolerantUnpickler(pickle.Unpickler):
    _ALLOWED_PREFIXES = ("torch", "collections", "numpy", "builtins", "_codecs")

    def find_class(self, module, name):
        if module.split(".")[0] in self._ALLOWED_PREFIXES:
            try:
                return super().find_class(module, name)
            except Exception:
                return _Stub
        return _Stub


class _TolerantPickle:
    Unpickler = _TolerantUnpickler
    __name__ = "tolerant_pickle"

    @staticmethod
    def load(f, **kw):
        return _TolerantUnpickler(f, **kw).load()


def torch_load(path):
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)


def _strip(sd, prefix="module."):
    return OrderedDict((k[len(prefix):] if k.startswith(prefix) else k, v) for k, v in sd.items() if torch.is_tensor(v))


def detector_state_dict(ckpt):
    """rtmdetl_e60.ckpt (or any mmdet RTMDet-Ins checkpoint / plain state_dict): `ckpt['state_dict']`, 'module.' stripped; EMA copies, the
    data_preprocessor buffers and BatchNorm counters carry no inference weights and are dropped."""
    obj = torch_load(ckpt) if isinstance(ckpt, str) else ckpt
    sd = obj.get("state_dict", obj) if isinstance(obj, dict) else obj
    sd = _strip(sd)
    return OrderedDict((k, v) for k, v in sd.items() if not k.startswith(("ema_", "data_preprocessor.")) and not k.endswith("num_batches_tracked"))


def leres_state_dict(ckpt):
    """res101.pth: `strip_prefix_if_present(checkpoint['depth_model'], 'module.')` loaded into RelDepthModel (keys `depth_model.*`)."""
    obj = torch_load(ckpt) if isinstance(ckpt, str) else ckpt
    sd = _strip(obj.get("depth_model", obj))
    return OrderedDict(((k if k.startswith("depth_model.") else "depth_model." + k), v) for k, v in sd.items())


def zoe_state_dict(ckpt):
    """ZoeD_M12_N.pt: `state_dict.get('model', state_dict)` with 'module.' stripped (model_io.py:27-47): `core.core.*` + the metric head."""
    obj = torch_load(ckpt) if isinstance(ckpt, str) else ckpt
    return _strip(obj.get("model", obj))


def plain_state_dict(ckpt):
    """refine_last.ckpt, the Inpaint / Refine net checkpoints: a plain state_dict (a {'state_dict': ...} wrapper is accepted too)."""
    obj = torch_load(ckpt) if isinstance(ckpt, str) else ckpt
    sd = obj.get("state_dict", obj) if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict) else obj
    return _strip(sd)
