"""Load a myria3d Lightning checkpoint without Lightning / omegaconf / torchmetrics.

The shipped checkpoint (``trained_model_assets/proto151_V2.0_epoch_100_Myria3DV3.1.0.ckpt``)
pickles omegaconf containers and torchmetrics objects inside ``hyper_parameters`` /
``callbacks``; the state dict itself is plain tensors with the ``model.`` prefix of
``myria3d/models/model.py:62``.  A stub unpickler replaces every non-torch class by an
inert placeholder (only an explicit whitelist of tensor-rebuilding helpers and plain containers is resolved for
real) so that ``state_dict`` can be read on a box that has none of them
(reference behaviour being replaced: ``Model.load_from_checkpoint`` at
``myria3d/predict.py:49`` / ``myria3d/train.py:167``).
"""
from __future__ import annotations

import pickle
from typing import Any, Dict

import torch

# Globals a Lightning checkpoint legitimately needs to rebuild its tensors and containers.  Everything else --
# omegaconf / torchmetrics classes, but also builtins.eval / exec / getattr / __import__ and the rest of torch, numpy
# and builtins -- becomes an inert stub: unpickling a checkpoint cannot call into arbitrary code.
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"),
    ("builtins", "bytearray"), ("builtins", "complex"), ("builtins", "slice"), ("builtins", "range"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_qtensor"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch._tensor", "_rebuild_from_type_v2"), ("torch.serialization", "_get_layout"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("_codecs", "encode"),
}
_ALLOWED_TORCH_SUFFIXES = ("Storage",)  # torch.FloatStorage, torch.LongStorage, torch.storage.UntypedStorage, ...


_TORCH_DTYPE_NAMES = {n for n in dir(torch) if isinstance(getattr(torch, n), torch.dtype)}


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state

    def __call__(self, *a, **k):
        return _Stub()


def _make_stub(module: str, name: str):
    return type(name, (_Stub,), {"__module__": module})


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if (module, name) in _ALLOWED_GLOBALS:
            return super().find_class(module, name)
        if module.split(".")[0] == "torch" and name.endswith(_ALLOWED_TORCH_SUFFIXES):
            return super().find_class(module, name)
        if module == "torch" and name in _TORCH_DTYPE_NAMES:
            return super().find_class(module, name)
        return _make_stub(module, name)


class _StubPickleModule:
    __name__ = "stub_pickle"
    Unpickler = _StubUnpickler
    load = staticmethod(pickle.load)


def load_lightning_checkpoint(path: str) -> Dict[str, Any]:
    """Raw checkpoint dict (``state_dict`` usable; other entries may hold stubs)."""
    return torch.load(path, map_location="cpu", pickle_module=_StubPickleModule, weights_only=False)


def net_state_dict(ckpt: Dict[str, Any], prefix: str = "model.") -> Dict[str, torch.Tensor]:
    """State dict of the neural net (``Model.model``), prefix stripped."""
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
