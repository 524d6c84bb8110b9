"""Load a myria3d Lightning checkpoint without Lightning / omegaconf / torchmetrics.

The shipped checkpoint (``trained_model_assets/proto151_V2.0_epoch_100_Myria3DV3.1.0.ckpt``)
pickles omegaconf containers and torchmetrics objects inside ``hyper_parameters`` /
``callbacks``; the state dict itself is plain tensors with the ``model.`` prefix of
``myria3d/models/model.py:62``.  A stub unpickler replaces every non-torch class by an
inert placeholder so that ``state_dict`` can be read on a box that has none of them
(reference behaviour being replaced: ``Model.load_from_checkpoint`` at
``myria3d/predict.py:49`` / ``myria3d/train.py:167``).
"""
from __future__ import annotations

import pickle
from typing import Any, Dict

import torch

_ALLOWED_PREFIXES = ("torch", "collections", "numpy", "builtins", "_codecs")


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
        if module.split(".")[0] in _ALLOWED_PREFIXES:
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
