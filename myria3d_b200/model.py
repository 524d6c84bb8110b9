"""LightningModule-shaped wrapper: drop-in for ``myria3d.models.model.Model`` (``models/model.py:32-198``).

Keeps the model-zoo factory (substring match, ``model.py:15-29``), ``forward(batch) -> (targets, logits)``,
``training_step / validation_step / test_step -> {"loss", "logits", "targets"}``, ``predict_step`` and
``configure_optimizers``.  Subclasses ``pytorch_lightning.LightningModule`` when Lightning is importable,
otherwise a plain ``nn.Module`` with the few Lightning facilities the class uses (``save_hyperparameters``,
``hparams``, ``log``), so the hot path can run on a box without Lightning.

Difference from the reference, on purpose: the eval-time interpolation of logits to the
full cloud (``model.py:86-98``) stays ON THE GPU -- ``b200_knn`` (k = interpolation_k) + ``b200_knn_interp``
-- instead of ``logits.cpu()`` + CPU kNN.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Optional

import torch
from torch import nn

from . import ops
from .randla_net import B200Block1Net, B200RandLANet

try:  # pragma: no cover - depends on the environment
    from pytorch_lightning import LightningModule as _Base

    _HAS_LIGHTNING = True
except Exception:  # Lightning absent: minimal shim
    _HAS_LIGHTNING = False

    class _Base(nn.Module):  # type: ignore[no-redef]
        def save_hyperparameters(self, ignore=()):
            pass

        def log(self, *args, **kwargs):
            pass


MODEL_ZOO = [B200RandLANet, B200Block1Net]  # the full net first: "RandLANet" keeps resolving to it


def get_neural_net_class(class_name: str) -> nn.Module:
    """Same factory as ``models/model.py:15-29``: first zoo class whose name CONTAINS ``class_name``.

    ``"B200RandLANet"`` resolves here; ``"RandLANet"`` does too (substring), ``"PyGRandLANet"`` does not,
    so a stock config keeps selecting the reference implementation.
    """
    for neural_net_class in MODEL_ZOO:
        if class_name in neural_net_class.__name__:
            return neural_net_class
    raise KeyError(f"Unknown class name {class_name}")


class Model(_Base):
    """See module docstring.  ``kwargs`` as in the reference: ``neural_net_class_name``,
    ``neural_net_hparams``, ``criterion``, ``interpolation_k``, ``num_workers``, ``lr``, ``optimizer``,
    ``lr_scheduler``, ``monitor`` ... (``configs/model/default.yaml``)."""

    def __init__(self, **kwargs: Any):
        super().__init__()
        if _HAS_LIGHTNING:
            self.save_hyperparameters(ignore=["criterion"])
        else:
            self._hparams_shim = SimpleNamespace(**{k: v for k, v in kwargs.items() if k != "criterion"})
        neural_net_class = get_neural_net_class(kwargs.get("neural_net_class_name"))
        self.model = neural_net_class(**kwargs.get("neural_net_hparams"))
        self.softmax = nn.Softmax(dim=1)
        self.criterion = kwargs.get("criterion")

    if not _HAS_LIGHTNING:

        @property
        def hparams(self):
            return self._hparams_shim

    def forward(self, batch) -> tuple:
        """``(targets, logits)`` like ``models/model.py:67-103``."""
        logits = self.model(batch.x, batch.pos, batch.batch, batch.ptr)
        if self.training or "copies" not in batch:
            return batch.y, logits
        # evaluation on the full cloud: k-NN inverse-distance interpolation (model.py:86-98), on the GPU
        copies = batch.copies
        dev = logits.device
        pos_sub = copies["pos_sampled_copy"].to(dev, torch.float32)
        pos_full = copies["pos_copy"].to(dev, torch.float32)
        sizes_y = [len(s) for s in batch.idx_in_original_cloud]  # == _get_batch_tensor_by_enumeration
        ptr_y_host = [0]
        for n in sizes_y:
            ptr_y_host.append(ptr_y_host[-1] + int(n))
        ptr_y = torch.tensor(ptr_y_host, dtype=torch.int64, device=dev)
        ptr_x = batch.ptr.to(dev, torch.int64)
        k = int(getattr(self.hparams, "interpolation_k", 10))
        ptr_x_host = batch.ptr.tolist()
        max_x = max((int(ptr_x_host[i + 1]) - int(ptr_x_host[i]) for i in range(len(ptr_x_host) - 1)), default=0)
        nbr, dist2 = ops.knn(pos_sub, ptr_x, pos_full, ptr_y, k, max(sizes_y) if sizes_y else 0, kt=k,
                             max_points_per_cloud=max_x)
        logits = ops.knn_interpolate_from_table(logits, nbr, dist2, k)
        targets = None
        if "transformed_y_copy" in copies:
            targets = copies["transformed_y_copy"].to(logits.device)
        return targets, logits

    def _step(self, batch, log_name: str, **log_kwargs) -> Dict[str, Any]:
        targets, logits = self.forward(batch)
        self.criterion = self.criterion.to(logits.device)
        loss = self._loss(logits, targets)
        self.log(log_name, loss, **log_kwargs)
        return {"loss": loss, "logits": logits, "targets": targets}

    def _loss(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """``self.criterion(logits, targets)`` (models/model.py:118); the configured criterion -- a plain
        ``torch.nn.CrossEntropyLoss(weight?, ignore_index=65, label_smoothing=0.0)`` (configs/model/criterion/*.yaml) --
        runs as the library's fused kernel pair (same value and gradient; torch's nll_loss kernels reduce with a single
        CTA and cost 0.33 ms per 204 800-point step), anything else is called as given."""
        crit = self.criterion
        if (type(crit) is nn.CrossEntropyLoss and crit.reduction == "mean" and crit.label_smoothing == 0.0
                and logits.is_cuda and logits.dim() == 2 and logits.shape[1] <= ops.CROSS_ENTROPY_MAX_CLASSES
                and targets is not None and targets.dtype == torch.int64 and targets.dim() == 1):
            return ops.cross_entropy(logits, targets, crit.weight, crit.ignore_index)
        return crit(logits, targets)

    def training_step(self, batch, batch_idx: int) -> dict:
        return self._step(batch, "train/loss", on_step=True, on_epoch=True, prog_bar=False)

    def validation_step(self, batch, batch_idx: int) -> dict:
        return self._step(batch, "val/loss", on_step=True, on_epoch=True)

    def test_step(self, batch, batch_idx: int) -> dict:
        return self._step(batch, "test/loss", on_step=False, on_epoch=True)

    def predict_step(self, batch, batch_idx: Optional[int] = None) -> dict:
        _, logits = self.forward(batch)
        return {"logits": logits.detach().cpu()}

    def configure_optimizers(self):
        self.lr = self.hparams.lr
        optimizer = self.hparams.optimizer(params=filter(lambda p: p.requires_grad, self.parameters()), lr=self.lr)
        if getattr(self.hparams, "lr_scheduler", None) is None:
            return optimizer
        return {
            "optimizer": optimizer,
            "lr_scheduler": self.hparams.lr_scheduler(optimizer),
            "monitor": self.hparams.monitor,
        }

    def _get_batch_tensor_by_enumeration(self, pos_x) -> torch.Tensor:
        return torch.cat([torch.full((len(sample_pos),), i) for i, sample_pos in enumerate(pos_x)])
