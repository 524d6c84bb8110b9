"""GPU mirror of the numeric half of ``myria3d/models/interpolation.py::Interpolator``.

Same constructor, ``store_predictions`` and ``reduce_predicted_logits`` as the reference (``interpolation.py:21-58,
93-121``); the scatter/softmax/argmax/entropy work runs in ``libb200randla`` (``stitch.cu``) instead of
torch_scatter + CPU torch.  Reading and writing the LAS file through PDAL (``load_full_las_for_update``,
the writer part of ``reduce_predictions_and_save``, ``:60-91,168-185``) is file I/O outside the hot path:
``reduce_predictions`` returns the arrays that the reference writes into the LAS dimensions, keyed by channel name.
"""
from __future__ import annotations

from typing import Dict, List, Literal, Optional, Tuple, Union

import numpy as np
import torch

from . import ops


class Interpolator:
    def __init__(
        self,
        interpolation_k: int = 10,
        classification_dict: Dict[int, str] = {},
        probas_to_save: Union[List[str], Literal["all"]] = "all",
        predicted_classification_channel: Optional[str] = "PredictedClassification",
        entropy_channel: Optional[str] = "entropy",
    ):
        self.k = interpolation_k
        self.classification_dict = classification_dict
        self.predicted_classification_channel = predicted_classification_channel
        self.entropy_channel = entropy_channel
        if probas_to_save == "all":
            self.probas_to_save = list(classification_dict.values())
        elif probas_to_save is None:
            self.probas_to_save = []
        else:
            self.probas_to_save = probas_to_save
        # ascending class index -> LAS classification code (interpolation.py:50-54)
        self.reverse_mapper: Dict[int, int] = {i: code for i, code in enumerate(classification_dict.keys())}
        self.logits: List[torch.Tensor] = []
        self.idx_in_full_cloud_list: List[np.ndarray] = []

    def store_predictions(self, logits: torch.Tensor, idx_in_original_cloud: List[np.ndarray]) -> None:
        """Keep the predictions made so far (``interpolation.py:93-96``); logits stay wherever they are (GPU or CPU)."""
        self.logits += [logits]
        self.idx_in_full_cloud_list += idx_in_original_cloud

    def _device(self) -> torch.device:
        for l in self.logits:
            if l.is_cuda:
                return l.device
        return torch.device("cuda", torch.cuda.current_device())

    @torch.no_grad()
    def _reduce(self, nb_points: int):
        dev = self._device()
        logits = torch.cat([l.to(dev, torch.float32) for l in self.logits])
        idx_np = np.concatenate(self.idx_in_full_cloud_list)
        del self.logits
        del self.idx_in_full_cloud_list
        idx = torch.from_numpy(np.ascontiguousarray(idx_np).astype(np.int64, copy=False)).to(dev)
        reduced = ops.stitch_scatter_sum(logits, idx, int(nb_points))
        return reduced, idx, idx_np

    @torch.no_grad()
    def reduce_predicted_logits(self, nb_points: int) -> Tuple[torch.Tensor, np.ndarray]:
        """Sum the logits of points predicted several times and re-select them in prediction order
        (``interpolation.py:98-121``).  Returns ``(logits[len(idx), C] on the GPU, idx_in_full_cloud)``."""
        reduced, idx, idx_np = self._reduce(nb_points)
        logits, _, _, _ = ops.stitch_finalize(reduced, idx, want_logits=True)
        return logits, idx_np

    @torch.no_grad()
    def reduce_predictions(self, nb_points: int) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
        """The values ``reduce_predictions_and_save`` assigns to ``las[channel][idx_in_full_cloud]``
        (``interpolation.py:139-167``): one probability array per saved class name, the predicted LAS codes under
        ``predicted_classification_channel`` and the entropy under ``entropy_channel``."""
        reduced, idx, idx_np = self._reduce(nb_points)
        _, probas, preds, entropy = ops.stitch_finalize(reduced, idx, want_logits=False)
        out: Dict[str, np.ndarray] = {}
        probas_np = probas.cpu().numpy()
        for i, class_name in enumerate(self.classification_dict.values()):
            if class_name in self.probas_to_save:
                out[class_name] = probas_np[:, i]
        if self.predicted_classification_channel:
            codes = np.array([self.reverse_mapper[i] for i in range(len(self.reverse_mapper))], dtype=np.int64)
            out[self.predicted_classification_channel] = codes[preds.cpu().numpy()]
        if self.entropy_channel:
            out[self.entropy_channel] = entropy.cpu().numpy()
        return out, idx_np
