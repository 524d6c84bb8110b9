"""CPU oracle of the sliding-window stitch (TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``; PARITY UNPINNED:
the reference holds no golden vectors for this step).

Restates ``myria3d/models/interpolation.py``:
  * ``reduce_predicted_logits`` (:98-121): ``torch.cat`` of the stored logits, ``torch_scatter.scatter_sum(logits, idx,
    out=zeros(nb_points, C), dim=0)`` -- for a 1-D index torch_scatter broadcasts it and calls
    ``out.scatter_add_(0, index, src)``, which the CPU backend executes sequentially in input order -- then
    ``reduced[idx]``;
  * ``reduce_predictions_and_save`` (:139-167): ``Softmax(dim=1)``, ``argmax(dim=1)`` mapped through
    ``reverse_mapper``, ``Categorical(probs=probas).entropy()``.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from torch.distributions import Categorical


def scatter_sum_rows(logits: torch.Tensor, idx: torch.Tensor, nb_points: int) -> torch.Tensor:
    """interpolation.py:115-116 (torch_scatter.scatter_sum with ``out=``)."""
    out = torch.zeros((nb_points, logits.size(1)), dtype=logits.dtype)
    index = idx.view(-1, 1).expand_as(logits)
    return out.scatter_add_(0, index, logits)


def scatter_sum_rows_loop(logits: torch.Tensor, idx: Sequence[int], nb_points: int) -> torch.Tensor:
    """The same sum as an explicit Python loop in input order (small cases; cross-check of ``scatter_add_``'s order)."""
    out = torch.zeros((nb_points, logits.size(1)), dtype=logits.dtype)
    for r, d in enumerate(idx):
        out[int(d)] = out[int(d)] + logits[r]
    return out


def reduce_predicted_logits(logits_list: List[torch.Tensor], idx_list: List[np.ndarray], nb_points: int):
    logits = torch.cat(logits_list).cpu()
    idx_np = np.concatenate(idx_list)
    reduced = scatter_sum_rows(logits, torch.from_numpy(idx_np), nb_points)
    return reduced[idx_np], idx_np


def reduce_predictions(logits_list, idx_list, nb_points: int, classification_dict: Dict[int, str]) -> Tuple[dict, np.ndarray]:
    logits, idx_np = reduce_predicted_logits(logits_list, idx_list, nb_points)
    probas = torch.nn.Softmax(dim=1)(logits)
    preds = torch.argmax(logits, dim=1)
    reverse_mapper = {i: code for i, code in enumerate(classification_dict.keys())}
    preds = np.vectorize(reverse_mapper.get)(preds)
    entropy = Categorical(probs=probas).entropy()
    return {"probas": probas, "preds": preds, "entropy": entropy, "logits": logits}, idx_np
