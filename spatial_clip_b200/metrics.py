"""In-batch retrieval metrics from the ranks the fused forward pass counts.

The reference's LightningModule re-materialises ``logits = image_features @ text_features.T * logit_scale`` every step
and feeds it to ``RecallAtK`` (``torch.topk`` + membership test) only to learn where the matching profile ranks
(/root/reference/src/models/spatial_clip_module.py:68,106,113,127; src/models/components/metrics.py:7-36).  With
``track_retrieval_ranks=True`` the loss modules expose that rank directly (``last_retrieval_ranks``: for every local
image row, how many gene profiles of the local batch score above its own), counted inside the similarity pass:

    loss_fn = SpatialLoss(..., track_retrieval_ranks=True)
    out = loss_fn(**loss_input)
    metrics.update(loss_fn.last_retrieval_ranks)         # instead of metrics(logits, arange)

``RecallAtKFromRanks`` keeps the reference metric's state (``correct`` / ``total`` counters, ``k`` capped by the batch
size exactly like ``k_eff = min(k, logits.size(1))``) without depending on torchmetrics; ties are counted in favour of
the matching pair (``torch.topk``'s tie order is unspecified).
"""
from __future__ import annotations

from typing import Dict, Iterable

import torch


def recall_at_k(ranks: torch.Tensor, k: int) -> torch.Tensor:
    """Fraction of rows whose matching column is among the k best of the local batch (metrics.py:22-36)."""
    k_eff = min(int(k), int(ranks.numel()))  # a batch of B rows has B candidate columns
    return (ranks < k_eff).float().mean()


class RecallAtKFromRanks:
    """Accumulating Recall@k over batches (``correct`` / ``total`` like the reference's torchmetrics Metric)."""

    def __init__(self, ks: Iterable[int] = (1, 5, 10), prefix: str = ""):
        self.ks = tuple(int(k) for k in ks)
        self.prefix = prefix
        self.reset()

    def reset(self) -> None:
        self.correct = {k: 0 for k in self.ks}
        self.total = 0

    @torch.no_grad()
    def update(self, ranks: torch.Tensor) -> None:
        n = int(ranks.numel())
        if n == 0:
            return
        hits = torch.stack([(ranks < min(k, n)).sum() for k in self.ks]).tolist()  # one device -> host read
        for k, h in zip(self.ks, hits):
            self.correct[k] += int(h)
        self.total += n

    def compute(self) -> Dict[str, float]:
        t = max(self.total, 1)
        return {f"{self.prefix}R@{k}": self.correct[k] / t for k in self.ks}
