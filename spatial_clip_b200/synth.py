"""Seeded synthetic inputs for the contrastive-loss hot path.

Everything is generated on the CPU with a ``torch.Generator`` so that the golden
fixtures, the CPU oracle, the GPU parity tests and ``bench.py`` all see the same
numbers for the same ``(N, D, K, seed)``.

Shapes follow the reference's collate contract
(``/root/reference/src/data/spatial_datamodule.py:111-137``,
``/root/reference/src/open_clip_train/spatial_data.py:74-77``):
tile ids int64 ``[N]``, neighbour ids int64 ``[N, K]`` padded with ``-1``,
neighbour alphas fp32 ``[N, K]`` padded with ``0.0``.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class SpotBatch:
    """One global batch of N (image tile, gene sentence) pairs."""

    image_features: torch.Tensor  # [N, D] fp32, L2-normalised
    text_features: torch.Tensor  # [N, D] fp32, L2-normalised
    tile_ids: torch.Tensor  # [N] int64
    neighbor_tile_ids: torch.Tensor  # [N, K] int64, -1 padded
    neighbor_alphas: torch.Tensor  # [N, K] fp32, 0 padded

    def rank_slice(self, rank: int, world_size: int) -> "SpotBatch":
        n = self.image_features.shape[0]
        assert n % world_size == 0
        b = n // world_size
        sl = slice(rank * b, (rank + 1) * b)
        return SpotBatch(
            self.image_features[sl],
            self.text_features[sl],
            self.tile_ids[sl],
            self.neighbor_tile_ids[sl],
            self.neighbor_alphas[sl],
        )


def make_spot_batch(
    n: int,
    d: int,
    k: int,
    seed: int,
    dup_frac: float = 0.0,
    self_loops: bool = False,
    negative_alphas: bool = False,
) -> SpotBatch:
    """SURVEY.md §8(d) generator.

    * ``I = normalize(randn)``, ``T = normalize(0.6 I + 0.8 normalize(randn))``
    * ids = ``randperm(10 N)[:N]``; ``dup_frac`` of them overwritten by another
      row's id (duplicate tile ids, as DistributedSampler padding produces)
    * each neighbour slot: p=0.5 an id of the batch (not guaranteed distinct),
      else an absent id ``>= 10 N``; valid count ``k_i ~ U{0..K}``, the rest
      padded ``-1`` / ``0.0``; alphas = softmax(randn) over the valid slots
    * ``self_loops`` forces slot 0 of every 7th row to the row's own id
      (reference dataset fixture has a ``1 -> 1`` edge,
      ``/root/reference/tests/test_spatial_datasets.py:45-51``)
    * ``negative_alphas`` flips the sign of every 5th valid alpha
    """
    g = torch.Generator().manual_seed(seed)
    img = F.normalize(torch.randn(n, d, generator=g), dim=-1)
    noise = F.normalize(torch.randn(n, d, generator=g), dim=-1)
    txt = F.normalize(0.6 * img + 0.8 * noise, dim=-1)

    ids = torch.randperm(10 * n, generator=g)[:n].to(torch.int64)
    if dup_frac > 0:
        n_dup = max(1, int(round(dup_frac * n)))
        dst = torch.randperm(n, generator=g)[:n_dup]
        src = torch.randint(0, n, (n_dup,), generator=g)
        ids = ids.clone()
        ids[dst] = ids[src]

    if k > 0:
        in_batch = torch.rand(n, k, generator=g) < 0.5
        pick = torch.randint(0, n, (n, k), generator=g)
        absent = 10 * n + torch.randint(0, n, (n, k), generator=g)
        nbr = torch.where(in_batch, ids[pick], absent).to(torch.int64)
        valid_cnt = torch.randint(0, k + 1, (n,), generator=g)
        valid = torch.arange(k).unsqueeze(0) < valid_cnt.unsqueeze(1)
        raw = torch.randn(n, k, generator=g)
        raw = raw.masked_fill(~valid, float("-inf"))
        alpha = torch.softmax(raw, dim=1)
        alpha = torch.where(valid, alpha, torch.zeros_like(alpha)).float()
        alpha = torch.nan_to_num(alpha, nan=0.0)
        nbr = torch.where(valid, nbr, torch.full_like(nbr, -1))
        if self_loops:
            rows = torch.arange(0, n, 7)
            nbr[rows, 0] = ids[rows]
            alpha[rows, 0] = torch.clamp(alpha[rows, 0], min=0.125)
        if negative_alphas:
            flat = alpha.view(-1)
            nz = torch.nonzero(flat > 0).squeeze(1)
            flat[nz[::5]] *= -1.0
    else:
        nbr = torch.empty(n, 0, dtype=torch.int64)
        alpha = torch.empty(n, 0, dtype=torch.float32)

    return SpotBatch(img, txt, ids, nbr.contiguous(), alpha.contiguous())


LOGIT_SCALES = {"init": 1.0 / 0.07, "capped": 55.0, "max": 100.0}


def shuffled_text_ids(tile_ids: torch.Tensor, seed: int, frac: float = 0.125) -> torch.Tensor:
    """Text-side tile ids that differ from the image-side ones: a seeded ``frac`` of the positions is permuted among
    itself (the reference API takes the two id vectors separately, losses.py:44-55, and resolves neighbours of image
    rows in the TEXT id map and vice versa, losses.py:92-108; its own data loader always passes equal vectors)."""
    g = torch.Generator().manual_seed(seed)
    n = tile_ids.shape[0]
    m = max(2, int(round(frac * n)))
    pos = torch.randperm(n, generator=g)[:m]
    out = tile_ids.clone()
    out[pos] = tile_ids[pos.roll(1)]
    return out
