"""Data-side resolution of the spatial-neighbour soft targets.

The reference resolves neighbour tile ids to columns of the global batch INSIDE every loss call: a Python dict over
all gathered ids and a double loop with ``.item()`` reads (/root/reference/src/models/components/losses.py:91-108),
after two id all-gathers (losses.py:63-68).  The kernels of this package do the same on the device
(``scl_build_positives``).  But the mapping only depends on what the sampler put into the global batch, so the data
pipeline can produce it ahead of the step, on the CPU workers that already build the neighbour lists
(/root/reference/src/data/spatial_datamodule.py:111-137, src/open_clip_train/spatial_data.py:37-85): the loss then
receives int32 columns + fp32 weights instead of int64 ids, and neither the id exchange nor the hash build is part
of the step any more (``SpatialLossFromColumns`` in losses.py).

``resolve_positive_columns`` is that producer: plain numpy on the host (this is collate-time integer work, not the
GPU hot path), vectorised over rows, and bit-exact against the reference's dense label rows -- same "last duplicate
id wins" map, same slot order, same fp32 accumulation order, same ``F.normalize(p=1)`` divide.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["resolve_positive_columns", "collate_positive_columns"]


def _last_index_lookup(all_ids: np.ndarray):
    """Sorted unique ids and, for each, the LAST position it occupies (dict comprehension, losses.py:92-93)."""
    n = all_ids.shape[0]
    uniq, first_in_reversed = np.unique(all_ids[::-1], return_index=True)
    return uniq, (n - 1 - first_in_reversed).astype(np.int64)


def resolve_positive_columns(all_tile_ids, neighbor_tile_ids, neighbor_alphas, neighbor_alpha_scale: float = 1.0,
                             rank: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """ELL soft-target lists of the local rows.

    all_tile_ids      int64 [N]       tile ids of the GLOBAL batch in rank-major order (what the id all-gather of
                                      losses.py:63-68 would return); for image rows pass the TEXT-side ids and vice
                                      versa (losses.py:102-108) -- the reference's loader makes them equal
    neighbor_tile_ids int64 [B_l, K]  padded with -1      (spatial_datamodule.py:127)
    neighbor_alphas   fp32  [B_l, K]  padded with 0.0     (spatial_datamodule.py:128)
    rank                              this rank's row block: local row i is global column rank * B_l + i

    Returns (columns int32 [B_l, K+1], weights fp32 [B_l, K+1], probs fp32 [B_l, K+1]): slot 0 is the row's own
    column with weight 1; further slots in first-touch order; unused slots are (-1, 0, 0).  ``weights`` are the
    reference's label values before ``F.normalize`` (bit-exact), ``probs`` after it.
    """
    ids = np.ascontiguousarray(torch.as_tensor(all_tile_ids).cpu().numpy().astype(np.int64, copy=False))
    nbr = np.ascontiguousarray(torch.as_tensor(neighbor_tile_ids).cpu().numpy().astype(np.int64, copy=False))
    alpha = np.ascontiguousarray(torch.as_tensor(neighbor_alphas).cpu().numpy().astype(np.float32, copy=False))
    if nbr.ndim != 2 or alpha.shape != nbr.shape:
        raise ValueError("neighbor_tile_ids / neighbor_alphas must both be [B_l, K]")
    b_local, k = nbr.shape
    if ids.ndim != 1 or ids.shape[0] < (rank + 1) * b_local:
        raise ValueError("all_tile_ids must hold the whole global batch (at least (rank + 1) * B_l ids)")
    kp1 = k + 1
    col = np.full((b_local, kp1), -1, dtype=np.int32)
    w = np.zeros((b_local, kp1), dtype=np.float32)
    col[:, 0] = rank * b_local + np.arange(b_local, dtype=np.int32)
    w[:, 0] = np.float32(1.0)
    cnt = np.ones(b_local, dtype=np.int64)
    rows = np.arange(b_local)
    if k > 0:
        uniq, last = _last_index_lookup(ids)
        a = np.maximum(alpha * np.float32(neighbor_alpha_scale), np.float32(0.0)).astype(np.float32)  # losses.py:100
        pos = np.searchsorted(uniq, nbr)
        pos_c = np.minimum(pos, uniq.shape[0] - 1)
        found = uniq[pos_c] == nbr
        target = np.where(found, last[pos_c], -1).astype(np.int32)
        slots = np.arange(kp1)[None, :]
        for s in range(k):  # slot order matters: fp32 accumulation and first-touch placement (losses.py:102-108)
            c = target[:, s]
            ok = (a[:, s] > 0) & (c >= 0)
            same = (col == c[:, None]) & (slots < cnt[:, None]) & ok[:, None]
            hit = same.any(axis=1)
            hit_slot = same.argmax(axis=1)
            r_hit = rows[hit]
            w[r_hit, hit_slot[hit]] = (w[r_hit, hit_slot[hit]] + a[r_hit, s]).astype(np.float32)
            new = ok & ~hit
            r_new = rows[new]
            col[r_new, cnt[new]] = c[new]
            w[r_new, cnt[new]] = a[r_new, s]
            cnt[new] += 1
    tot = np.zeros(b_local, dtype=np.float32)
    for t in range(kp1):  # sequential fp32 sum over the used slots (unused slots add +0.0)
        tot = (tot + w[:, t]).astype(np.float32)
    inv = (np.float32(1.0) / np.maximum(tot, np.float32(1e-12))).astype(np.float32)  # F.normalize(p=1) eps
    q = (w * inv[:, None]).astype(np.float32)
    return torch.from_numpy(col), torch.from_numpy(w), torch.from_numpy(q)


def collate_positive_columns(batch: dict, neighbor_alpha_scale: float = 1.0, all_tile_ids=None,
                             rank: Optional[int] = None) -> dict:
    """Add ``positive_columns`` / ``positive_weights`` / ``positive_probs`` to a collated batch dictionary
    (the dictionary ``SpatialDataModule._collate_fn`` returns, spatial_datamodule.py:111-137).

    Single process: the local batch IS the global batch, so the ids come from the batch itself.  Several ranks: the
    columns are positions in the GLOBAL batch, so ``all_tile_ids`` (the sampler knows the global batch order) and
    ``rank`` are required -- when torch.distributed is initialised with more than one rank, leaving them out raises
    instead of silently producing rank 0's columns on every rank."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if multi and (all_tile_ids is None or rank is None):
        raise ValueError("collate_positive_columns: torch.distributed is initialised with world_size > 1 -- pass "
                         "all_tile_ids (tile ids of the global batch, rank-major) and rank")
    ids = batch["text_tile_ids"] if all_tile_ids is None else all_tile_ids
    col, w, q = resolve_positive_columns(ids, batch["neighbor_tile_ids"], batch["neighbor_alphas"],
                                         neighbor_alpha_scale, 0 if rank is None else rank)
    out = dict(batch)
    out.update(positive_columns=col, positive_weights=w, positive_probs=q)
    return out
