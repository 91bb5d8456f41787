"""Blockwise fp64 CPU oracle for BASELINE-size batches (N = 16384 / 32768).  TEST INFRASTRUCTURE ONLY.

Same arithmetic as ``oracle/contrastive_oracle.py`` (the dense restatement pinned to the reference goldens;
``tests/test_oracle.py::test_blockwise_oracle_matches_dense_oracle`` checks the two against each other on every
flag combination), organised so that nothing of size N x N is ever held: ONE pass over row blocks of the similarity
matrix yields the row statistics of both directions (the text-row direction is the column direction of the same
matrix, reference: src/models/components/losses.py:78-79), every rank's loss / gap / d logit_scale follow from
them, and gradients are evaluated in closed form for a SAMPLE of rows (the full [N, D] gradient would be a second
N x N x D pass per modality).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s parity block import
this file; the product never does.

Reference lines restated (relative to /root/reference): src/models/components/losses.py:73-122 (cap, logits, soft
labels, soft CE both directions, temperature regulariser), src/open_clip/loss.py:21-65 (which gathered tensors
carry gradient), :91-155 (plain CLIP).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from .contrastive_oracle import soft_label_triples


@dataclass
class BlockwiseResult:
    loss: np.ndarray  # [W]  per-rank loss
    gap: np.ndarray  # [W]
    d_scale: np.ndarray  # [W]  d loss_r / d logit_scale
    rows: np.ndarray  # [S]  global indices of the sampled rows
    d_image_rows: np.ndarray  # [S, D]  what image_features.grad holds for those rows on their owner rank
    d_text_rows: np.ndarray  # [S, D]
    lse_img: np.ndarray  # [N] nats
    lse_txt: np.ndarray


def _ell(ids_cols, nbr_ids, nbr_alpha, alpha_scale, world):
    """Soft-target lists of ALL rows as flat (row, col, q) arrays (losses.py:91-111), rank by rank."""
    n = nbr_ids.shape[0]
    b = n // world
    rr, cc, qq = [], [], []
    for r in range(world):
        sl = slice(r * b, (r + 1) * b)
        rows, sums = soft_label_triples(ids_cols, nbr_ids[sl], nbr_alpha[sl], alpha_scale, r)
        for i, lst in enumerate(rows):
            tot = max(float(sums[i]), 1e-12)
            for col, w in lst:
                rr.append(r * b + i)
                cc.append(col)
                qq.append(float(w) / tot)
    return np.asarray(rr, dtype=np.int64), np.asarray(cc, dtype=np.int64), np.asarray(qq, dtype=np.float64)


def blockwise_oracle(image: np.ndarray, text: np.ndarray, logit_scale: float, image_ids: Optional[np.ndarray],
                     text_ids: Optional[np.ndarray], nbr_ids: Optional[np.ndarray], nbr_alpha: Optional[np.ndarray],
                     world_size: int = 1, cap_logit_scale: Optional[float] = None, temp_reg_weight: float = 0.0,
                     neighbor_alpha_scale: float = 1.0, local_loss: bool = True, gather_with_grad: bool = True,
                     sample_rows: Sequence[int] = (), block: int = 2048, threads: Optional[int] = None,
                     kind: str = "spatial") -> BlockwiseResult:
    """SpatialLoss (kind="spatial") or local-loss ClipLoss (kind="clip", ids / neighbours ignored) at any N.

    ``image`` / ``text``: GLOBAL [N, D] arrays in rank-major order (what the kernels see: pass bf16-rounded values
    to compare against the bf16 mode).  Gradients are those of ``sum_r loss_r`` with respect to the features of each
    row's owner rank, with the reference's gradient routing for ``local_loss`` / ``gather_with_grad``."""
    if threads:
        torch.set_num_threads(threads)
    n, d = image.shape
    assert n % world_size == 0
    b = n // world_size
    img = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float64))
    txt = torch.from_numpy(np.ascontiguousarray(text, dtype=np.float64))
    s = float(logit_scale)
    if cap_logit_scale is not None:
        s = min(s, float(cap_logit_scale))
    w = float(temp_reg_weight)
    c = 0.5 / b

    # ---- one pass: row (image-direction) and column (text-direction) statistics
    lse_i = torch.empty(n, dtype=torch.float64)
    mu_i = torch.empty(n, dtype=torch.float64)
    m2_i = torch.empty(n, dtype=torch.float64)
    cmax = torch.full((n,), -np.inf, dtype=torch.float64)
    c0 = torch.zeros(n, dtype=torch.float64)
    c1 = torch.zeros(n, dtype=torch.float64)
    c2 = torch.zeros(n, dtype=torch.float64)
    for r0 in range(0, n, block):
        z = img[r0:r0 + block] @ txt.T
        l = s * z
        m = l.max(dim=1, keepdim=True).values
        e = torch.exp(l - m)
        s0 = e.sum(1)
        lse_i[r0:r0 + block] = m[:, 0] + torch.log(s0)
        mu = (e * z).sum(1) / s0
        mu_i[r0:r0 + block] = mu
        m2_i[r0:r0 + block] = (e * z * z).sum(1) / s0
        bm = l.max(dim=0).values
        new = torch.maximum(cmax, bm)
        sc = torch.exp(cmax - new)
        sc[torch.isinf(cmax)] = 0.0
        e = torch.exp(l - new[None, :])
        c0 = c0 * sc + e.sum(0)
        e *= z
        c1 = c1 * sc + e.sum(0)
        e *= z
        c2 = c2 * sc + e.sum(0)
        cmax = new
        del z, l, e
    lse_t = cmax + torch.log(c0)
    mu_t = c1 / c0
    m2_t = c2 / c0
    var_i = m2_i - mu_i * mu_i
    var_t = m2_t - mu_t * mu_t

    # ---- soft targets (sparse) and their logits
    if kind == "spatial":
        ri, ci, qi = _ell(text_ids, nbr_ids, nbr_alpha, neighbor_alpha_scale, world_size)  # image rows -> text cols
        if image_ids is text_ids or np.array_equal(image_ids, text_ids):
            rt, ct, qt = ri, ci, qi
        else:
            rt, ct, qt = _ell(image_ids, nbr_ids, nbr_alpha, neighbor_alpha_scale, world_size)
    else:
        ri = ci = rt = ct = np.arange(n, dtype=np.int64)
        qi = qt = np.ones(n, dtype=np.float64)
    img_np, txt_np = img.numpy(), txt.numpy()

    def pair_dots(a, rows, bmat, cols):
        out = np.empty(len(rows), dtype=np.float64)
        for k0 in range(0, len(rows), 65536):
            sl = slice(k0, k0 + 65536)
            out[sl] = np.einsum("ij,ij->i", a[rows[sl]], bmat[cols[sl]])
        return out

    z_qi = pair_dots(img_np, ri, txt_np, ci)  # z[row, col] of every image-direction soft-target entry
    z_qt = pair_dots(txt_np, rt, img_np, ct)
    zq_i = np.bincount(ri, weights=qi * z_qi, minlength=n)
    zq_t = np.bincount(rt, weights=qt * z_qt, minlength=n)

    lse_i_n, lse_t_n = lse_i.numpy(), lse_t.numpy()
    mu_i_n, mu_t_n = mu_i.numpy(), mu_t.numpy()
    loss = np.empty(world_size)
    gap = np.zeros(world_size)
    d_scale = np.empty(world_size)
    for r in range(world_size):
        sl = slice(r * b, (r + 1) * b)
        ce = c * ((lse_i_n[sl] - s * zq_i[sl]).sum() + (lse_t_n[sl] - s * zq_t[sl]).sum())
        gsum = c * ((mu_i_n[sl] - zq_i[sl]).sum() + (mu_t_n[sl] - zq_t[sl]).sum())
        if w > 0:
            gap[r] = gsum
        loss[r] = ce + w * gap[r] * gap[r]
        d_scale[r] = gsum + 2.0 * w * gap[r] * c * (var_i.numpy()[sl].sum() + var_t.numpy()[sl].sum())
    k2 = 2.0 * w * gap  # per owner rank

    # ---- sampled gradient rows (closed forms of contrastive_oracle.spatial_loss_oracle)
    rows = np.asarray(list(sample_rows), dtype=np.int64)
    d_img = np.zeros((len(rows), d))
    d_txt = np.zeros((len(rows), d))
    if len(rows):
        owner_all = torch.from_numpy(np.arange(n) // b)
        k2_all = torch.from_numpy(k2)[owner_all]  # k2 of every row's owner rank
        rows_t = torch.from_numpy(rows)
        own_r = owner_all[rows_t]

        def col_mask():
            # which "other-direction" rows j reach a sampled row through the gathered tensor (loss.py:49-61)
            if gather_with_grad or world_size == 1:
                return None  # all
            if not local_loss:
                return owner_all[None, :] == own_r[:, None]  # only the re-spliced local slab
            return False  # none

        def dense_part(x, y, lse_row, mu_row, lse_col, mu_col):
            """sum_j [g_row(i, j) + g_col(j, i)] y_j for the sampled rows of x (dense, soft targets excluded)."""
            z = x[rows_t] @ y.T  # [S, N]
            k2_row = k2_all[rows_t][:, None]
            p = torch.exp(s * z - lse_row[rows_t][:, None])
            g = c * p * (s + k2_row * (1.0 + s * (z - mu_row[rows_t][:, None])))
            mask = col_mask()
            if mask is not False:
                pc = torch.exp(s * z - lse_col[None, :])
                gc = c * pc * (s + k2_all[None, :] * (1.0 + s * (z - mu_col[None, :])))
                if mask is not None:
                    gc = gc * mask
                g = g + gc
            return (g @ y).numpy()

        def sparse_part(out, y_np, r_own, c_own, q_own, r_opp, c_opp, q_opp):
            """- sum over soft targets: own lists (row i -> col) and the opposite lists that name column i."""
            pos = {int(g): k for k, g in enumerate(rows)}
            sel = np.isin(r_own, rows)
            for i, col, q in zip(r_own[sel], c_own[sel], q_own[sel]):
                out[pos[int(i)]] -= c * (s + k2[i // b]) * q * y_np[col]
            mask = col_mask()
            if mask is False:
                return
            sel = np.isin(c_opp, rows)
            for j, col, q in zip(r_opp[sel], c_opp[sel], q_opp[sel]):
                if mask is not None and j // b != col // b:
                    continue
                out[pos[int(col)]] -= c * (s + k2[j // b]) * q * y_np[j]

        d_img += dense_part(img, txt, lse_i, mu_i, lse_t, mu_t)
        sparse_part(d_img, txt_np, ri, ci, qi, rt, ct, qt)
        d_txt += dense_part(txt, img, lse_t, mu_t, lse_i, mu_i)
        sparse_part(d_txt, img_np, rt, ct, qt, ri, ci, qi)

    return BlockwiseResult(loss, gap, d_scale, rows, d_img, d_txt, lse_i_n, lse_t_n)


def sample_rows_for(n: int, world_size: int, per_rank: int, seed: int = 0) -> List[int]:
    """A fixed pseudo-random sample: ``per_rank`` rows of every rank's block, always including its first and last row."""
    rng = np.random.RandomState(seed)
    b = n // world_size
    out: List[int] = []
    for r in range(world_size):
        pick = {0, b - 1}
        while len(pick) < min(per_rank, b):
            pick.add(int(rng.randint(0, b)))
        out.extend(r * b + p for p in sorted(pick))
    return out
