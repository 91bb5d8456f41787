"""CPU oracle for the Spatial-Clip contrastive-loss hot path.  TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of the reference's loss arithmetic, written from the
algorithm (not copied), used solely as the checker in ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg.  Nothing in
``spatial_clip_b200/`` imports it; the product path is CUDA-only and fails loudly
without its extension.

Pinning status: the reference repository holds NO golden vector / known-answer
test for its loss.  The oracle is therefore pinned against
outputs of the *reference modules themselves*, imported unmodified in the build
container by ``tests/golden/make_golden.py`` and committed as
``tests/golden/*.npz`` (loss, dI, dT, d logit_scale and the dense soft-label
matrices).  ``tests/test_oracle.py`` checks every function below against those.

Reference lines restated here (all relative to /root/reference):
  * feature / id gather, rank-major concatenation  src/open_clip/loss.py:21-65,
    src/models/components/losses.py:58-71
  * STE cap of the logit scale                      src/models/components/losses.py:73-76
  * local-rows x global-cols similarity blocks       src/models/components/losses.py:78-89
  * soft-label construction (dict lookup, k-order)   src/models/components/losses.py:91-111
  * soft cross-entropy, both directions              src/models/components/losses.py:113-115
  * temperature regulariser                          src/models/components/losses.py:117-122
  * plain CLIP loss and its label offset             src/open_clip/loss.py:91-155
  * legacy twin with identical arithmetic            src/open_clip_train/spatial_loss.py:37-155

Gradients are closed forms (SURVEY.md §8a "closed forms"), i.e. what autograd
produces for ``sum_r loss_r`` when every rank calls ``backward()`` on its own
loss and the feature all-gather is differentiable (reduce-scatter SUM in
``torch.distributed.nn.functional._AllGather.backward``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# integer path: id -> column, soft-label triples
# --------------------------------------------------------------------------------------
def last_index_map(ids: np.ndarray) -> Dict[int, int]:
    """id -> global column; for duplicated ids the LAST position wins.

    Restates the dict comprehension at losses.py:92-93 (later keys overwrite).
    """
    table: Dict[int, int] = {}
    for col, tid in enumerate(ids.tolist()):
        table[int(tid)] = col
    return table


def soft_label_triples(
    all_ids: np.ndarray,
    nbr_ids: np.ndarray,
    nbr_alpha: np.ndarray,
    alpha_scale: float,
    rank: int,
    scale_always: bool = True,
) -> Tuple[List[List[Tuple[int, np.float32]]], np.ndarray]:
    """Un-normalised soft labels of the local rows as per-row (col, weight) lists.

    Follows losses.py:94-108: the row's own column ``rank*B_l + i`` gets 1.0, then
    each neighbour slot in k order adds ``float(alpha)`` (fp32 accumulation) to
    the column its id maps to; alpha <= 0 and unmapped ids are skipped.  The
    alphas are scaled and clamped first (losses.py:100); the legacy twin skips
    the multiply when the scale is exactly 1.0 (spatial_loss.py:110-112), which
    is numerically the same thing.

    Returns (rows, row_sum) where rows[i] is ordered by first touch and
    row_sum[i] is the fp32 L1 norm the reference divides by (losses.py:110).
    """
    b_local, k = nbr_ids.shape
    table = last_index_map(all_ids)
    a = nbr_alpha.astype(np.float32)
    if scale_always or alpha_scale != 1.0:
        a = a * np.float32(alpha_scale)
    a = np.maximum(a, np.float32(0.0)).astype(np.float32)
    rows: List[List[Tuple[int, np.float32]]] = []
    sums = np.zeros(b_local, dtype=np.float32)
    for i in range(b_local):
        acc: Dict[int, np.float32] = {rank * b_local + i: np.float32(1.0)}
        for slot in range(k):
            w = a[i, slot]
            if not (w > 0):
                continue
            col = table.get(int(nbr_ids[i, slot]))
            if col is None:
                continue
            acc[col] = np.float32(acc.get(col, np.float32(0.0)) + w)
        rows.append(list(acc.items()))
        # F.normalize(p=1) reduces the dense fp32 row; the row has <= K+1 non-zeros so
        # any summation order of exactly representable partials differs by <= 1 ulp.
        tot = np.float32(0.0)
        for _, w in acc.items():
            tot = np.float32(tot + w)
        sums[i] = tot
    return rows, sums


def dense_labels(rows, n_global: int) -> np.ndarray:
    out = np.zeros((len(rows), n_global), dtype=np.float32)
    for i, lst in enumerate(rows):
        for col, w in lst:
            out[i, col] = w
    return out


# --------------------------------------------------------------------------------------
# floating-point path
# --------------------------------------------------------------------------------------
def _lse_rows(logits: np.ndarray) -> np.ndarray:
    m = logits.max(axis=1, keepdims=True)
    return (m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True)))[:, 0]


@dataclass
class RankResult:
    loss: float
    gap: float
    d_scale: float  # d loss_r / d logit_scale  (local, before any DDP averaging)
    lse_img: np.ndarray = field(repr=False, default=None)  # LSE of logits_per_image rows
    lse_txt: np.ndarray = field(repr=False, default=None)
    mu_img: np.ndarray = field(repr=False, default=None)  # E_p[z] per image row
    mu_txt: np.ndarray = field(repr=False, default=None)


@dataclass
class OracleResult:
    ranks: List[RankResult]
    d_image: np.ndarray  # [N, D]; rows of rank r = what rank r's image_features.grad holds
    d_text: np.ndarray


def spatial_loss_oracle(
    image: np.ndarray,
    text: np.ndarray,
    logit_scale: float,
    image_ids: np.ndarray,
    text_ids: np.ndarray,
    nbr_ids: np.ndarray,
    nbr_alpha: np.ndarray,
    world_size: int = 1,
    cap_logit_scale: Optional[float] = None,
    temp_reg_weight: float = 0.0,
    neighbor_alpha_scale: float = 1.0,
    local_loss: bool = True,
    gather_with_grad: bool = True,
    dtype=np.float64,
) -> OracleResult:
    """SpatialLoss / GlobalMappingMultiPositiveClipLoss over a W-rank emulation.

    ``image``/``text`` are the GLOBAL [N, D] arrays (rank-major order, i.e. what
    gather_features returns, loss.py:51-52); rank r owns rows r*B_l..(r+1)*B_l.
    Every similarity block is local-rows x global-cols regardless of
    ``local_loss`` (losses.py:78-79); ``local_loss``/``gather_with_grad`` only
    decide which gathered tensors carry gradient (loss.py:49-61).
    """
    n, _ = image.shape
    assert n % world_size == 0
    b = n // world_size
    img = image.astype(dtype)
    txt = text.astype(dtype)
    s_eff = float(logit_scale)
    if cap_logit_scale is not None:
        s_eff = min(s_eff, float(cap_logit_scale))  # forward value of the STE, losses.py:74-76
    w = float(temp_reg_weight)
    c = 0.5 / b

    d_img = np.zeros_like(img)
    d_txt = np.zeros_like(txt)
    out: List[RankResult] = []
    for r in range(world_size):
        sl = slice(r * b, (r + 1) * b)
        z_it = img[sl] @ txt.T  # losses.py:78
        z_ti = txt[sl] @ img.T  # losses.py:79
        l_it = s_eff * z_it
        l_ti = s_eff * z_ti
        rows_it, sum_it = soft_label_triples(text_ids, nbr_ids[sl], nbr_alpha[sl], neighbor_alpha_scale, r)
        rows_ti, sum_ti = soft_label_triples(image_ids, nbr_ids[sl], nbr_alpha[sl], neighbor_alpha_scale, r)
        q_it = dense_labels(rows_it, n).astype(dtype) / np.maximum(sum_it.astype(dtype), 1e-12)[:, None]
        q_ti = dense_labels(rows_ti, n).astype(dtype) / np.maximum(sum_ti.astype(dtype), 1e-12)[:, None]

        lse_i = _lse_rows(l_it)
        lse_t = _lse_rows(l_ti)
        loss = 0.5 * ((lse_i - (q_it * l_it).sum(1)).mean() + (lse_t - (q_ti * l_ti).sum(1)).mean())

        p_it = np.exp(l_it - lse_i[:, None])
        p_ti = np.exp(l_ti - lse_t[:, None])
        mu_i = (p_it * z_it).sum(1)
        mu_t = (p_ti * z_ti).sum(1)
        gap = 0.0
        if w > 0:
            gap = 0.5 * ((mu_i.mean() - (q_it * z_it).sum(1).mean()) + (mu_t.mean() - (q_ti * z_ti).sum(1).mean()))
            loss = loss + w * gap * gap

        # d loss_r / d z  for both blocks
        k2 = 2.0 * w * gap
        g_it = c * (s_eff * (p_it - q_it) + k2 * (p_it * (1.0 + s_eff * (z_it - mu_i[:, None])) - q_it))
        g_ti = c * (s_eff * (p_ti - q_ti) + k2 * (p_ti * (1.0 + s_eff * (z_ti - mu_t[:, None])) - q_ti))
        # d loss_r / d s_eff  (STE: d s_eff / d s == 1 even when clipped)
        var_i = (p_it * z_it * z_it).sum(1) - mu_i * mu_i
        var_t = (p_ti * z_ti * z_ti).sum(1) - mu_t * mu_t
        d_s = c * (((p_it - q_it) * z_it).sum() + ((p_ti - q_ti) * z_ti).sum())
        d_s += k2 * c * (var_i.sum() + var_t.sum())

        # local tensors always carry gradient
        d_img[sl] += g_it @ txt
        d_txt[sl] += g_ti @ img
        # gathered tensors: all rows with gather_with_grad, else only the re-spliced local
        # slab when local_loss is False (loss.py:58-61), else nothing
        if gather_with_grad:
            d_txt += g_it.T @ img[sl]
            d_img += g_ti.T @ txt[sl]
        elif not local_loss and world_size > 1:
            d_txt[sl] += g_it[:, sl].T @ img[sl]
            d_img[sl] += g_ti[:, sl].T @ txt[sl]
        elif world_size == 1:
            # W == 1: "gathered" tensors ARE the local tensors (losses.py:70-71)
            d_txt += g_it.T @ img[sl]
            d_img += g_ti.T @ txt[sl]

        out.append(RankResult(float(loss), float(gap), float(d_s), lse_i, lse_t, mu_i, mu_t))
    return OracleResult(out, d_img, d_txt)


def clip_loss_oracle(
    image: np.ndarray,
    text: np.ndarray,
    logit_scale: float,
    world_size: int = 1,
    local_loss: bool = False,
    gather_with_grad: bool = False,
    dtype=np.float64,
) -> OracleResult:
    """open_clip ClipLoss (loss.py:104-155) over a W-rank emulation.

    ``local_loss=True`` : rank r scores its B_l rows against all N columns, labels
    offset by ``rank*B_l`` (loss.py:95-96, :117-118).
    ``local_loss=False``: every rank scores the full N x N matrix (loss.py:120-121).
    """
    n, _ = image.shape
    b = n // world_size
    img = image.astype(dtype)
    txt = text.astype(dtype)
    s = float(logit_scale)
    d_img = np.zeros_like(img)
    d_txt = np.zeros_like(txt)
    out: List[RankResult] = []
    if world_size == 1 or local_loss:
        ident = [[(r * b + i, np.float32(1.0)) for i in range(b)] for r in range(world_size)]
        for r in range(world_size):
            sl = slice(r * b, (r + 1) * b)
            z_it = img[sl] @ txt.T
            z_ti = txt[sl] @ img.T
            q = np.zeros((b, n), dtype=dtype)
            q[np.arange(b), r * b + np.arange(b)] = 1.0
            lse_i = _lse_rows(s * z_it)
            lse_t = _lse_rows(s * z_ti)
            loss = 0.5 * ((lse_i - s * (q * z_it).sum(1)).mean() + (lse_t - s * (q * z_ti).sum(1)).mean())
            p_it = np.exp(s * z_it - lse_i[:, None])
            p_ti = np.exp(s * z_ti - lse_t[:, None])
            c = 0.5 / b
            g_it = c * s * (p_it - q)
            g_ti = c * s * (p_ti - q)
            d_s = c * (((p_it - q) * z_it).sum() + ((p_ti - q) * z_ti).sum())
            d_img[sl] += g_it @ txt
            d_txt[sl] += g_ti @ img
            if gather_with_grad or world_size == 1:
                d_txt += g_it.T @ img[sl]
                d_img += g_ti.T @ txt[sl]
            out.append(RankResult(float(loss), 0.0, float(d_s), lse_i, lse_t, (p_it * z_it).sum(1), (p_ti * z_ti).sum(1)))
        del ident
    else:
        z = img @ txt.T
        lse_r = _lse_rows(s * z)
        lse_c = _lse_rows(s * z.T)
        diag = np.einsum("ij,ij->i", img, txt)
        loss = 0.5 * ((lse_r - s * diag).mean() + (lse_c - s * diag).mean())
        p_r = np.exp(s * z - lse_r[:, None])
        p_c = np.exp(s * z - lse_c[None, :])
        c = 0.5 / n
        g = c * s * (p_r + p_c - 2.0 * np.eye(n, dtype=dtype))
        d_s = c * (((p_r + p_c - 2.0 * np.eye(n, dtype=dtype)) * z).sum())
        gi_full = g @ txt
        gt_full = g.T @ img
        for r in range(world_size):
            sl = slice(r * b, (r + 1) * b)
            # every rank evaluates the same full loss; with a differentiable gather the
            # reduce-scatter sums W identical copies, otherwise only the re-spliced local
            # slab carries gradient (loss.py:58-61)
            mult = world_size if gather_with_grad else 1
            d_img[sl] = mult * gi_full[sl]
            d_txt[sl] = mult * gt_full[sl]
            out.append(RankResult(float(loss), 0.0, float(d_s), lse_r[sl], lse_c[sl], None, None))
    return OracleResult(out, d_img, d_txt)
