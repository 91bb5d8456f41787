"""CPU baseline port of the reference loss.  TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

A torch-CPU restatement of the reference's per-rank SpatialLoss / ClipLoss step that keeps the reference's
*cost structure*: dense [B_l, N] fp32 logits from two matmuls, the dict + Python double loop with
``.item()`` reads that builds dense soft labels, L1 normalisation, two log-softmaxes, the temperature
regulariser's softmaxes, and autograd backward (reference: src/models/components/losses.py:73-122,
src/open_clip/loss.py:117-153).  ``bench.py`` times it on the GPU box's host cores as the
``cpu_baseline`` / ``--impl reference`` leg because /root/reference itself does not travel to the box.
tests/test_oracle.py checks it against the goldens minted from the real reference modules.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def spatial_rank_step(img_l, txt_l, img_all, txt_all, scale, ids_all, nbr_ids, nbr_alpha, rank,
                      cap=40.0, temp_reg_weight=0.05, alpha_scale=0.5):
    """One rank's fwd+bwd.  img_l/txt_l: [B_l, D] leaf tensors requiring grad; *_all: [N, D]."""
    b_local, n_global = img_l.shape[0], img_all.shape[0]
    s_eff = scale
    if cap is not None:
        s_eff = scale + (torch.clamp(scale, max=cap) - scale).detach()
    z_it = img_l @ txt_all.T
    z_ti = txt_l @ img_all.T
    l_it = (s_eff * z_it).float()
    l_ti = (s_eff * z_ti).float()

    lookup = {int(t.item()): j for j, t in enumerate(ids_all)}
    own = torch.arange(b_local) + b_local * rank
    lab_it = torch.zeros(b_local, n_global)
    lab_it[torch.arange(b_local), own] = 1.0
    lab_ti = lab_it.clone()
    a = (nbr_alpha * alpha_scale).clamp_min(0)
    for i in range(b_local):
        for slot in range(nbr_ids.shape[1]):
            w = a[i, slot].item()
            if w <= 0:
                continue
            j = lookup.get(int(nbr_ids[i, slot].item()))
            if j is not None:
                lab_it[i, j] += w
                lab_ti[i, j] += w
    lab_it = F.normalize(lab_it, p=1, dim=1)
    lab_ti = F.normalize(lab_ti, p=1, dim=1)

    loss = 0.5 * (-(F.log_softmax(l_it, 1) * lab_it).sum(1).mean() - (F.log_softmax(l_ti, 1) * lab_ti).sum(1).mean())
    if temp_reg_weight > 0:
        p_it, p_ti = F.softmax(l_it, 1), F.softmax(l_ti, 1)
        gap = 0.5 * (((p_it * z_it).sum(1).mean() - (lab_it * z_it).sum(1).mean())
                     + ((p_ti * z_ti).sum(1).mean() - (lab_ti * z_ti).sum(1).mean()))
        loss = loss + temp_reg_weight * gap * gap
    loss.backward()
    return loss.detach()


def clip_rank_step(img_l, txt_l, img_all, txt_all, scale, rank):
    b_local = img_l.shape[0]
    l_it = scale * img_l @ txt_all.T
    l_ti = scale * txt_l @ img_all.T
    labels = torch.arange(b_local) + b_local * rank
    loss = 0.5 * (F.cross_entropy(l_it, labels) + F.cross_entropy(l_ti, labels))
    loss.backward()
    return loss.detach()
