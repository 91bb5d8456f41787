"""TEST-ONLY checker backend: a torch-CPU statement of each C-ABI op's contract (include/scl_b200.h).

It lets the host-side logic of spatial_clip_b200/losses.py (rank resolution, collective order, which
column terms carry gradient, packing of the statistics exchange) run under gloo on CPU, where the
CUDA kernels cannot.  It is injected explicitly by tests through ``losses._set_ops_for_testing``;
nothing in the package imports it and the package has no CPU path of its own.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

from oracle.contrastive_oracle import soft_label_triples

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


class EmulatedOps:
    name = "emulated"

    def __init__(self, round_bf16=True):
        self.round_bf16 = round_bf16
        self.calls = []

    def cast_bf16(self, x, normalize=False):
        self.calls.append("cast_bf16")
        xf = x.float()
        if normalize:
            xf = torch.nn.functional.normalize(xf, dim=-1)
        return xf.to(torch.bfloat16) if self.round_bf16 else xf.clone()

    def check_positives(self, col, q, n_global, rank):
        self.calls.append("check_positives")
        b_local = col.shape[0]
        bad = (col < -1) | (col >= n_global)
        flag = 0
        if bool(bad.any()):
            flag |= 1
        col_out = torch.where(bad, torch.full_like(col, -1), col)
        own = torch.arange(b_local, dtype=col.dtype) + rank * b_local
        if not torch.equal(col_out[:, 0], own):
            flag |= 2
        unused = col_out < 0
        if bool(((col == -1) & (q != 0)).any()):
            flag |= 4
        q_out = torch.where(unused, torch.zeros_like(q), q)
        return col_out, q_out, torch.tensor([flag], dtype=torch.int32)

    def prep_scalars(self, logit_scale, cap):
        s = float(logit_scale[0])
        s_eff = min(s, cap) if cap is not None and cap > 0 else s
        return torch.tensor([s_eff, s_eff * LOG2E, s], dtype=torch.float32)

    def build_positives(self, all_ids, nbr_ids, nbr_alpha, b_local, k, alpha_scale, rank, like):
        self.calls.append("build_positives")
        kp1 = k + 1
        col = torch.full((b_local, kp1), -1, dtype=torch.int32)
        w = torch.zeros(b_local, kp1)
        q = torch.zeros(b_local, kp1)
        if k == 0:
            col[:, 0] = torch.arange(b_local, dtype=torch.int32) + rank * b_local
            w[:, 0] = 1.0
            q[:, 0] = 1.0
            return col, w, q
        rows, sums = soft_label_triples(all_ids.numpy(), nbr_ids.numpy(), nbr_alpha.numpy(), alpha_scale, rank)
        for i, lst in enumerate(rows):
            inv = np.float32(1.0) / max(np.float32(sums[i]), np.float32(1e-12))  # the kernel multiplies by 1/sum in fp32
            for t, (c, ww) in enumerate(lst):
                col[i, t] = c
                w[i, t] = float(ww)
                q[i, t] = float(np.float32(ww) * inv)
        return col, w, q

    def fwd_rowstats(self, x_rows, y_cols, scalars, debug_z=False):
        self.calls.append("fwd_rowstats")
        z = x_rows.double() @ y_cols.double().t()
        y2 = z * float(scalars[1])
        m = y2.max(dim=1).values
        e = torch.exp2(y2 - m[:, None])
        part = torch.stack([m, e.sum(1), (e * z).sum(1), (e * z * z).sum(1)], dim=1).float()
        plan = SimpleNamespace(n_slots=1, m_pad=x_rows.shape[0])
        return (part, plan, z.float()) if debug_z else (part, plan)

    def row_finalize(self, partial, plan, x_rows, y_all, pos_col, pos_q):
        self.calls.append("row_finalize")
        m, s0, s1, s2 = partial.double().unbind(1)
        mu = s1 / s0
        var = s2 / s0 - mu * mu
        zq = torch.zeros_like(mu)
        xd, yd = x_rows.double(), y_all.double()
        for t in range(pos_col.shape[1]):
            c = pos_col[:, t].long()
            ok = c >= 0
            zz = (xd * yd[c.clamp(min=0)]).sum(1)
            zq += torch.where(ok, pos_q[:, t].double() * zz, torch.zeros_like(zz))
        return torch.stack([m + torch.log2(s0), mu, var, zq], dim=1).float()

    def reduce_rows(self, a, b, scalars):
        s_eff = float(scalars[0])
        ad, bd = a.double(), b.double()
        return torch.tensor([
            (ad[:, 0] * LN2 - s_eff * ad[:, 3]).sum(), (bd[:, 0] * LN2 - s_eff * bd[:, 3]).sum(),
            (ad[:, 1] - ad[:, 3]).sum(), (bd[:, 1] - bd[:, 3]).sum(), ad[:, 2].sum(), bd[:, 2].sum()],
            dtype=torch.float32)

    def loss_scalars(self, sums6, scalars, c, w):
        s = sums6.double()
        gsum = c * (s[2] + s[3])
        gap = gsum if w > 0 else 0.0 * gsum
        loss = c * (s[0] + s[1]) + w * gap * gap
        ds = gsum + 2 * w * gap * c * (s[4] + s[5])
        return torch.stack([loss, gap, ds, 2 * w * gap]).float()

    # ---- composite phases, assembled from the per-op contracts above
    def prepare(self, image, text, logit_scale, cap, split=False):
        # split (the fp32-accurate mode): the checker keeps the unrounded values in one [rows, D] tensor -- the
        # hi/lo layout is a device detail; what the host logic must get right is which tensor goes where
        self.calls.append("prepare_split" if split else "prepare")
        rb, self.round_bf16 = self.round_bf16, self.round_bf16 and not split
        try:
            img = self.cast_bf16(image)
            txt = self.cast_bf16(text)
        finally:
            self.round_bf16 = rb
        return img, txt, img, txt, self.prep_scalars(logit_scale, cap)

    def forward_all(self, img_l, txt_l, img_all, txt_all, scalars, ids, b_local, rank, k, alpha_scale, c, w,
                    finalize_scalars, want_ranks=False, waits=None, positives=None):
        # waits (overlapped exchanges): each phase may only touch the operand whose wait has been called -- the
        # checker poisons nothing, but it calls the waits exactly where the CUDA backend does
        wait_ids, wait_txt, wait_img = waits if waits is not None else (None, None, None)
        if waits is not None:
            self.calls.append("forward_all_phased")
        if wait_ids is not None:
            wait_ids()
        if positives is not None:  # resolved on the data side: the builder is skipped
            self.calls.append("forward_all_precomputed")
            it = tuple(positives[:3])
            ti = tuple(positives[3:6]) if len(positives) >= 6 else it
        elif ids is None:
            it = ti = self.build_positives(None, None, None, b_local, 0, 1.0, rank, img_l)
        else:
            img_ids_all, txt_ids_all, nbr, alpha, same = ids
            it = self.build_positives(txt_ids_all, nbr, alpha, b_local, k, alpha_scale, rank, img_l)
            ti = it if same else self.build_positives(img_ids_all, nbr, alpha, b_local, k, alpha_scale, rank, img_l)
        if wait_txt is not None:
            wait_txt()
        part_i, plan_i = self.fwd_rowstats(img_l, txt_all, scalars)
        stats_i = self.row_finalize(part_i, plan_i, img_l, txt_all, it[0], it[2])
        if wait_img is not None:
            wait_img()
        part_t, plan_t = self.fwd_rowstats(txt_l, img_all, scalars)
        stats_t = self.row_finalize(part_t, plan_t, txt_l, img_all, ti[0], ti[2])
        sums6 = self.reduce_rows(stats_i, stats_t, scalars)
        out4 = self.loss_scalars(sums6, scalars, c, w) if finalize_scalars else None
        ranks = None
        if want_ranks:  # contract of scl_fwd_rowstats_ranks: local columns scoring above the row's own pair
            lo = rank * b_local
            z = img_l.double() @ txt_all[lo:lo + b_local].double().t()
            above = z > z.diagonal()[:, None]
            above.fill_diagonal_(False)
            ranks = above.sum(1).to(torch.int32)
        return it, ti, stats_i, stats_t, sums6, out4, ranks

    def backward_dir(self, *args, split=False):
        return self.bwd_rows(*args[:-1], opp_q_local=args[-1])

    def exchange_records(self, parts, world, gather_fn):
        self.calls.append("exchange_records")
        flat = torch.cat([p.reshape(-1).view(torch.float32) for p in parts])
        gathered = gather_fn(flat.reshape(1, -1))
        outs, o = [], 0
        for p in parts:
            n = p.numel()
            outs.append(gathered[:, o:o + n].contiguous().view(p.dtype).reshape((world * p.shape[0],) + tuple(p.shape[1:])))
            o += n
        return outs

    def bwd_rows(self, x_rows, y_all, row_stats, col_stats, pos_col, pos_q, opp_col_all, opp_q_all,
                 b_local, rank, gaps, scalars, grad_out, c, w, mult, col_mode, out_dtype, opp_q_local=None):
        self.calls.append("bwd_rows")
        m, d = x_rows.shape
        n = y_all.shape[0]
        s_eff, s2 = float(scalars[0]), float(scalars[1])
        g = float(grad_out[0]) * mult * c
        xd, yd = x_rows.double(), y_all.double()
        z = xd @ yd.t()
        rs, cs = row_stats.double(), col_stats.double()
        gaps = gaps.double()
        k2r = 2 * w * gaps[rank]
        u = g * (s_eff + k2r * (1 - s_eff * rs[:, 1]))
        v = g * k2r * s_eff * torch.ones(m, dtype=torch.float64)
        owner = torch.arange(n) // b_local
        on = torch.ones(n, dtype=torch.float64) if col_mode == 2 else (
            (owner == rank).double() if col_mode == 1 else torch.zeros(n, dtype=torch.float64))
        k2c = 2 * w * gaps[owner]
        uc = on * g * (s_eff + k2c * (1 - s_eff * cs[:, 1]))
        vc = on * g * k2c * s_eff
        p = torch.exp2(z * s2 - rs[:, 0:1])
        pc = torch.exp2(z * s2 - cs[None, :, 0])
        G = p * (u[:, None] + v[:, None] * z) + pc * (uc[None, :] + vc[None, :] * z)
        dx = G @ yd
        # sparse soft-target terms
        coef = g * (s_eff + k2r)
        for t in range(pos_col.shape[1]):
            cc = pos_col[:, t].long()
            ok = (cc >= 0).double()
            dx -= (coef * ok * pos_q[:, t].double())[:, None] * yd[cc.clamp(min=0)]
        if col_mode != 0:
            lo, hi = rank * b_local, (rank + 1) * b_local
            for t in range(opp_col_all.shape[1]):
                cc = opp_col_all[:, t].long()
                sel = (cc >= lo) & (cc < hi)
                if col_mode == 1:
                    sel &= owner == rank
                idx = torch.nonzero(sel).squeeze(1)
                if idx.numel() == 0:
                    continue
                coefj = g * (s_eff + k2c[idx]) * opp_q_all[idx, t].double()
                dx.index_add_(0, cc[idx] - lo, -(coefj[:, None] * yd[idx]))
        return dx.to(out_dtype)


def _split2(x):
    h = x.float().bfloat16().float()
    return h, (x.float() - h).bfloat16().float()


class SplitArithmeticOps(EmulatedOps):
    """TEST-ONLY: the ARITHMETIC of the fp32-accurate device mode, restated on the CPU in fp32 -- operands as bf16
    hi/lo pairs, z = xh.yh + xh.yl + xl.yh, dL/dz split into two bf16 tiles, dX = G1.Yh + G1.Yl + G2.Yh -- so the
    error budget of that scheme can be pinned against the reference goldens without a GPU
    (tests/test_fp32_mode_numerics.py)."""
    def cast_bf16(self, x, normalize=False):
        return x.float().clone()
    def _z(self, x, y):
        xh, xl = _split2(x); yh, yl = _split2(y)
        return (xh @ yh.t() + xh @ yl.t() + xl @ yh.t())   # fp32
    def fwd_rowstats(self, x_rows, y_cols, scalars, debug_z=False):
        z = self._z(x_rows, y_cols)
        y2 = z * float(scalars[1])
        m = y2.max(dim=1).values
        e = torch.exp2(y2 - m[:, None])
        part = torch.stack([m, e.sum(1), (e * z).sum(1), (e * z * z).sum(1)], dim=1).float()
        return part, SimpleNamespace(n_slots=1, m_pad=x_rows.shape[0])
    def row_finalize(self, partial, plan, x_rows, y_all, pos_col, pos_q):
        m, s0, s1, s2 = partial.float().unbind(1)
        mu = s1 / s0; var = s2 / s0 - mu * mu
        zq = torch.zeros_like(mu)
        xh, xl = _split2(x_rows); yh, yl = _split2(y_all)
        for t in range(pos_col.shape[1]):
            c = pos_col[:, t].long(); ok = c >= 0; cc = c.clamp(min=0)
            zz = (xh * yh[cc]).sum(1) + (xh * yl[cc]).sum(1) + (xl * yh[cc]).sum(1)
            zq += torch.where(ok, pos_q[:, t] * zz, torch.zeros_like(zz))
        return torch.stack([m + torch.log2(s0), mu, var, zq], dim=1).float()
    def bwd_rows(self, x_rows, y_all, row_stats, col_stats, pos_col, pos_q, opp_col_all, opp_q_all,
                 b_local, rank, gaps, scalars, grad_out, c, w, mult, col_mode, out_dtype, opp_q_local=None):
        m, d = x_rows.shape; n = y_all.shape[0]
        s_eff, s2 = float(scalars[0]), float(scalars[1])
        g = float(grad_out[0]) * mult * c
        z = self._z(x_rows, y_all)
        rs, cs = row_stats.float(), col_stats.float(); gaps = gaps.float()
        k2r = 2 * w * gaps[rank]
        u = g * (s_eff + k2r * (1 - s_eff * rs[:, 1])); v = g * k2r * s_eff * torch.ones(m)
        owner = torch.arange(n) // b_local
        on = torch.ones(n) if col_mode == 2 else ((owner == rank).float() if col_mode == 1 else torch.zeros(n))
        k2c = 2 * w * gaps[owner]
        uc = on * g * (s_eff + k2c * (1 - s_eff * cs[:, 1])); vc = on * g * k2c * s_eff
        p = torch.exp2(z * s2 - rs[:, 0:1]); pc = torch.exp2(z * s2 - cs[None, :, 0])
        G = p * (u[:, None] + v[:, None] * z) + pc * (uc[None, :] + vc[None, :] * z)
        # own-column term subtracted before rounding (as the kernel does)
        qd = pos_q[:, 0].float() + (opp_q_local[:, 0].float() if col_mode != 0 else 0)
        tdiag = g * (s_eff + k2r) * qd
        idx = torch.arange(m); G[idx, rank * b_local + idx] -= tdiag
        G1, G2 = _split2(G); yh, yl = _split2(y_all)
        dx = G1 @ yh + G1 @ yl + G2 @ yh
        yd = yh + yl
        coef = g * (s_eff + k2r)
        for t in range(1, pos_col.shape[1]):
            cc = pos_col[:, t].long(); ok = (cc >= 0).float()
            dx -= (coef * ok * pos_q[:, t])[:, None] * yd[cc.clamp(min=0)]
        if col_mode != 0:
            lo, hi = rank * b_local, (rank + 1) * b_local
            for t in range(1, opp_col_all.shape[1]):
                cc = opp_col_all[:, t].long(); sel = (cc >= lo) & (cc < hi)
                if col_mode == 1: sel &= owner == rank
                ix = torch.nonzero(sel).squeeze(1)
                if ix.numel() == 0: continue
                coefj = g * (s_eff + k2c[ix]) * opp_q_all[ix, t]
                dx.index_add_(0, cc[ix] - lo, -(coefj[:, None] * yd[ix]))
        return dx.to(out_dtype)
