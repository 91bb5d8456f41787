"""Drop-in loss modules for Spatial-Clip's Lightning/Hydra ``train.py`` backed by sm_100a kernels.

Host-side mirror of the reference's loss-module API (same class names, constructor keywords,
``forward`` parameter NAMES -- the LightningModule dispatches by ``inspect.signature``,
/root/reference/src/models/spatial_clip_module.py:44,55-64 -- and the same
``{"contrastive_loss": scalar}`` result):

  * ``SpatialLoss``  <- /root/reference/src/models/components/losses.py:11-124
  * ``ClipLoss``     <- /root/reference/src/models/components/losses.py:126-141 over
                        /root/reference/src/open_clip/loss.py:68-155
  * ``GlobalMappingMultiPositiveClipLoss`` (legacy positional order, bare-tensor return)
                     <- /root/reference/src/open_clip_train/spatial_loss.py:10-155
  * feature / id exchange with ``local_loss`` / ``gather_with_grad`` semantics
                     <- /root/reference/src/open_clip/loss.py:21-65

What differs from the reference, by design:
  * the [B_l, N] logits, soft labels and softmaxes never exist in HBM; the work is done by the fused
    tcgen05 kernels reached through the C ABI (``include/scl_b200.h``);
  * ``gather_with_grad``'s backward reduce-scatter of [N, D] gradients is replaced by an all-gather of
    per-row statistics (a few floats per row): each rank recomputes its row block and its column
    block and forms ``d(sum_r loss_r)/d(local features)`` itself -- the same tensor the reference's
    reduce-scatter delivers;
  * rank / world size are resolved lazily at the first ``forward`` when not given explicitly (the
    reference freezes them in ``__init__``, before Lightning creates the process group; pass
    ``world_size=1`` to reproduce that quirk);
  * a scalar ``logit_bias`` cancels in both softmaxes and is ignored (zero gradient);
  * arithmetic: bf16 operands, fp32 accumulation / logits / statistics on chip ("float32_logits" is
    always on); the extra constructor keyword ``precision="fp32"`` switches to bf16 hi/lo operand pairs
    (three tensor-core products per GEMM) for fp32-grade results.

The modules hold no parameters and no buffers (checkpoints stay interchangeable).
No CPU fallback exists: without libscl_b200.so, or with CPU tensors, ``forward`` raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

__all__ = ["SpatialLoss", "ClipLoss", "GlobalMappingMultiPositiveClipLoss", "SpatialLossFromColumns"]

_OPS = None


def _ops():
    """The compute backend: CUDA kernels behind the C ABI.  (tests may inject a checker backend.)"""
    global _OPS
    if _OPS is None:
        from ._cuda import CudaOps

        _OPS = CudaOps()
    return _OPS


def _set_ops_for_testing(ops):
    global _OPS
    prev = _OPS
    _OPS = ops
    return prev


# ------------------------------------------------------------------------------------------------
# distributed plumbing (torch.distributed: NCCL on the GPU box, gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------
def _all_gather_rows(t: torch.Tensor, world: int, group) -> torch.Tensor:
    """Rank-major concatenation along dim 0 (== torch.cat(all_gather(...)), loss.py:51-57).

    Moved as raw bytes so that bf16 / int32-bit-pattern payloads work on every backend (gloo has no
    16-bit integer or bf16 all-gather)."""
    t = t.contiguous()
    carrier = t.view(torch.uint8).reshape(-1)
    out = torch.empty(world * carrier.numel(), dtype=torch.uint8, device=t.device)
    try:
        dist.all_gather_into_tensor(out, carrier, group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(carrier) for _ in range(world)]
        dist.all_gather(parts, carrier, group=group)
        out = torch.cat(parts, dim=0)
    return out.view(t.dtype).reshape((world * t.shape[0],) + tuple(t.shape[1:]))


def _all_gather_rows_async(t: torch.Tensor, world: int, group):
    """Non-blocking form of ``_all_gather_rows``: returns (gathered tensor, wait).  ``wait()`` makes the current
    stream (NCCL) / the caller (gloo) wait for THIS exchange only, so later exchanges keep running underneath the
    kernels that need only this one."""
    t = t.contiguous()
    carrier = t.view(torch.uint8).reshape(-1)
    out = torch.empty(world * carrier.numel(), dtype=torch.uint8, device=t.device)
    work = dist.all_gather_into_tensor(out, carrier, group=group, async_op=True)
    return out.view(t.dtype).reshape((world * t.shape[0],) + tuple(t.shape[1:])), work.wait


@dataclass
class _Cfg:
    kind: str  # "spatial" | "clip"
    rank: int
    world: int
    local_loss: bool
    gather_with_grad: bool
    cap: Optional[float]
    temp_reg_weight: float
    alpha_scale: float
    group: object = None
    split: bool = False  # fp32-accurate mode: operands carried as bf16 hi/lo pairs
    want_ranks: bool = False  # also count every local image row's in-batch retrieval rank


def _col_mode(cfg: _Cfg) -> int:
    """Which column-direction terms reach the local features (loss.py:49-61).

    2: every column (world 1, differentiable gather, or the global ClipLoss matrix);
    1: only this rank's own columns (non-differentiable gather with the local slab re-spliced);
    0: none (non-differentiable gather, local_loss)."""
    if cfg.world == 1 or cfg.gather_with_grad:
        return 2
    if cfg.kind == "clip" and not cfg.local_loss:
        return 2
    return 0 if cfg.local_loss else 1


class _ContrastiveLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, image_tile_ids, text_tile_ids, neighbor_tile_ids,
                neighbor_alphas, cfg: _Cfg, positives=None):
        ops = _ops()
        b_local, d = image_features.shape
        world, rank = cfg.world, cfg.rank
        n = world * b_local
        dev = image_features.device

        scale = logit_scale.detach().reshape(-1)[:1].to(device=dev, dtype=torch.float32).contiguous()
        if hasattr(ops, "check_shapes"):
            ops.check_shapes(b_local, n, d, cfg.split, any(ctx.needs_input_grad[:2]))

        # ---- cap, bf16 copies (gather_features operands, loss.py:21-65).  Single rank + backward wanted: the
        # transposed copies the backward GEMMs need come out of the same pass (dImage needs text^T and vice versa)
        ld_t = (n + 7) // 8 * 8
        # (grad mode is off inside Function.forward; needs_input_grad carries the intent)
        no_t = getattr(ops, "mn_major", False) and not cfg.split  # developer knob: no transposed copies at all
        fuse_t = world == 1 and not no_t
        # *_l: the local rows as row operands, *_c: the same rows as column operands (identical tensors unless
        # cfg.split, where rows are laid out (h|h|l) and columns (h|l|h), 3 D wide)
        img_l, txt_l, img_c, txt_c, img_t, txt_t, scalars = ops.prepare(
            image_features.detach().contiguous(), text_features.detach().contiguous(), scale, cfg.cap,
            fuse_t and ctx.needs_input_grad[1], fuse_t and ctx.needs_input_grad[0], ld_t, split=cfg.split)
        # Developer knob (SCL_OVERLAP_GATHER=1, W > 1): the exchanges are issued back to back without waiting -- gene
        # features, tile ids, image features -- and the forward runs in three phases, each waiting only for the
        # operand it reads, so the image-feature exchange runs underneath the image-rows pass (which needs only the
        # gathered gene features).  Same collectives in the same order on every rank; off until run on a B200.
        overlap = world > 1 and getattr(ops, "overlap_gather", False)
        waits = None
        if overlap:
            txt_all, wait_txt = _all_gather_rows_async(txt_c, world, cfg.group)
        elif world > 1:
            img_all = _all_gather_rows(img_c, world, cfg.group)
            txt_all = _all_gather_rows(txt_c, world, cfg.group)
        else:
            img_all, txt_all = img_c, txt_c

        # ---- tile ids (losses.py:63-68); plain CLIP has only the diagonal
        ids = None
        k = 0
        if positives is not None:
            # soft targets resolved on the data side (positives.py): (columns int32, weights, probs) [B_l, K+1] of the
            # image rows, optionally followed by the same three for the text rows.  No id exchange, no hash build.
            k = positives[0].shape[1] - 1
        elif cfg.kind == "spatial":
            k = neighbor_tile_ids.shape[1]
            same_ids = (image_tile_ids.data_ptr() == text_tile_ids.data_ptr()
                        and image_tile_ids.shape == text_tile_ids.shape)
            img_ids = image_tile_ids.to(torch.int64).contiguous()
            txt_ids = img_ids if same_ids else text_tile_ids.to(torch.int64).contiguous()
            if overlap:
                img_ids_all, wait_ids = _all_gather_rows_async(img_ids, world, cfg.group)
                if same_ids:
                    txt_ids_all = img_ids_all
                else:
                    txt_ids_all, wait_ids2 = _all_gather_rows_async(txt_ids, world, cfg.group)
                    wait_ids = (lambda a=wait_ids, b=wait_ids2: (a(), b()))
            elif world > 1:
                img_ids_all = _all_gather_rows(img_ids, world, cfg.group)
                txt_ids_all = img_ids_all if same_ids else _all_gather_rows(txt_ids, world, cfg.group)
            else:
                img_ids_all, txt_ids_all = img_ids, txt_ids
            ids = (img_ids_all, txt_ids_all, neighbor_tile_ids.to(torch.int64).contiguous(),
                   neighbor_alphas.to(torch.float32).contiguous(), same_ids)
        if overlap:
            img_all, wait_img = _all_gather_rows_async(img_c, world, cfg.group)
            waits = (wait_ids if ids is not None else None, wait_txt, wait_img)

        # ---- soft targets (losses.py:91-111), both fused similarity + online-LSE passes (losses.py:78-89,
        # 113-121), row reductions and the loss scalars: one host call
        global_clip = cfg.kind == "clip" and world > 1 and not cfg.local_loss
        c = 0.5 / (n if global_clip else b_local)
        (col_it, w_it, q_it), (col_ti, w_ti, q_ti), stats_i, stats_t, sums6, out4, ranks = ops.forward_all(
            img_l, txt_l, img_all, txt_all, scalars, ids, b_local, rank, k, cfg.alpha_scale, c, cfg.temp_reg_weight,
            finalize_scalars=not global_clip, want_ranks=cfg.want_ranks, **({"waits": waits} if waits else {}),
            **({"positives": positives} if positives is not None else {}))
        if ranks is None:
            ranks = torch.empty((0,), dtype=torch.int32, device=dev)
        if global_clip:  # every rank evaluates the full N x N loss (loss.py:120-121)
            # all-gather + fixed-order sum (bitwise identical on every rank, unlike an all-reduce tree)
            sums6 = _all_gather_rows(sums6.reshape(1, 6), world, cfg.group).sum(dim=0)
            out4 = ops.loss_scalars(sums6, scalars, c, cfg.temp_reg_weight)

        ctx.cfg = cfg
        ctx.no_t = no_t
        ctx.d = d
        ctx.c = c
        ctx.in_dtypes = (image_features.dtype, text_features.dtype, logit_scale.dtype)
        ctx.scale_shape = logit_scale.shape
        ctx.scale_device = logit_scale.device
        ctx.transposed = (img_t, txt_t)  # None unless produced above
        ctx.save_for_backward(img_l, txt_l, img_all, txt_all, scalars, stats_i, stats_t, out4, col_it, q_it, col_ti,
                              q_ti)
        if positives is not None:  # the caller already holds them: do not hand inputs back as outputs
            col_it, w_it, q_it = (t.new_empty((0,)) for t in (col_it, w_it, q_it))
        ctx.mark_non_differentiable(col_it, w_it, q_it, ranks)
        return out4[0].clone(), col_it, w_it, q_it, ranks

    @staticmethod
    def backward(ctx, grad_loss, *_unused):
        ops = _ops()
        cfg: _Cfg = ctx.cfg
        (img_l, txt_l, img_all, txt_all, scalars, stats_i, stats_t, out4, col_it, q_it, col_ti, q_ti) = ctx.saved_tensors
        world, rank = cfg.world, cfg.rank
        b_local, d = img_l.shape[0], ctx.d
        n = world * b_local
        go = grad_loss.detach().reshape(1).to(torch.float32).contiguous()

        # ---- exchange per-row statistics instead of reduce-scattering [N, D] gradients
        if world > 1:
            gap_rows = out4[1:2].reshape(1, 1)
            (stats_i_all, stats_t_all, col_it_all, q_it_all, col_ti_all, q_ti_all, gaps) = ops.exchange_records(
                [stats_i, stats_t, col_it, q_it, col_ti, q_ti, gap_rows], world,
                lambda t: _all_gather_rows(t, world, cfg.group))
            gaps = gaps.reshape(world)
        else:
            stats_i_all, stats_t_all = stats_i, stats_t
            col_it_all, q_it_all, col_ti_all, q_ti_all = col_it, q_it, col_ti, q_ti
            gaps = out4[1:2].contiguous()

        mode = _col_mode(cfg)
        global_clip = cfg.kind == "clip" and world > 1 and not cfg.local_loss
        mult = float(world) if (global_clip and cfg.gather_with_grad) else 1.0
        w = cfg.temp_reg_weight
        ld_t = (n + 7) // 8 * 8
        need_i, need_t, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_img = d_txt = d_scale = None
        img_all_t, txt_all_t = ctx.transposed

        def grad_image():
            t = txt_all_t
            if t is None and not ctx.no_t:
                t = ops.transpose_split(txt_all, d, ld_t) if cfg.split else \
                    ops.cast_bf16(txt_all, want_rows=False, want_t=True, ld_t=ld_t)[1]
            return ops.backward_dir(img_l, txt_all, t, stats_i, stats_t_all, col_it, q_it, col_ti_all, q_ti_all,
                                    b_local, rank, gaps, scalars, go, ctx.c, w, mult, mode, ctx.in_dtypes[0], q_ti,
                                    split=cfg.split)

        def grad_text():
            t = img_all_t
            if t is None and not ctx.no_t:
                t = ops.transpose_split(img_all, d, ld_t) if cfg.split else \
                    ops.cast_bf16(img_all, want_rows=False, want_t=True, ld_t=ld_t)[1]
            return ops.backward_dir(txt_l, img_all, t, stats_t, stats_i_all, col_ti, q_ti, col_it_all, q_it_all,
                                    b_local, rank, gaps, scalars, go, ctx.c, w, mult, mode, ctx.in_dtypes[1], q_it,
                                    split=cfg.split)

        side = ops.side_stream(img_l.device) if (need_i and need_t and getattr(ops, "two_streams", False)) else None
        if side is not None:
            # Developer knob (SCL_BWD_STREAMS=1): the two directions are independent, so the gene-side chain
            # (transposed copy, coefficients, tensor-core pass, finish) runs on a second stream and its short
            # kernels fill the wave tails of the image-side pass.  Both streams start behind everything issued so
            # far and the caller's stream continues behind both.
            cur = torch.cuda.current_stream(img_l.device)
            fork = torch.cuda.Event()
            fork.record(cur)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                d_txt = grad_text()
                join = torch.cuda.Event()
                join.record(side)
            d_img = grad_image()
            cur.wait_event(join)
            d_txt.record_stream(cur)
        else:
            if need_i:
                d_img = grad_image()
            if need_t:
                d_txt = grad_text()
        if need_s:
            # straight-through cap: d s_eff / d s == 1 even when clipped (losses.py:73-76)
            d_scale = (go * out4[2]).to(device=ctx.scale_device, dtype=ctx.in_dtypes[2]).reshape(ctx.scale_shape)
        return d_img, d_txt, d_scale, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------------
class _LossBase(nn.Module):
    def __init__(self, local_loss, gather_with_grad, rank, world_size, use_horovod, precision="bf16",
                 track_retrieval_ranks=False):
        super().__init__()
        if use_horovod:
            raise NotImplementedError("horovod exchange is out of scope; use torch.distributed (NCCL)")
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        # "bf16": operands rounded to bf16 (loss rel 1e-3, grads 3e-2 of max vs the fp32 reference).
        # "fp32": operands carried as bf16 hi/lo pairs, three tensor-core products per GEMM, dL/dz as two bf16 tiles
        #         (loss rel 1e-5, grads 1e-4 of max vs the fp32 reference; ~3x the tensor work; CTA-pair kernels)
        self.precision = precision
        # In-pass retrieval metrics: after each forward, ``last_retrieval_ranks`` holds, per local image
        # row, how many gene profiles of the local batch score above the matching one -- the quantity the reference's
        # LightningModule recovers with a [B_l, B_l] logits matmul + topk (spatial_clip_module.py:68,
        # metrics.py:22-36); Recall@k = (ranks < k).float().mean(), see spatial_clip_b200/metrics.py
        self.track_retrieval_ranks = bool(track_retrieval_ranks)
        self.last_retrieval_ranks = None
        self.local_loss = bool(local_loss)
        self.gather_with_grad = bool(gather_with_grad)
        self.use_horovod = False
        self._rank_arg = rank
        self._world_arg = world_size
        self.process_group = None

    # rank / world resolution: explicit ctor values win; otherwise ask torch.distributed lazily
    @property
    def world_size(self) -> int:
        if self._world_arg is not None:
            return int(self._world_arg)
        return dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        if self._rank_arg is not None:
            return int(self._rank_arg)
        return dist.get_rank(self.process_group) if dist.is_available() and dist.is_initialized() else 0

    @staticmethod
    def _check_features(image_features, text_features, logit_bias):
        if image_features.dim() != 2 or image_features.shape != text_features.shape:
            raise ValueError(f"image_features {tuple(image_features.shape)} and text_features "
                             f"{tuple(text_features.shape)} must both be [B, D]")
        if logit_bias is not None and torch.as_tensor(logit_bias).numel() != 1:
            raise NotImplementedError("only a scalar logit_bias is supported (it cancels in both softmaxes)")

    _KERNEL_DTYPES = (torch.float32, torch.bfloat16, torch.float16)

    @classmethod
    def _kernel_dtype(cls, t: torch.Tensor) -> torch.Tensor:
        """The kernels read fp32 / bf16 / fp16 features; anything else (e.g. float64, which the reference's torch ops
        would accept) goes through an autograd-tracked cast to fp32, so its gradient comes back in the caller's dtype."""
        return t if t.dtype in cls._KERNEL_DTYPES else t.to(torch.float32)

    @staticmethod
    def _scale_tensor(logit_scale, like):
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=like.device, dtype=torch.float32)
        return logit_scale


class SpatialLoss(_LossBase):
    """Multi-positive spatial-neighbour CLIP loss (reference: losses.py:11-124)."""

    def __init__(self, local_loss: bool = False, gather_with_grad: bool = False, rank: Optional[int] = None,
                 world_size: Optional[int] = None, use_horovod: bool = False,
                 cap_logit_scale: Optional[float] = None, temp_reg_weight: float = 0.0,
                 float32_logits: bool = False, neighbor_alpha_scale: float = 1.0, precision: str = "bf16",
                 track_retrieval_ranks: bool = False):
        super().__init__(local_loss, gather_with_grad, rank, world_size, use_horovod, precision, track_retrieval_ranks)
        self.cap_logit_scale = cap_logit_scale
        self.temp_reg_weight = float(temp_reg_weight or 0.0)
        self.float32_logits = float32_logits  # logits are always fp32 on chip
        self.neighbor_alpha_scale = float(neighbor_alpha_scale)
        self.last_positives = None  # (col, raw weight, normalised weight) ELL lists of the last call

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
                image_tile_ids: torch.Tensor, text_tile_ids: torch.Tensor, neighbor_tile_ids: torch.Tensor,
                neighbor_alphas: torch.Tensor, logit_bias: Optional[torch.Tensor] = None,
                output_dict: bool = True) -> Dict[str, torch.Tensor]:
        self._check_features(image_features, text_features, logit_bias)
        image_features, text_features = self._kernel_dtype(image_features), self._kernel_dtype(text_features)
        b = image_features.shape[0]
        if neighbor_tile_ids.dim() != 2 or neighbor_tile_ids.shape[0] != b or \
                neighbor_alphas.shape != neighbor_tile_ids.shape or image_tile_ids.shape[0] != b or \
                text_tile_ids.shape[0] != b:
            raise ValueError("tile ids must be [B] and neighbour ids / alphas [B, K]")
        cfg = _Cfg("spatial", self.rank, self.world_size, self.local_loss, self.gather_with_grad,
                   self.cap_logit_scale, self.temp_reg_weight, self.neighbor_alpha_scale, self.process_group,
                   self.precision == "fp32", self.track_retrieval_ranks)
        loss, col, w, q, ranks = _ContrastiveLossFn.apply(image_features, text_features,
                                                          self._scale_tensor(logit_scale, image_features),
                                                          image_tile_ids, text_tile_ids, neighbor_tile_ids,
                                                          neighbor_alphas, cfg, None)
        self.last_positives = (col, w, q)
        self.last_retrieval_ranks = ranks if self.track_retrieval_ranks else None
        return {"contrastive_loss": loss}


class SpatialLossFromColumns(SpatialLoss):
    """``SpatialLoss`` fed with soft targets that the data pipeline already resolved to global columns
    (producer: ``spatial_clip_b200.positives.resolve_positive_columns`` at collate time).

    Same constructor and same arithmetic as ``SpatialLoss`` (reference: losses.py:11-124); ``forward`` takes
    ``positive_columns`` int32 / ``positive_probs`` fp32 ``[B_l, K+1]`` (slot 0 = the row's own column) instead of the
    four id / neighbour tensors, so the two id all-gathers (losses.py:63-68), the id -> column map and the label loop
    (losses.py:91-111) leave the step altogether.  ``neighbor_alpha_scale`` is applied by the producer, not here.
    ``positive_*_text`` are the text-row lists when the two id vectors differ (the reference's loader makes them
    equal).  The LightningModule dispatches by parameter name (spatial_clip_module.py:44,58-61), so a collate that
    adds these keys (``positives.collate_positive_columns``) is all the integration needs."""

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
                positive_columns: torch.Tensor, positive_probs: torch.Tensor,
                positive_weights: Optional[torch.Tensor] = None, positive_columns_text: Optional[torch.Tensor] = None,
                positive_probs_text: Optional[torch.Tensor] = None, logit_bias: Optional[torch.Tensor] = None,
                output_dict: bool = True) -> Dict[str, torch.Tensor]:
        self._check_features(image_features, text_features, logit_bias)
        image_features, text_features = self._kernel_dtype(image_features), self._kernel_dtype(text_features)
        b = image_features.shape[0]
        if positive_columns.dim() != 2 or positive_columns.shape[0] != b or \
                positive_probs.shape != positive_columns.shape:
            raise ValueError("positive_columns / positive_probs must both be [B, K+1]")
        if (positive_columns_text is None) != (positive_probs_text is None):
            raise ValueError("positive_columns_text and positive_probs_text go together")
        dev = image_features.device

        def prep(col, q, w):
            col = col.to(device=dev, dtype=torch.int32).contiguous()
            q = q.to(device=dev, dtype=torch.float32).contiguous()
            w = q if w is None else w.to(device=dev, dtype=torch.float32).contiguous()
            return col, w, q

        pos = prep(positive_columns, positive_probs, positive_weights)
        if positive_columns_text is not None:
            if positive_columns_text.shape != positive_columns.shape or \
                    positive_probs_text.shape != positive_columns.shape:
                raise ValueError("text-row lists must have the shape of the image-row lists")
            pos = pos + prep(positive_columns_text, positive_probs_text, None)
        cfg = _Cfg("spatial", self.rank, self.world_size, self.local_loss, self.gather_with_grad,
                   self.cap_logit_scale, self.temp_reg_weight, self.neighbor_alpha_scale, self.process_group,
                   self.precision == "fp32", self.track_retrieval_ranks)
        loss, _, _, _, ranks = _ContrastiveLossFn.apply(image_features, text_features,
                                                        self._scale_tensor(logit_scale, image_features), None, None,
                                                        None, None, cfg, pos)
        self.last_positives = pos[:3]
        self.last_retrieval_ranks = ranks if self.track_retrieval_ranks else None
        return {"contrastive_loss": loss}


class ClipLoss(_LossBase):
    """Symmetric InfoNCE (reference: losses.py:126-141 wrapping open_clip loss.py:68-155)."""

    def __init__(self, local_loss: bool = False, gather_with_grad: bool = False, cache_labels: bool = False,
                 rank: Optional[int] = None, world_size: Optional[int] = None, use_horovod: bool = False,
                 precision: str = "bf16", track_retrieval_ranks: bool = False):
        super().__init__(local_loss, gather_with_grad, rank, world_size, use_horovod, precision, track_retrieval_ranks)
        self.cache_labels = cache_labels  # labels are implicit (the diagonal); nothing to cache

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
                logit_bias: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        self._check_features(image_features, text_features, logit_bias)
        image_features, text_features = self._kernel_dtype(image_features), self._kernel_dtype(text_features)
        cfg = _Cfg("clip", self.rank, self.world_size, self.local_loss, self.gather_with_grad, None, 0.0, 1.0,
                   self.process_group, self.precision == "fp32", self.track_retrieval_ranks)
        loss, _, _, _, ranks = _ContrastiveLossFn.apply(image_features, text_features,
                                                        self._scale_tensor(logit_scale, image_features), None, None,
                                                        None, None, cfg, None)
        self.last_retrieval_ranks = ranks if self.track_retrieval_ranks else None
        return {"contrastive_loss": loss}


class GlobalMappingMultiPositiveClipLoss(SpatialLoss):
    """Legacy twin with the open_clip_train positional order (reference: spatial_loss.py:37-155)."""

    def __init__(self, local_loss: bool = False, gather_with_grad: bool = False, cache_labels: bool = False,
                 rank: Optional[int] = 0, world_size: Optional[int] = 1, use_horovod: bool = False,
                 cap_logit_scale: Optional[float] = None, temp_reg_weight: float = 0.0,
                 float32_logits: bool = False, neighbor_alpha_scale: float = 1.0, precision: str = "bf16",
                 track_retrieval_ranks: bool = False):
        super().__init__(local_loss, gather_with_grad, rank, world_size, use_horovod, cap_logit_scale,
                         temp_reg_weight, float32_logits, neighbor_alpha_scale, precision, track_retrieval_ranks)
        self.cache_labels = cache_labels

    def forward(self, image_features, text_features, image_tile_ids, text_tile_ids, neighbor_tile_ids,
                neighbor_alphas, logit_scale, logit_bias=None, output_dict: bool = False):
        out = super().forward(image_features, text_features, logit_scale, image_tile_ids, text_tile_ids,
                              neighbor_tile_ids, neighbor_alphas, logit_bias)
        return out if output_dict else out["contrastive_loss"]
