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

import atexit
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

__all__ = ["SpatialLoss", "ClipLoss", "GlobalMappingMultiPositiveClipLoss", "SpatialLossFromColumns",
           "release_cuda_graphs"]

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
def _all_gather_rows_async(t: torch.Tensor, world: int, group):
    """Rank-major concatenation along dim 0 (== torch.cat(all_gather(...)), loss.py:51-57), issued without waiting:
    returns (gathered tensor, wait).  ``wait()`` makes the current stream (NCCL) / the caller (gloo) wait for THIS
    exchange only, so later exchanges keep running underneath the kernels that need only this one.

    Moved as raw bytes so that bf16 / int32-bit-pattern payloads work on every backend (gloo has no 16-bit integer or
    bf16 all-gather)."""
    t = t.contiguous()
    carrier = t.view(torch.uint8).reshape(-1)
    out = torch.empty(world * carrier.numel(), dtype=torch.uint8, device=t.device)
    work = dist.all_gather_into_tensor(out, carrier, group=group, async_op=True)
    return out.view(t.dtype).reshape((world * t.shape[0],) + tuple(t.shape[1:])), work.wait


def _all_gather_rows(t: torch.Tensor, world: int, group) -> torch.Tensor:
    out, wait = _all_gather_rows_async(t, world, group)
    wait()
    return out


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
    use_graphs: bool = True  # replay forward / backward as CUDA graphs after two eager calls (_GraphedStep)


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


_NEUTRAL_LSE = 1.0e30  # column statistics of rows that carry no gradient: exp2(z s - 1e30) == 0


def _forward_impl(ops, cfg: _Cfg, image_features, text_features, scale, image_tile_ids, text_tile_ids,
                  neighbor_tile_ids, neighbor_alphas, positives, need_backward):
    """Everything the forward launches, on the current stream, from plain tensors (no autograd state): used directly
    (eager) and under CUDA-graph capture.  Returns (out4, lists_it, ranks, saved) with saved = the tensors backward
    needs."""
    b_local, d = image_features.shape
    world, rank = cfg.world, cfg.rank
    n = world * b_local
    dev = image_features.device
    mark = getattr(ops, "mark", lambda name, device: None)
    mark("fwd:start", dev)
    # ---- cap + bf16 copies of the local rows (the operands of gather_features, loss.py:21-65): one launch.
    # *_l: the local rows as row operands, *_c: the same rows as column operands (identical tensors unless
    # cfg.split, where rows are laid out (h|h|l) and columns (h|l|h), 3 D wide)
    img_l, txt_l, img_c, txt_c, scalars = ops.prepare(image_features, text_features, scale, cfg.cap, split=cfg.split)
    mark("fwd:prepared", dev)

    # ---- exchanges (W > 1), all issued up front without waiting, in the order their consumers run: tile ids (one
    # packed [2, B_l] record: both id vectors, losses.py:63-68) -> soft targets; gene features -> image-rows pass;
    # image features -> text-rows pass.  Every phase waits only for the operand it reads, so the later exchanges
    # run underneath the earlier phases' kernels.  The sequence of collectives depends on the configuration only.
    ids = None
    k = 0
    waits = None
    wait_ids = None
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
        if world > 1:
            both, wait_ids = _all_gather_rows_async(torch.stack([img_ids, txt_ids]).reshape(1, 2, b_local), world,
                                                    cfg.group)
            ids_all = (both, None)  # [W, 2, B_l]; de-interleaved after the wait
        else:
            ids_all = (img_ids, txt_ids)
        ids = [ids_all[0], ids_all[1], neighbor_tile_ids.to(torch.int64).contiguous(),
               neighbor_alphas.to(torch.float32).contiguous(), same_ids]
    if world > 1:
        txt_all, wait_txt = _all_gather_rows_async(txt_c, world, cfg.group)
        img_all, wait_img = _all_gather_rows_async(img_c, world, cfg.group)
        if wait_ids is not None:
            def wait_ids_and_split(w=wait_ids, rec=ids):
                w()
                rec[0], rec[1] = rec[0][:, 0].reshape(-1), rec[0][:, 1].reshape(-1)  # rank-major [N] each
                if rec[4]:
                    rec[1] = rec[0]
            waits = (wait_ids_and_split, wait_txt, wait_img)
        else:
            waits = (None, wait_txt, wait_img)
    else:
        img_all, txt_all = img_c, txt_c

    # ---- soft targets (losses.py:91-111), both fused similarity + online-LSE passes (losses.py:78-89,
    # 113-121), row reductions and the loss scalars: one host call (one per phase when exchanges are in flight)
    global_clip = cfg.kind == "clip" and world > 1 and not cfg.local_loss
    c = 0.5 / (n if global_clip else b_local)
    (col_it, w_it, q_it), (col_ti, w_ti, q_ti), stats_i, stats_t, sums6, out4, ranks = ops.forward_all(
        img_l, txt_l, img_all, txt_all, scalars, ids, b_local, rank, k, cfg.alpha_scale, c, cfg.temp_reg_weight,
        finalize_scalars=not global_clip, want_ranks=cfg.want_ranks, **({"waits": waits} if waits else {}),
        **({"positives": positives} if positives is not None else {}))
    mark("fwd:passes", dev)
    if ranks is None:
        ranks = torch.empty((0,), dtype=torch.int32, device=dev)
    if global_clip:  # every rank evaluates the full N x N loss (loss.py:120-121)
        # all-gather + fixed-order sum (bitwise identical on every rank, unlike an all-reduce tree)
        sums6 = _all_gather_rows(sums6.reshape(1, 6), world, cfg.group).sum(dim=0)
        out4 = ops.loss_scalars(sums6, scalars, c, cfg.temp_reg_weight)
    saved = (img_l, txt_l, img_all, txt_all, scalars, stats_i, stats_t, out4, col_it, q_it, col_ti, q_ti)
    return out4, (col_it, w_it, q_it), ranks, saved, c


def _backward_impl(ops, cfg: _Cfg, saved, go, c, d, need_i, need_t, out_dtypes):
    """Everything the backward launches (see _forward_impl).  Returns (d_image or None, d_text or None)."""
    (img_l, txt_l, img_all, txt_all, scalars, stats_i, stats_t, out4, col_it, q_it, col_ti, q_ti) = saved
    world, rank = cfg.world, cfg.rank
    b_local = img_l.shape[0]
    n = world * b_local
    mark = getattr(ops, "mark", lambda name, device: None)
    mark("bwd:start", img_l.device)
    mode = _col_mode(cfg)
    global_clip = cfg.kind == "clip" and world > 1 and not cfg.local_loss

    # ---- exchange per-row statistics instead of reduce-scattering [N, D] gradients.  Only when other ranks'
    # rows reach the local features at all: with a non-differentiable gather and local_loss (mode 0) the
    # reference's backward has no collective either, and none is issued here.
    if world > 1 and mode != 0:
        gap_rows = out4[1:2].reshape(1, 1)
        (stats_i_all, stats_t_all, col_it_all, q_it_all, col_ti_all, q_ti_all, gaps) = ops.exchange_records(
            [stats_i, stats_t, col_it, q_it, col_ti, q_ti, gap_rows], world,
            lambda t: _all_gather_rows(t, world, cfg.group))
        gaps = gaps.reshape(world)
    elif world > 1:
        # column statistics that switch the column-direction terms off (bwd_coeffs multiplies them by zero; the
        # neutral LSE keeps the exponentials finite), this rank's gap in its slot, no opposite-direction lists
        neutral = torch.zeros((n, 4), dtype=torch.float32, device=img_l.device)
        neutral[:, 0] = _NEUTRAL_LSE
        stats_i_all = stats_t_all = neutral
        col_it_all = q_it_all = col_ti_all = q_ti_all = None
        gaps = torch.zeros((world,), dtype=torch.float32, device=img_l.device)
        gaps[rank:rank + 1] = out4[1:2]
    else:
        stats_i_all, stats_t_all = stats_i, stats_t
        col_it_all, q_it_all, col_ti_all, q_ti_all = col_it, q_it, col_ti, q_ti
        gaps = out4[1:2].contiguous()
    mark("bwd:exchanged", img_l.device)

    mult = float(world) if (global_clip and cfg.gather_with_grad) else 1.0
    w = cfg.temp_reg_weight
    d_img = d_txt = None
    if need_i:
        d_img = ops.backward_dir(img_l, txt_all, stats_i, stats_t_all, col_it, q_it, col_ti_all, q_ti_all,
                                 b_local, rank, gaps, scalars, go, c, w, mult, mode, out_dtypes[0], q_ti,
                                 split=cfg.split)
        mark("bwd:d_image", img_l.device)
    if need_t:
        d_txt = ops.backward_dir(txt_l, img_all, stats_t, stats_i_all, col_ti, q_ti, col_it_all, q_it_all,
                                 b_local, rank, gaps, scalars, go, c, w, mult, mode, out_dtypes[1], q_it,
                                 split=cfg.split)
        mark("bwd:d_text", img_l.device)
    return d_img, d_txt


# ------------------------------------------------------------------------------------------------
# CUDA graphs: one replay per forward, one per backward
# ------------------------------------------------------------------------------------------------
class _GraphedStep:
    """Forward and backward of ONE configuration (shapes, dtypes, flags, process group) captured as two CUDA graphs.

    A step is ~25 kernel launches, four collectives and a few dozen small tensor operations: ~0.6 ms of host time at
    one rank and ~1.1 ms with the exchanges (measured, bench.py host_enqueue_ms_per_step) -- more than the ~1 ms of
    kernels a rank of an 8-GPU job runs.  The first ``WARMUP`` calls run eagerly (they also initialise the NCCL
    communicator); the next one is captured -- inputs are copied into static buffers, results are cloned out -- and
    every later call is two copies and one ``replay()``.  NCCL collectives are captured with the kernels.

    The saved activations live in the graph's private pool and are overwritten by the next forward, so a backward must
    follow ITS forward (the training loop's order); anything else raises."""

    WARMUP = 2

    def __init__(self):
        self.calls = 0
        self.generation = 0
        self.pool = None
        self.fwd = None  # (graph, static inputs, outputs)
        self.bwd = None
        self.fwd_launches = self.bwd_launches = 0
        self.broken = False

    @staticmethod
    def _copy_in(static, fresh):
        for s_t, f_t in zip(static, fresh):
            if s_t is not None and s_t.data_ptr() != f_t.data_ptr():
                s_t.copy_(f_t, non_blocking=True)


_GRAPHS: Dict[tuple, _GraphedStep] = {}
# Every CAPTURED configuration owns a private memory pool (static inputs, gathered operands, partials).  A job has a
# handful of configurations (train / eval batch, the ragged last batch); beyond this many captures, new ones run
# eagerly.  (Configurations still in their eager warm-up calls do not count.)
_MAX_GRAPHED_CONFIGS = 16


def release_cuda_graphs():
    """Drop every captured step (graphs, static buffers).  Captured NCCL kernels keep their communicator busy, so this
    must run BEFORE ``torch.distributed.destroy_process_group()`` (which otherwise waits for them forever); it is also
    registered with ``atexit``, ahead of torch's own teardown."""
    if not _GRAPHS:
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    for step in _GRAPHS.values():
        # autograd nodes of the last step may still hold the step object: reset the graphs themselves
        step.broken = True
        for captured in (step.fwd, step.bwd):
            if captured is not None:
                try:
                    captured[0].reset()
                except Exception:
                    pass
        step.fwd = step.bwd = step.pool = None
    _GRAPHS.clear()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


atexit.register(release_cuda_graphs)


def _graph_key(cfg: _Cfg, tensors, flags):
    sig = tuple((tuple(t.shape), t.dtype, t.device.index) if t is not None else None for t in tensors)
    return (cfg.kind, cfg.rank, cfg.world, cfg.local_loss, cfg.gather_with_grad, cfg.cap, cfg.temp_reg_weight,
            cfg.alpha_scale, id(cfg.group), cfg.split, cfg.want_ranks, sig, flags)


def _graphs_usable(ops, cfg: _Cfg, tensors) -> bool:
    if not cfg.use_graphs or not getattr(ops, "graph_capable", False):
        return False
    if not all(t is None or t.is_cuda for t in tensors):
        return False
    if ops.kernel_events is not None or getattr(ops, "timeline", None) is not None:
        return False  # developer timing modes record events around individual launches
    if cfg.world > 1 and dist.get_backend(cfg.group) != "nccl":
        return False  # only NCCL collectives can be captured
    return not torch.cuda.is_current_stream_capturing()  # (else: the caller is capturing the whole step already)


class _ContrastiveLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, image_tile_ids, text_tile_ids, neighbor_tile_ids,
                neighbor_alphas, cfg: _Cfg, positives=None):
        ops = _ops()
        b_local, d = image_features.shape
        n = cfg.world * b_local
        dev = image_features.device
        need_backward = any(ctx.needs_input_grad[:3])
        scale = logit_scale.detach().reshape(-1)[:1].to(device=dev, dtype=torch.float32).contiguous()
        if hasattr(ops, "check_shapes"):
            ops.check_shapes(b_local, n, d, cfg.split, any(ctx.needs_input_grad[:2]))
        img_in = image_features.detach().contiguous()
        txt_in = text_features.detach().contiguous()
        if txt_in.dtype != img_in.dtype:
            txt_in = txt_in.to(img_in.dtype)
        ins = [img_in, txt_in, scale, image_tile_ids, text_tile_ids, neighbor_tile_ids, neighbor_alphas]
        ins += list(positives) if positives is not None else []

        step = None
        if positives is None and _graphs_usable(ops, cfg, ins):
            same_ids = cfg.kind == "spatial" and image_tile_ids.data_ptr() == text_tile_ids.data_ptr()
            key = _graph_key(cfg, ins, (need_backward, same_ids, tuple(ctx.needs_input_grad[:3])))
            step = _GRAPHS.setdefault(key, _GraphedStep())
            if step.broken:
                step = None
        if step is not None:
            step.calls += 1
            if step.calls <= step.WARMUP:
                step = None
            elif step.fwd is None and sum(st.fwd is not None for st in _GRAPHS.values()) >= _MAX_GRAPHED_CONFIGS:
                step.broken = True  # this configuration stays eager
                step = None
        if step is None:
            out4, lists, ranks, saved, c = _forward_impl(ops, cfg, img_in, txt_in, scale, image_tile_ids, text_tile_ids,
                                                         neighbor_tile_ids, neighbor_alphas, positives, need_backward)
        else:
            if step.fwd is None:  # capture
                static = [t.clone() if t is not None else None for t in ins]
                if cfg.kind == "spatial" and image_tile_ids.data_ptr() == text_tile_ids.data_ptr():
                    static[4] = static[3]  # keep the aliasing the capture-time control flow saw
                launches0 = ops.launches
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(dev)
                with torch.cuda.graph(graph, pool=step.pool, capture_error_mode="thread_local"):
                    outs = _forward_impl(ops, cfg, static[0], static[1], static[2], static[3], static[4], static[5],
                                         static[6], None, need_backward)
                step.pool = graph.pool()
                step.fwd = (graph, static, outs)
                step.fwd_launches = ops.launches - launches0
            else:
                step._copy_in(step.fwd[1], ins)
                ops.launches += step.fwd_launches
            step.fwd[0].replay()
            step.generation += 1
            out4, lists, ranks, saved, c = step.fwd[2]  # (lists / ranks: "of the last call", static buffers)

        ctx.cfg = cfg
        ctx.d = d
        ctx.c = c
        ctx.in_dtypes = (image_features.dtype, text_features.dtype, logit_scale.dtype)
        ctx.scale_shape = logit_scale.shape
        ctx.scale_device = logit_scale.device
        ctx.step = step
        ctx.generation = step.generation if step is not None else 0
        if step is None:
            ctx.save_for_backward(*saved)
        else:
            ctx.saved_static = saved  # graph-pool buffers: valid until the next forward of this configuration
        col_it, w_it, q_it = lists
        if positives is not None:  # the caller already holds them: do not hand inputs back as outputs
            col_it, w_it, q_it = (t.new_empty((0,)) for t in (col_it, w_it, q_it))
        ctx.mark_non_differentiable(col_it, w_it, q_it, ranks)
        return out4[0].clone(), col_it, w_it, q_it, ranks

    @staticmethod
    def backward(ctx, grad_loss, *_unused):
        ops = _ops()
        cfg: _Cfg = ctx.cfg
        step: Optional[_GraphedStep] = ctx.step
        need_i, need_t, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        go = grad_loss.detach().reshape(1).to(torch.float32).contiguous()
        if step is None:
            saved = ctx.saved_tensors
            d_img, d_txt = _backward_impl(ops, cfg, saved, go, ctx.c, ctx.d, need_i, need_t, ctx.in_dtypes)
            out4 = saved[7]
        else:
            if step.broken or ctx.generation != step.generation:
                raise RuntimeError("spatial_clip_b200: backward() of a loss whose CUDA-graph buffers were overwritten by a "
                                   "later forward of the same configuration; call backward before the next forward, or "
                                   "construct the loss with cuda_graphs=False")
            saved = ctx.saved_static
            out4 = saved[7]
            if step.bwd is None:  # capture
                go_static = go.clone()
                launches0 = ops.launches
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(go.device)
                with torch.cuda.graph(graph, pool=step.pool, capture_error_mode="thread_local"):
                    outs = _backward_impl(ops, cfg, saved, go_static, ctx.c, ctx.d, need_i, need_t, ctx.in_dtypes)
                step.bwd = (graph, [go_static], outs)
                step.bwd_launches = ops.launches - launches0
            else:
                step._copy_in(step.bwd[1], [go])
                ops.launches += step.bwd_launches
            step.bwd[0].replay()
            d_img, d_txt = (t.clone() if t is not None else None for t in step.bwd[2])
        d_scale = None
        if need_s:
            # straight-through cap: d s_eff / d s == 1 even when clipped (losses.py:73-76)
            d_scale = (go * out4[2]).to(device=ctx.scale_device, dtype=ctx.in_dtypes[2]).reshape(ctx.scale_shape)
        return d_img, d_txt, d_scale, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------------
class _LossBase(nn.Module):
    def __init__(self, local_loss, gather_with_grad, rank, world_size, use_horovod, precision="bf16",
                 track_retrieval_ranks=False, cuda_graphs=True):
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
        # After two eager calls per configuration, forward and backward are replayed as CUDA graphs (kernels AND the
        # NCCL exchanges): the step's host cost drops from ~1 ms to a few copies and two replays.  Requires the training
        # loop's order (each backward before the next forward of the same shapes); set False otherwise.
        self.cuda_graphs = bool(cuda_graphs)
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
                 track_retrieval_ranks: bool = False, cuda_graphs: bool = True):
        super().__init__(local_loss, gather_with_grad, rank, world_size, use_horovod, precision, track_retrieval_ranks,
                         cuda_graphs)
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
                   self.precision == "fp32", self.track_retrieval_ranks, self.cuda_graphs)
        loss, col, w, q, ranks = _ContrastiveLossFn.apply(image_features, text_features,
                                                          self._scale_tensor(logit_scale, image_features),
                                                          image_tile_ids, text_tile_ids, neighbor_tile_ids,
                                                          neighbor_alphas, cfg, None)
        self.last_positives = (col, w, q)
        self.last_retrieval_ranks = ranks if self.track_retrieval_ranks else None
        return {"contrastive_loss": loss}


class _ListCheck:
    """Outcome of scl_check_positives for caller-resolved soft-target lists: a device flag that is read without
    stalling the step.  The first ``sync_calls`` forwards block on it (a mis-configured producer -- e.g. local columns
    on every rank -- fails on the first batch); afterwards the flag of step i is copied to pinned memory and examined at
    step i + 1, so a bad batch still raises, one step late, at no synchronisation cost."""

    BITS = ((1, "a column outside [-1, N) (such entries were dropped)"),
            (2, "slot 0 of a row is not the row's own column rank * B_l + i (lists built for another rank / batch?)"),
            (4, "an unused slot (column -1) carried weight"))

    def __init__(self, sync_calls: int = 1):
        self.sync_calls = sync_calls
        self.calls = 0
        self.pending = None  # (pinned host int32[1], event)

    @classmethod
    def _raise(cls, bits: int):
        what = "; ".join(msg for bit, msg in cls.BITS if bits & bit)
        raise ValueError(f"SpatialLossFromColumns: invalid positive_columns / positive_probs: {what}")

    def submit(self, flags):
        self.poll(block=False)
        self.calls += 1
        flag = flags[0] if len(flags) == 1 else flags[0] | flags[1]
        if not flag.is_cuda or self.calls <= self.sync_calls:
            bits = int(flag.reshape(-1)[0])
            if bits:
                self._raise(bits)
            return
        host = torch.empty((1,), dtype=torch.int32).pin_memory()
        host.copy_(flag.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(flag.device))
        self.pending = (host, ev)

    def poll(self, block: bool):
        if self.pending is None:
            return
        host, ev = self.pending
        if block:
            ev.synchronize()
        elif not ev.query():
            return
        self.pending = None
        if int(host[0]):
            self._raise(int(host[0]))


class SpatialLossFromColumns(SpatialLoss):
    """``SpatialLoss`` fed with soft targets that the data pipeline already resolved to global columns
    (producer: ``spatial_clip_b200.positives.resolve_positive_columns`` at collate time).

    Same constructor and same arithmetic as ``SpatialLoss`` (reference: losses.py:11-124); ``forward`` takes
    ``positive_columns`` int32 / ``positive_probs`` fp32 ``[B_l, K+1]`` (slot 0 = the row's own column) instead of the
    four id / neighbour tensors, so the two id all-gathers (losses.py:63-68), the id -> column map and the label loop
    (losses.py:91-111) leave the step altogether.  ``neighbor_alpha_scale`` is applied by the producer, not here.
    ``positive_*_text`` are the text-row lists when the two id vectors differ (the reference's loader makes them
    equal).  The LightningModule dispatches by parameter name (spatial_clip_module.py:44,58-61), so a collate that
    adds these keys (``positives.collate_positive_columns``) is all the integration needs.

    The lists come from outside the library, so every call passes them through ``scl_check_positives``: columns
    outside [-1, N) are dropped before any kernel indexes with them, and a violated contract (range, slot 0 = own
    column ``rank * B_l + i``, weight on an unused slot) raises ``ValueError`` -- synchronously on the first call, one
    step late afterwards (``_ListCheck``)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._list_check = _ListCheck()

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
        if positive_columns.shape[1] > 32:
            raise ValueError("at most 31 neighbour slots per row")
        if (positive_columns_text is None) != (positive_probs_text is None):
            raise ValueError("positive_columns_text and positive_probs_text go together")
        dev = image_features.device
        rank, world = self.rank, self.world_size
        ops = _ops()
        flags = []

        def prep(col, q, w):
            col = col.to(device=dev, dtype=torch.int32).contiguous()
            q = q.to(device=dev, dtype=torch.float32).contiguous()
            col, q, flag = ops.check_positives(col, q, world * b, rank)
            flags.append(flag)
            w = q if w is None else w.to(device=dev, dtype=torch.float32).contiguous()
            return col, w, q

        pos = prep(positive_columns, positive_probs, positive_weights)
        if positive_columns_text is not None:
            if positive_columns_text.shape != positive_columns.shape or \
                    positive_probs_text.shape != positive_columns.shape:
                raise ValueError("text-row lists must have the shape of the image-row lists")
            pos = pos + prep(positive_columns_text, positive_probs_text, None)
        self._list_check.submit(flags)
        cfg = _Cfg("spatial", rank, world, self.local_loss, self.gather_with_grad,
                   self.cap_logit_scale, self.temp_reg_weight, self.neighbor_alpha_scale, self.process_group,
                   self.precision == "fp32", self.track_retrieval_ranks, self.cuda_graphs)
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
                 precision: str = "bf16", track_retrieval_ranks: bool = False, cuda_graphs: bool = True):
        super().__init__(local_loss, gather_with_grad, rank, world_size, use_horovod, precision, track_retrieval_ranks,
                         cuda_graphs)
        self.cache_labels = cache_labels  # labels are implicit (the diagonal); nothing to cache

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
                logit_bias: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        self._check_features(image_features, text_features, logit_bias)
        image_features, text_features = self._kernel_dtype(image_features), self._kernel_dtype(text_features)
        cfg = _Cfg("clip", self.rank, self.world_size, self.local_loss, self.gather_with_grad, None, 0.0, 1.0,
                   self.process_group, self.precision == "fp32", self.track_retrieval_ranks, self.cuda_graphs)
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
                 track_retrieval_ranks: bool = False, cuda_graphs: bool = True):
        super().__init__(local_loss, gather_with_grad, rank, world_size, use_horovod, cap_logit_scale,
                         temp_reg_weight, float32_logits, neighbor_alpha_scale, precision, track_retrieval_ranks,
                         cuda_graphs)
        self.cache_labels = cache_labels

    def forward(self, image_features, text_features, image_tile_ids, text_tile_ids, neighbor_tile_ids,
                neighbor_alphas, logit_scale, logit_bias=None, output_dict: bool = False):
        out = super().forward(image_features, text_features, logit_scale, image_tile_ids, text_tile_ids,
                              neighbor_tile_ids, neighbor_alphas, logit_bias)
        return out if output_dict else out["contrastive_loss"]
