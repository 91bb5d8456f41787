"""CPU test of the ctypes layer (spatial_clip_b200/_cuda.py) for every mode the modules can select: each C entry
point is replaced by a ctypes callback with the SAME prototype, so argument counts / types / struct layouts are
checked by ctypes exactly as in a real call, while nothing is launched.  Host-only entry points (plans, work-area
sizes, error strings) are forwarded to the real library.  What this cannot see is the device code."""
import ctypes as C

import pytest
import torch

from spatial_clip_b200 import ClipLoss, SpatialLoss, _cuda, losses
from spatial_clip_b200.synth import make_spot_batch

HOST_ONLY = {"scl_abi_version", "scl_error_string", "scl_fwd_plan", "scl_bwd_plan", "scl_positives_workspace_bytes",
             "scl_fwd_workspace_bytes", "scl_bwd_workspace_bytes", "scl_bwd_finish_workspace_bytes"}


class _CallbackLib:
    def __init__(self, real):
        self.calls = []
        self.fwd_phases = []
        self._keep = []
        for name, (res, args) in _cuda.EXPORTS.items():
            if name in HOST_ONLY:
                setattr(self, name, getattr(real, name))
                continue
            proto = C.CFUNCTYPE(res, *args)

            def make(nm):
                def cb(*a):
                    self.calls.append(nm)
                    if nm == "scl_fwd_all":  # which phases of the composite forward were asked for
                        self.fwd_phases.append(C.cast(a[0], C.POINTER(_cuda.FwdArgs)).contents.phases)
                    return 0
                return cb

            fn = proto(make(name))
            self._keep.append(fn)
            setattr(self, name, fn)


class _PlumbingOps(_cuda.CudaOps):
    def __init__(self, lib):
        self.lib = lib
        self._checked = set()
        self.launches = 0
        self.kernel_events = None

    def _stream(self, t):
        return 0

    def _timed(self, name, device, fn):
        return fn()


class _NoGuard:
    def __init__(self, device):
        pass

    def __enter__(self):
        pass

    def __exit__(self, *exc):
        pass


@pytest.fixture()
def plumbing(monkeypatch):
    from spatial_clip_b200 import build

    build.build()
    real = _cuda.load_library()
    monkeypatch.setattr(_cuda, "_DeviceGuard", _NoGuard)
    lib = _CallbackLib(real)

    def install(**kw):
        ops = _PlumbingOps(lib, **kw)
        losses._set_ops_for_testing(ops)
        return ops, lib

    yield install
    losses._set_ops_for_testing(None)


def _step(mod, b, spatial=True, d_dtype=torch.float32):
    img = b.image_features.to(d_dtype).requires_grad_(True)
    txt = b.text_features.to(d_dtype).requires_grad_(True)
    s = torch.tensor(30.0, requires_grad=True)
    if spatial:
        out = mod(img, txt, s, b.tile_ids, b.tile_ids.clone(), b.neighbor_tile_ids, b.neighbor_alphas)
    else:
        out = mod(img, txt, s)
    out["contrastive_loss"].backward()
    assert img.grad.shape == img.shape and img.grad.dtype == d_dtype and txt.grad.shape == txt.shape
    return out


@pytest.mark.parametrize("mode", ["default", "fp32", "ranks", "bf16_inputs", "timed_kernels", "d768", "d1280"])
def test_every_mode_reaches_the_library_with_well_formed_calls(plumbing, mode):
    d = {"d768": 768, "d1280": 1280}.get(mode, 128)
    b = make_spot_batch(n=300, d=d, k=8, seed=3)
    ops, lib = plumbing()
    kw = {}
    if mode == "fp32":
        kw["precision"] = "fp32"
    if mode == "ranks":
        kw["track_retrieval_ranks"] = True
    cfg = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
               neighbor_alpha_scale=0.5, float32_logits=True)
    mod = SpatialLoss(**cfg, **kw)
    if mode == "timed_kernels":  # bench.py's roofline / developer timing mode: the backward launches go out one by one
        ops.kernel_events = {}
        _step(mod, b)
        assert {"scl_bwd_coeffs", "scl_bwd_rows", "scl_bwd_finish"} <= set(lib.calls) and "scl_bwd_dir" not in lib.calls
        return
    _step(mod, b, d_dtype=torch.bfloat16 if mode == "bf16_inputs" else torch.float32)
    calls = lib.calls
    assert calls.count("scl_fwd_all") == 1 and calls.count("scl_bwd_dir") == 2
    if mode == "fp32":
        assert calls.count("scl_split_bf16") == 2
        assert "scl_prepare" not in calls and "scl_cast_bf16" not in calls
    else:
        assert calls.count("scl_prepare") == 1 and "scl_cast_bf16" not in calls  # no transposed copies anywhere
    if mode == "ranks":
        assert mod.last_retrieval_ranks is not None and mod.last_retrieval_ranks.shape == (300,)
    # plain CLIP through the same layer
    lib.calls.clear()
    _step(ClipLoss(**({"precision": "fp32"} if mode == "fp32" else {})), b, spatial=False)
    assert lib.calls.count("scl_fwd_all") == 1 and lib.calls.count("scl_bwd_dir") == 2


def test_unsupported_width_is_refused_before_any_launch(plumbing):
    ops, lib = plumbing()
    b = make_spot_batch(n=64, d=96, k=0, seed=1)  # D % 64 != 0
    with pytest.raises(_cuda.SclError) as ei:
        ClipLoss()(b.image_features, b.text_features, torch.tensor(10.0))
    assert "unsupported shape" in str(ei.value) and not lib.calls


def _gloo_worker(rank, world, port, mode, q):
    import os
    import traceback

    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        _cuda._DeviceGuard = _NoGuard
        lib = _CallbackLib(_cuda.load_library())
        losses._set_ops_for_testing(_PlumbingOps(lib))
        b = make_spot_batch(n=256, d=128, k=8, seed=5).rank_slice(rank, world)
        cfg = dict(local_loss=True, gather_with_grad=(mode != "local_only"), cap_logit_scale=40.0, temp_reg_weight=0.05,
                   neighbor_alpha_scale=0.5, float32_logits=True)
        _step(SpatialLoss(**cfg, **({"precision": "fp32"} if mode == "fp32" else {})), b)
        q.put((rank, list(lib.calls)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("mode", ["default", "fp32", "local_only"])
def test_two_ranks_over_gloo_issue_the_same_call_sequence(mode):
    """world_size 2: gathered operands, one scl_fwd_all call per forward phase and the single statistics-record
    exchange (scl_unpack_records with its pointer tables) -- none at all when no other rank's rows reach the local
    features (local_loss without a differentiable gather, like the reference's backward)."""
    import socket

    import torch.multiprocessing as mp

    from spatial_clip_b200 import build

    build.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, mode, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    for r in range(2):
        assert isinstance(got[r], list), got[r]
    assert got[0] == got[1], "ranks must issue identical call sequences (collective order)"
    calls = got[0]
    # exchanges in flight: one scl_fwd_all call per phase (soft targets / image-rows pass / text-rows pass)
    assert calls.count("scl_fwd_all") == 3 and calls.count("scl_bwd_dir") == 2
    assert calls.count("scl_unpack_records") == (0 if mode == "local_only" else 1)
    assert "scl_cast_bf16" not in calls


def test_precomputed_columns_skip_the_builder_phase(plumbing):
    """SpatialLossFromColumns: one scl_fwd_all call with phases = 6 (both similarity passes, no soft-target builder);
    the caller's int32 / fp32 lists reach it through scl_check_positives."""
    from spatial_clip_b200 import SpatialLossFromColumns
    from spatial_clip_b200.positives import resolve_positive_columns

    ops, lib = plumbing()
    b = make_spot_batch(n=300, d=128, k=8, seed=3)
    col, w, q = resolve_positive_columns(b.tile_ids, b.neighbor_tile_ids, b.neighbor_alphas, 0.5)
    img = b.image_features.clone().requires_grad_(True)
    txt = b.text_features.clone().requires_grad_(True)
    mod = SpatialLossFromColumns(cap_logit_scale=40.0, temp_reg_weight=0.05)
    mod(img, txt, torch.tensor(30.0, requires_grad=True), positive_columns=col, positive_probs=q,
        positive_weights=w)["contrastive_loss"].backward()
    assert lib.calls.count("scl_fwd_all") == 1 and lib.fwd_phases == [6] and lib.calls.count("scl_bwd_dir") == 2
    assert lib.calls.count("scl_check_positives") == 1
    assert img.grad.shape == img.shape and mod.last_positives[0].dtype == torch.int32
    # the ordinary route asks for everything in one call
    lib.calls.clear()
    lib.fwd_phases.clear()
    _step(SpatialLoss(), b)
    assert lib.fwd_phases == [0]
