"""ctypes binding of libscl_b200.so (include/scl_b200.h) -> ``CudaOps``.

This is the ONLY compute backend of the package.  There is no CPU or eager-PyTorch implementation:
if the shared library is missing, cannot be loaded, or a tensor is not on a CUDA device, the call
raises.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(__file__).resolve().parent / "libscl_b200.so"
_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class SclPlan(C.Structure):
    _fields_ = [("chunks", C.c_int), ("tiles_per_chunk", C.c_int), ("n_slots", C.c_int), ("m_pad", C.c_int),
                ("n_pad", C.c_int), ("d_split", C.c_int), ("split", C.c_int)]


class SclError(RuntimeError):
    pass


_VP, _I, _F, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class PrepareArgs(C.Structure):
    _fields_ = [("image", _VP), ("text", _VP), ("src_dtype", _I), ("logit_scale", _VP), ("cap", _F), ("rows", _I),
                ("d", _I), ("image_bf16", _VP), ("text_bf16", _VP), ("scalars3", _VP)]


class FwdArgs(C.Structure):
    _fields_ = [("img_l", _VP), ("txt_l", _VP), ("img_all", _VP), ("txt_all", _VP), ("b_local", _I), ("n_global", _I),
                ("d", _I), ("rank", _I), ("scalars3", _VP), ("img_ids_all", _VP),
                ("txt_ids_all", _VP), ("nbr_ids", _VP), ("nbr_alpha", _VP), ("k", _I), ("alpha_scale", _F),
                ("same_ids", _I), ("c", _F), ("w", _F), ("finalize_scalars", _I), ("col_it", _VP), ("w_it", _VP),
                ("q_it", _VP), ("col_ti", _VP), ("w_ti", _VP), ("q_ti", _VP), ("stats_i", _VP), ("stats_t", _VP),
                ("sums6", _VP), ("out4", _VP), ("workspace", _VP), ("workspace_bytes", _SZ), ("ranks_out", _VP),
                ("phases", _I)]


class BwdArgs(C.Structure):
    _fields_ = [("x_rows", _VP), ("y_all", _VP), ("b_local", _I), ("n_global", _I),
                ("d", _I), ("rank", _I), ("row_stats", _VP), ("col_stats_all", _VP), ("pos_col", _VP),
                ("pos_q", _VP), ("opp_q_local", _VP), ("opp_col_all", _VP), ("opp_q_all", _VP), ("k_plus_1", _I),
                ("gaps", _VP), ("scalars3", _VP), ("grad_out", _VP), ("c", _F), ("w", _F), ("mult", _F),
                ("col_mode", _I), ("dx_out", _VP), ("out_dtype", _I), ("workspace", _VP), ("workspace_bytes", _SZ),
                ("split", _I)]


EXPORTS = {
    # name: (restype, argtypes)
    "scl_abi_version": (C.c_int, []),
    "scl_error_string": (C.c_char_p, [C.c_int]),
    "scl_check_device": (C.c_int, [C.POINTER(C.c_int)]),
    "scl_fwd_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(SclPlan)]),
    "scl_bwd_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(SclPlan)]),
    "scl_split_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "scl_cast_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "scl_check_positives": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "scl_prep_scalars": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "scl_positives_workspace_bytes": (C.c_size_t, [C.c_int]),
    "scl_build_positives": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                      C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "scl_fwd_rowstats": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.POINTER(SclPlan), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "scl_fwd_rowstats_ranks": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.POINTER(SclPlan), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "scl_row_finalize": (C.c_int, [C.c_void_p, C.POINTER(SclPlan), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "scl_reduce_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scl_loss_scalars": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "scl_bwd_coeffs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SclPlan), C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scl_bwd_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.POINTER(SclPlan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scl_bwd_finish_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "scl_bwd_finish": (C.c_int, [C.c_void_p, C.POINTER(SclPlan), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p,
                                 C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]),
    "scl_prepare": (C.c_int, [C.POINTER(PrepareArgs), C.c_void_p]),
    "scl_fwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "scl_fwd_all": (C.c_int, [C.POINTER(FwdArgs), C.c_void_p]),
    "scl_bwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "scl_bwd_dir": (C.c_int, [C.POINTER(BwdArgs), C.c_void_p]),
    "scl_unpack_records": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load libscl_b200.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise SclError(
            f"{_LIB_PATH} is missing. Build it with `python -m spatial_clip_b200.build` "
            "(nvcc, sm_100a). spatial_clip_b200 has no CPU / eager fallback.")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.scl_abi_version() != 7:
        raise SclError("libscl_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


class _DeviceGuard:
    """`with torch.cuda.device(d)` costs ~10 us per use; skip it when d is already the current device (the
    normal case: DDP ranks, and the autograd engine sets the device on its threads)."""

    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


class CudaOps:
    """Tensor-level wrappers over the C ABI.  Every method launches on the current stream of the
    tensors' device and returns without synchronising."""

    name = "cuda"
    graph_capable = True  # every method only enqueues work on the current stream: safe under CUDA-graph capture

    def __init__(self):
        self.lib = load_library()
        self._checked = set()
        self.launches = 0  # kernels launched through this object (bench.py reports it)
        self.kernel_events = None  # set to {} to record CUDA events around the tensor-core kernels

    def mark(self, name, device):
        """Developer timeline (bench.py --timeline): a CUDA event on the current stream at a named point of the step;
        consecutive marks bracket the phases (exchanges, passes, finishes).  No-op unless self.timeline is a list."""
        tl = getattr(self, "timeline", None)
        if tl is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(device))
        tl.append((name, ev))

    def _timed(self, name, device, fn):
        """Run one C-ABI launch, optionally bracketed by CUDA events on its stream (bench.py roofline)."""
        if self.kernel_events is None:
            return fn()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream(device)
        a.record(st)
        rc = fn()
        b.record(st)
        self.kernel_events.setdefault(name, []).append((a, b))
        return rc

    # ---------------------------------------------------------------- helpers
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.scl_error_string(rc).decode()
            raise SclError(f"{what} failed: {msg} (code {rc})")

    def _stream(self, t: torch.Tensor) -> int:
        if not t.is_cuda:
            raise SclError("spatial_clip_b200 kernels need CUDA tensors (sm_100a); got a CPU tensor and there is "
                           "no CPU fallback")
        idx = t.device.index
        if idx not in self._checked:
            with torch.cuda.device(idx):
                n = C.c_int(0)
                self._check(self.lib.scl_check_device(C.byref(n)), "scl_check_device")
            self._checked.add(idx)
        return torch.cuda.current_stream(t.device).cuda_stream

    @staticmethod
    def empty(shape, dtype, like: torch.Tensor) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=like.device)

    # ---------------------------------------------------------------- plans
    def fwd_plan(self, m_rows: int, n_cols: int, d: int) -> SclPlan:
        p = SclPlan()
        self._check(self.lib.scl_fwd_plan(m_rows, n_cols, d, C.byref(p)), "scl_fwd_plan")
        return p

    def bwd_plan(self, m_rows: int, n_cols: int, d: int, split: bool = False) -> SclPlan:
        p = SclPlan()
        self._check(self.lib.scl_bwd_plan(m_rows, n_cols, d, int(split), C.byref(p)), "scl_bwd_plan")
        return p

    def check_shapes(self, b_local: int, n_global: int, d: int, split: bool, need_backward: bool):
        """Raise for an unsupported [rows, D] / precision combination before anything is launched."""
        self.fwd_plan(b_local, n_global, 3 * d if split else d)
        if need_backward or split:
            self.bwd_plan(b_local, n_global, d, split)

    # ---------------------------------------------------------------- ops
    def cast_bf16(self, x, normalize=False):
        st = self._stream(x)
        rows, d = x.shape
        y = self.empty((rows, d), torch.bfloat16, x)
        with _DeviceGuard(x.device):
            self._check(self.lib.scl_cast_bf16(_ptr(x), _DTYPE_CODE[x.dtype], _ptr(y), rows, d, int(normalize), st),
                        "scl_cast_bf16")
        self.launches += 1
        return y

    def split_cast(self, x, want_rows=True, want_cols=True):
        """fp32-accurate mode operands: x [rows, D] -> (h|h|l) and/or (h|l|h) bf16 rows of width 3 D (scl_split_bf16)."""
        st = self._stream(x)
        rows, d = x.shape
        r = self.empty((rows, 3 * d), torch.bfloat16, x) if want_rows else None
        c = self.empty((rows, 3 * d), torch.bfloat16, x) if want_cols else None
        with _DeviceGuard(x.device):
            self._check(self.lib.scl_split_bf16(_ptr(x), _DTYPE_CODE[x.dtype], _ptr(r), _ptr(c), rows, d, st),
                        "scl_split_bf16")
        self.launches += 1
        return r, c

    def prep_scalars(self, logit_scale, cap):
        st = self._stream(logit_scale)
        out = self.empty((3,), torch.float32, logit_scale)
        with _DeviceGuard(logit_scale.device):
            self._check(self.lib.scl_prep_scalars(_ptr(logit_scale), float(cap) if cap is not None else -1.0,
                                                  _ptr(out), st), "scl_prep_scalars")
        self.launches += 1
        return out

    def build_positives(self, all_ids, nbr_ids, nbr_alpha, b_local, k, alpha_scale, rank, like):
        st = self._stream(like)
        n_global = all_ids.shape[0] if all_ids is not None else 0
        kp1 = k + 1
        col = self.empty((b_local, kp1), torch.int32, like)
        w = self.empty((b_local, kp1), torch.float32, like)
        q = self.empty((b_local, kp1), torch.float32, like)
        ws = None
        ws_bytes = 0
        if k > 0:
            ws_bytes = self.lib.scl_positives_workspace_bytes(n_global)
            ws = self.empty((ws_bytes,), torch.uint8, like)
        with _DeviceGuard(like.device):
            self._check(self.lib.scl_build_positives(_ptr(all_ids), max(n_global, b_local), _ptr(nbr_ids),
                                                     _ptr(nbr_alpha), b_local, k, float(alpha_scale), rank, _ptr(ws),
                                                     ws_bytes, _ptr(col), _ptr(w), _ptr(q), st),
                        "scl_build_positives")
        self.launches += 3 if k > 0 else 1
        return col, w, q

    def check_positives(self, col, q, n_global, rank):
        """Sanitised copies of caller-resolved soft-target lists + a device flag (int32[1]) of what was wrong with
        them (scl_check_positives; bit 0 column out of range, bit 1 slot 0 is not the own column, bit 2 weight on an
        unused slot)."""
        st = self._stream(col)
        b_local, kp1 = col.shape
        col_out = torch.empty_like(col)
        q_out = torch.empty_like(q)
        flag = torch.zeros((1,), dtype=torch.int32, device=col.device)
        with _DeviceGuard(col.device):
            self._check(self.lib.scl_check_positives(_ptr(col), _ptr(q), b_local, kp1, n_global, rank, _ptr(col_out),
                                                     _ptr(q_out), _ptr(flag), st), "scl_check_positives")
        self.launches += 1
        return col_out, q_out, flag

    def fwd_rowstats(self, x_rows, y_cols, scalars, debug_z=False):
        st = self._stream(x_rows)
        m, d = x_rows.shape
        n = y_cols.shape[0]
        plan = self.fwd_plan(m, n, d)
        partial = self.empty((plan.n_slots * plan.m_pad, 4), torch.float32, x_rows)
        dbg = torch.zeros((m, n), dtype=torch.float32, device=x_rows.device) if debug_z else None
        with _DeviceGuard(x_rows.device):
            self._check(self._timed("fwd_rowstats", x_rows.device, lambda: self.lib.scl_fwd_rowstats(
                _ptr(x_rows), m, _ptr(y_cols), n, d, _ptr(scalars), C.byref(plan), _ptr(partial), _ptr(dbg), n, st)),
                "scl_fwd_rowstats")
        self.launches += 1
        return (partial, plan, dbg) if debug_z else (partial, plan)

    def row_finalize(self, partial, plan, x_rows, y_all, pos_col, pos_q):
        st = self._stream(x_rows)
        m, d = x_rows.shape
        stats = self.empty((m, 4), torch.float32, x_rows)
        with _DeviceGuard(x_rows.device):
            self._check(self.lib.scl_row_finalize(_ptr(partial), C.byref(plan), m, d, _ptr(x_rows), _ptr(y_all),
                                                  _ptr(pos_col), _ptr(pos_q), pos_col.shape[1], _ptr(stats), st),
                        "scl_row_finalize")
        self.launches += 1
        return stats

    def reduce_rows(self, stats_a, stats_b, scalars):
        st = self._stream(stats_a)
        sums = self.empty((6,), torch.float32, stats_a)
        with _DeviceGuard(stats_a.device):
            self._check(self.lib.scl_reduce_rows(_ptr(stats_a), _ptr(stats_b), stats_a.shape[0], _ptr(scalars),
                                                 _ptr(sums), st), "scl_reduce_rows")
        self.launches += 1
        return sums

    def loss_scalars(self, sums6, scalars, c, w):
        st = self._stream(sums6)
        out = self.empty((4,), torch.float32, sums6)
        with _DeviceGuard(sums6.device):
            self._check(self.lib.scl_loss_scalars(_ptr(sums6), _ptr(scalars), float(c), float(w), _ptr(out), st),
                        "scl_loss_scalars")
        self.launches += 1
        return out

    # ---------------------------------------------------------------- composite phases (one host call each)
    def prepare(self, image, text, logit_scale, cap, split=False):
        """cap + bf16 casts of both modalities, one launch: scl_prepare.  Returns the row operands, the column
        operands (the same tensors unless split) and the scalars."""
        st = self._stream(image)
        rows, d = image.shape
        if text.dtype != image.dtype:
            text = text.to(image.dtype)
        if split:  # fp32-accurate mode: bf16 hi/lo pairs, K-concatenated
            img_r, img_c = self.split_cast(image)
            txt_r, txt_c = self.split_cast(text)
            return img_r, txt_r, img_c, txt_c, self.prep_scalars(logit_scale, cap)
        both = torch.empty((2, rows, d), dtype=torch.bfloat16, device=image.device)
        img, txt = both[0], both[1]
        scal = self.empty((3,), torch.float32, image)
        a = PrepareArgs(_ptr(image), _ptr(text), _DTYPE_CODE[image.dtype], _ptr(logit_scale),
                        float(cap) if cap is not None else -1.0, rows, d, _ptr(img), _ptr(txt), _ptr(scal))
        with _DeviceGuard(image.device):
            self._check(self.lib.scl_prepare(C.byref(a), st), "scl_prepare")
        self.launches += 1
        return img, txt, img, txt, scal

    def forward_all(self, img_l, txt_l, img_all, txt_all, scalars, ids, b_local, rank, k, alpha_scale, c, w,
                    finalize_scalars, want_ranks=False, waits=None, positives=None):
        """Soft targets + both fused similarity/LSE passes + reductions: scl_fwd_all.
        ids = None (plain CLIP) or (img_ids_all, txt_ids_all, nbr_ids, nbr_alpha, same_ids).
        want_ranks: also count, in the image-rows pass, each row's in-batch retrieval rank (self.last_ranks).
        waits = (wait_ids or None, wait_txt_all, wait_img_all): the gathered operands are still in flight; the call is
        then issued once per phase (soft targets / image-rows pass / text-rows pass + reductions), each right after
        the wait for the one operand that phase reads.
        positives = (col, w, q) int32 / fp32 / fp32 [B_l, K+1] of the image rows (+ the same three of the text rows):
        soft targets resolved on the data side (already passed through check_positives); the builder phase is
        skipped."""
        if self.kernel_events is not None and not want_ranks:  # bench.py roofline: time the tensor-core launches
            return self._forward_all_unfused(img_l, txt_l, img_all, txt_all, scalars, ids, b_local, rank, k,
                                             alpha_scale, c, w, finalize_scalars, waits, positives)
        st = self._stream(img_l)
        n, d = img_all.shape
        kp1 = k + 1
        dev = img_l.device
        same = ids is None or ids[4]
        if positives is not None:
            for t in positives:
                if t.device != dev or not t.is_contiguous() or tuple(t.shape) != (b_local, kp1):
                    raise SclError("precomputed soft-target lists must be contiguous [B_l, K+1] tensors on the "
                                   "features' device")
            col_it, w_it, q_it = positives[:3]
            col_ti, w_ti, q_ti = positives[3:6] if len(positives) >= 6 else positives[:3]
            same = len(positives) < 6
        else:
            lists = torch.empty((2 if not same else 1, 3, b_local, kp1), dtype=torch.float32, device=dev)
            col_it, w_it, q_it = lists[0, 0].view(torch.int32), lists[0, 1], lists[0, 2]
            if same:
                col_ti, w_ti, q_ti = col_it, w_it, q_it
            else:
                col_ti, w_ti, q_ti = lists[1, 0].view(torch.int32), lists[1, 1], lists[1, 2]
        small = torch.empty((2 * b_local + 3, 4), dtype=torch.float32, device=dev)
        stats_i, stats_t = small[:b_local], small[b_local:2 * b_local]
        sums6 = small[2 * b_local:2 * b_local + 2].reshape(-1)[:6]
        out4 = small[2 * b_local + 2]
        ws_bytes = self.lib.scl_fwd_workspace_bytes(b_local, n, d, k)
        if ws_bytes == 0:
            self._check(-2, "scl_fwd_workspace_bytes")
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        a = FwdArgs(_ptr(img_l), _ptr(txt_l), _ptr(img_all), _ptr(txt_all), b_local, n, d, rank,
                    _ptr(scalars), _ptr(ids[0]) if ids else None, _ptr(ids[1]) if ids else None,
                    _ptr(ids[2]) if ids else None, _ptr(ids[3]) if ids else None, k, float(alpha_scale), int(same),
                    float(c), float(w), int(finalize_scalars), _ptr(col_it), _ptr(w_it), _ptr(q_it), _ptr(col_ti),
                    _ptr(w_ti), _ptr(q_ti), _ptr(stats_i), _ptr(stats_t), _ptr(sums6), _ptr(out4), _ptr(ws), ws_bytes,
                    None, 0)
        ranks = None
        if want_ranks:
            ranks = torch.empty((b_local,), dtype=torch.int32, device=dev)
            a.ranks_out = _ptr(ranks)
        with _DeviceGuard(dev):
            if waits is None:
                a.phases = 6 if positives is not None else 0  # 6: both similarity passes, no soft-target builder
                self._check(self.lib.scl_fwd_all(C.byref(a), st), "scl_fwd_all")
            else:
                for phase, wait in zip((1, 2, 4), waits):
                    if wait is not None:
                        wait()
                    self.mark(f"fwd:wait{phase}", dev)
                    if phase == 1 and positives is not None:
                        continue
                    if phase == 1 and ids:  # the gathered id vectors exist only now (the wait de-interleaves them)
                        a.img_ids_all, a.txt_ids_all = _ptr(ids[0]), _ptr(ids[1])
                    a.phases = phase
                    self._check(self.lib.scl_fwd_all(C.byref(a), st), "scl_fwd_all")
        built = 0 if positives is not None else (3 if k > 0 else 1) * (1 if same else 2)
        self.launches += built + 5 + (2 if want_ranks else 0)
        return (col_it, w_it, q_it), (col_ti, w_ti, q_ti), stats_i, stats_t, sums6, out4, ranks

    def _forward_all_unfused(self, img_l, txt_l, img_all, txt_all, scalars, ids, b_local, rank, k, alpha_scale, c, w,
                             finalize_scalars, waits, positives):
        """The same sequence as scl_fwd_all, one C-ABI call per launch, so that CUDA events can bracket the two
        tensor-core passes (kernel_events mode only)."""
        wait_ids, wait_txt, wait_img = waits if waits is not None else (None, None, None)
        if wait_ids is not None:
            wait_ids()
        if positives is not None:
            it = tuple(positives[:3])
            ti = tuple(positives[3:6]) if len(positives) >= 6 else it
        elif ids is None:
            it = ti = self.build_positives(None, None, None, b_local, 0, 1.0, rank, img_l)
        else:
            it = self.build_positives(ids[1], ids[2], ids[3], b_local, k, alpha_scale, rank, img_l)
            ti = it if ids[4] else self.build_positives(ids[0], ids[2], ids[3], b_local, k, alpha_scale, rank, img_l)
        if wait_txt is not None:
            wait_txt()
        part, plan = self.fwd_rowstats(img_l, txt_all, scalars)
        stats_i = self.row_finalize(part, plan, img_l, txt_all, it[0], it[2])
        if wait_img is not None:
            wait_img()
        part, plan = self.fwd_rowstats(txt_l, img_all, scalars)
        stats_t = self.row_finalize(part, plan, txt_l, img_all, ti[0], ti[2])
        sums6 = self.reduce_rows(stats_i, stats_t, scalars)
        out4 = self.loss_scalars(sums6, scalars, c, w) if finalize_scalars else self.empty((4,), torch.float32, img_l)
        return it, ti, stats_i, stats_t, sums6, out4, None

    def backward_dir(self, x_rows, y_all, row_stats, col_stats, pos_col, pos_q, opp_col_all, opp_q_all,
                     b_local, rank, gaps, scalars, grad_out, c, w, mult, col_mode, out_dtype, opp_q_local, split=False):
        """dX of the local rows for one direction: scl_bwd_dir (coefficients + fused tensor-core pass + sparse finish).
        With kernel timing on (bench.py roofline) the three launches are issued separately.
        split: x_rows [m, 3 D], y_all [n, 3 D] (fp32-accurate mode); the result is [m, D]."""
        if self.kernel_events is not None:
            return self.bwd_rows(x_rows, y_all, row_stats, col_stats, pos_col, pos_q, opp_col_all, opp_q_all,
                                 b_local, rank, gaps, scalars, grad_out, c, w, mult, col_mode, out_dtype,
                                 opp_q_local=opp_q_local, split=split)
        st = self._stream(x_rows)
        m, d = x_rows.shape
        if split:
            d //= 3
        n = y_all.shape[0]
        kp1 = pos_col.shape[1]
        ws_bytes = self.lib.scl_bwd_workspace_bytes(m, n, d, kp1)
        if ws_bytes == 0:
            self._check(-2, "scl_bwd_workspace_bytes")
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x_rows.device)
        out = self.empty((m, d), out_dtype, x_rows)
        a = BwdArgs(_ptr(x_rows), _ptr(y_all), m, n, d, rank,
                    _ptr(row_stats), _ptr(col_stats), _ptr(pos_col), _ptr(pos_q), _ptr(opp_q_local),
                    _ptr(opp_col_all), _ptr(opp_q_all), kp1, _ptr(gaps), _ptr(scalars), _ptr(grad_out),
                    float(c), float(w), float(mult), col_mode, _ptr(out), _DTYPE_CODE[out_dtype], _ptr(ws), ws_bytes,
                    int(split))
        with _DeviceGuard(x_rows.device):
            self._check(self.lib.scl_bwd_dir(C.byref(a), st), "scl_bwd_dir")
        self.launches += 3 + (3 if col_mode != 0 else 0)
        return out

    def exchange_records(self, parts, world, gather_fn):
        """All-gather a list of per-rank [B_l, k] fp32/int32 tensors as ONE flat record per rank and split the
        result into contiguous rank-major [world*B_l, k] tensors (one collective + two launches)."""
        like = parts[0]
        st = self._stream(like)
        flat = torch.cat([p.reshape(-1).view(torch.float32) for p in parts])
        rec = flat.numel()
        gathered = gather_fn(flat.reshape(1, rec))  # [world, rec]
        outs = [self.empty((world * p.shape[0],) + tuple(p.shape[1:]), p.dtype, like) for p in parts]
        n = len(parts)
        ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        offs = (C.c_int * n)()
        lens = (C.c_int * n)()
        o = 0
        for k, p in enumerate(parts):
            offs[k] = o
            lens[k] = p.numel()
            o += p.numel()
        with _DeviceGuard(like.device):
            self._check(self.lib.scl_unpack_records(_ptr(gathered), world, rec, n, ptrs, offs, lens, st),
                        "scl_unpack_records")
        self.launches += 2
        return outs

    def bwd_rows(self, x_rows, y_all, row_stats, col_stats, pos_col, pos_q, opp_col_all, opp_q_all,
                 b_local, rank, gaps, scalars, grad_out, c, w, mult, col_mode, out_dtype, opp_q_local=None,
                 split=False):
        """dX for the local rows: coefficients + fused tensor-core pass + sparse finish."""
        st = self._stream(x_rows)
        m, d = x_rows.shape
        if split:
            d //= 3
        n = y_all.shape[0]
        kp1 = pos_col.shape[1]
        plan = self.bwd_plan(m, n, d, split)
        row_coef = self.empty((plan.m_pad, 4), torch.float32, x_rows)
        col_coef = self.empty((plan.n_pad, 4), torch.float32, x_rows)
        partial = self.empty((plan.chunks, plan.m_pad, d), torch.float32, x_rows)
        out = self.empty((m, d), out_dtype, x_rows)
        fin_bytes = self.lib.scl_bwd_finish_workspace_bytes(n, b_local, kp1) if col_mode != 0 else 0
        fin_ws = self.empty((fin_bytes,), torch.uint8, x_rows) if fin_bytes else None
        if opp_q_local is None:  # single rank: the gathered opposite list IS the local one
            opp_q_local = opp_q_all[rank * b_local:(rank + 1) * b_local]
        with _DeviceGuard(x_rows.device):
            self._check(self.lib.scl_bwd_coeffs(_ptr(row_stats), m, _ptr(col_stats), n, C.byref(plan), b_local, rank,
                                                _ptr(gaps), _ptr(scalars), _ptr(grad_out), float(c), float(w),
                                                float(mult), col_mode, _ptr(pos_q), _ptr(opp_q_local),
                                                kp1, _ptr(row_coef), _ptr(col_coef), st),
                        "scl_bwd_coeffs")
            self._check(self._timed("bwd_rows", x_rows.device, lambda: self.lib.scl_bwd_rows(
                _ptr(x_rows), m, _ptr(y_all), n, d, rank * b_local, _ptr(scalars), C.byref(plan),
                _ptr(row_coef), _ptr(col_coef), _ptr(partial), st)),
                "scl_bwd_rows")
            self._check(self.lib.scl_bwd_finish(_ptr(partial), C.byref(plan), m, d, _ptr(y_all), _ptr(pos_col),
                                                _ptr(pos_q), kp1, _ptr(opp_col_all), _ptr(opp_q_all), n,
                                                b_local, rank, _ptr(gaps), _ptr(scalars), _ptr(grad_out), float(c),
                                                float(w), float(mult), col_mode, _ptr(fin_ws), fin_bytes, _ptr(out),
                                                _DTYPE_CODE[out_dtype], st), "scl_bwd_finish")
        self.launches += 3 + (3 if col_mode != 0 else 0)
        return out
