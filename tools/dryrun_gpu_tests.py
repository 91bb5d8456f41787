#!/usr/bin/env python
"""Developer tool: execute the Python side of GPU test files on the CPU.  `.cuda()` becomes a copy, the CUDA runtime
calls become no-ops and the compute backend is the TEST-ONLY emulated ops (tests/emulated_ops.py), so that fixtures,
argument plumbing and assertions of tests that cannot run here are at least exercised once before they cost GPU time.
Only tests that go through the module API (not raw kernel entry points) can pass this way, and not the multi-rank
ones (their spawned workers do not inherit the stand-ins): deselect those with -k "not multi_rank".

    python tools/dryrun_gpu_tests.py tests/test_gpu_positive_columns.py [-k expr]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from emulated_ops import EmulatedOps  # noqa: E402
from spatial_clip_b200 import _cuda, losses  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self.clone()
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.is_available = lambda: True
_real_tensor, _real_zeros, _real_full, _real_empty = torch.tensor, torch.zeros, torch.full, torch.empty


def _cpu(fn):
    def wrapped(*a, **k):
        k.pop("device", None)
        return fn(*a, **k)
    return wrapped


torch.tensor, torch.zeros, torch.full, torch.empty = map(_cpu, (_real_tensor, _real_zeros, _real_full, _real_empty))


class DryOps(EmulatedOps):
    def __init__(self):
        super().__init__(round_bf16=True)


_cuda.CudaOps = DryOps
losses._set_ops_for_testing(DryOps())

if __name__ == "__main__":  # (spawned workers re-import this module: they must not start pytest again)
    import pytest

    sys.exit(pytest.main(sys.argv[1:] + ["-q", "-m", "gpu", "-p", "no:cacheprovider"]))
