"""The C-ABI library loads and exports every symbol include/scl_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from spatial_clip_b200 import build
    from spatial_clip_b200._cuda import load_library

    build.build()
    return load_library()


def _declared():
    text = (ROOT / "include/scl_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scl_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from spatial_clip_b200._cuda import EXPORTS

    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in scl_b200.h but not exported"
    assert set(names) == set(EXPORTS), "ctypes table and header disagree"


def test_plans_and_errors_without_gpu(lib):
    from spatial_clip_b200._cuda import SclPlan

    assert lib.scl_abi_version() == 7
    p = SclPlan()
    assert lib.scl_fwd_plan(4096, 32768, 512, ctypes.byref(p)) == 0
    assert p.m_pad == 4096 and p.n_slots == 4 * p.chunks and p.chunks * p.tiles_per_chunk >= 32768 // 256
    assert lib.scl_bwd_plan(300, 300, 512, 0, ctypes.byref(p)) == 0
    assert p.m_pad == 384 and p.n_pad == 512 and p.d_split == 1 and p.split == 0
    assert lib.scl_fwd_plan(128, 128, 96, ctypes.byref(p)) == -2  # D % 64 != 0
    assert lib.scl_fwd_plan(128, 128, 1024, ctypes.byref(p)) == 0
    for d, slices in ((640, 2), (768, 2), (1024, 2), (1152, 3), (1280, 4), (1536, 3)):
        assert lib.scl_bwd_plan(128, 128, d, 0, ctypes.byref(p)) == 0 and p.d_split == slices, d
    assert lib.scl_bwd_plan(128, 128, 1600, 0, ctypes.byref(p)) == -2  # beyond 1536
    assert lib.scl_fwd_plan(128, 128, 3 * 1536, ctypes.byref(p)) == 0  # K-concatenated fp32-mode operands
    assert lib.scl_fwd_plan(128, 128, 3 * 1536 + 64, ctypes.byref(p)) == -2
    assert lib.scl_bwd_plan(300, 300, 512, 1, ctypes.byref(p)) == 0 and p.split == 1 and p.d_split == 1
    assert b"unsupported shape" in lib.scl_error_string(-2)
    assert lib.scl_fwd_plan(300, 300, 512, ctypes.byref(p)) == 0 and p.m_pad == 512
    assert lib.scl_positives_workspace_bytes(1000) >= 2048 * 12
    # the backward's chunk picker charges every chunk's partial slab: few chunks for a rank of an 8-GPU job
    assert lib.scl_bwd_plan(4096, 32768, 512, 0, ctypes.byref(p)) == 0 and p.chunks <= 5
    assert lib.scl_bwd_workspace_bytes(4096, 32768, 512, 9) > 0 and lib.scl_bwd_finish_workspace_bytes(32768, 4096, 9) > 0


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of every argument struct as gcc sees include/scl_b200.h == the ctypes mirror in _cuda.py."""
    import shutil
    import subprocess

    from spatial_clip_b200 import _cuda

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    mirrors = {"scl_plan": _cuda.SclPlan, "scl_prepare_args": _cuda.PrepareArgs, "scl_fwd_args": _cuda.FwdArgs,
               "scl_bwd_args": _cuda.BwdArgs}
    lines = []
    for cname, cls in mirrors.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void) {\n%s\nreturn 0;\n}\n'
                   % (ROOT / "include/scl_b200.h", "\n".join(lines)))
    exe = tmp_path / "layout"
    subprocess.run([gcc, str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
