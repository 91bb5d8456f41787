#!/usr/bin/env python
"""Compile the reference's own loss modules into ``oracle/_ref/`` (bytecode only).  TEST / BENCH INFRASTRUCTURE.

The reference (Biogod2020/Spatial-Clip) is pure Python, so the analogue of "compile the reference's few source files
into oracle/_ref/*.so" is ``py_compile``: the three files on the hot path are compiled WHERE THEY LIE under
/root/reference into bytecode files (``.bin``: a ``*.pyc`` pattern is commonly excluded when trees are copied) under ``oracle/_ref/`` (git-ignored, not gpurun-ignored: the bytecode travels to
the GPU box like our own built ``.so``; no reference source is copied into the repository).  ``oracle/ref_loader.py``
imports them behind a stub ``open_clip`` package (the real ``open_clip/__init__`` needs ftfy / timm, absent here).

    python oracle/make_ref.py          # run in the build container; __graft_entry__.build() calls it too

Used by: ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference") and
``tests/test_reference_arm.py``.  Never imported by the product.
"""
from __future__ import annotations

import py_compile
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
# module name in oracle/_ref -> reference source (the files SURVEY.md 8a cites for the path)
SOURCES = {
    "open_clip_loss": "src/open_clip/loss.py",
    "components_losses": "src/models/components/losses.py",
    "legacy_spatial_loss": "src/open_clip_train/spatial_loss.py",
}


def make_ref(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref/ holds the compiled reference modules afterwards."""
    if not REF.exists():
        return all((OUT / f"{name}.bin").exists() for name in SOURCES)
    OUT.mkdir(exist_ok=True)
    for name, rel in SOURCES.items():
        src = REF / rel
        py_compile.compile(str(src), cfile=str(OUT / f"{name}.bin"), dfile=f"<reference>/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print(f"compiled {src} -> {OUT / (name + '.bin')}")
    (OUT / "PYTHON_VERSION").write_text("%d.%d\n" % sys.version_info[:2])
    return True


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    print("oracle/_ref ready" if ok else "no /root/reference here and no prebuilt oracle/_ref")
