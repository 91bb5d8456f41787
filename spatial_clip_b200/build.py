"""In-tree build of libscl_b200.so (nvcc, sm_100a only).  No JIT cache: the .so sits next to this file so it
travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libscl_b200.so"
SOURCES = ["scl_api.cu", "scl_fwd2.cu", "scl_bwd2.cu", "scl_aux.cu", "scl_split.cu", "scl_rank.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libscl_b200.so")


STAMP = PKG / "libscl_b200.so.sha"  # digest of the sources + flags the .so was built from (travels with the .so)


def _source_digest() -> str:
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    deps = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))) + \
        [PKG.parent / "include/scl_b200.h"]
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def needs_build() -> bool:
    """Stale when the library is missing or was built from other sources.  Compared by content, not by mtime: a
    snapshot copied to another box keeps the bytes but not necessarily the order of the timestamps."""
    if not LIB.exists():
        return True
    if not STAMP.exists():  # library from before the stamp existed: fall back to timestamps
        t = LIB.stat().st_mtime
        deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
            [PKG.parent / "include/scl_b200.h"]
        return any(p.stat().st_mtime > t for p in deps)
    return STAMP.read_text().strip() != _source_digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []
    logs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        logs.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    STAMP.write_text(_source_digest() + "\n")
    (objdir / "ptxas.log").write_text("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
