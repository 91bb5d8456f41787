#!/usr/bin/env python
"""Developer lab (not part of the product): which resource bounds bwd_rows_pair_kernel?

    python tools/bwd_lab.py build     # here: compile variant libraries into tools/lab/ (they travel with gpurun)
    python tools/bwd_lab.py run       # on the GPU box: time every variant at N = 32768, D = 512

Variants are compile-time switches of csrc/scl_bwd2.cu (SCL_LAB_* / SCL_B2_STAGES): without the epilogue math, without
the TMA reloads of the ring, with a shallower ring.  Results are wrong by construction in the NO_* variants; only the
time matters."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LAB = ROOT / "tools" / "lab"
sys.path.insert(0, str(ROOT))

VARIANTS = {
    "base": [],
    "no_fence": ["-DSCL_LAB_NO_FENCE"],
    "no_sts": ["-DSCL_LAB_NO_STS"],
    "no_epi": ["-DSCL_LAB_NO_EPI"],
}


def build():
    from spatial_clip_b200 import build as b

    LAB.mkdir(exist_ok=True)
    objdir = b.CSRC / "build"
    b.build()
    others = [str(objdir / (s + ".o")) for s in b.SOURCES if s != "scl_bwd2.cu"]
    procs = []
    for name, flags in VARIANTS.items():
        obj = LAB / f"bwd2_{name}.o"
        cmd = [b._nvcc(), *[f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c",
               str(b.CSRC / "scl_bwd2.cu"), "-o", str(obj)]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise SystemExit(f"{name}: {out}")
        lib = LAB / f"libscl_{name}.so"
        subprocess.run([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), str(obj),
                        *others], check=True)
        obj.unlink()
        print("built", lib)


def run_one(name, m=32768, n=32768, d=512, reps=5):
    import torch

    from spatial_clip_b200 import _cuda

    _cuda._LIB_PATH = LAB / f"libscl_{name}.so"
    ops = _cuda.CudaOps()
    g = torch.Generator().manual_seed(1)
    x = torch.nn.functional.normalize(torch.randn(m, d, generator=g), dim=-1).cuda().bfloat16()
    y = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda().bfloat16()
    scal = ops.prep_scalars(torch.tensor([40.0], device="cuda"), None)
    rs = torch.zeros(m, 4, device="cuda")
    rs[:, 0] = 60.0
    cs = torch.zeros(n, 4, device="cuda")
    cs[:, 0] = 60.0
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    gaps = torch.tensor([0.1], device="cuda")
    go = torch.tensor([1.0], device="cuda")
    ops.kernel_events = {}
    for _ in range(reps + 1):
        ops.bwd_rows(x, y, rs, cs, col, q, col, q, m, 0, gaps, scal, go, 0.5 / m, 0.05, 1.0, 2, torch.float32,
                     opp_q_local=q)
    torch.cuda.synchronize()
    ev = ops.kernel_events["bwd_rows"][1:]
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    plan = ops.bwd_plan(m, n, d)
    return {"variant": name, "m": m, "n": n, "ms_min": ms[0], "ms_med": ms[len(ms) // 2],
            "chunks": plan.chunks, "tflops_executed": 4.0 * m * n * d / (ms[len(ms) // 2] * 1e-3) / 1e12}


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "run":
        for name in VARIANTS:
            for shape in ((32768, 32768),):
                r = subprocess.run([sys.executable, __file__, "one", name, str(shape[0]), str(shape[1])],
                                   capture_output=True, text=True, timeout=300)
                print(r.stdout.strip() or r.stderr[-800:], flush=True)
    elif sys.argv[1] == "one":
        print(json.dumps(run_one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))))
