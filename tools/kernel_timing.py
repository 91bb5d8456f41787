#!/usr/bin/env python
"""Developer tool: per-role wait-cycle breakdown of the CTA-pair kernels at the bench shape (needs a B200)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from spatial_clip_b200._cuda import CudaOps  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = int(sys.argv[2]) if len(sys.argv) > 2 else n
d = 512
ops = CudaOps()
ops.variant = 1
g = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(m, d, generator=g), dim=-1).cuda().bfloat16()
y = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda().bfloat16()
scal = ops.prep_scalars(torch.tensor([40.0], device="cuda"), None)
for _ in range(2):
    ops.fwd_rowstats(x, y, scal)
ops.cycle_buffers = {}
ops.kernel_events = {}
part, plan = ops.fwd_rowstats(x, y, scal)
col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
q = torch.zeros((m, 1), device="cuda")
stats = ops.row_finalize(part, plan, x, y, col, q)
cstats = torch.zeros(n, 4, device="cuda")
cstats[:, 0] = 60.0
_, y_t = ops.cast_bf16(y, want_rows=False, want_t=True, ld_t=n)
gaps = torch.zeros(1, device="cuda")
go = torch.ones(1, device="cuda")
ocol = torch.full((n, 1), -1, dtype=torch.int32, device="cuda")
oq = torch.zeros((n, 1), device="cuda")
ops.bwd_rows(x, y, y_t, stats, cstats, col, q, ocol, oq, max(m, n), 0, gaps, scal, go, 0.5 / m, 0.0, 1.0, 2,
             torch.float32, opp_q_local=torch.zeros((m, 1), device="cuda"))
torch.cuda.synchronize()
for name, evs in ops.kernel_events.items():
    print(name, "ms:", [round(a.elapsed_time(b), 3) for a, b in evs])


def report(name, labels):
    buf = ops.cycle_buffers[name][-1].view(-1, 16).cpu()
    used = buf[(buf != 0).any(dim=1)]
    lead = used[used[:, labels["_lead_col"]] != 0]
    peer = used[used[:, labels["_lead_col"]] == 0]
    print(f"== {name}: {len(used)} CTAs ({len(lead)} leaders)")
    for k, i in labels.items():
        if k.startswith("_"):
            continue
        src = lead
        print(f"   {k:34s} mean {src[:, i].double().mean():12.0f}  max {src[:, i].max():10d}")
    if len(peer):
        print("   peer producer life/empty-wait     ", peer[:, 0].double().mean(), peer[:, 1].double().mean())


report("fwd", {"_lead_col": 2, "producer lifetime": 0, "producer wait empty": 1, "mma lifetime": 2, "mma wait a_full": 3,
               "mma wait tmem_empty": 4, "mma wait full": 5, "epi(w4) lifetime": 6, "epi wait tmem_full": 7, "tiles": 8})
report("bwd", {"_lead_col": 3, "producer lifetime": 0, "producer wait empty": 1, "producer wait coef_empty": 2,
               "mma lifetime": 3, "mma wait x_full": 4, "mma wait tmem_empty": 5, "mma wait full(z)": 6,
               "mma wait g_full": 7, "mma wait full(yT)": 8, "epi(w4) lifetime": 9, "epi wait coef_full": 10,
               "epi wait tmem_full": 11, "epi wait g_empty": 12, "steps": 13})
