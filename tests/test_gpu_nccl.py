"""One process per GPU over NCCL against the reference's multi-rank goldens: every local_loss / gather_with_grad
combination of ClipLoss (reference: src/open_clip/loss.py:21-65, :120-155) and the SpatialLoss fixtures
(src/models/components/losses.py:73-122), five steps per module so that the eager route (steps 0-1) and the CUDA-graph
replay with its captured all-gathers (steps 2-4) are both compared.  The two-rank fixtures only (4 / 8 ranks over
NCCL are covered by bench.py's parity block, profiles/r2_bench_{4,8}gpu.json); skipped on a one-GPU box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`).
"""
import socket

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, text_ids_for
from spatial_clip_b200.synth import make_spot_batch

pytestmark = pytest.mark.gpu

STEPS = 5


def _nccl_worker(rank, world, port, name, q):
    import os
    import traceback

    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from spatial_clip_b200 import ClipLoss, SpatialLoss, losses, release_cuda_graphs

        meta, _ = load_golden(name)
        full = make_spot_batch(**meta["gen"])
        b = full.rank_slice(rank, world)
        bl = b.tile_ids.shape[0]
        txt_ids = text_ids_for(meta, full)[rank * bl:(rank + 1) * bl].cuda()
        c = dict(meta["ctor"])
        if meta["kind"] == "spatial":
            c.pop("cache_labels", None)
            mod = SpatialLoss(**c)
        else:
            mod = ClipLoss(**c)
        assert (mod.rank, mod.world_size) == (rank, world)
        ids, nbr, alpha = b.tile_ids.cuda(), b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda()
        res = []
        for _ in range(STEPS):
            img = b.image_features.cuda().requires_grad_(True)
            txt = b.text_features.cuda().requires_grad_(True)
            s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
            out = mod(img, txt, s, ids, txt_ids, nbr, alpha) if meta["kind"] == "spatial" else mod(img, txt, s)
            loss = out["contrastive_loss"]
            loss.backward()
            torch.cuda.synchronize()
            res.append((float(loss.detach()), img.grad.cpu().numpy(), txt.grad.cpu().numpy(), float(s.grad)))
            del out, loss
        graphed = any(st.fwd is not None and st.bwd is not None for st in losses._GRAPHS.values())
        q.put((rank, res, graphed))
        release_cuda_graphs()  # captured NCCL work must be gone before the communicator is
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # surface the traceback in the parent instead of hanging the other ranks
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("name", golden_names(world=2))
def test_nccl_ranks_match_reference(name):
    import torch.multiprocessing as mp

    meta, gold = load_golden(name)
    world = meta["world"]
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, name, q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    got = []
    try:
        for _ in range(world):
            item = q.get(timeout=150)
            assert len(item) == 3, f"rank {item[0]} raised:\n{item[1]}"
            got.append(item)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    bl = meta["gen"]["n"] // world
    for rank, steps, graphed in got:
        assert graphed, "the replay route was not reached"
        sl = slice(rank * bl, (rank + 1) * bl)
        for k, (loss, gi, gt, ds) in enumerate(steps):
            rep = (name, rank, k, loss, gold["loss"][rank], ds, gold["d_scale"][rank],
                   np.abs(gi - gold["d_image"][sl]).max() / np.abs(gold["d_image"][sl]).max(),
                   np.abs(gt - gold["d_text"][sl]).max() / np.abs(gold["d_text"][sl]).max())
            # vs the reference's fp32 goldens (inputs NOT pre-rounded to bf16): bf16-mode tolerances
            assert abs(loss - gold["loss"][rank]) <= 1e-3 * abs(gold["loss"][rank]) + 2e-6 * meta["scale"], rep
            assert abs(ds - gold["d_scale"][rank]) <= 3e-2 * abs(gold["d_scale"][rank]) + 1e-5, rep
            assert rep[7] <= 3e-2 and rep[8] <= 3e-2, rep
            # same inputs every step: eager and replayed steps must agree bit for bit
            assert loss == steps[0][0] and ds == steps[0][3], rep
            assert np.array_equal(gi, steps[0][1]) and np.array_equal(gt, steps[0][2]), rep
    for p in procs:
        assert p.exitcode == 0, "a rank did not shut down cleanly"
