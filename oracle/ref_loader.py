"""Import the reference's own loss modules from the bytecode in ``oracle/_ref/`` (see oracle/make_ref.py) and run
ONE rank of a W-rank emulation through them.  TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product.

The classes returned are the reference's, unmodified: ``SpatialLoss`` / ``ClipLoss`` of
src/models/components/losses.py, ``ClipLoss`` / ``gather_features`` of src/open_clip/loss.py and the legacy
``GlobalMappingMultiPositiveClipLoss`` of src/open_clip_train/spatial_loss.py.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
import types
from pathlib import Path
from unittest import mock

REF_DIR = Path(__file__).resolve().parent / "_ref"
_CACHE = None


def available() -> bool:
    ver = REF_DIR / "PYTHON_VERSION"
    return (REF_DIR / "components_losses.bin").exists() and ver.exists() and \
        ver.read_text().strip() == "%d.%d" % sys.version_info[:2]


def _load(name: str, file: str):
    loader = importlib.machinery.SourcelessFileLoader(name, str(REF_DIR / file))
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


def load_reference():
    """(open_clip.loss module, components.losses module, legacy spatial_loss module) or None when oracle/_ref is absent."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        return None
    had = sys.modules.get("open_clip")
    pkg = types.ModuleType("open_clip")  # stub package: the modules only need `from open_clip.loss import ...`
    pkg.__path__ = []
    sys.modules["open_clip"] = pkg
    try:
        oc_loss = _load("open_clip.loss", "open_clip_loss.bin")
        pkg.loss = oc_loss
        pkg.ClipLoss = oc_loss.ClipLoss
        comp = _load("scl_reference_components_losses", "components_losses.bin")
        legacy = _load("scl_reference_legacy_spatial_loss", "legacy_spatial_loss.bin")
    finally:
        if had is not None:
            sys.modules["open_clip"] = had
    _CACHE = (oc_loss, comp, legacy)
    return _CACHE


def reference_rank_step(img_l, txt_l, img_all, txt_all, scale, ids_all, ids_l, nbr_ids, nbr_alpha, rank, world, ctor):
    """The reference's SpatialLoss forward + backward for ONE rank of a `world`-rank job, single process.

    img_l / txt_l: this rank's [B_l, D] leaf tensors (requires_grad); img_all / txt_all / ids_all: what the
    all-gathers would deliver (rank-major).  ``gather_features`` and ``dist.all_gather`` are replaced by stand-ins
    that hand back those tensors with the reference's gradient routing for local_loss / gather_with_grad
    (src/open_clip/loss.py:49-61); everything else -- logits, the id dicts, the Python label loop, both soft
    cross-entropies, the temperature regulariser, autograd -- is the reference's own code."""
    import torch

    _, comp, _ = load_reference()
    b = img_l.shape[0]
    sl = slice(rank * b, (rank + 1) * b)

    def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0,
                        world_size=1, use_horovod=False):
        gi, gt = img_all.detach().clone(), txt_all.detach().clone()
        if gather_with_grad or not local_loss:
            gi = torch.cat([gi[:sl.start], image_features, gi[sl.stop:]], dim=0)
            gt = torch.cat([gt[:sl.start], text_features, gt[sl.stop:]], dim=0)
        return gi, gt

    def all_gather(out_list, t):
        for q in range(world):
            out_list[q] = ids_all[q * b:(q + 1) * b]

    mod = comp.SpatialLoss(rank=rank, world_size=world, **ctor)
    mod.rank, mod.world_size = rank, world  # the constructor prefers a live process group's values (losses.py:29-31)
    with mock.patch.object(comp, "gather_features", gather_features), \
            mock.patch.object(comp.dist, "all_gather", all_gather):
        out = mod(image_features=img_l, text_features=txt_l, logit_scale=scale, image_tile_ids=ids_l,
                  text_tile_ids=ids_l, neighbor_tile_ids=nbr_ids, neighbor_alphas=nbr_alpha)
    loss = out["contrastive_loss"]
    loss.backward()
    return loss.detach()
