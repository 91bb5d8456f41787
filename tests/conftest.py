import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(kind=None, world=None):
    idx = json.loads((GOLDEN / "index.json").read_text())
    out = []
    for name, meta in sorted(idx.items()):
        if kind is not None and meta["kind"] != kind:
            continue
        if world is not None and meta["world"] != world:
            continue
        out.append(name)
    return out


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    data = {k: z[k] for k in z.files if k != "meta"}
    return meta, data


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def text_ids_for(meta, batch):
    """Text-side tile ids of a fixture: the image-side ids unless the fixture was minted with different ones."""
    if "text_ids_seed" in meta:
        from spatial_clip_b200.synth import shuffled_text_ids

        return shuffled_text_ids(batch.tile_ids, meta["text_ids_seed"])
    return batch.tile_ids.clone()
