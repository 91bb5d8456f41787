"""Pin the CPU oracle against the golden vectors minted from the reference modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, text_ids_for
from oracle.contrastive_oracle import (clip_loss_oracle, dense_labels, soft_label_triples,
                                       spatial_loss_oracle)
from spatial_clip_b200.synth import make_spot_batch


def _inputs(meta):
    b = make_spot_batch(**meta["gen"])
    return (b.image_features.numpy(), b.text_features.numpy(), b.tile_ids.numpy(),
            b.neighbor_tile_ids.numpy(), b.neighbor_alphas.numpy())


def _text_ids(meta):
    return text_ids_for(meta, make_spot_batch(**meta["gen"])).numpy()


@pytest.mark.parametrize("name", golden_names("spatial"))
def test_spatial_oracle_matches_reference(name):
    meta, gold = load_golden(name)
    img, txt, ids, nbr, alpha = _inputs(meta)
    c = meta["ctor"]
    res = spatial_loss_oracle(img, txt, meta["scale"], ids, _text_ids(meta), nbr, alpha, world_size=meta["world"],
                              cap_logit_scale=c.get("cap_logit_scale"), temp_reg_weight=c.get("temp_reg_weight", 0.0),
                              neighbor_alpha_scale=c.get("neighbor_alpha_scale", 1.0),
                              local_loss=c["local_loss"], gather_with_grad=c["gather_with_grad"])
    loss = np.array([r.loss for r in res.ranks])
    ds = np.array([r.d_scale for r in res.ranks])
    # the reference ran in fp32; the oracle in fp64 -> tolerances are the reference's rounding
    np.testing.assert_allclose(loss, gold["loss"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(ds, gold["d_scale"], rtol=2e-4, atol=2e-6)
    scale_i = np.abs(gold["d_image"]).max()
    scale_t = np.abs(gold["d_text"]).max()
    assert np.abs(res.d_image - gold["d_image"]).max() <= 2e-5 * scale_i + 1e-9
    assert np.abs(res.d_text - gold["d_text"]).max() <= 2e-5 * scale_t + 1e-9


@pytest.mark.parametrize("name", [n for n in golden_names("spatial", world=1) if "bias" not in n and "legacy" not in n])
def test_soft_label_triples_bit_exact(name):
    meta, gold = load_golden(name)
    if "labels_i_t" not in gold:
        pytest.skip("no dense labels stored")
    _, _, ids, nbr, alpha = _inputs(meta)
    scale = meta["ctor"].get("neighbor_alpha_scale", 1.0)
    # the reference's labels BEFORE F.normalize(p=1): bit-exact fp32 equality.  Image rows resolve their neighbours in
    # the TEXT id map, text rows in the IMAGE id map (losses.py:92-93,102-108)
    for id_map, key in ((_text_ids(meta), "labels_i_t"), (ids, "labels_t_i")):
        rows, _ = soft_label_triples(id_map, nbr, alpha, scale, rank=0)
        dense = dense_labels(rows, len(ids))
        assert dense.dtype == np.float32
        assert np.array_equal(dense.view(np.uint32), gold[key].view(np.uint32)), key


def test_label_invariants_from_reference_notebook():
    """The three invariants of /root/reference/notebooks/test1_loss_test.ipynb (cells 1-3)."""
    b = make_spot_batch(n=48, d=64, k=8, seed=7)
    ids, nbr, alpha = b.tile_ids.numpy(), b.neighbor_tile_ids.numpy(), b.neighbor_alphas.numpy()
    world, bl = 3, 16
    for r in range(world):
        sl = slice(r * bl, (r + 1) * bl)
        rows, sums = soft_label_triples(ids, nbr[sl], alpha[sl], 1.0, rank=r)
        for i, lst in enumerate(rows):
            cols = dict(lst)
            # check 2: own column is rank*B_l + i with weight >= 1
            assert cols[r * bl + i] >= 1.0
            # check 1: normalised row sums to 1
            tot = sum(float(w) for w in cols.values())
            assert abs(tot / float(sums[i]) - 1.0) < 1e-6
            # check 3: absent neighbours contribute nothing
            present = set(ids.tolist())
            expected = 1.0 + sum(float(a) for nid, a in zip(nbr[sl][i], alpha[sl][i]) if a > 0 and int(nid) in present)
            assert abs(tot - expected) < 1e-5


@pytest.mark.parametrize("name", golden_names("clip"))
def test_clip_oracle_matches_reference(name):
    meta, gold = load_golden(name)
    img, txt, *_ = _inputs(meta)
    c = meta["ctor"]
    res = clip_loss_oracle(img, txt, meta["scale"], world_size=meta["world"], local_loss=c["local_loss"],
                           gather_with_grad=c["gather_with_grad"])
    loss = np.array([r.loss for r in res.ranks])
    ds = np.array([r.d_scale for r in res.ranks])
    np.testing.assert_allclose(loss, gold["loss"], rtol=5e-6, atol=1e-6)
    np.testing.assert_allclose(ds, gold["d_scale"], rtol=5e-4, atol=2e-6)  # fp32 reference
    # floor: the fp32 reference forms (p - 1) with p ~ 1, an absolute error of ~6e-8 * s * c per logit
    floor = 2e-7 * meta["scale"] / (len(img) // meta["world"])
    assert np.abs(res.d_image - gold["d_image"]).max() <= 2e-5 * np.abs(gold["d_image"]).max() + floor
    assert np.abs(res.d_text - gold["d_text"]).max() <= 2e-5 * np.abs(gold["d_text"]).max() + floor


def test_torch_port_matches_reference_goldens():
    """The CPU-baseline port (oracle/torch_port.py) reproduces the reference's numbers."""
    import torch

    from oracle.torch_port import clip_rank_step, spatial_rank_step

    meta, gold = load_golden("spatial_n256_w4")
    b = make_spot_batch(**meta["gen"])
    world, bl = 4, 64
    for r in range(world):
        sl = slice(r * bl, (r + 1) * bl)
        img_l = b.image_features[sl].clone().requires_grad_(True)
        txt_l = b.text_features[sl].clone().requires_grad_(True)
        s = torch.tensor(meta["scale"], requires_grad=True)
        loss = spatial_rank_step(img_l, txt_l, b.image_features, b.text_features, s, b.tile_ids,
                                 b.neighbor_tile_ids[sl], b.neighbor_alphas[sl], r)
        np.testing.assert_allclose(float(loss), gold["loss"][r], rtol=2e-6)
        np.testing.assert_allclose(float(s.grad), gold["d_scale"][r], rtol=1e-4, atol=1e-6)
    meta, gold = load_golden("clip_n128_w1_s14")
    b = make_spot_batch(**meta["gen"])
    img = b.image_features.clone().requires_grad_(True)
    txt = b.text_features.clone().requires_grad_(True)
    s = torch.tensor(meta["scale"], requires_grad=True)
    loss = clip_rank_step(img, txt, img, txt, s, 0)
    np.testing.assert_allclose(float(loss), gold["loss"][0], rtol=2e-6)
    assert np.abs(img.grad.numpy() - gold["d_image"]).max() <= 1e-5 * np.abs(gold["d_image"]).max()


# ---------------------------------------------------------------- the blockwise (large-N) oracle
@pytest.mark.parametrize("name", [n for n in golden_names("spatial") if "legacy" not in n])
def test_blockwise_oracle_matches_reference_goldens(name):
    """oracle/blockwise_oracle.py (the checker of the BASELINE-size GPU tests and of bench.py's parity block) against
    the reference's own outputs: per-rank loss / d_scale and a sample of gradient rows, every flag combination."""
    from oracle.blockwise_oracle import blockwise_oracle, sample_rows_for

    meta, gold = load_golden(name)
    img, txt, ids, nbr, alpha = _inputs(meta)
    c = meta["ctor"]
    world = meta["world"]
    rows = sample_rows_for(img.shape[0], world, 5, seed=3)
    res = blockwise_oracle(img, txt, meta["scale"], ids, _text_ids(meta), nbr, alpha, world_size=world,
                           cap_logit_scale=c.get("cap_logit_scale"), temp_reg_weight=c.get("temp_reg_weight", 0.0),
                           neighbor_alpha_scale=c.get("neighbor_alpha_scale", 1.0), local_loss=c["local_loss"],
                           gather_with_grad=c["gather_with_grad"], sample_rows=rows, block=48)
    np.testing.assert_allclose(res.loss, gold["loss"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(res.d_scale, gold["d_scale"], rtol=2e-4, atol=2e-6)
    assert np.abs(res.d_image_rows - gold["d_image"][rows]).max() <= 2e-5 * np.abs(gold["d_image"]).max() + 1e-9
    assert np.abs(res.d_text_rows - gold["d_text"][rows]).max() <= 2e-5 * np.abs(gold["d_text"]).max() + 1e-9


@pytest.mark.parametrize("name", [n for n in golden_names("clip") if "ll1" in n or "_w1_" in n or "n300" in n])
def test_blockwise_oracle_clip_matches_reference_goldens(name):
    from oracle.blockwise_oracle import blockwise_oracle, sample_rows_for

    meta, gold = load_golden(name)
    img, txt, _, _, _ = _inputs(meta)
    c = meta["ctor"]
    world = meta["world"]
    rows = sample_rows_for(img.shape[0], world, 5, seed=4)
    res = blockwise_oracle(img, txt, meta["scale"], None, None, None, None, world_size=world, local_loss=True,
                           gather_with_grad=c.get("gather_with_grad", False), sample_rows=rows, block=40, kind="clip")
    np.testing.assert_allclose(res.loss, gold["loss"], rtol=2e-6, atol=2e-6)
    assert np.abs(res.d_image_rows - gold["d_image"][rows]).max() <= 2e-5 * np.abs(gold["d_image"]).max() + 1e-9
    assert np.abs(res.d_text_rows - gold["d_text"][rows]).max() <= 2e-5 * np.abs(gold["d_text"]).max() + 1e-9
