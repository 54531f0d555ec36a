"""Latent data path on CPU: CustomDataset layout (dataset.py:18-85) and the double-buffered batch loader's host logic."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

REF_DATASET = "/root/reference/image/dataset.py"


def _make_tree(root, n=7, size=4, text_dir=None, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "images", "00000"))
    os.makedirs(os.path.join(root, "vae-sd", "00000"))
    if text_dir:
        os.makedirs(os.path.join(root, text_dir, "00000"))
    labels = []
    for i in range(n):
        name = f"00000/img{i:08d}"
        if i % 2:
            np.save(os.path.join(root, "images", name + ".npy"), rng.integers(0, 255, (3, 8, 8), dtype=np.uint8))
        else:
            import PIL.Image
            PIL.Image.fromarray(rng.integers(0, 255, (8, 8, 3), dtype=np.uint8)).save(os.path.join(root, "images", name + ".png"))
        np.save(os.path.join(root, "vae-sd", f"00000/img-mean-std-{i:08d}.npy"),
                rng.standard_normal((1, 8, size, size)).astype(np.float32))
        labels.append([f"00000/img-mean-std-{i:08d}.npy", int(rng.integers(0, 1000))])
        if text_dir:
            np.save(os.path.join(root, text_dir, name + ".npy"), rng.standard_normal((12,)).astype(np.float32))
    rng.shuffle(labels)                                   # the json is a lookup table, not an ordering
    with open(os.path.join(root, "vae-sd", "dataset.json"), "w") as f:
        json.dump({"labels": labels}, f)
    return root


def test_custom_dataset_items(tmp_path):
    from reed_b200.image.dataset import CustomDataset
    root = _make_tree(str(tmp_path), n=5)
    ds = CustomDataset(root)
    assert len(ds) == 5 and ds.labels.dtype == np.int64
    table = dict(json.load(open(os.path.join(root, "vae-sd", "dataset.json")))["labels"])
    for i in range(5):
        image, moments, label, text = ds[i]
        assert image.shape == (3, 8, 8) and image.dtype == torch.uint8
        assert moments.shape == (1, 8, 4, 4) and moments.dtype == torch.float32
        assert int(label) == table[ds.feature_fnames[i]]
        assert torch.equal(text, torch.zeros_like(moments))
    lean = CustomDataset(root, load_images=False)
    assert lean[2][0].numel() == 0 and torch.equal(lean[2][1], ds[2][1])
    with pytest.raises(AssertionError):
        CustomDataset(root, text_embeds_dir="text_embeds_missing")


@pytest.mark.skipif(not os.path.exists(REF_DATASET), reason="reference checkout not present")
def test_custom_dataset_matches_reference_class(tmp_path):
    """Same items as the reference's own CustomDataset on the same directory tree (images, moments, labels, text)."""
    from reed_b200.image.dataset import CustomDataset
    spec = importlib.util.spec_from_file_location("ref_dataset", REF_DATASET)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for text_dir in (None, "text_embeds_qwenvl"):
        root = _make_tree(str(tmp_path / f"t{text_dir}"), n=6, text_dir=text_dir, seed=3)
        a, b = ref.CustomDataset(root, text_embeds_dir=text_dir), CustomDataset(root, text_embeds_dir=text_dir)
        assert len(a) == len(b) and a.image_fnames == b.image_fnames and a.feature_fnames == b.feature_fnames
        assert np.array_equal(a.labels, b.labels)
        for i in range(len(a)):
            for x, y in zip(a[i], b[i]):
                assert x.dtype == y.dtype and torch.equal(x, y)


def test_batch_loader_shards_and_stages(tmp_path):
    from reed_b200.image.dataset import CustomDataset, LatentBatchLoader
    root = _make_tree(str(tmp_path), n=13, text_dir="text_embeds_qwenvl")
    ds = CustomDataset(root, text_embeds_dir="text_embeds_qwenvl", load_images=False)
    seen = []
    for rank in range(2):
        loader = LatentBatchLoader(ds, 3, "cpu", rank=rank, world=2, generator=torch.Generator().manual_seed(5), depth=2)
        assert len(loader) == 2                                   # 13 // 2 = 6 per rank, drop_last -> 2 batches of 3
        order = LatentBatchLoader(ds, 3, "cpu", rank=rank, world=2, generator=torch.Generator().manual_seed(5)).epoch_indices()
        got = []
        for moments, labels, text in loader:
            assert moments.shape == (3, 8, 4, 4) and labels.shape == (3,) and labels.dtype == torch.int64
            assert text.shape == (3, 12)
            got.append((moments.clone(), labels.clone(), text.clone()))    # slots are reused: clone before the next batch
        assert len(got) == 2
        flat = [i for i in order]
        for b, (moments, labels, text) in enumerate(got):
            for j in range(3):
                item = ds[flat[b * 3 + j]]
                assert torch.equal(moments[j], item[1][0]) and int(labels[j]) == int(item[2]) and torch.equal(text[j], item[3])
        seen.append(set(order))
    assert not (seen[0] & seen[1])                                # ranks read disjoint shards
    plain = LatentBatchLoader(ds, 4, "cpu", shuffle=False, with_text=False, depth=3)
    batches = [(m.clone(), y.clone(), t) for m, y, t in plain]
    assert len(batches) == 3 and batches[0][2] is None
    assert torch.equal(torch.cat([y for _, y, _ in batches]), torch.from_numpy(ds.labels[:12]))


def test_sample_posterior_refuses_cpu_tensors():
    from reed_b200.image.dataset import sample_posterior
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sample_posterior(torch.zeros(1, 8, 4, 4))


def test_default_shuffle_gives_disjoint_shards_and_leaves_the_global_generator_alone(tmp_path):
    """Without a caller generator every rank must still stride THE SAME permutation (seed + epoch), and the draw must not
    advance the global CPU generator SILoss takes its time draws from (each rank seeds that one differently, train.py:176)."""
    from reed_b200.image.dataset import CustomDataset, LatentBatchLoader
    root = _make_tree(str(tmp_path), n=20, text_dir=None)
    ds = CustomDataset(root, load_images=False)
    loaders = [LatentBatchLoader(ds, 2, "cpu", rank=r, world=4, seed=11) for r in range(4)]
    for epoch in range(3):
        shards = []
        for r, ld in enumerate(loaders):
            torch.manual_seed(100 + r)                            # per-rank global seed, as the reference does
            before = torch.get_rng_state()
            shards.append(ld.epoch_indices())
            assert torch.equal(torch.get_rng_state(), before)
        flat = [i for s in shards for i in s]
        assert len(flat) == len(set(flat)) == 16                  # 20 // 4 = 5 per rank -> 2 batches of 2
        if epoch == 0:
            first = shards
        else:
            assert shards != first                                # a fresh permutation per epoch
    again = LatentBatchLoader(ds, 2, "cpu", rank=2, world=4, seed=11).epoch_indices()
    assert again == first[2]
