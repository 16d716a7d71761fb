import os

import numpy as np
import pytest
import torch

import fluxdist as fd
from fluxdist import data as D
from fluxdist.parallel import prepare_training, shard_indices


def test_shard_indices_drop_remainder_and_disjoint():
    ix = shard_indices(103, 4, np.random.default_rng(0))
    assert [len(i) for i in ix] == [25] * 4
    allv = np.concatenate(ix)
    assert len(set(allv.tolist())) == 100 and allv.max() < 100
    for d, i in enumerate(ix):  # contiguous partitions, shuffled inside
        assert i.min() == d * 25 and i.max() == d * 25 + 24


def test_synthetic_tree_labels_and_solutions():
    tree = D.SyntheticTree(nclasses=20, ntrain=50, nval=10, size=(64, 80))
    ls = D.labels(tree)
    assert len(ls) == 20 and ls["label"][0] == "n00000001"
    key = D.train_solutions(tree, "LOC_train_solution.csv", classes=range(10))
    assert set(key["class_idx"]) <= set(range(10)) and len(key) > 0
    assert D.makepaths("n00000003_7", "train").endswith("train/n00000003/n00000003_7.JPEG")
    assert D.makepaths("ILSVRC2012_val_00000001", "val").endswith("val/ILSVRC2012_val_00000001.JPEG")


def test_minibatch_decodes_and_onehots():
    tree = D.SyntheticTree(nclasses=5, ntrain=20, nval=4, size=(300, 280))
    key = D.train_solutions(tree, "LOC_train_solution.csv", classes=range(5))
    x, y = D.minibatch(tree, key, nsamples=6, class_idx=range(5), generator=np.random.default_rng(0))
    assert x.shape == (6, 3, 224, 224) and x.dtype == torch.float32
    assert y.shape == (6, 5) and torch.all(y.sum(1) == 1)
    xu, yi = D.minibatch(tree, key, nsamples=2, class_idx=range(5), u8=True, onehot=False)
    assert xu.shape == (2, 224, 224, 3) and xu.dtype == torch.uint8 and yi.dtype == torch.int64


def test_preprocess_geometry():
    im = (np.random.rand(300, 500, 3) * 255).astype(np.uint8)
    r = D.resize_smallest_dimension(im, 256)
    assert min(r.shape[:2]) == 256
    c = D.center_crop(r, 224)
    assert c.shape == (224, 224, 3)
    t = D.preprocess(im)
    assert t.shape == (3, 224, 224)
    flat = D.preprocess(np.full((256, 256, 3), 124, np.uint8))
    assert abs(flat[0].mean().item() - (124 / 255 - 0.485) / 0.229) < 1e-3


def test_prefetch_loader_order_and_errors():
    dl = D.DataLoader(lambda n: n * 2, ([1, 2, 3, 4],), buffersize=2)
    assert list(dl) == [2, 4, 6, 8]
    assert len(dl) == 4

    def boom(n):
        raise ValueError("bad")

    with pytest.raises(ValueError):
        list(D.DataLoader(boom, ([1],), buffersize=1))


def test_batch_loader():
    x, y = torch.arange(10.0).view(10, 1), torch.arange(10)
    dl = D.DataLoader(lambda b: b, (x, y), batchsize=4)
    bs = list(dl)
    assert len(bs) == 2 and bs[0][0].shape == (4, 1) and len(dl) == 2


def test_staged_loader_cpu():
    def produce(hx, hy):
        hx.fill_(7)
        hy.fill_(1)

    sl = D.StagedLoader(produce, 3, "cpu", (2, 4), (2,), x_dtype=torch.uint8)
    out = [(x.clone(), y.clone()) for x, y in sl]
    assert len(out) == 3 and int(out[0][0].sum()) == 7 * 8
    assert sl.h2d_bytes == 2 * 4 + 2 * 8


def test_prepare_training_step_count_and_shards():
    key = D.SyntheticKey(1000, nclasses=10, shape=(3, 32, 32))
    m = fd.tiny_convnet(10)
    setup, buffer = prepare_training(m, key, fd.fake_devices(4), fd.Momentum(), 10, epochs=2, buffersize=2,
                                     engine="staged")
    assert setup.cycle == 1000 * 2 // 4 // 10
    assert len(setup.dls) == 4 and len(buffer) == 4 and all(len(s) == 250 for s in setup.shards)
    assert all(isinstance(v["conv"]["weight"], torch.Tensor) for v in buffer.values())
    setup2, _ = prepare_training(m, key, fd.fake_devices(2), fd.Momentum(), 10, engine="staged")
    assert setup2.cycle == 5  # default cycle (ddp_tasks.jl:252)


def test_registry_lookup():
    tree = D.dataset("imagenet_synthetic")
    assert isinstance(tree, D.SyntheticTree)
    remote = D.dataset("imagenet")  # S3 entry of Data.toml: a URL-backed tree; nothing is fetched until a blob is read
    assert type(remote).__name__ == "RemoteTree" and "jc-datasets" in remote.base_url
    D.register("tmpds", path="/nonexistent")
    with pytest.raises(FileNotFoundError):
        D.dataset("tmpds")


def test_assemble():
    a = D.assemble([torch.zeros(3, 4), torch.ones(3, 4)])
    assert a.shape == (2, 3, 4)


def test_dataset_helper_stubs_and_legacy_alias():
    import resnetimagenet

    assert resnetimagenet.prepare_training is fd.prepare_training
    tree = D.SyntheticTree(nclasses=5, ntrain=20, nval=6)
    v = D.valimgs(tree, classes=range(5))
    assert len(v) == 6 and D.ground_truth(v, v["ImageId"][0]) == int(v["class_idx"][0])
    assert D.make_fname("ILSVRC2012_val_00000001").endswith("val/ILSVRC2012_val_00000001.JPEG")
    assert len(D.testimgs(tree, 3)) == 3


def test_remote_tree_driver_fetches_and_caches(tmp_path):
    """The remote (S3 / URL) storage driver of the registry (reference Data.toml:14-26): blobs come over a URL once and
    from the local cache afterwards (file:// stands in for the object store: no network here)."""
    from fluxdist.data import registry as R

    store = tmp_path / "store" / "imagenet"
    (store / "sub").mkdir(parents=True)
    (store / "LOC_synset_mapping.txt").write_text("n01440764 tench\n")
    (store / "sub" / "blob.bin").write_bytes(b"\x01\x02\x03")
    R.register("remote_test", driver="S3", endpoint=f"file://{tmp_path / 'store'}", prefix="imagenet",
               cache=str(tmp_path / "cache"))
    tree = R.dataset("remote_test")
    assert isinstance(tree, R.RemoteTree)
    assert tree["LOC_synset_mapping.txt"] == b"n01440764 tench\n"
    assert tree.exists("sub/blob.bin") and not tree.exists("missing.bin")
    (store / "sub" / "blob.bin").unlink()                      # the "network" goes away: the cache still serves it
    assert tree["sub/blob.bin"] == b"\x01\x02\x03"
    with pytest.raises(FileNotFoundError):
        tree["never_there.bin"]


def test_getweights_downloads_into_cache(tmp_path, monkeypatch):
    import torch

    import importlib

    P = importlib.import_module("fluxdist.data.preprocess")
    src = tmp_path / "srv"
    src.mkdir()
    torch.save({"w": torch.arange(3)}, src / "tiny.pt")
    monkeypatch.setenv("FLUXDIST_WEIGHTS", str(tmp_path / "wcache"))
    with pytest.raises(FileNotFoundError):
        P.getweights("tiny.pt")                                  # nothing cached, no store configured
    p = P.getweights("tiny.pt", url=f"file://{src / 'tiny.pt'}")
    assert os.path.exists(p) and torch.equal(P.weights("tiny.pt")["w"], torch.arange(3))
