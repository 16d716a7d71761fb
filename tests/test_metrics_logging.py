import torch

import fluxdist as fd
from fluxdist.utils import MemoryLogger, kacc, maxk, showpreds, topk, topkaccuracy, with_logger


def test_topk_accuracy():
    out = torch.tensor([[0.1, 0.7, 0.2], [0.5, 0.3, 0.2], [0.2, 0.3, 0.5]])
    y = torch.tensor([1, 1, 0])
    assert abs(topkaccuracy(out, y, k=1) - 1 / 3) < 1e-6
    assert abs(topkaccuracy(out, y, k=2) - 2 / 3) < 1e-6
    assert topkaccuracy(out, fd.onehot(y, 3), k=3) == 1.0


def test_maxk_kacc():
    mk = maxk(torch.tensor([0.1, 0.9, 0.5]), 2)
    assert [i for i, _ in mk] == [1, 2]
    assert kacc([mk], [2]) == 1.0 and kacc([mk], [0]) == 0.0


def test_showpreds_marks_hit():
    out = torch.tensor([[0.1, 0.7, 0.2]])
    s = showpreds(topk(out, 2), [1], {0: "cat", 1: "dog", 2: "eel"})
    assert "dog  ✅" in s and "Sample 1" in s


def test_log_loss_and_acc_records():
    m = fd.tiny_convnet(10)
    x, y = torch.rand(8, 3, 32, 32), fd.onehot(torch.randint(0, 10, (8,)), 10)
    lg = MemoryLogger()
    with with_logger(lg):
        assert fd.log_loss_and_acc(fd.logitcrossentropy, (fd.FakeDevice(0), m), None) is None
        l, accs = fd.log_loss_and_acc(fd.logitcrossentropy, (fd.FakeDevice(0), m), (x, y), "val")
    names = [r["name"] for r in lg.records]
    assert "val_fake:0" in names and "val_1_fake:0" in names and "val_10_fake:0" in names
    assert set(accs) == {1, 5, 10}


def test_wandb_adapter_config_dict():
    lg = fd.WandbLogger(project="DDP", name="Resnet-34", config={"saveweights": False, "cycles": 100},
                        fallback=MemoryLogger())
    assert lg.get_config("cycles") == 100 and lg.get_config("missing") is None
    lg.update_config({"eta": 0.1})
    lg.info("x", a=1)
    assert lg.step == 1 and lg.config["eta"] == 0.1
