import threading
import time

import pytest
import torch

from fluxdist.parallel import Channel, RemoteChannel, StoreChannel, syncgrads, wait_all_ready


def test_capacity_one_backpressure():
    c = Channel(1)
    c.put(1)
    assert c.isready()
    with pytest.raises(Exception):
        c.put(2, timeout=0.05)
    assert c.take() == 1
    assert not c.isready()


def test_none_is_a_message():
    c = Channel(1)
    c.put(None)
    assert c.isready()
    assert c.take() is None


def test_remote_channel_roundtrip_tree():
    c = RemoteChannel(1)
    tree = {"a": {"w": torch.arange(4.0)}, "b": None}
    c.put(tree)
    time.sleep(0.05)
    assert c.isready()
    out = c.take(timeout=5)
    assert out["b"] is None and torch.equal(out["a"]["w"], torch.arange(4.0))


def test_store_channel():
    import torch.distributed as dist

    store = dist.HashStore()
    a, b = StoreChannel(store, "ch"), StoreChannel(store, "ch")
    assert not b.isready()
    a.put({"x": torch.ones(2)})
    assert b.isready()
    with pytest.raises(TimeoutError):
        a.put(1, timeout=0.05)  # capacity 1
    assert torch.equal(b.take()["x"], torch.ones(2))
    a.put(None)
    assert b.take() is None


def test_syncgrads_mean_and_abort():
    n = 3
    ip = [Channel(1) for _ in range(n)]
    op = [Channel(1) for _ in range(n)]
    t = threading.Thread(target=syncgrads, args=(ip, op), kwargs={"device": "cpu"}, daemon=True)
    t.start()
    for i, c in enumerate(ip):
        c.put({"w": torch.full((2,), float(i)), "pool": None})
    outs = [o.take(timeout=10) for o in op]
    for o in outs:
        assert torch.allclose(o["w"], torch.full((2,), 1.0))  # mean of 0,1,2 (Q1 fixed: /n not /4)
        assert o["pool"] is None
    for c in ip:
        c.put(None)
    t.join(timeout=10)
    assert not t.is_alive()


def test_wait_all_ready_timeout():
    assert not wait_all_ready([Channel(1)], timeout=0.05)
