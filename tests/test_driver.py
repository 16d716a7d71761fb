import torch

from fluxdist.bin import driver, infer_demo


def test_task_driver_runs_on_fake_devices(capsys):
    rc = driver.main(["--mode", "task", "--model", "tiny", "--dataset", "synthetic", "--classes", "10", "--devices", "2",
                      "--batch", "4", "--cycle", "2", "--engine", "staged"])
    assert rc == 0
    assert "trained 2 replicas" in capsys.readouterr().out


def test_process_driver_threads(capsys):
    rc = driver.main(["--mode", "process", "--model", "tiny", "--dataset", "synthetic", "--classes", "10", "--devices", "2",
                      "--batch", "4", "--nsamples", "8", "--cycles", "1", "--no-spawn"])
    assert rc == 0
    assert "trained 2 workers" in capsys.readouterr().out


def test_rgba_bytes():
    buf = bytes(range(16)) * 4
    im = infer_demo.rgba_bytes_to_rgb(buf, 4, 4)
    assert im.shape == (4, 4, 3) and im[0, 0].tolist() == [0, 1, 2]
