"""The runnable recipes under examples/ (the reference's README / bin/driver.jl usage) work on CPU fake devices."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_task_mode_example_tiny():
    reps = _load("task_mode").main(["--tiny", "--steps", "2", "--devices", "2"])
    assert len(reps) == 2


def test_process_mode_example_tiny_threads():
    models, states = _load("process_mode").main(["--tiny", "--workers", "2", "--no-spawn", "--nsamples", "32"])
    assert len(models) == 2 and len(states) == 2
