"""Host-side checks of the in-tree extension that need no GPU: it imports (built by ``__graft_entry__.build()`` or on first
use), reports the architecture it was compiled for, exposes the run-time switches, and its host-side planning helpers
behave.  Kernel numerics live in the gpu-marked tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ext():
    from fluxdist import ops

    try:
        return ops.extension()
    except RuntimeError as e:  # no nvcc on this machine and no prebuilt .so
        pytest.skip(f"extension not available: {e}")


def test_extension_reports_sm100a_and_counts_launches(ext):
    assert ext.arch == "sm_100a"
    assert ext.launch_count() >= 0


def test_pdl_mask_switch_round_trips(ext):
    """FLUXDIST_PDL / set_pdl: bit mask of kernel classes launched as programmatic dependents (csrc/common.cuh)."""
    old = ext.pdl_mask()
    try:
        ext.set_pdl(0x05)
        assert ext.pdl_mask() == 0x05
        ext.set_pdl(0)
        assert ext.pdl_mask() == 0
    finally:
        ext.set_pdl(old)
    assert ext.pdl_mask() == old


def test_pdl_default_mask_is_the_tensor_core_classes():
    """A fresh process without FLUXDIST_PDL: igemm fprop/dgrad, igemm wgrad and halo fprop/dgrad (bits 0-2); with the
    variable set, its value (any base) wins."""
    code = "from fluxdist import ops; print(ops.extension().pdl_mask())"
    env = {k: v for k, v in os.environ.items() if k != "FLUXDIST_PDL"}
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    if out.returncode != 0:
        pytest.skip(f"extension not importable in a subprocess: {out.stderr[-300:]}")
    assert out.stdout.strip().splitlines()[-1] == "7"
    env["FLUXDIST_PDL"] = "0x4f"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.stdout.strip().splitlines()[-1] == str(0x4F)


def test_sass_contains_the_blackwell_instructions():
    """The built .so really carries tcgen05 / TMA / multimem / PDL code for sm_100a (cuobjdump runs without a GPU)."""
    import glob
    import shutil

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    so = glob.glob(os.path.join(ROOT, "fluxdistributed.jl_b200", "_C.*.so"))
    if not so:
        pytest.skip("extension not built")
    sass = subprocess.run(["cuobjdump", "-sass", so[0]], capture_output=True, text=True, timeout=600).stdout
    assert "EF_CUDA_SM100" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "LDGMC", "ACQBULK", "PREEXIT"):
        assert mnemonic in sass, mnemonic
