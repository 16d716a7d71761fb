"""Op layer: Python entry points of the hand-written sm_100a kernels in ``csrc/``
plus their PyTorch reference implementations (the numerics oracle and the
CPU fallback).  SURVEY §2.5 maps each kernel to the reference call site it
replaces.

The extension is built in-tree (``_C*.so`` next to this package, see
``build.py``); on a CUDA box the native path *must* be present — callers that ask
for it get a loud error instead of a silent PyTorch fallback.
"""
from __future__ import annotations

import importlib
import os
from typing import Optional

import torch

_C = None
_C_ERR: Optional[BaseException] = None


def _load():
    global _C, _C_ERR
    if _C is not None or _C_ERR is not None:
        return _C
    try:
        _C = importlib.import_module("fluxdist._C")
    except ModuleNotFoundError as e:
        # the in-tree .so is missing (fresh checkout: it is git-ignored).  Build it once, under a file lock so that
        # the ranks of a torchrun job do not compile concurrently; FLUXDIST_AUTOBUILD=0 turns this off.
        _C, _C_ERR = _autobuild(e)
    except BaseException as e:  # noqa: BLE001 - recorded and re-raised by require_native
        _C_ERR = e
        _C = None
    return _C


def _autobuild(first_error: BaseException):
    if os.environ.get("FLUXDIST_AUTOBUILD", "1") == "0":
        return None, first_error
    try:
        import fcntl
        import importlib.util

        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location("fluxdist_build", os.path.join(here, "build.py"))
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        with open(os.path.join(here, ".build.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            try:
                if not os.path.exists(b.ext_path()):  # another rank may have built it while we waited
                    b.build()
            finally:
                fcntl.flock(lk, fcntl.LOCK_UN)
        importlib.invalidate_caches()
        return importlib.import_module("fluxdist._C"), None
    except BaseException as e:  # noqa: BLE001 - nvcc missing, compile error, ...: report both
        return None, RuntimeError(f"{first_error!r}; automatic in-tree build failed: {e!r}")


def native_available() -> bool:
    """True when the sm_100a extension is importable *and* a CUDA device exists."""
    if os.environ.get("FLUXDIST_DISABLE_NATIVE") == "1":
        return False
    return torch.cuda.is_available() and _load() is not None


def extension():
    """The loaded extension module (imports on CPU too, for build checks)."""
    m = _load()
    if m is None:
        raise RuntimeError(f"fluxdist native extension not built/loadable: {_C_ERR!r}; run "
                           f"`python -c 'import __graft_entry__ as g; g.build()'`")
    return m


def probe_extension():
    """Hardware-probe module (``_Cprobe``: descriptor-shift, MMA-rate and TMA-bandwidth probes).  Built next to the
    product extension, used by ``scripts/probe_*.py`` only."""
    return importlib.import_module("fluxdist._Cprobe")


def require_native() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("native sm_100a path requested but no CUDA device is visible")
    extension()


from .reference import (bn_apply_ref, bn_bwd_apply_ref, bn_bwd_ref, bn_relu_pool_ref, bn_stats_ref,  # noqa: E402
                        conv_bn_act_ref, conv_dgrad_ref, conv_fprop_ref, conv_wgrad_ref, softmax_ce_ref)
from .native import (bn_apply, bn_bwd, bn_bwd_apply, bn_bwd_fused, bn_fwd_fused, bn_relu_pool, bn_stats,  # noqa: E402
                     bn_workspace, conv3x3_dgrad, conv3x3_fprop, conv3x3_wgrad, conv_bn_act, conv_dgrad, conv_fprop,
                     conv_wgrad, fused_update_local, gap_linear, gemm_nn, gemm_nt, gemm_tn, maxpool2d, normalize_u8,
                     softmax_ce, stem_fprop, stem_s2d, stem_wgrad)
