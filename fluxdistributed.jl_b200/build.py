"""In-tree build of the sm_100a extension (``fluxdistributed.jl_b200/_C*.so``).

Kernels live in ``csrc/*.cu`` and are compiled with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (they include no torch
headers, so each file compiles in seconds and cross-compiles without a GPU);
``csrc/bind.cpp`` is the only translation unit that sees torch/pybind and is
compiled by the host compiler.  A generated ninja file gives incremental,
parallel rebuilds.  The built ``.so`` is git-ignored but travels to the GPU box
with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
    "-DNDEBUG",
]


def _cuda_home() -> str:
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "bin", "nvcc")):
            return c
    raise RuntimeError("nvcc not found")


def ext_path(name: str = "_C") -> str:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(HERE, name + suffix)


def build(verbose: bool = False, force: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    cuda = _cuda_home()
    os.makedirs(BUILD, exist_ok=True)
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cpp = sorted(glob.glob(os.path.join(CSRC, "*.cpp")))
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", f"-I{CSRC}",
                                                   f"-I{cuda}/include"]
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}",
                 "-DTORCH_API_INCLUDE_EXTENSION_H", "-Wno-deprecated-declarations"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    out = ext_path()
    lines = [
        "ninja_required_version = 1.3",
        f"nvcc = {cuda}/bin/nvcc",
        "cxx = g++",
        f"nvflags = {' '.join(NVCC_FLAGS)} -I{CSRC}",
        f"cxxflags = {' '.join(cxx_flags)} {' '.join(inc)}",
        f"ldflags = -shared -L{torch_lib} -L{cuda}/lib64 -lc10 -ltorch_cpu -ltorch -ltorch_python -lc10_cuda "
        f"-ltorch_cuda -lcudart_static -lrt -lpthread -ldl -Wl,-rpath,{torch_lib}",
        "rule nvcc",
        "  command = $nvcc $nvflags -MD -MF $out.d -c $in -o $out > $out.log 2>&1 || (cat $out.log; exit 1)",
        "  depfile = $out.d",
        "  deps = gcc",
        "rule cxx",
        "  command = $cxx $cxxflags -DTORCH_EXTENSION_NAME=$modname -MMD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "rule link",
        "  command = $cxx $in $ldflags -o $out",
    ]
    objs = []
    for s in cu:
        o = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(o)
        lines.append(f"build {o}: nvcc {s}")
    for s in cpp:
        o = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(o)
        lines.append(f"build {o}: cxx {s}")
        lines.append("  modname = _C")
    lines.append(f"build {out}: link {' '.join(objs)}")
    # hardware probes (scripts/probe_*.py): their own module, never linked into the product extension
    pobjs = []
    pout = ext_path("_Cprobe")
    for s in sorted(glob.glob(os.path.join(CSRC, "probes", "*.cu"))):
        o = os.path.join(BUILD, "probe_" + os.path.basename(s) + ".o")
        pobjs.append(o)
        lines.append(f"build {o}: nvcc {s}")
    for s in sorted(glob.glob(os.path.join(CSRC, "probes", "*.cpp"))):
        o = os.path.join(BUILD, "probe_" + os.path.basename(s) + ".o")
        pobjs.append(o)
        lines.append(f"build {o}: cxx {s}")
        lines.append("  modname = _Cprobe")
    if pobjs:
        lines.append(f"build {pout}: link {' '.join(pobjs)}")
    lines.append(f"default {out}" + (f" {pout}" if pobjs else ""))
    with open(os.path.join(BUILD, "build.ninja"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if force:
        for o in objs + [out]:
            if os.path.exists(o):
                os.remove(o)
    cmd = [sys.executable, "-m", "ninja", "-C", BUILD] + (["-v"] if verbose else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-8000:] + "\n" + r.stderr[-4000:])
        raise RuntimeError("extension build failed")
    if verbose:
        print(r.stdout[-4000:])
    return out


def build_staging_mock(tsan: bool = False, force: bool = False) -> str:
    """Host-only build of the pinned staging ring over the mock CUDA runtime (``csrc/mock``): module ``_Cstaging_mock``
    for ``tests/test_staging_mock.py``.  g++ only — no nvcc, no torch.  ``tsan=True`` instruments it with
    ThreadSanitizer (``_Cstaging_mock_tsan``; run the test with ``LD_PRELOAD=$(gcc -print-file-name=libtsan.so)``,
    see scripts/sanitize.sh)."""
    import pybind11

    name = "_Cstaging_mock_tsan" if tsan else "_Cstaging_mock"
    out = ext_path(name)
    srcs = [os.path.join(CSRC, "staging.cpp"), os.path.join(CSRC, "mock", "staging_mock_bind.cpp")]
    deps = srcs + [os.path.join(CSRC, "mock", "staging_mock_cuda.h"), os.path.join(CSRC, "staging.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-DFD_STAGING_MOCK_CUDA",
           f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", f"-I{CSRC}"]
    if tsan:
        cmd += ["-fsanitize=thread", "-DFD_MOCK_MODNAME=_Cstaging_mock_tsan"]
    cmd += srcs + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("staging mock build failed:\n" + r.stderr[-4000:])
    return out


def ptxas_report() -> str:
    """Concatenate the per-file ``-Xptxas -v`` logs (registers / spills / smem)."""
    out = []
    for lg in sorted(glob.glob(os.path.join(BUILD, "*.cu.o.log"))):
        with open(lg) as f:
            out.append(f"== {os.path.basename(lg)}\n" + f.read())
    return "\n".join(out)


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="--force" in sys.argv)
    print("built", p)
    if "--ptxas" in sys.argv:
        print(ptxas_report())
