"""Build libvgpt_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m videogpt_b200.build [--force]

The library is plain C ABI (include/vgpt_b200.h), statically links the CUDA runtime and resolves
the one driver symbol it needs (cuTensorMapEncodeTiled) at run time, so it builds on the
driver-less build container and travels to the GPU box as a single .so.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvgpt_b200.so")
SOURCES = ["runtime.cu", "gemm_pair_tcgen05.cu", "attention.cu", "attention_pair_tcgen05.cu", "elementwise.cu", "peer.cu", "api.cu"]
# descriptor / issue-rate probes (tests/test_umma_layouts.py, tools/umma_rate.py): their own library, not the product's
PROBE_LIB = os.path.join(HERE, "libvgpt_b200_probe.so")
PROBE_SOURCES = ["runtime.cu", "probe/umma_probe.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(PROBE_LIB):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(PROBE_LIB))
    deps = [os.path.join(d, f) for d, _, fs in os.walk(CSRC) for f in fs]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "vgpt_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in list(dict.fromkeys(SOURCES + PROBE_SOURCES)):
        obj = os.path.join(objdir, src.replace("/", "_").replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = {}
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
        objs[src] = obj
    for lib, sources in ((LIB, SOURCES), (PROBE_LIB, PROBE_SOURCES)):
        cmd = [nvcc, "-shared", "-o", lib, *[objs[s] for s in sources], "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
