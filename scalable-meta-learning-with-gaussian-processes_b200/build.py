"""In-tree builds of the native code (explicit nvcc / g++ command lines, no JIT cache).

build_cuda(): csrc/*.cu -> csrc/libscaml_b200.so for sm_100a (the product library).
build_emu():  the same kernel sources with -DSCAML_EMU via g++ -> csrc/libscaml_emu.so,
              a CPU *logic emulation* used only by tests/ (see csrc/emu/cuda_emu.h).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CUDA_LIB = os.path.join(CSRC, "libscaml_b200.so")
EMU_LIB = os.path.join(CSRC, "libscaml_emu.so")

CUDA_SOURCES = ["scaml_capi.cu", "scaml_microbench.cu"]
HEADERS = ["scaml_device.cuh", "scaml_fit.cuh", "scaml_fit8.cuh", "scaml_kmat.cuh", "scaml_predict.cuh", "scaml_cond.cuh", "scaml_cross.cuh", "scaml_target.cuh", "scaml_lbfgs.cuh", "scaml_grad.cuh", "scaml_gradval.cuh", "scaml_tile256.cuh",
           os.path.join("emu", "cuda_emu.h"), os.path.join("..", "..", "include", "scaml_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HEADERS]
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", CUDA_LIB] + CUDA_SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return CUDA_LIB


def build_prof(force: bool = False) -> str:
    """Diagnostics variant with per-phase clock64 counters (-DSCAML_PROF); not used by the product path."""
    out = os.path.join(CSRC, "libscaml_b200_prof.so")
    deps = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HEADERS]
    if not force and _newer(out, deps):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DSCAML_PROF", "-o", out] + CUDA_SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (prof build) failed:\n" + res.stdout + res.stderr)
    return out


def build_ablate(force: bool = False) -> str:
    """Timing-only ablation variant (-DSCAML_ABLATE, wrong results by design); experiments only."""
    out = os.path.join(CSRC, "libscaml_b200_abl.so")
    deps = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HEADERS]
    if not force and _newer(out, deps):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DSCAML_ABLATE", "-o", out] + CUDA_SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (ablation build) failed:\n" + res.stdout + res.stderr)
    return out


def build_emu(force: bool = False) -> str:
    deps = [os.path.join(CSRC, s) for s in ["scaml_capi.cu"] + HEADERS]
    if not force and _newer(EMU_LIB, deps):
        return EMU_LIB
    cmd = ["g++", "-std=c++17", "-O2", "-mfma", "-DSCAML_EMU", "-x", "c++", "-fPIC", "-shared", "-pthread",
           "-o", EMU_LIB, "scaml_capi.cu"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (emulation build) failed:\n" + res.stdout + res.stderr)
    return EMU_LIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--emu" in sys.argv:
        print(build_emu(force=True))
