"""Test-only CPU *logic emulation* build of the kernel sources.

Compiles scaml_capi.cu (and with it every .cuh kernel) with g++ and -DSCAML_EMU against csrc/emu/cuda_emu.h, a
minimal host re-implementation of the CUDA constructs the kernels use (thread blocks as fibres, shared memory,
warp shuffles, mma.sync fragments, cp.async).  The result, tests/_build/libscaml_emu.so, lets the GPU-less CI
container check tile layouts / indexing / host logic through the same C ABI.  It is NOT part of the product:
the package neither builds nor loads it (asserted by tests/test_capi_symbols.py), and nothing here is timed.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "scalable-meta-learning-with-gaussian-processes_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
EMU_LIB = os.path.join(OUT_DIR, "libscaml_emu.so")


def _sources():
    deps = [os.path.join(ROOT, "include", "scaml_b200.h"), os.path.join(CSRC, "emu", "cuda_emu.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    return deps


def build_emu(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(EMU_LIB):
        t = os.path.getmtime(EMU_LIB)
        if all(os.path.getmtime(d) <= t for d in _sources()):
            return EMU_LIB
    tmp = EMU_LIB + f".{os.getpid()}.tmp"  # several test processes may build at once: write, then rename
    extra = os.environ.get("SCAML_EMU_DEFS", "").split()  # A/B experiment switches, e.g. "-DSCAML_FIT_BLOCKINV"
    cmd = ["g++", "-std=c++17", "-O2", "-mfma", "-DSCAML_EMU"] + extra + ["-x", "c++", "-fPIC", "-shared", "-pthread",
           "-o", tmp, "scaml_capi.cu"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (emulation build) failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, EMU_LIB)
    return EMU_LIB


if __name__ == "__main__":
    print(build_emu(force=True))
