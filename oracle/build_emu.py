"""Build oracle/_build/libvc_emu.so (TEST INFRASTRUCTURE ONLY): the product's host orchestration
(videocad_b200/csrc/model_*.cpp, c_api.cpp) linked against the plain-C++ kernel restatements in
oracle/csrc/kernels_cpu.cpp.  Used by the `-m "not gpu"` tests to check host logic without a GPU.
The product never loads this library (videocad_b200.lib refuses non-CUDA builds)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "videocad_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libvc_emu.so")

SOURCES = [
    os.path.join(HERE, "csrc", "kernels_cpu.cpp"),
    os.path.join(CSRC, "model_vit.cpp"),
    os.path.join(CSRC, "model_seq.cpp"),
    os.path.join(CSRC, "c_api.cpp"),
    os.path.join(CSRC, "host_util.cpp"),
]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = SOURCES + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [
        os.path.join(ROOT, "include", "videocad_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-DVC_CUDA_BUILD=0", "-I", os.path.join(ROOT, "include"),
           "-I", CSRC, "-o", OUT] + SOURCES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building the CPU emulation library")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
