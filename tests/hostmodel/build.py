"""Builds tests/hostmodel/_build/libplsvo_hostmodel.so: the product's host code (pl-svo_b200/csrc/plsvo_abi.cu, compiled
unchanged as C++) + the model CUDA runtime + the digest kernels.  TEST INFRASTRUCTURE ONLY (fake_cuda.h)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libplsvo_hostmodel.so")
SOURCES = [os.path.join(ROOT, "pl-svo_b200", "csrc", "plsvo_abi.cu"), os.path.join(HERE, "fake_cudart.cpp"),
           os.path.join(HERE, "fake_kernels.cpp")]
DEPS = SOURCES + [os.path.join(HERE, "fake_cuda.h"), os.path.join(ROOT, "pl-svo_b200", "csrc", "internal.h"),
                  os.path.join(ROOT, "include", "plsvo_b200.h")]


def cuda_include() -> str:
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    raise RuntimeError("cuda_runtime.h not found (the host model compiles against the real CUDA headers)")


def build(force: bool = False, abi_source: str | None = None, out: str | None = None) -> str:
    """abi_source / out: build a variant from another copy of plsvo_abi.cu (the seeded-fault tests mutate one)."""
    out = out or OUT
    sources = [abi_source or SOURCES[0]] + SOURCES[1:]
    deps = sources + DEPS[len(SOURCES):]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-I" + cuda_include(),
           "-I" + os.path.join(ROOT, "pl-svo_b200", "csrc"),  # a mutated copy lives elsewhere but includes "internal.h"
           "-x", "c++", *sources, "-o", out + ".tmp", "-lpthread", "-ldl",
           "-Wl,-Bsymbolic"]  # bind the model runtime inside the library even when a real libcudart is already loaded (torch)
    subprocess.run(cmd, check=True)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    print(build(force=True))
