"""Builds examples/_build/c_abi_render against libwr_b200.so with the host compiler only (no nvcc needed for the
caller: the C ABI hides every CUDA kernel)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build", "c_abi_render")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c_abi_render.cpp")
    lib_dir = os.path.join(ROOT, "worldrenderer_b200", "lib")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", src, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", lib_dir, "-lwr_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
