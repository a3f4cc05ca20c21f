"""Builds libbcmpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbcmpc.so")
SRC = os.path.join(HERE, "csrc", "api.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in
        ("api.cu", "common.cuh", "ptx.cuh", "rollout_verify.cuh", "rollout_tmem.cuh", "aux_kernels.cuh")]
DEPS.append(os.path.join(HERE, "..", "include", "bcmpc.h"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """nvcc cross-compiles without a GPU; the .so travels to the GPU box with the snapshot."""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
