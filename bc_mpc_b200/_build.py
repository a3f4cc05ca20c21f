"""Builds libbcmpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a: one object per translation unit,
compiled in parallel (the tcgen05 kernel has ~60 instantiations), then one link."""
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbcmpc.so")
OBJ = os.path.join(HERE, "_obj")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
NVCC_FLAGS += os.environ.get("BCMPC_NVCC_EXTRA", "").split()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "bcmpc.h")]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force=False, verbose=False):
    """nvcc cross-compiles without a GPU; the .so travels to the GPU box with the snapshot."""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas=-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([nvcc, *NVCC_FLAGS, *extra, "-c", "-o", obj, src], cwd=HERE, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    for src, _, r in results:
        if verbose or r.returncode:
            print(f"== {os.path.basename(src)}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.run([nvcc, *NVCC_FLAGS, "-shared", "-o", LIB, *[o for _, o, _ in results]], check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
