"""Builds libbcmpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a: one object per translation unit,
compiled in parallel (the tcgen05 kernels have ~200 instantiations), then one link.  Only the translation units whose
source or (transitively) included headers changed are recompiled."""
import glob
import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbcmpc.so")
OBJ = os.path.join(HERE, "_obj")
NVCC_BASE = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
NVCC_EXTRA = os.environ.get("BCMPC_NVCC_EXTRA", "").split()
# -lineinfo (ncu source pages) on the translation units whose kernels are profiled - the C ABI, the bench workload's kernel
# (variant 0), the policy-guided one (variant 1), the narrow-net, cluster-split and two-tile kernels; the other variants and
# every debug flavour (tests / bcmpc_rollout_debug only) go without it: line tables triple an object's size and the library is
# pushed to the GPU box on every run (100 MB -> ~65 MB).  BCMPC_LINEINFO=all / none overrides.
LINEINFO = os.environ.get("BCMPC_LINEINFO", "production")
LINEINFO_TUS = {"api.cu", "tm_var0.cu", "tm_var1.cu", "tm_narrow.cu", "tm_deep.cu", "tm_duo.cu"}


def nvcc_flags(src):
    li = LINEINFO == "all" or (LINEINFO == "production" and os.path.basename(src) in LINEINFO_TUS)
    return NVCC_BASE + (["-lineinfo"] if li else []) + NVCC_EXTRA


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def includes(path, seen=None):
    """Transitive closure of the quoted includes of `path` (project headers only)."""
    seen = set() if seen is None else seen
    try:
        txt = open(path).read()
    except OSError:
        return seen
    for inc in _INC.findall(txt):
        p = os.path.normpath(os.path.join(os.path.dirname(path), inc))
        if p not in seen and os.path.exists(p):
            seen.add(p)
            includes(p, seen)
    return seen


def deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "bcmpc.h")]


def obj_of(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def flags_stamp():
    return " ".join(NVCC_BASE + NVCC_EXTRA + [LINEINFO] + sorted(LINEINFO_TUS))


def obj_stale(src):
    obj = obj_of(src)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src, *includes(src)])


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
    stamp = os.path.join(OBJ, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != flags_stamp():
        force = True                       # different flags: every object is stale
    extra = ["-Xptxas=-v"] if verbose else []
    todo = [s for s in sources() if force or obj_stale(s)]
    live = {obj_of(s) for s in sources()}
    for o in glob.glob(os.path.join(OBJ, "*.o")):       # objects of deleted translation units must not be linked
        if o not in live:
            os.remove(o)

    def compile_one(src):
        r = subprocess.run([nvcc, *nvcc_flags(src), *extra, "-c", "-o", obj_of(src), src], cwd=HERE, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, todo))
    for src, r in results:
        if verbose or r.returncode:
            print(f"== {os.path.basename(src)}\n{r.stdout}{r.stderr}")
        if r.returncode:
            if os.path.exists(obj_of(src)):
                os.remove(obj_of(src))
            raise RuntimeError(f"nvcc failed on {src}")
    open(stamp, "w").write(flags_stamp())
    subprocess.run([nvcc, *NVCC_BASE, "-shared", "-o", LIB, *sorted(live)], check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--incremental" not in sys.argv, verbose="--quiet" not in sys.argv))
