"""Build the native parts in-tree (the .so files travel to the GPU box with the repo snapshot).

  scicone_b200/lib/libscicone_score.so   CUDA kernels + C ABI (include/scicone_score.h), sm_100a only
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "lib")
CSRC = os.path.join(HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC",
              "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (and there is no CPU fallback)")
    return exe


def build_score_lib(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libscicone_score.so")
    srcs = [os.path.join(CSRC, "scicone_score.cu")]
    deps = srcs + sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.hpp")))
    deps.append(os.path.join(ROOT, "include", "scicone_score.h"))
    if force or _newer(out, deps):
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out, srcs[0], "-ldl"]
        subprocess.check_call(cmd)
    return out


def build_host(force=False):
    """The MCMC host: scicone_b200/bin/inference (CLI of the reference's `inference`) and lib/libscicone_host.so."""
    host = os.path.join(HERE, "host")
    bin_dir = os.path.join(HERE, "bin")
    os.makedirs(bin_dir, exist_ok=True)
    srcs = [os.path.join(host, "inference_main.cpp")] + sorted(glob.glob(os.path.join(host, "*.hpp")))
    srcs.append(os.path.join(ROOT, "include", "scicone_score.h"))
    srcs.append(os.path.join(LIB_DIR, "libscicone_score.so"))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    common = [cxx, "-std=c++17", "-O2", "-fno-fast-math", "-ffp-contract=off", "-Wall", "-Wno-sign-compare", srcs[0],
              "-L" + LIB_DIR, "-lscicone_score"]
    exe = os.path.join(bin_dir, "inference")
    lib = os.path.join(LIB_DIR, "libscicone_host.so")
    if force or _newer(exe, srcs):
        subprocess.check_call(common + ["-Wl,-rpath,$ORIGIN/../lib", "-o", exe])
    score = os.path.join(bin_dir, "score")  # the reference's `score` executable (score_tree.cpp), same source file
    if force or _newer(score, srcs):
        subprocess.check_call(common + ["-DSCICONE_SCORE_MAIN", "-Wl,-rpath,$ORIGIN/../lib", "-o", score])
    if force or _newer(lib, srcs):
        subprocess.check_call(common + ["-DSCICONE_HOST_NO_MAIN", "-fPIC", "-shared", "-Wl,-rpath,$ORIGIN", "-o", lib])
    # restarts over the GPUs of a box + best-of-N / robustness / retries (pyscicone's learn_tree_parallel): no library needed
    learn = os.path.join(bin_dir, "learn_tree")
    learn_src = os.path.join(host, "learn_tree_main.cpp")
    if force or _newer(learn, [learn_src]):
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-Wall", learn_src, "-o", learn])
    return [exe, score, lib, learn]


def build_all(force=False, verbose=False):
    return [build_score_lib(force, verbose)] + build_host(force)


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
