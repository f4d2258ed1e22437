"""In-tree build of the two shared libraries (no JIT cache, no pip install).

  libycnr_host.so  g++   csrc/host_frontend.cc          (CPU front end, include/ycnr_host.h)
  libycnr_als.so   nvcc  csrc/*.cu for sm_100a only     (CUDA hot path, include/ycnr_als.h)

The built .so files live next to this file so that they travel with the repo
snapshot to the GPU box and show up as in-tree native code when loaded.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
HOST_LIB = os.path.join(PKG_DIR, "libycnr_host.so")
CUDA_LIB = os.path.join(PKG_DIR, "libycnr_als.so")

CUDA_SOURCES = ["ycnr_als.cu"]
CUDA_HEADERS = ["als_kernels.cuh", "rmse_kernels.cuh", "common.cuh", "gram_tc.cuh", "portion_kernels.cuh", "recommend_kernels.cuh", "ingest_kernels.cuh"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return proc.stdout


def build_host(force=False):
    src = os.path.join(CSRC, "host_frontend.cc")
    deps = [src, os.path.join(INCLUDE, "ycnr_host.h")]
    if force or _newer(HOST_LIB, deps):
        _run(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", INCLUDE,
              src, "-o", HOST_LIB])
    return HOST_LIB


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in CUDA_HEADERS] + [os.path.join(INCLUDE, "ycnr_als.h")]
    if force or _newer(CUDA_LIB, deps):
        nvcc = find_nvcc()
        if nvcc is None:
            raise RuntimeError("nvcc not found: the CUDA hot path cannot be built (there is no CPU fallback)")
        cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo"] + NVCC_ARCH + [
            "-Xcompiler", "-fPIC", "-shared", "-I", INCLUDE, "-I", CSRC]
        cmd += os.environ.get("YCNR_NVCC_FLAGS", "").split()       # tuning experiments: -DYCNR_...=n
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += srcs + ["-o", CUDA_LIB, "-lcudart"]
        out = _run(cmd)
        if verbose:
            print(out)
    return CUDA_LIB


def build_all(force=False, verbose=False):
    return build_host(force), build_cuda(force, verbose)


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_all(force=force, verbose="-v" in sys.argv))
