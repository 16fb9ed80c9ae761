"""Builds the sm_100a CUDA library in-tree: selalib_b200/lib/libsllb200.so.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libsllb200.so")
SOURCES = ["sllb_kernels.cu", "sllb_capi.cu", "sllb_sims.cu", "sllb_dd6d.cu", "sllb_compat6d.cu", "sllb_spline_dd.cu", "sllb_splitting.cu", "sllb_hermite.cu", "sllb_sim4d_nml.cu", "sllb_lagrange_plane.cu", "sllb_diag.cu", "sllb_poisson_direct.cu", "sllb_sim2d_nml.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-ccbin", HOST_CXX, "-fmad=true"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sll_b200.h"),
                                                         os.path.join(HERE, "..", "include", "sll_b200_sim6d_compat.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    objs = [os.path.join(LIBDIR, src.replace(".cu", ".o")) for src in SOURCES]

    def compile_one(pair):
        src, obj = pair
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.check_call(cmd)

    # the translation units are independent: compile them side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=1 if verbose else min(len(SOURCES), os.cpu_count() or 1)) as pool:
        list(pool.map(compile_one, zip(SOURCES, objs)))
    cmd = [NVCC, "-shared", "-ccbin", HOST_CXX, "-o", SO] + objs + ["-lcufft", "-lnccl", "-lcudart"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
