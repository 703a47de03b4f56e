"""Builds the sm_100a shared library ``_lib/libwd_b200.so`` in-tree with nvcc (no torch extension, no JIT cache).

``python -m worddiffusion_b200.build`` or ``__graft_entry__.build()``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "libwd_b200.so")
SOURCES = ["gemm_tc.cu", "gemm_pair.cu", "ops.cu", "attn_flash.cu", "attn_tc.cu", "engine.cu", "wgrad_tc.cu", "ops_bwd.cu", "train.cu", "f32_path.cu", "f32_gemm_tc.cu", "tblock.cu"]
HEADERS = ["common.cuh", "gemm_tc.cuh", "ops.cuh", "wgrad_tc.cuh", "ops_bwd.cuh", "engine_internal.h", "epilogue.cuh", "f32_tc.h", "tblock.cuh", os.path.join("..", "..", "include", "wd_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build_library(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([_nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB_PATH, objs):
        # static cudart (nvcc default): the .so must not depend on a libcudart.so being on the loader path of the GPU box
        run([_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"])
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
