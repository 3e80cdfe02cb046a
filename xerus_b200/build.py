"""Builds libxb200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m xerus_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repository snapshot.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libxb200.so")
SOURCES = ["runtime.cu", "gemm_f64.cu", "movement.cu", "qr_f64.cu", "small_f64.cu", "svd_f64.cu", "solve_f64.cu", "blas_api.cu", "tt.cu", "als.cu", "fileio.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libxb200.so cannot be built")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "xb200.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        p = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, p.returncode, p.stdout + p.stderr

    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for s, rc, out in ex.map(compile_one, jobs):
            logs.append("== %s\n%s" % (os.path.basename(s), out))
            if rc != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (s, out))
    with open(os.path.join(BUILD, "ptxas.log"), "a" if not force else "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    if jobs or force or not os.path.exists(LIB):
        p = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
