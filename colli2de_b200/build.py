"""In-tree build of the native libraries (no JIT cache: the .so files travel to the GPU box with the repo).

    libc2d_gpu.so          nvcc, sm_100a only, -fmad=false (the analogue of the reference's
                           -ffp-contract=off, CMakeLists.txt:17-20) — kernels + the C ABI of include/c2d_gpu.h
    libc2d_b200_runner.so  g++ -std=c++23: tests/cpp/runner.cpp compiled against THIS repo's
                           include/colli2de/Registry.hpp (the same runner source is compiled against the
                           reference headers for oracle/_ref) — the flat rn_* interface bench/tests drive.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
GPU_LIB = os.path.join(HERE, "libc2d_gpu.so")
RUNNER_LIB = os.path.join(HERE, "libc2d_b200_runner.so")

CU_SOURCES = ["c2d_sort.cu", "c2d_gpu.cu", "c2d_batch.cu", "c2d_rays.cu", "c2d_grid.cu", "c2d_query.cu", "c2d_comm.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-Xcompiler", "-fPIC",
]
if os.path.exists(os.path.join(CSRC, "c2d_rays.cu")):
    NVCC_FLAGS.append("-DC2D_HAVE_RAYS")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _run(cmd: list[str]):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]}")


def build_gpu_lib(force: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    headers.append(os.path.join(REPO, "include", "c2d_gpu.h"))
    objs = []
    jobs = []
    for src in CU_SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or not _newer(obj, [path] + headers):
            jobs.append(["nvcc", *NVCC_FLAGS, "-c", path, "-o", obj])
        objs.append(obj)
    rebuilt = bool(jobs)
    if jobs:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(_run, jobs))
    if rebuilt or not os.path.exists(GPU_LIB):
        _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", GPU_LIB, *objs, "-ldl"])
    return GPU_LIB


def build_runner(force: bool = False) -> str:
    src = os.path.join(REPO, "tests", "cpp", "runner.cpp")
    inc = os.path.join(REPO, "include")
    deps = [src, os.path.join(REPO, "tests", "cpp", "runner_api.h"), GPU_LIB]
    for root, _, files in os.walk(inc):
        deps += [os.path.join(root, f) for f in files]
    if force or not _newer(RUNNER_LIB, deps):
        _run(["g++", "-std=c++23", "-O2", "-fPIC", "-shared", "-DRN_BACKEND_B200", "-I" + inc,
              "-I" + os.path.join(REPO, "tests", "cpp"), src, "-o", RUNNER_LIB,
              "-L" + HERE, "-lc2d_gpu", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic"])
    return RUNNER_LIB


API_CHECK = os.path.join(HERE, "api_check")


def build_api_check(force: bool = False) -> str:
    """tests/cpp/api_check.cpp: an "application" using the drop-in headers with non-uint32 EntityIds."""
    src = os.path.join(REPO, "tests", "cpp", "api_check.cpp")
    inc = os.path.join(REPO, "include")
    deps = [src, GPU_LIB, os.path.join(inc, "colli2de", "Registry.hpp"), os.path.join(inc, "c2d_gpu.h")]
    if force or not _newer(API_CHECK, deps):
        _run(["g++", "-std=c++23", "-O2", "-DENABLE_ASSERT", "-I" + inc, src, "-o", API_CHECK, "-L" + HERE, "-lc2d_gpu",
              "-Wl,-rpath,$ORIGIN"])
    return API_CHECK


SNAPSHOT_CHECK = os.path.join(HERE, "libc2d_snapshot_check.so")


def build_snapshot_check(force: bool = False) -> str:
    """tests/cpp/snapshot_check.cpp: the host-only snapshot codec (no CUDA) behind a C interface, for the CPU tests."""
    src = os.path.join(REPO, "tests", "cpp", "snapshot_check.cpp")
    inc = os.path.join(REPO, "include")
    deps = [src, os.path.join(inc, "colli2de", "internal", "utils", "Snapshot.hpp")]
    if force or not _newer(SNAPSHOT_CHECK, deps):
        _run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-I" + inc, src, "-o", SNAPSHOT_CHECK])
    return SNAPSHOT_CHECK


def build_oracles():
    targets = ["ref"] + (["oracle"] if os.path.exists(os.path.join(REPO, "oracle", "c2d_oracle.c")) else [])
    _run(["make", "-C", os.path.join(REPO, "oracle"), *targets])


def build_all(force: bool = False):
    build_gpu_lib(force)
    build_runner(force)
    build_api_check(force)
    build_snapshot_check(force)
    build_oracles()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", GPU_LIB, RUNNER_LIB)
