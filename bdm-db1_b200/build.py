"""Build libdb1_sm100.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

Usage:  python bdm-db1_b200/build.py [--force] [--verbose]
The .so lands next to this file so that it travels with the repo snapshot to the GPU box.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdb1_sm100.so")
HOSTLIB = os.path.join(HERE, "libdb1_host.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(HERE, "..", "include"),
]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "db1_sm100.h"))
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    for f in cus:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        if force or _newer([src] + headers, obj):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((f, cmd))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = {ex.submit(subprocess.run, cmd, capture_output=True, text=True): f for f, cmd in jobs}
            for fu in cf.as_completed(futs):
                r = fu.result()
                if verbose or r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s" % futs[fu])
    objs = [os.path.join(OBJ, f[:-3] + ".o") for f in cus]
    if force or jobs or _newer(objs, LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    # host-only helpers (discretiser, RL token layout, index builders): plain g++
    host_srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cpp"))
    if host_srcs and (force or _newer(host_srcs + headers, HOSTLIB)):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I",
               os.path.join(HERE, "..", "include"), "-o", HOSTLIB] + host_srcs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("host lib build failed")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
