"""
Builds the C-ABI shared library `libxitorch_b200.so` (in-tree, next to the sources) for sm_100a.

    python -m xitorch_b200.csrc.build [--force]

nvcc cross-compiles without a GPU.  Objects are rebuilt only when a source or header is newer.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["matvec.cu", "solve.cu", "gmres.cu", "symeig.cu", "linop.cu", "peer.cu", "small_solve.cu"]
HEADERS = ["common.cuh", "matvec.cuh", "solve_common.cuh", os.path.join(ROOT, "include", "xitorch_b200.h")]
LIB = os.path.join(HERE, "libxitorch_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            log = open(o + ".log", "w")
            procs.append((src, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log, o))
    failed = []
    for src, p, log, o in procs:
        rc = p.wait()
        log.close()
        if verbose or rc != 0:
            sys.stderr.write(open(o + ".log").read())
        if rc != 0:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed for: %s" % ", ".join(failed))
    if force or procs or _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
