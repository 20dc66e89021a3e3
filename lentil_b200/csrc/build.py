"""Build liblentil_b200.so in-tree with nvcc for sm_100a (no torch headers needed: the boundary
is a plain C ABI).  Usage: python -m lentil_b200.csrc.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["capi.cu", "mft_c128.cu", "mft_folded.cu", "mft_czt.cu", "mft_c64.cu", "pupil_prep.cu", "accum.cu", "fit_tilt.cu", "detector_ops.cu", "rescale.cu"]
HEADERS = ["lfd_common.cuh", os.path.join(ROOT, "include", "lentil_b200.h")]
LIB = os.path.join(os.path.dirname(HERE), "liblentil_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "--use_fast_math" if False else "-DLFD_NO_FAST_MATH",  # FP64 path: never fast-math
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [
        h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = res.stdout
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building liblentil_b200.so")
    if verbose:
        for line in log.splitlines():
            if "registers" in line or "spill" in line or "error" in line or "warning" in line:
                print(line)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
