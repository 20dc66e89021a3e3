"""Build liblentil_b200.so in-tree with nvcc for sm_100a (no torch headers needed: the boundary
is a plain C ABI).  Every .cu is compiled to its own object (in parallel, rebuilt only when it or a
header changed) and the objects are linked into the shared library.

Usage: python -m lentil_b200.csrc.build [--force] [--tag NAME -DMACRO=VALUE ...]
A tagged build (development aid: kernel-variant experiments) writes liblentil_b200_NAME.so next to the
production library; LFD_LIB=<path> makes lentil_b200._lib load it instead."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["capi.cu", "mft_c128.cu", "mft_folded.cu", "mft_czt.cu", "mft_czt_f32.cu", "mft_c64.cu", "pupil_prep.cu", "accum.cu", "fit_tilt.cu", "detector_ops.cu", "rescale.cu"]
HEADERS = ["lfd_common.cuh", "mft_czt_common.cuh", "mft_czt_body.cuh", os.path.join(ROOT, "include", "lentil_b200.h")]
LIB = os.path.join(os.path.dirname(HERE), "liblentil_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-DLFD_NO_FAST_MATH",  # FP64 path: never fast-math
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _header_paths():
    return [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]


def _lib_path(tag):
    return LIB if not tag else LIB[:-3] + "_" + tag + ".so"


def needs_build(tag=None):
    lib = _lib_path(tag)
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(HERE, s) for s in SOURCES] + _header_paths()
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, defines, log):
    cmd = [_nvcc()] + NVCC_FLAGS + list(defines) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append((src, " ".join(cmd), res.stdout, res.returncode))
    return res.returncode


def build(force=False, verbose=True, tag=None, defines=()):
    lib = _lib_path(tag)
    if not force and not needs_build(tag):
        return lib
    obj_dir = os.path.join(OBJ_DIR, tag or "prod")
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _header_paths())
    jobs, objs, log = [], [], []
    for s in SOURCES:
        src, obj = os.path.join(HERE, s), os.path.join(obj_dir, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    if verbose:
        print(f"nvcc {' '.join(NVCC_FLAGS + list(defines))}: compiling {len(jobs)} of {len(SOURCES)} sources", flush=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        rcs = list(ex.map(lambda j: _compile(j[0], j[1], defines, log), jobs))
    text = "".join(f"$ {cmd}\n{out}\n" for _, cmd, out, _ in log)
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs
    if not any(rcs):
        res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        text += f"$ {' '.join(link)}\n{res.stdout}\n"
        rcs.append(res.returncode)
    with open(os.path.join(HERE, "build.log" if not tag else f"build_{tag}.log"), "w") as fh:
        fh.write(text)
    if any(rcs):
        sys.stderr.write(text)
        raise RuntimeError("nvcc failed building " + os.path.basename(lib))
    if verbose:
        for line in text.splitlines():
            if "spill" in line and " 0 bytes spill stores, 0 bytes spill loads" not in line or "error" in line or "warning" in line:
                print(line)
    return lib


if __name__ == "__main__":
    args = sys.argv[1:]
    tag = args[args.index("--tag") + 1] if "--tag" in args else None
    build(force="--force" in args, tag=tag, defines=[a for a in args if a.startswith("-D")])
