"""Build libcwt_b200.so in-tree with nvcc for sm_100a (no torch headers: pure C ABI).

    python -m few_shot_seg_cwt_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcwt_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _code_only(text: str) -> str:
    """Source text without comments and blank lines: the fingerprint names the CODE a profile was taken on, so editing a
    comment does not orphan the evidence under profiles/ (string literals in these sources hold no comment markers)."""
    import re
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return "\n".join(l.rstrip() for l in text.splitlines() if l.strip())


def _fingerprint() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(_code_only(open(os.path.join(root, f), encoding="utf-8").read()).encode())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB


def fingerprint() -> str:
    """sha256 over the code (comments stripped) of every CUDA source and header and the compiler flags: names the build a
    profile was taken on."""
    return _fingerprint()


# the sources a kernel is compiled from (its .cu and every header that file includes, directly or not): a capture of that
# kernel stays valid while THESE are unchanged, whatever happens to the rest of the library
KERNEL_SOURCES = {
    "k_fit_resident": ["fit_resident.cu", "resident_common.cuh", "hires.cuh", "common.cuh", "tma_pipe.cuh"],
    "k_fit_l2": ["fit_l2.cu", "resident_common.cuh", "hires.cuh", "common.cuh", "tma_pipe.cuh"],
    "k_logits_iou_stream": ["iou.cu", "iou_stream.cuh", "skinny.cuh", "hires.cuh", "tma_pipe.cuh", "common.cuh"],
    "k_rtf_stream": ["transformer.cu", "skinny_stream.cuh", "skinny.cuh", "tma_pipe.cuh", "common.cuh"],
    "k_ftr_stream": ["transformer.cu", "skinny_stream.cuh", "skinny.cuh", "tma_pipe.cuh", "common.cuh"],
    "k_kproj_scores": ["kproj_tcgen05.cu", "common.cuh"],
}


def kernel_fingerprint(kernel: str) -> str:
    """sha256 over the code (comments stripped) of the sources ``kernel`` is compiled from, the public header and the flags."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in KERNEL_SOURCES[kernel]] + [os.path.join(os.path.dirname(HERE), "include", "cwt_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(_code_only(open(f, encoding="utf-8").read()).encode())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def kernel_fingerprints() -> dict:
    return {k: kernel_fingerprint(k) for k in KERNEL_SOURCES}


if __name__ == "__main__":
    if "--fingerprint" in sys.argv:
        print(_fingerprint())
    elif "--kernel-fingerprints" in sys.argv:
        import json
        print(json.dumps(kernel_fingerprints(), indent=1))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
