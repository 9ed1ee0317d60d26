"""Builds the CUDA library of the encode path for sm_100a, in-tree.

  python -m dmmt_jpeg_encoder_b200.build [--force] [--verbose]

Outputs (git-ignored, shipped to the GPU box by gpurun):
  dmmt_jpeg_encoder_b200/lib/libdmmt_cuda.so   C-ABI shared library (include/dmmt_cuda.h)
  dmmt_jpeg_encoder_b200/lib/libdmmt_cuda.a    the same objects as a static archive (what a Rust
                                               build.rs links)
  dmmt_jpeg_encoder_b200/lib/dmmt-jpeg-encoder the CLI front-end (mirrors src/cli.rs + src/main.rs)
  dmmt_jpeg_encoder_b200/lib/dmmt-jpeg-batch   many files per process: parallel ingest, pipelined batches, direct
                                               file writes from the pinned output arena (csrc/cli_batch.cpp)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(LIB, "obj")
SO = os.path.join(LIB, "libdmmt_cuda.so")
AR = os.path.join(LIB, "libdmmt_cuda.a")
CLI = os.path.join(LIB, "dmmt-jpeg-encoder")
CLI_BATCH = os.path.join(LIB, "dmmt-jpeg-batch")

CU_SOURCES = ["k1_transform.cu", "k2_entropy.cu", "dmmt_api.cu", "dmmt_batch.cu", "dmmt_shard.cu"]
# -fmad=false: the reference's f32 arithmetic never contracts a*b+c; the kernels also use
# explicit __fmul_rn/__fadd_rn, the flag is the belt to those braces.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _deps() -> list[str]:
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(HERE, "..", "include", "dmmt_cuda.h"))
    d.append(os.path.abspath(__file__))
    return d


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    deps = _deps()
    have_cli = os.path.exists(os.path.join(CSRC, "cli_main.cpp"))
    if not force and _newer(SO, deps) and _newer(AR, deps) and (not have_cli or (_newer(CLI, deps) and _newer(CLI_BATCH, deps))):
        return SO
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in CU_SOURCES:
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, *objs])
    if os.path.exists(AR):
        os.remove(AR)
    subprocess.check_call(["ar", "rcs", AR, *objs])
    if have_cli:
        for src, exe in (("cli_main.cpp", CLI), ("cli_batch.cpp", CLI_BATCH)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(CSRC, src), "-o", exe,
                                   "-L" + LIB, "-ldmmt_cuda", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
