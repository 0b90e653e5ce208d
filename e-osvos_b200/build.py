"""Builds libeosvos_b200.so in-tree with nvcc for sm_100a (the only supported target)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeosvos_b200.so")
SOURCES = ["common.cu", "conv_gemm.cu", "conv_fprop.cu", "conv_api.cu", "gn.cu", "roi_align.cu", "mask_loss.cu", "mask_tail.cu", "jf_measure.cu",
           "meta_update.cu", "misc.cu", "nms.cu", "rpn.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]
# activation / operand storage: fp16 (default) or bf16 (EOSVOS_ACT=bf16 at build time)
if os.environ.get("EOSVOS_ACT", "fp16").lower() == "bf16":
    NVCC_FLAGS.append("-DEOSVOS_ACT_BF16")


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "eosvos_b200.h"))
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stdout.write(f"--- {s}\n{out}")
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed building libeosvos_b200.so")
    if procs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
