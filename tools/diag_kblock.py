"""Which of A-operand TMA loads, B-operand TMA loads and MMAs bounds a K block of conv_fprop_kernel?

`--build` (here, no GPU needed) compiles timing-only variants of the library with -DEOSVOS_DIAG=<bits> into
tools/_diag/ (git-ignored; travels to the GPU box).  Without arguments (on the GPU) every variant times the same
launches with CUDA events, L2-warm (a CUDA graph of back-to-back repeats) and prints microseconds per launch and per K block.
Results of the variants are garbage by construction; only the product library (bits 0) is ever shipped.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIAG = os.path.join(ROOT, "tools", "_diag")
VARIANTS = {0: "product", 1: "no A loads", 2: "no B loads", 3: "no loads", 4: "no MMA", 7: "ring handshake only"}
CASES = [  # (N, H, W, Cin, Cout, k, bn_hint)
    (1, 12, 21, 256, 256, 3, 0), (1, 48, 84, 256, 256, 3, 0), (1, 48, 84, 256, 256, 3, 256),
    (3, 48, 84, 256, 256, 3, 0), (3, 48, 84, 1024, 256, 1, 0), (3, 192, 336, 256, 256, 3, 0),
    (3, 192, 336, 64, 64, 3, 0), (3, 96, 168, 128, 128, 3, 0),
]


def build():
    sys.path.insert(0, ROOT)
    import importlib
    b = importlib.import_module("e-osvos_b200.build")
    os.makedirs(DIAG, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(os.path.dirname(b.LIB), "build")
    for bits in VARIANTS:
        if bits == 0:
            continue
        obj = os.path.join(DIAG, f"conv_fprop_d{bits}.o")
        subprocess.check_call([nvcc] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] +
                              [f"-DEOSVOS_DIAG={bits}", "-c", os.path.join(b.CSRC, "conv_fprop.cu"), "-o", obj])
        objs = [obj if s == "conv_fprop.cu" else os.path.join(objdir, s.replace(".cu", ".o")) for s in b.SOURCES]
        subprocess.check_call([nvcc, "-shared", "-o", os.path.join(DIAG, f"libeosvos_d{bits}.so")] + objs + ["-lcudart"])
        os.remove(obj)
    print("built", sorted(os.listdir(DIAG)))


def child(bits):
    sys.path.insert(0, ROOT)
    import torch
    from eosvos_b200 import _lib
    if bits:
        _lib.LIB_PATH = os.path.join(DIAG, f"libeosvos_d{bits}.so")
    from eosvos_b200 import kernels as k
    dev = torch.device("cuda:0")
    row = []
    for (N, H, W, Ci, Co, ks, hint) in CASES:
        x = torch.randn(N, H, W, Ci, device=dev).to(k.ACT_DTYPE)
        w = (torch.randn(Co, ks, ks, Ci, device=dev) * 0.05).to(k.ACT_DTYPE)
        kw = dict(stride=1, pad=ks // 2)
        if hint:
            kw["bn_hint"] = hint
        for _ in range(5):
            k.conv2d_fprop(x, w, **kw)
        torch.cuda.synchronize()
        reps = 40
        g = torch.cuda.CUDAGraph()                   # replayed launches: no host time between them
        with torch.cuda.graph(g):
            for _ in range(reps):
                k.conv2d_fprop(x, w, **kw)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) * 1e3 / reps)
    print("RESULT", bits, " ".join(f"{v:.2f}" for v in row), flush=True)


def main():
    print("cases:", CASES)
    for pdl in ("1", "0"):
        print(f"EOSVOS_PDL={pdl} (programmatic dependent launch {'on' if pdl == '1' else 'off'})")
        for bits, name in VARIANTS.items():
            out = subprocess.run([sys.executable, __file__, "--child", str(bits)], capture_output=True, text=True,
                                 env=dict(os.environ, EOSVOS_PDL=pdl))
            line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
            print(f"{name:>22s}:", line[0].split(" ", 2)[2] if line else out.stderr[-400:])


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    elif "--child" in sys.argv:
        child(int(sys.argv[sys.argv.index("--child") + 1]))
    else:
        main()
