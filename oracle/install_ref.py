"""TEST INFRASTRUCTURE ONLY.  Recipe that places an UNMODIFIED copy of the reference's Python sources, configs and
split lists under the git-ignored directory oracle/_ref/ so that they travel to the GPU box with the working tree
(like the built .so; `/root/reference` itself does not exist there):

    python -m oracle.install_ref            # /root/reference -> oracle/_ref   (no-op when the source is absent)

What is taken: src/**/*.py, cfgs/*.yaml, data/<DATASET>/*.txt, LICENSE.  Nothing is edited; the API drift of the
installed torch/torchvision is bridged at import time by oracle/ref_shims.py + oracle/ref_harness.py.
oracle/_ref/ is listed in .gitignore (never committed) and NOT in .gpurunignore.
Used by: tests (drop-in and end-to-end parity), bench.py --impl reference, bench.py's gpu_eager_baseline.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("EOSVOS_REFERENCE_SOURCE", "/root/reference")


def _wanted(rel):
    parts = rel.split(os.sep)
    if parts[0] == "src":
        return rel.endswith(".py")
    if parts[0] == "cfgs":
        return rel.endswith(".yaml")
    if parts[0] == "data":
        return rel.endswith(".txt") and len(parts) == 3
    return rel == "LICENSE"


def install(source=SOURCE, dest=DEST, verbose=False):
    """-> number of files copied (0 when up to date), or None when there is no reference checkout to copy from."""
    if not os.path.isdir(os.path.join(source, "src", "networks")):
        return None
    copied = 0
    for dirpath, _, files in os.walk(source):
        for f in files:
            full = os.path.join(dirpath, f)
            rel = os.path.relpath(full, source)
            if not _wanted(rel):
                continue
            out = os.path.join(dest, rel)
            if os.path.exists(out) and filecmp.cmp(full, out, shallow=False):
                continue
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(full, out)
            copied += 1
            if verbose:
                print("  ", rel)
    return copied


def installed(dest=DEST):
    return os.path.isdir(os.path.join(dest, "src", "networks"))


if __name__ == "__main__":
    n = install(verbose="-v" in sys.argv)
    print(f"reference not found at {SOURCE}" if n is None else f"{n} file(s) copied into {DEST}")
