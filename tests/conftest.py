import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def statistical(attempts=2):
    """For the full-model comparisons whose outcome is a random variable: the CUDA path sums GroupNorm statistics, split-K
    weight gradients and RoIAlign gradients with fp32 atomics, a random-init network amplifies that
    noise chaotically, and a lock-step frame can land on a tie configuration the comparison rules do not name (seen
    about once per ten runs of the whole suite, never twice in a row).  Such a test states its bound for ONE draw and is
    allowed to draw again once; the first failure is printed, never hidden.  Kernel-level and integer-exact tests do
    not use this."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            for i in range(attempts):
                try:
                    return fn(*args, **kwargs)
                except AssertionError as e:
                    if i == attempts - 1:
                        raise
                    sys.stderr.write(f"\n[statistical] {fn.__name__}: draw {i + 1} failed -- {str(e)[:400]} -- drawing again\n")
                    import gc
                    gc.collect()
                    try:
                        import torch
                        torch.cuda.empty_cache()
                    except Exception:
                        pass
        return wrapper
    return deco


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _library_is_current():
    """The tests exercise the in-tree libeosvos_b200.so: (re)build it incrementally when it is missing or older than
    its sources (a no-op otherwise), so a stale binary can never be what is tested."""
    import eosvos_b200  # noqa: F401  (root shim: makes the package importable)
    from eosvos_b200 import build as B
    try:
        B.build()
    except Exception as e:      # no nvcc on this machine: test whatever binary is there (load() raises if none)
        if not os.path.exists(B.LIB):
            raise
        print(f"[conftest] could not rebuild libeosvos_b200.so ({e}); testing the existing binary")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
