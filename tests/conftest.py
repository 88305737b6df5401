"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol
export.  `-m gpu` runs on a B200 and calls the CUDA path through the C-ABI.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "smart-crossover_b200"), os.path.join(ROOT, "tests"),
          os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # SX_SORT_TUNING=<code> runs the suite under another radix-downsweep shape (sx_sort_set_tuning)
    tuning = os.environ.get("SX_SORT_TUNING")
    if tuning:
        from smart_crossover._native import lib
        assert lib.sx_sort_set_tuning(int(tuning)) == 0, f"bad SX_SORT_TUNING={tuning}"
