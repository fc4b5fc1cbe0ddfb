import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    """Cases of the kernels the product supports (fixtures named next_* belong to the not-yet-supported families)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("next_"))


def next_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("next_"))


def load_golden(name):
    import torch
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["kind"] = str(g["kind"])
    for k in ("C", "Q", "D"):
        g[k] = int(g[k])
    g["jitter"] = float(g["jitter"])
    g["params"] = {k[2:]: torch.tensor(v, dtype=torch.float64) for k, v in g.items() if k.startswith("p_")}
    g["sigma_t"] = torch.tensor(g["sigma"], dtype=torch.float64)
    g["name"] = name
    return g


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (build it if it is not there yet)."""
    from mogptk_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.load()


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mogptk_b200.engine import Engine
    eng = Engine(device=0, max_n=8192)
    yield eng
    eng.close()
