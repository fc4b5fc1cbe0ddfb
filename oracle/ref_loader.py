"""Import the unmodified reference (GAMES-UChile/mogptk) for tests and baselines.

TEST / BASELINE INFRASTRUCTURE -- only tests/, bench.py's reference legs and
__graft_entry__.smoke() may import this; the product (mogptk_b200/) never does.

Where it comes from: ``oracle/_ref/mogptk`` (byte-for-byte copy made by oracle/build_ref.py; this is
what exists on the GPU box) or, in the build container, /root/reference directly.

The reference imports matplotlib and IPython at module import (mogptk/model.py:9-12, gpr/plot.py:2-4,
data.py:11-14, util.py:3, gpr/model.py:5); neither is installed here and there is no network, so inert
stand-ins are registered in ``sys.modules`` first.  Plotting is never exercised.
"""
import json
import os
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_CHECKOUT = "/root/reference"

_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.dates",
          "matplotlib.units", "mpl_toolkits", "mpl_toolkits.axes_grid1", "IPython", "IPython.display"]


def reference_root():
    """Directory that contains the reference's ``mogptk`` package, or None."""
    for root in (REF_COPY, REF_CHECKOUT):
        if os.path.isfile(os.path.join(root, "mogptk", "__init__.py")):
            return root
    return None


def available():
    return reference_root() is not None


def verify_copy():
    """True when oracle/_ref matches the MANIFEST written by build_ref.py (i.e. the copy is unmodified)."""
    mpath = os.path.join(REF_COPY, "MANIFEST.json")
    if not os.path.exists(mpath):
        return False
    from oracle.build_ref import manifest
    with open(mpath) as f:
        want = json.load(f)["files"]
    return manifest(os.path.join(REF_COPY, "mogptk")) == want


def import_reference(device="cpu"):
    """Returns the reference's ``mogptk`` module with ``gpr.config.device`` set to `device`
    ("cpu", "cuda" or "cuda:n"; gpr/config.py:36-62)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not available: run `python oracle/build_ref.py` in the build container")
    for m in _STUBS:
        sys.modules.setdefault(m, MagicMock())
    import pandas.plotting
    pandas.plotting.register_matplotlib_converters = lambda *a, **k: None
    if root not in sys.path:
        sys.path.insert(0, root)
    import mogptk
    if os.path.dirname(os.path.dirname(os.path.abspath(mogptk.__file__))) != os.path.abspath(root):
        raise RuntimeError("a different `mogptk` is already imported from %s" % mogptk.__file__)
    set_device(mogptk, device)
    return mogptk


def set_device(mogptk, device):
    device = str(device)
    if device == "cpu":
        mogptk.gpr.use_cpu()
    else:
        idx = int(device.split(":")[1]) if ":" in device else None
        mogptk.gpr.use_gpu(idx)
