"""Import the UNMODIFIED reference package from /root/reference with stub third-party modules.

TEST INFRASTRUCTURE ONLY.  This file is used by `oracle/gen_golden.py` (and by ad-hoc
validation in the build container) to run the real fancompute/wavetorch CPU path.
/root/reference does not exist on the GPU box; there the only copy of the reference is the pip-installed one under
baseline/_ref/ (see DESIGN.md section 5), which `bench.py --impl reference` (and only that arm) loads through this
module.  Nothing under tests/ -m gpu, the product arm of bench.py or __graft_entry__.smoke() imports it.

`import wavetorch` fails out of the box because wavetorch/__init__.py:1 pulls in
skimage (geom.py:7, source.py:1), librosa (data/vowels.py:6), matplotlib/seaborn (plot.py:6-11),
none of which is installed here.  The two functions that matter numerically are restated:

* skimage.draw.circle(r, c, radius)  -- all integer (rr, cc) with
  ((rr-r)/radius)^2 + ((cc-c)/radius)^2 < 1   (scikit-image <= 0.18 semantics, used at
  geom.py:149 for the blur kernel and study/propagate.py:24 for the lens disk).
* skimage.draw.line(r0, c0, r1, c1)  -- Bresenham, end points inclusive (source.py:31).
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _disk(r, c, radius, shape=None):
    radius = float(radius)
    lo_r, hi_r = int(np.floor(r - radius)), int(np.ceil(r + radius))
    lo_c, hi_c = int(np.floor(c - radius)), int(np.ceil(c + radius))
    rr, cc = np.mgrid[lo_r:hi_r + 1, lo_c:hi_c + 1]
    inside = ((rr - r) / radius) ** 2 + ((cc - c) / radius) ** 2 < 1.0
    rr, cc = rr[inside], cc[inside]
    if shape is not None:
        keep = (rr >= 0) & (rr < shape[0]) & (cc >= 0) & (cc < shape[1])
        rr, cc = rr[keep], cc[keep]
    return rr, cc


def _bresenham(r0, c0, r1, c1):
    r0, c0, r1, c1 = int(r0), int(c0), int(r1), int(c1)
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    sr = 1 if r1 >= r0 else -1
    sc = 1 if c1 >= c0 else -1
    rr, cc = [], []
    r, c = r0, c0
    if dc >= dr:
        err = dc // 2
        for _ in range(dc + 1):
            rr.append(r); cc.append(c)
            err -= dr
            if err < 0:
                r += sr
                err += dc
            c += sc
    else:
        err = dr // 2
        for _ in range(dr + 1):
            rr.append(r); cc.append(c)
            err -= dc
            if err < 0:
                c += sc
                err += dr
            r += sr
    return np.asarray(rr, dtype=np.int64), np.asarray(cc, dtype=np.int64)


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "skimage" not in sys.modules:
        draw = mod("skimage.draw", circle=_disk, line=_bresenham)
        mod("skimage", draw=draw)
    for name in ("librosa", "librosa.display", "seaborn", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.animation", "matplotlib.gridspec", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.axes_divider"):
        if name not in sys.modules:
            mod(name)
    sys.modules["mpl_toolkits.axes_grid1.axes_divider"].make_axes_locatable = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.modules["librosa"].display = sys.modules["librosa.display"]


def reference_root(prefer_installed=False):
    """Directory holding the reference package: the source tree (build container) or the pip-installed copy under
    baseline/_ref (the only one that exists on the GPU box); None when neither is there."""
    roots = (INSTALLED_ROOT, REFERENCE_ROOT) if prefer_installed else (REFERENCE_ROOT, INSTALLED_ROOT)
    for r in roots:
        if os.path.isfile(os.path.join(r, "wavetorch", "rnn.py")):
            return r
    return None


def load_reference(root=None):
    """Return the reference `wavetorch` module (real code, stubbed third-party deps)."""
    _install_stubs()
    root = root or reference_root()
    if root is None:
        raise ImportError("no copy of the reference: neither %s nor %s exists" % (REFERENCE_ROOT, INSTALLED_ROOT))
    if not os.path.isdir(os.path.join(root, "wavetorch", "data")) and "wavetorch.data" not in sys.modules:
        # the reference's setup.py lists packages=['wavetorch'] only, so a pip install drops the wavetorch.data
        # sub-package that wavetorch/__init__.py:1 imports (the librosa vowel loader: out of scope here)
        sys.modules["wavetorch.data"] = types.ModuleType("wavetorch.data")
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("wavetorch")


draw_disk = _disk
draw_line = _bresenham
