"""TEST INFRASTRUCTURE — loader for the *unmodified* reference sources.

This file loads ``/root/reference/{chroma,dtw,otw_eran,livenote,livenote_v2,wtw}.py``
(Python-2 sources) into Python 3 through a load-time text shim.  It does not
restate any algorithm: the reference's own code executes.  It only exists in
the build container (``/root/reference`` is absent on the GPU box), so it is
used for exactly two things:

* ``tests/golden/make_golden.py`` — generate the committed golden vectors;
* ``tests/test_oracle_cpu.py`` — pin the C/numpy restatement in
  ``oracle/`` against the real reference (skipped when the tree is absent).

Nothing in the product package may import this module.

Shim edits (all syntactic):
  * ``print x``            -> ``print(x)``       otw_eran.py:54,71  livenote_v2.py:57,64,81  wtw.py:153
  * ``L/2``                -> ``L//2``           chroma.py:49,53    wtw.py:142,146
  * ``dtw_win_size/hop_size`` etc -> ``//``      wtw.py:96-107,127-128
  * ``numpy.int``          -> ``int``            dtw.py:17
Third-party modules the reference imports but which are absent here
(matplotlib, IPython, pyaudio, librosa) are stubbed; the three librosa
functions that carry arithmetic are restated in ``oracle/librosa_restated.py``.
"""
import os
import re
import sys
import types

import numpy as np

from . import librosa_restated as _lr

REFERENCE_ROOT = os.environ.get("AFS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dtw.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_installed = False


def _install_stubs():
    global _installed
    if _installed:
        return
    if not hasattr(np, "int"):
        np.int = int  # dtw.py:17 uses the alias removed in numpy 1.24
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot", rcParams={})
        _stub("matplotlib", pyplot=plt)
    if "IPython" not in sys.modules:
        disp = _stub("IPython.display")
        _stub("IPython", display=disp)
    if "pyaudio" not in sys.modules:
        _stub("pyaudio")
    if "librosa" not in sys.modules:
        _stub(
            "librosa",
            load=_lr.load,
            display=_stub("librosa.display"),
            filters=_stub("librosa.filters", chroma=_lr.filters_chroma),
            util=_stub("librosa.util", normalize=_lr.util_normalize),
            feature=_stub("librosa.feature", chroma_stft=lambda **kw: None),  # dead call, wtw.py:85
        )
    _installed = True


_cache = {}


def load(name):
    """Return the reference module ``name`` (e.g. ``'dtw'``) executed under the shim."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, name + ".py")
    with open(path) as fh:
        src = fh.read()
    src = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", src, flags=re.M)
    src = re.sub(r"\bL/2\b", "L//2", src)
    src = src.replace("(self.dtw_win_size/self.hop_size)", "(self.dtw_win_size//self.hop_size)")
    src = src.replace("self.dtw_hop_size / self.hop_size", "self.dtw_hop_size // self.hop_size")
    src = src.replace("(self.dtw_hop_size/self.hop_size)", "(self.dtw_hop_size//self.hop_size)")
    mod = types.ModuleType("_afs_reference_" + name)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    _cache[name] = mod
    return mod


def song(rel):
    """Path of a file under the reference's Songs/ directory."""
    return os.path.join(REFERENCE_ROOT, "Songs", rel)
