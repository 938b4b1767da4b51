"""CPU suite: the C-ABI library loads without a GPU and exports every symbol that
include/afsync.h declares; compute entry points fail loudly without a device."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "afsync.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(entry):
    entry.build()
    nat = entry.submodule("_native")
    L = nat.lib()
    declared = header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), "libafsync.so does not export %s" % name
    # and the binding table covers the header exactly
    assert sorted(nat.exported_symbols()) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name


def test_library_is_sm100a_only(entry):
    nat = entry.submodule("_native")
    out = subprocess.run(["cuobjdump", "-lelf", nat.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_version_and_errors(entry):
    nat = entry.submodule("_native")
    L = nat.lib()
    assert b"sm_100a" in L.afs_version()
    import ctypes as C
    h = C.c_void_p()
    rc = L.afs_dtw_plan_create(C.byref(h), 0, None, None, None, None, 12, 0)
    assert rc == -1 and b"afs_dtw_plan_create" in L.afs_last_error()


def test_no_cpu_fallback(entry):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dtw = entry.submodule("dtw")
    nat = entry.submodule("_native")
    import numpy as np
    with pytest.raises(nat.AfsError):
        dtw.DTW(np.zeros((12, 4)), np.zeros((12, 5)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "real-time-audio-sync_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), "%s mentions the oracle" % f


def test_host_side_filterbank_matches_oracle_restatement(entry):
    """chroma.chroma_filterbank (product, host side, feeds afs_chroma_plan_create) vs the oracle's restatement of
    librosa.filters.chroma and the independent port in transformers."""
    import numpy as np
    chroma = entry.submodule("chroma")
    from oracle import librosa_restated as lr
    fb = chroma.chroma_filterbank(22050, 4096)
    assert fb.shape == (12, 2049) and np.array_equal(fb, lr.filters_chroma(22050, 4096))
    assert chroma.fft_len == 4096 and chroma.hop_size == 2048 and chroma.fs == 22050      # chroma.py:20-22


def test_drop_in_module_names_and_signatures(entry):
    """The reference's import lines keep working when the package directory is on sys.path."""
    import inspect
    import subprocess
    import sys
    import os
    pkg = entry.PKG_DIR
    code = ("import sys; sys.path.insert(0, %r); "
            "from chroma import wav_to_chroma, wav_to_chroma_col, wav_to_chroma_diff, create_stft, create_chroma; "
            "from dtw import DTW; from otw_eran import OnlineTimeWarping; from livenote_v2 import LiveNoteV2; "
            "from livenote import LiveNote; from wtw import WTW; import inspect; "
            "print(list(inspect.signature(DTW).parameters)[:2], list(inspect.signature(OnlineTimeWarping.__init__).parameters)[:3], "
            "list(inspect.signature(LiveNoteV2.__init__).parameters)[:5], list(inspect.signature(WTW.__init__).parameters)[:4])") % pkg
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "['seq_a', 'seq_b'] ['self', 'ref', 'params'] ['self', 'ref', 'params', 'debug_params', 'chroma_diff'] ['self', 'ref_recording', 'params', 'debug_params']" in out.stdout
