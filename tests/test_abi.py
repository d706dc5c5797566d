"""CPU: the C-ABI library loads and exports every symbol include/chbin_b200.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

import chbin_b200
from chbin_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "chbin_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = capi.load()
    names = _declared()
    assert len(names) >= 25
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/chbin_b200.h but not exported by libchbin_b200.so"
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    assert lib.chb_abi_version() == 1


def test_library_is_sm100a_cuda():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", capi.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        capi.Context(0)
    import numpy as np

    X, bins, _ = chbin_b200.synth.make_contig_features(100, 2, 1, 5)
    with pytest.raises(RuntimeError):
        chbin_b200.fit_cluster(X, 2, bins, None, 5, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ch-bin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
