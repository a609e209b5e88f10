"""CPU: the C-ABI library builds, loads and exports every symbol include/bonsai_b200.h declares, and refuses
to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bonsai_b200 import build, capi
    build.build()
    return capi.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bonsai_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bns_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from bonsai_b200 import capi
    syms = declared_symbols()
    assert len(syms) >= 25
    assert sorted(capi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_struct_layouts(lib):
    from bonsai_b200 import capi
    assert ctypes.sizeof(capi.Config) == 4 + 4 + 64 + 4 * 4 + 4 + 28
    assert ctypes.sizeof(capi.DbHeader) == 128
    assert lib.bns_b200_version().startswith(b"bonsai_b200")
    assert lib.bns_b200_strerror(-2) == b"CUDA error"


def test_open_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bonsai_b200 import capi
    with pytest.raises(capi.BnsError) as ei:
        capi.Context(k=31)
    assert ei.value.code == -2


def test_no_oracle_in_product():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "bonsai_b200")
    banned = ("pyoracle", "bns_oracle", "libbns_ref", "from oracle", "import oracle", "oracle/")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                for b in banned:
                    assert b not in txt, (fn, b)


def test_value_dictionary_host_code(tmp_path):
    """The loader's distinct-value pass (bonsai_b200/csrc/bns_host_util.h: bitmap for values < 2^24, sort-merge above) against
    sort + unique over khash arrays with empty / deleted / occupied buckets: tests/host/distinct_values.cpp, host code only."""
    import shutil
    import subprocess
    gxx = shutil.which("g++") or "/usr/bin/g++"
    exe = str(tmp_path / "distinct_values")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "distinct_values.cpp")
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe, src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout and r.stdout.count(" ok") == 9, r.stdout


def test_bns_repack_layout():
    """repack (python/bns.cpp:130-150): list entry `kind` lands at flat positions kind * pairs + [0, pairs) of a (pairs, len) array"""
    import numpy as np
    from bonsai_b200 import bns
    r = bns.repack([np.arange(3, dtype=np.float32), 10 + np.arange(3, dtype=np.float32)], 3)
    assert r.shape == (3, 2) and r.dtype == np.float32
    assert r.reshape(-1).tolist() == [0, 1, 2, 10, 11, 12]
