"""CPU check of the minimizer-bucketed table layout's bijection (no GPU: the functions are __host__ __device__)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_minimizer_layout_bijection(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "loc_roundtrip")
    src = os.path.join(ROOT, "tests", "host", "loc_roundtrip.cu")
    r = subprocess.run([nvcc, "-std=c++17", "-O2", "-w", "-o", exe, src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    lines = r.stdout.strip().splitlines()
    checked = [ln for ln in lines if ln.startswith("k=")]
    assert len(checked) >= 10 and all(ln.endswith("bad=0") for ln in checked), r.stdout
    lpr = float([ln for ln in lines if ln.startswith("lines_per_read=")][0].split("=")[1])
    assert lpr < 45.0          # the 120 consecutive 31-mers of a read have their homes in ~30 lines (15 groups of two) instead of 120
