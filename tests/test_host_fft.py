"""CPU: the Stockham butterflies of csrc/tru_fft.cuh (the exact code the CUDA kernels run)
compiled for the host and checked against a double-precision DFT."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_fft_butterflies_on_host(tmp_path):
    exe = str(tmp_path / "host_fft_test")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", "host_fft_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "OK" in out.stdout
