"""The register-resident FFT passes of the resampler (multimodalsignal_b200/csrc/fft_fast.cuh) on the CPU: the pass bodies are
__host__ __device__, tests/fft_fast_host.cpp emulates a CTA by two loops over the thread index and checks, for power-of-two,
3 * 2^a and 9 * 2^a lengths: the register DFTs against the definition, inverse(forward(x)) = M x, the convolution theorem in the
transform's own frequency order (what the chirp-z resampler relies on), and the fused hooks (chirp loads, filter loads,
multiplied stores, pruned stores) against their unfused equivalents.  The GPU parity of the whole resampler against the
reference's scipy.signal.resample fixtures is tests/test_gpu_preprocess.py."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fft_fast_pass_bodies_on_the_host(tmp_path):
    exe = tmp_path / "fft_fast_host"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "fft_fast_host.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "all ok" in out.stdout
    assert "9x16" in out.stdout and "12x16" in out.stdout and "16x16" in out.stdout and "16x8" in out.stdout and "8x8" in out.stdout
