"""The fused kernel's FFT core (auditory_b200/csrc/aud_fft_core.cuh: packed DFT-20s, exchange layout, pass-2
lane assignment, real-pair split, self-paired rows) emulated lane by lane on the CPU against a float64 DFT.
No GPU needed: the header is __host__ __device__ and tests/cpp/fft_core_emul.cu runs its host side."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "fft_core_emul")


@pytest.fixture(scope="module")
def exe():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I", os.path.join(ROOT, "auditory_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "fft_core_emul.cu"), "-o", EXE])
    return EXE


@pytest.mark.parametrize("seed", [1, 2, 3, 99])
def test_one_warp_round_matches_a_float64_dft(exe, seed):
    out = subprocess.run([exe, str(seed)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
