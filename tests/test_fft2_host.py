"""CPU unit test of the in-place large-radix FFT core (polyblur_b200/csrc/fft2.cuh): the
__host__ __device__ stage functions are compiled with g++ and run with one simulated thread,
then compared with a float64 naive DFT (forward, scrambled order) and round-tripped through
the inverse-by-forward transform."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "polyblur_b200", "csrc")
SIZES = [2, 3, 4, 6, 16, 30, 64, 500, 700, 1080, 1152, 1920, 2000, 2016, 2048, 2160, 2240, 2304,
         3840, 4000, 4096, 9000, 12000, 1001, 1331, 2197]


@pytest.fixture(scope="module")
def exe():
    out = os.path.join(CSRC, "build", "fft2_host_test")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O2", "-x", "c++", "-std=c++17", "-I/usr/local/cuda/include",
                    os.path.join(CSRC, "fft2_host_test.cu"), "-o", out], check=True)
    return out


def test_fft2_core_against_naive_dft(exe):
    res = subprocess.run([exe] + [str(n) for n in SIZES], capture_output=True, text=True)
    lines = [l.split() for l in res.stdout.strip().splitlines()]
    assert len(lines) == len(SIZES)
    for n, f in zip(SIZES, lines):
        assert int(f[0]) == n and f[1] != "noplan", f
        assert int(f[1]) <= 4, f"{n}: {f[1]} stages"
        assert float(f[2]) < 5e-7 and float(f[3]) < 5e-7 and f[4] == "1", f
    assert res.returncode == 0


def test_fft2_rejects_large_primes(exe):
    res = subprocess.run([exe, "97", "1104", "9973"], capture_output=True, text=True)
    assert all(l.split()[1] == "noplan" for l in res.stdout.strip().splitlines())
