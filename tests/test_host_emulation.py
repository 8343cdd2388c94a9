"""Host-side emulation of the device FFT plans and of the phase stepper (no GPU): the device
headers are compiled for the host with nvcc/g++ and every "thread" is run in a loop, so the
index arithmetic, shared-memory layouts and twiddle logic of fft_core.cuh / fft32_core.cuh are
checked against a float64 FFT, and phase.h against the reference's running float sum."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host")


def _build_run(src, tmp_path, compiler, extra=()):
    exe = str(tmp_path / (os.path.basename(src) + ".bin"))
    if compiler == "nvcc":
        cmd = ["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", *extra, "-o", exe, src]
    else:
        cmd = ["g++", "-O2", "-fno-fast-math", "-o", exe, src, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return subprocess.run([exe], capture_output=True, text=True)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_fft_core_plans_on_host(tmp_path):
    r = _build_run(os.path.join(HOST, "emu_fft_core.cu"), tmp_path, "nvcc")
    assert r.returncode == 0, r.stdout[-2000:]
    assert "WORST" in r.stdout


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_fft32_core_plans_on_host(tmp_path):
    r = _build_run(os.path.join(HOST, "emu_fft32_core.cu"), tmp_path, "nvcc")
    assert r.returncode == 0, r.stdout
    lines = [l.split() for l in r.stdout.strip().splitlines()]
    assert [int(l[0]) for l in lines] == [10, 11, 12, 13, 14], r.stdout
    for l in lines:
        assert len(l) == 3 and float(l[1]) < 3e-7, r.stdout      # rel rms vs float64


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_fft1_pipe_math_and_queue_on_host(tmp_path):
    """fft1_pipe.cuh: two-pass column / row transforms through both exchanges, the four-step index
    mapping through the transposed intermediate, and the dependency order of the work queue."""
    r = _build_run(os.path.join(HOST, "emu_fft1_pipe.cu"), tmp_path, "nvcc",
                   extra=("-gencode", "arch=compute_100a,code=sm_100a"))
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "queue ok", r.stdout
    sizes = [l.split() for l in lines[:-1]]
    assert [int(l[0]) for l in sizes] == [15, 16, 17, 18, 19, 20], r.stdout
    for l in sizes:
        assert float(l[1]) < 4e-7, r.stdout                      # rel rms vs float64


def test_phase_stepper_equals_running_sum(tmp_path):
    r = _build_run(os.path.join(HOST, "phase_stepper.cpp"), tmp_path, "g++")
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout
