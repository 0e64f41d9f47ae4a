"""Host unit tests of the packed FFT engine (nw_bfly2.cuh, nw_fft2.cuh): the same headers the CUDA kernels use are
compiled with g++ (scalar stand-ins for the packed fp32 instructions) and checked against a naive DFT:
every radix butterfly, decimation-in-frequency and -in-time transforms for a range of 2-3-5-smooth lengths with
run-time plans, and the compile-time-plan drivers of the hot lengths."""
import os
import re
import subprocess

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "engine")


def _run(name):
    exe = os.path.join(HERE, name)
    src = exe + ".cpp"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", exe, src])
    try:
        return subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout
    finally:
        os.remove(exe)


def _errors(text):
    return [float(x) for x in re.findall(r"\d\.\d+e[-+]\d+", text)]


def test_butterflies_match_naive_dft():
    errs = _errors(_run("t_bfly"))
    assert len(errs) >= 44 and max(errs) < 2e-6          # fp32 columns ~1e-7, fp64 ~1e-14 (relative to sums of ~R terms)


def test_runtime_plan_transforms():
    out = _run("t_fft2")
    errs = _errors(out)
    assert len(errs) >= 150 and max(errs) < 5e-6, out


def test_compile_time_plan_transforms():
    out = _run("t_fft2s")
    errs = _errors(out)
    assert len(errs) == 8 and max(errs) < 5e-6, out
