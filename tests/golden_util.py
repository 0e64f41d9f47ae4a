"""Helpers shared by the golden-fixture tests (CPU oracle and GPU parity)."""
import numpy as np


def broadband(n, seed):
    """Same generator as oracle/gen_golden.py:broadband."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 1000.0
    return rng.standard_normal(n) + np.sin(2 * np.pi * 10 * t) + 0.5 * np.sin(2 * np.pi * 60 * t + 1.0)


def case_wave(case):
    if "wave" in case:
        return np.asarray(case["wave"])
    return broadband(int(case["n"]), int(case["seed"]))


# Tolerance metrics (BASELINE.json north_star / SURVEY.md 8d).
#
# Both are per (signal, frequency) ROW and relative to that row, with a floor on the
# denominator: a row whose peak is below `floor` x the largest row of the same transform is
# measured against floor x that largest row instead.  Such rows hold nothing but the rounding
# noise of the forward FFT (e.g. the README case, a pure 60 Hz sine: every row away from 60 Hz
# is ~1e-17 in the reference itself, and numpy.fft vs scipy.fftpack already disagree there by
# O(1) relative) - no implementation can match them to 1e-12 of their own peak.  With the
# default floors the absolute bar on those rows is 1e-15 (fp64) / 1e-7 (fp32) of the transform's
# peak, i.e. the rounding floor of the arithmetic.
def peak_rel_err(out, ref, floor=1e-3):
    """fp64 metric: max|out-ref| / max|ref| per row."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    num = np.abs(out - ref).max(axis=-1)
    den = np.abs(ref).max(axis=-1)
    den = np.maximum(den, floor * den.max()) if den.size else den
    return num / np.where(den > 0, den, 1.0)


def l2_rel_err(out, ref, floor=1e-2):
    """fp32 metric: ||out-ref||_2 / ||ref||_2 per row."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    num = np.sqrt((np.abs(out - ref) ** 2).sum(axis=-1))
    den = np.sqrt((np.abs(ref) ** 2).sum(axis=-1))
    den = np.maximum(den, floor * den.max()) if den.size else den
    return num / np.where(den > 0, den, 1.0)
