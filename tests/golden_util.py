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
# Both are per (signal, frequency) ROW and relative to that row ITSELF: no floor by default.  The only cases that pass
# a floor are the pure-sine README inputs (`readme_*`): every row away from the sine's frequency holds nothing but the
# rounding noise of the forward FFT (~1e-17 of the transform's peak in the reference itself; numpy.fft and
# scipy.fftpack already disagree there by O(1) relative), so those rows are measured against `floor` x the largest row
# of the same transform - README_FLOOR_F64 / README_FLOOR_F32, i.e. an absolute bar of 1e-15 / 1e-7 of the peak.
README_FLOOR_F64 = 1e-3
README_FLOOR_F32 = 1e-2


def case_floor(name, f32=False):
    """Denominator floor of a golden case: zero except for the pure-sine README cases."""
    if str(name).startswith("readme"):
        return README_FLOOR_F32 if f32 else README_FLOOR_F64
    return 0.0


# Representability: a row whose reference values lie below ~1e-30 (fp32) / ~1e-290 (fp64) cannot be held to a RELATIVE
# precision by the arithmetic type at all (float32 has no normal numbers below 1.2e-38; the golden cases contain rows
# of peak 6e-84 and 8e-304 - a Morse wavelet at 1 Hz applied to 0.3 s of signal).  Such rows are held to the same
# absolute bar instead: |out - ref| <= tol * ABS_TINY.
ABS_TINY_F32 = 1e-30
ABS_TINY_F64 = 1e-290


def peak_rel_err(out, ref, floor=0.0, tiny=ABS_TINY_F64):
    """fp64 metric: max|out-ref| / max|ref| per row."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    num = np.abs(out - ref).max(axis=-1)
    den = np.abs(ref).max(axis=-1)
    if floor > 0 and den.size:
        den = np.maximum(den, floor * den.max())
    return num / np.maximum(den, tiny)


def l2_rel_err(out, ref, floor=0.0, tiny=ABS_TINY_F32):
    """fp32 metric: ||out-ref||_2 / ||ref||_2 per row."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    num = np.sqrt((np.abs(out - ref) ** 2).sum(axis=-1))
    den = np.sqrt((np.abs(ref) ** 2).sum(axis=-1))
    if floor > 0 and den.size:
        den = np.maximum(den, floor * den.max())
    return num / np.maximum(den, tiny * np.sqrt(ref.shape[-1]))
