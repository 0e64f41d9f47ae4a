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


def peak_rel_err(out, ref):
    """fp64 tolerance metric: max|out-ref| / max|ref| per row."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    num = np.abs(out - ref).max(axis=-1)
    den = np.abs(ref).max(axis=-1)
    return num / np.where(den > 0, den, 1.0)


def l2_rel_err(out, ref):
    """fp32 tolerance metric: ||out-ref||_2 / ||ref||_2 per row."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    num = np.sqrt((np.abs(out - ref) ** 2).sum(axis=-1))
    den = np.sqrt((np.abs(ref) ** 2).sum(axis=-1))
    return num / np.where(den > 0, den, 1.0)
