"""CPU oracle for the frequency-domain CWT path of ninwavelets.

TEST INFRASTRUCTURE ONLY.  This module is the checker, never the product:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  `ninwavelets_b200` never does - its
compute path is the CUDA library and it fails loudly without it.

It is a numpy/scipy restatement (not a copy) of the reference's algorithm,
written as plain functions over a small `Family` record.  Each function cites
the reference lines (relative to `/root/reference/ninwavelets/`) it follows.
The arithmetic itself lives in third-party numpy / scipy.fftpack (unpinned
upstream, `setup.py:6`); conventions restated here: forward DFT unnormalised
with e^{-2 pi i k n / N}, inverse with 1/N and e^{+...}.

PARITY PINNING.  The reference ships no golden vectors, KATs or tests with
assertions, so the oracle is pinned against *outputs of
the reference itself run in the authoring container*: `oracle/gen_golden.py`
imports the unmodified reference through `oracle/refload.py`, writes
`tests/golden/*.npz`, and `tests/test_oracle_golden.py` checks this module
against those fixtures (bit-exact for spectra, <= 4 ulp-of-peak for
transforms) on every CPU run.
"""
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
from scipy.fftpack import fft as _fft, ifft as _ifft


# --------------------------------------------------------------------------
# family description
# --------------------------------------------------------------------------
@dataclass
class Family:
    """Parameters of one wavelet family instance (ctor kwargs of the reference).

    kind: 'morse' | 'morlet' | 'shannon' | 'mexicanhat' | 'haar'
    Defaults follow wavelets.py:38-40 (Morse), :110-113 (Morlet),
    :210-212 (MexicanHat), :247-249 (Shannon), :266-269 (Haar).
    """
    kind: str
    sfreq: float = 1000.0
    b: float = 17.5            # Morse beta
    r: float = 3.0             # Morse gamma
    sigma: float = 7.0         # Morlet / MexicanHat / Shannon (unused by Shannon)
    gabor: bool = False
    real_wave_length: float = 1.0
    interpolate: bool = False
    # Morlet constants, wavelets.py:118-122
    c: float = field(init=False, default=0.0)
    k: float = field(init=False, default=0.0)

    def __post_init__(self):
        if self.kind == "morlet":
            s2 = np.square(self.sigma)
            self.c = np.float_power(1 + np.exp(-s2) - 2 * np.exp(-3 / 4 * s2), -1 / 2)
            self.k = 0 if self.gabor else np.exp(-np.float_power(self.sigma, 2) / 2)

    # reference `mode` (base.py:126-142): Reverse/Both use the analytic
    # spectrum, Normal uses the time-domain formula + FFT.
    @property
    def analytic(self) -> bool:
        return self.kind in ("morse", "morlet", "shannon")


def peak_freq(fam: Family, freq: float) -> float:
    """peak_freq(): base.py:218-219 (1.0), wavelets.py:143-144 (Morlet),
    wavelets.py:227-228 (MexicanHat)."""
    if fam.kind == "morlet":
        return fam.sigma / (1.0 - np.exp(-fam.sigma * freq))
    if fam.kind == "mexicanhat":
        return np.sqrt(6) / np.pi / np.pi
    return 1.0


# --------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------
def trans_grid(fam: Family, freq: float, real_wave_length: float) -> np.ndarray:
    """DFT frequency grid in Hz, base.py:173-194 (`_setup_trans_shape`)."""
    one = 1 / freq
    total = fam.sfreq / freq * real_wave_length
    return np.arange(0, total, one)


def wavelet_timeline(fam: Family, freq: float) -> np.ndarray:
    """Zero-centred time grid of a time-domain wavelet, base.py:196-216 called
    with real_length=1, zero_mean=True from base.py:357."""
    pk = peak_freq(fam, freq)
    total = 1 / pk * freq * 2 * np.pi
    one = 1 / fam.sfreq * 2 * np.pi * freq / pk
    return np.arange(-total / 2, total / 2, one)


# --------------------------------------------------------------------------
# formulas
# --------------------------------------------------------------------------
def analytic_spectrum(fam: Family, grid: np.ndarray, freq: float) -> np.ndarray:
    """trans_formula of the analytic families."""
    if fam.kind == "morse":  # wavelets.py:65-74
        x = grid / freq
        step = np.heaviside(x, x)
        return 2.0 * (step * np.float_power(x, fam.b)
                      * np.exp((fam.b / fam.r) * (1.0 - np.float_power(x, fam.r))))
    if fam.kind == "morlet":  # wavelets.py:132-136
        x = grid / freq * peak_freq(fam, freq)
        return (fam.c * np.float_power(np.pi, -1 / 4)
                * (np.exp(-np.square(fam.sigma - x) / 2) - fam.k * np.exp(-np.square(x) / 2)))
    if fam.kind == "shannon":  # wavelets.py:256-262: 1 where f_k <= 1.0 Hz else 0; freq ignored
        return np.where(grid <= 1.0, 1.0, 0.0)
    raise ValueError("no analytic spectrum for %r" % fam.kind)


def time_formula(fam: Family, timeline: np.ndarray, freq: float) -> np.ndarray:
    """formula() of the time-domain families."""
    if fam.kind == "mexicanhat":  # wavelets.py:219-221
        return ((1 - np.power(timeline / fam.sigma, 2))
                * np.exp(-np.square(timeline) / np.square(fam.sigma) / 2))
    if fam.kind == "haar":  # wavelets.py:272-280
        out = np.zeros_like(timeline)
        out[(0.0 < timeline) & (timeline <= 1.0)] = 1.0
        out[(-1.0 < timeline) & (timeline <= 0.0)] = -1.0
        return out
    raise ValueError("no time-domain formula for %r" % fam.kind)


# --------------------------------------------------------------------------
# spectrum bank (make_fft_wavelet / make_fft_wavelets)
# --------------------------------------------------------------------------
def interpolate_alias(spec: np.ndarray) -> np.ndarray:
    """Zero every bin >= int(N/2), base.py:107-123."""
    half = int(spec.shape[0] / 2)
    out = np.zeros_like(spec)
    out[:half] = spec[:half]
    return out


def pad_to(spec: np.ndarray, n: int) -> np.ndarray:
    """Length fix-up, base.py:75-82: truncate, or centre zero-pad with the
    smaller half in front."""
    m = spec.shape[0]
    if m > n:
        return spec[:n]
    front = (n - m) // 2
    out = np.zeros(n, dtype=spec.dtype)
    out[front:front + m] = spec
    return out


def make_fft_wavelet(fam: Family, freq: float, real_length: float = 1.0) -> np.ndarray:
    """One spectrum, base.py:221-256."""
    if freq == 0:
        raise ZeroDivisionError  # base.py:234-235
    if fam.analytic:
        if fam.interpolate:  # base.py:239-242
            grid = trans_grid(fam, real_length, real_length / 2)
            return np.hstack((analytic_spectrum(fam, grid, freq), np.zeros(len(grid))))
        grid = trans_grid(fam, real_length, real_length)  # base.py:244-246
        return analytic_spectrum(fam, grid, freq)
    # Normal mode, base.py:250-255 (+ 356-358 for the wavelet itself)
    wavelet = time_formula(fam, wavelet_timeline(fam, freq), freq)
    half = int((fam.sfreq * fam.real_wave_length - wavelet.shape[0]) / 2)
    padded = np.hstack((np.zeros(half), wavelet, np.zeros(half)))
    spec = _fft(padded)
    return np.abs(spec.real) + 1j * np.abs(spec.imag)


def make_fft_wavelets(fam: Family, freqs: Sequence[float], real_wave_length: float = 1.0):
    """Spectrum bank, base.py:258-279 (raises IndexError/TypeError like the
    reference when fewer than two indexable freqs are given)."""
    _ = freqs[1] - freqs[0]  # freq_dist, base.py:272
    bank = [make_fft_wavelet(fam, f, real_wave_length) for f in freqs]
    if fam.interpolate:
        bank = [interpolate_alias(w) for w in bank]
    return bank


# --------------------------------------------------------------------------
# transform
# --------------------------------------------------------------------------
def cwt(fam: Family, wave: np.ndarray, freqs: Sequence[float],
        bank: Optional[list] = None) -> np.ndarray:
    """CWT of one 1-D signal, base.py:378-407.  `bank` plays the role of the
    reference's `self.fft_wavelets` cache (reuse=True)."""
    n = wave.shape[0]
    if bank is None:
        bank = make_fft_wavelets(fam, freqs, n / fam.sfreq)
    stack = np.array([pad_to(w, n) for w in bank])
    spectrum = _fft(wave)
    if fam.interpolate:
        spectrum = interpolate_alias(spectrum)
    return _ifft(stack * spectrum)


def cwt_abs(fam, wave, freqs, bank=None):
    """base.py:427-443."""
    return np.abs(cwt(fam, wave, freqs, bank))


def power(fam, wave, freqs, bank=None):
    """base.py:409-425."""
    return cwt_abs(fam, wave, freqs, bank) ** 2


# --------------------------------------------------------------------------
# Baseline, base.py:18-20, 23-68 (1-D)
# --------------------------------------------------------------------------
BASELINE_MODES = ("mean", "ratio", "percent", "log", "zscore", "zlog")


def baseline(wave: np.ndarray, sfreq: float, start: float, stop: float, mode: str) -> np.ndarray:
    seg = wave[int(start * sfreq): int(stop * sfreq)]  # base.py:49
    m = seg.mean()                                     # base.py:50
    if mode == "mean":
        return wave - m                                # :52-53
    if mode == "ratio":
        return wave / m                                # :55-56
    if mode == "percent":
        return (wave - m) / m                          # :58-59
    if mode == "log":
        return np.log10(wave / m)                      # :61-62
    if mode == "zscore":
        return (wave - m) / np.std(seg)                # :64-65
    if mode == "zlog":
        return np.log10(wave / m) / np.std(seg)        # :67-68
    raise ValueError(mode)


def baseline_rows(rows: np.ndarray, sfreq, start, stop, mode) -> np.ndarray:
    """Apply `baseline` independently to every row of a (..., T) array - the
    per-(signal, frequency)-row reading used by BASELINE.json config 3."""
    flat = rows.reshape(-1, rows.shape[-1])
    out = np.stack([baseline(r, sfreq, start, stop, mode) for r in flat])
    return out.reshape(rows.shape)


# --------------------------------------------------------------------------
# epochs, mneutils.py:22-24, 37-40, 53-55, 68-71
# --------------------------------------------------------------------------
def epochs_cwt(fam: Family, epochs_1ch: np.ndarray, freqs) -> np.ndarray:
    """(E,T) -> (E,F,T) complex; spectra built once on the first epoch."""
    bank = make_fft_wavelets(fam, freqs, epochs_1ch.shape[1] / fam.sfreq)
    return np.array([cwt(fam, w, freqs, bank) for w in epochs_1ch])


def epochs_power(fam, epochs_1ch, freqs) -> np.ndarray:
    return np.mean(np.abs(epochs_cwt(fam, epochs_1ch, freqs)) ** 2, axis=0)


def epochs_itc(fam, epochs_1ch, freqs) -> np.ndarray:
    z = epochs_cwt(fam, epochs_1ch, freqs)
    return np.abs(np.mean(z / np.abs(z), axis=0))


# --------------------------------------------------------------------------
# synthetic workloads shared by tests and bench
# --------------------------------------------------------------------------
def readme_sine(n: int = 300, sfreq: float = 1000.0, f0: float = 60.0) -> np.ndarray:
    t = np.arange(n) / sfreq
    return np.sin(2 * np.pi * f0 * t)


def eeg_like(n_ch: int, n: int, sfreq: float = 1000.0, seed: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sfreq
    x = rng.standard_normal((n_ch, n))
    for f0 in (10.0, 40.0, 60.0):
        ph = rng.uniform(0, 2 * np.pi, size=(n_ch, 1))
        x += np.sin(2 * np.pi * f0 * t[None, :] + ph)
    return x


def meg_epochs_like(n_sig: int, n: int = 1500, sfreq: float = 1000.0, seed: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sfreq
    x = rng.standard_normal((n_sig, n))
    burst = ((t >= 0.5) & (t < 1.0)) * np.sin(2 * np.pi * 10.0 * t)
    return x + 2.0 * burst[None, :]
