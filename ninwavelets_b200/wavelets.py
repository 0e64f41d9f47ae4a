"""Wavelet families of the reference (`ninwavelets/wavelets.py`) on the B200 backend.

Constructor signatures and attributes are the reference's (wavelets.py:38-40,
110-113, 210-212, 247-249, 266-269).  The numpy `trans_formula` / `formula`
methods are kept because they are the package's plugin API (README.md:342-355)
and are what a user subclass overrides; the transform itself never calls them
for the built-in analytic families - Morse, Morlet and Shannon spectra are
generated in registers by the CUDA kernels from the parameters `_native` hands
to the plan.
"""
from typing import Optional

import numpy as np

from . import _backend as _be
from .base import WaveletBase, WaveletMode


class Morse(WaveletBase):
    """Generalized Morse wavelets, beta `b`, gamma `r` (reference wavelets.py:7-74)."""

    def __init__(self, sfreq: float = 1000, b: float = 17.5, r: float = 3, real_wave_length: float = 1.,
                 interpolate: bool = False, cuda: bool = False, **kw) -> None:
        super().__init__(sfreq, real_wave_length, interpolate, cuda, **kw)
        self.r: float = r
        self.b: float = b
        self.mode = WaveletMode.Reverse
        self.help = 'Generalized Morse wavelets are defined in the Fourier domain.'

    def _native(self, freqs):
        return dict(family=_be.MORSE, p0=self.b, p1=self.r)

    def trans_formula(self, freqs: np.ndarray, freq: float = 1.) -> np.ndarray:
        # reference wavelets.py:65-74
        x = freqs / freq
        return 2. * (np.heaviside(x, x) * np.float_power(x, self.b)
                     * np.exp((self.b / self.r) * (1. - np.float_power(x, self.r))))


class Morlet(WaveletBase):
    """Morlet / Gabor wavelets (reference wavelets.py:77-144)."""

    def __init__(self, sfreq: float = 1000, sigma: float = 7., real_wave_length: float = 1., gabor: bool = False,
                 interpolate: bool = False, cuda: bool = False, **kw) -> None:
        super().__init__(sfreq, real_wave_length, interpolate, cuda, **kw)
        self.mode = WaveletMode.Both
        self.sigma = sigma
        s2 = np.square(self.sigma)
        self.c = np.float_power(1 + np.exp(-s2) - 2 * np.exp(-3 / 4 * s2), -1 / 2)   # wavelets.py:118-121
        self.k = 0 if gabor else np.exp(-np.float_power(self.sigma, 2) / 2)            # wavelets.py:122

    def _native(self, freqs):
        return dict(family=_be.MORLET, p0=self.sigma, p1=self.c * np.float_power(np.pi, -1 / 4), p2=self.k,
                    aux=np.array([self.peak_freq(f) for f in freqs], dtype=np.float64))

    def trans_formula(self, freqs: np.ndarray, freq: float = 1) -> np.ndarray:
        # reference wavelets.py:132-136
        x = freqs / freq * self.peak_freq(freq)
        return (self.c * np.float_power(np.pi, -1 / 4)
                * (np.exp(-np.square(self.sigma - x) / 2) - self.k * np.exp(-np.square(x) / 2)))

    def formula(self, timeline: np.ndarray, freq: float = 1) -> np.ndarray:
        # reference wavelets.py:138-141
        return (self.c * np.float_power(np.pi, (-1 / 4)) * np.exp(-np.square(timeline) / 2)
                * (np.exp(self.sigma * 1j * timeline) - self.k))

    def peak_freq(self, freq: float) -> float:
        return self.sigma / (1. - np.exp(-self.sigma * freq))   # reference wavelets.py:143-144


class MorseMNE(Morse):
    """Reference wavelets.py:147-191: Morse through `mne.time_frequency.tfr.cwt`.  Needs mne and the
    time-domain synthesis that is outside this package's device path; kept for the export list only."""

    def cwt(self, wave, freqs, use_fft: bool = True, mode: str = 'same', decim: float = 1):
        raise NotImplementedError("MorseMNE (deprecated upstream, wavelets.py:149-153) is not part of the "
                                  "device path; use Morse(cuda=True)")


class MexicanHat(WaveletBase):
    """Mexican-hat wavelets, time-domain formula (reference wavelets.py:194-228)."""

    def __init__(self, sfreq: float = 1000, sigma: float = 7, real_wave_length: float = 1.,
                 interpolate: bool = False, cuda: bool = False, **kw) -> None:
        super().__init__(sfreq, real_wave_length, interpolate, cuda, **kw)
        self.sigma: float = sigma
        self.mode = WaveletMode.Normal
        self.help = ''

    def formula(self, tc: np.ndarray, freq: float = 1) -> np.ndarray:
        # reference wavelets.py:219-221
        return (1 - np.power(tc / self.sigma, 2)) * np.exp(-np.square(tc) / np.square(self.sigma) / 2)

    def peak_freq(self, freq: float) -> float:
        return np.sqrt(6) / np.pi / np.pi   # reference wavelets.py:227-228


class Shannon(WaveletBase):
    """Shannon "wavelets" as implemented upstream: a 1 Hz brick-wall low-pass that ignores `freq`
    and `sigma` (reference wavelets.py:231-262)."""

    def __init__(self, sfreq: float = 1000, sigma: float = 7, real_wave_length: float = 1.,
                 interpolate: bool = False, cuda: bool = False, **kw) -> None:
        super().__init__(sfreq, real_wave_length, interpolate, cuda, **kw)
        self.sigma: float = sigma
        self.mode = WaveletMode.Reverse
        self.help = ''

    def _native(self, freqs):
        return dict(family=_be.SHANNON)

    def trans_formula(self, tc: np.ndarray, freq: float = 1) -> np.ndarray:
        return np.where(tc <= 1., 1., 0.)   # reference wavelets.py:256-262


class Haar(WaveletBase):
    """Haar wavelets, time-domain formula (reference wavelets.py:265-280)."""

    def __init__(self, sfreq: float = 1000, real_wave_length: float = 1., interpolate: bool = False, **kw) -> None:
        kw.setdefault("cuda", True)   # the reference's Haar has no cuda switch (wavelets.py:266-269)
        cuda = kw.pop("cuda")
        super().__init__(sfreq, real_wave_length, interpolate, cuda, **kw)
        self.mode = WaveletMode.Normal

    def formula(self, timeline: np.ndarray, freq: float = 1) -> np.ndarray:
        out = np.zeros_like(timeline)
        out[(0. < timeline) & (timeline <= 1.)] = 1.
        out[(-1. < timeline) & (timeline <= 0.)] = -1.
        return out
