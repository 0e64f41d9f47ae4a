"""Host-side mirror of the reference's `ninwavelets/base.py` for the CWT path.

Same class names, constructor arguments, attributes and error behaviour as the
reference (`WaveletBase`, `WaveletMode`, `Baseline`, `plot_tf`; reference
base.py:23-68, 126-142, 145-446), but `cwt / abs / power` make exactly one call
into the CUDA library (`_backend.Plan`) instead of the numpy/cupy sequence of
base.py:394-406.  Only `cuda=True` objects can transform: this package has no
CPU implementation and raises rather than fall back.

Extensions (keyword-only, all default to the reference's behaviour):
  dtype      'float64' (reference precision) or 'float32' (the fast mode)
  device     CUDA ordinal (default: current torch device)
  prune_eps  spectrum bins below prune_eps * peak are skipped (None: library default)
  resample   long rows, abs / power: None or True lets the library compute band-limited rows at a
             decimated length and interpolate them ("resampled rows"); False forces the
             exact length-N inverse transform for every row
  resample_tol  bound on the interpolation error of a resampled row relative to the wavelet's peak
             gain times the input's spectral mass (0: library default, 1e-6 in float32, 5e-14 in float64)
and `wave` may be a 2-D `[S, N]` batch or a torch CUDA tensor (then the result
stays on the device).
"""
from enum import Enum
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _backend as _be

Numbers = Union[List[float], np.ndarray, range]


class WaveletMode(Enum):
    """Which formula a family provides (reference base.py:126-142)."""
    Normal = 0            # time-domain formula only
    Both = 1              # time-domain and Fourier-domain formulas
    Reverse = 2           # Fourier-domain formula only
    Indifferentiable = 3
    Twice = 4


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _window(n: int, sfreq: float, start: float, stop: float):
    """Sample range of `wave[int(start*sfreq):int(stop*sfreq)]` (reference base.py:49)."""
    lo, hi, _ = slice(int(start * sfreq), int(stop * sfreq)).indices(n)
    return lo, max(lo, hi)


class Baseline:
    """Baseline correction of one 1-D wave (reference base.py:23-68) on the device.

    `wave` may be a numpy array or a torch CUDA tensor; rows of a 2-D array are
    corrected independently (the per-(signal, frequency)-row use of config 3).
    """

    def __init__(self, wave, sfreq: float, start: float, stop: float, *, device: Optional[int] = None) -> None:
        self.wave = wave
        self.sfreq = sfreq
        self.start, self.stop = start, stop
        self._device = device

    def _run(self, mode: str):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ninwavelets_b200.Baseline needs a CUDA device; there is no CPU path")
        from ._rows import baseline_rows
        return baseline_rows(self.wave, self.sfreq, self.start, self.stop, mode, self._device)

    def mean(self):
        return self._run("mean")

    def ratio(self):
        return self._run("ratio")

    def percent(self):
        return self._run("percent")

    def log(self):
        return self._run("log")

    def zscore(self):
        return self._run("zscore")

    def zlog(self):
        return self._run("zlog")


class _LazyBank(Sequence):
    """`WaveletBase.fft_wavelets`: list-like view of the plan's spectra.

    The reference keeps F host arrays (base.py:276-278); here the spectra live
    only as per-frequency parameters on the device and are materialised (by the
    device, `nwcwt_spectrum_bank`) when somebody actually indexes the cache.
    """

    def __init__(self, plan: "_be.Plan", real: bool):
        self.plan = plan
        self._real = real
        self._host = None

    def _materialise(self):
        if self._host is None:
            bank = self.plan.spectrum_bank_device().cpu().numpy().astype(np.complex128)
            self._host = bank.real.copy() if self._real else bank
        return self._host

    def __len__(self):
        return self.plan.n_freqs

    def __getitem__(self, i):
        return self._materialise()[i]


class WaveletBase:
    """Base class of the wavelet families (reference base.py:145-446).

    Subclasses set `self.mode` and override `trans_formula` / `formula` /
    `peak_freq` exactly as in the reference (README.md:342-355).  The built-in
    families are evaluated in registers on the device; a user subclass whose
    numpy formula the device cannot know is tabulated once per plan on the host
    (by the user's own code) and uploaded.
    """

    _native = None   # set by built-in families: callable(self) -> dict(family=..., p0=..., ...)

    def __init__(self, sfreq: float = 1000, real_wave_length: float = 1., interpolate: bool = True,
                 cuda: bool = False, *, dtype="float64", device: Optional[int] = None,
                 prune_eps: Optional[float] = None, resample: Optional[bool] = None,
                 resample_tol: float = 0.0) -> None:
        self.mode: WaveletMode = WaveletMode.Normal
        self.sfreq: float = sfreq
        self.help: str = ''
        self.real_wave_length: float = real_wave_length
        self.freq_dist: float
        self.interpolate = interpolate
        self.cuda = cuda
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64")
        self.device = device
        self.prune_eps = prune_eps
        self.resample = resample
        self.resample_tol = resample_tol
        self._plan: Optional[_be.Plan] = None

    # ---- formulas a family may override (reference base.py:218-219, 281-344) ---
    def peak_freq(self, freq: float) -> float:
        return 1.

    def formula(self, timeline: np.ndarray, freq: float) -> np.ndarray:
        return timeline

    def trans_formula(self, freqs: np.ndarray, freq: float = 1.) -> np.ndarray:
        return freqs

    def cp_trans_formula(self, freqs, freq: float = 1.):
        # the reference's cupy twin (base.py:324-344); the device path does not use it
        return self.trans_formula(freqs, freq)

    # ---- grids (reference base.py:173-216) ----------------------------------------
    def _setup_trans_shape(self, freq: float, real_wave_length: float, cuda=False) -> np.ndarray:
        one = 1 / freq
        total = self.sfreq / freq * real_wave_length
        return np.arange(0, total, one)

    def _setup_waveletshape(self, freq: float, real_length: float = 1, zero_mean: bool = False) -> np.ndarray:
        pk = self.peak_freq(freq)
        total = real_length / pk * freq * 2 * np.pi
        one = 1 / self.sfreq * 2 * np.pi * freq / pk
        if zero_mean:
            return np.arange(-total / 2, total / 2, one)
        return np.arange(0, total, one)

    # ---- time-domain wavelets (reference base.py:346-376) ----------------------------
    def make_wavelet(self, freq: float) -> np.ndarray:
        if freq == 0:
            raise ZeroDivisionError
        if self.mode in (WaveletMode.Reverse, WaveletMode.Twice):
            raise NotImplementedError(
                "time-domain synthesis of Fourier-only families (reference base.py:349-355) is plot/MNE "
                "interop and outside the device path of this package")
        timeline = self._setup_waveletshape(freq, 1, zero_mean=True)
        return self.formula(timeline, freq)

    def make_wavelets(self, freqs: Numbers):
        self.wavelets = [self.make_wavelet(f) for f in freqs]
        return self.wavelets

    # ---- plan construction -----------------------------------------------------------
    def _device_index(self) -> int:
        if self.device is not None:
            return int(self.device)
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ninwavelets_b200 needs a CUDA device (B200); there is no CPU path")
        return torch.cuda.current_device()

    def _require_cuda(self):
        if not self.cuda:
            raise RuntimeError(
                "%s(cuda=False): ninwavelets_b200 implements only the device path of the reference "
                "(cuda=True); it has no numpy/scipy fallback" % type(self).__name__)

    def _uses_native(self) -> bool:
        """The family's spectrum is one of the built-in analytic formulas (real valued), evaluated by the device."""
        return self.mode in (WaveletMode.Reverse, WaveletMode.Both) and self._native is not None \
            and not self._is_overridden("trans_formula") and not self._is_overridden("peak_freq")

    def _is_overridden(self, name: str) -> bool:
        """True if a user subclass replaced a formula of a built-in family."""
        for klass in type(self).__mro__:
            if name in klass.__dict__:
                return not klass.__module__.startswith(__package__)
        return False

    def _normal_mode_tables(self, freqs):
        """Spectra of time-domain families (reference base.py:250-255): wavelet, symmetric zero
        padding to sfreq*real_wave_length samples, forward FFT, |re| + i|im|."""
        rows = []
        for f in freqs:
            if f == 0:
                raise ZeroDivisionError
            w = np.asarray(self.make_wavelet(f), dtype=np.float64)
            half = int((self.sfreq * self.real_wave_length - w.shape[0]) / 2)
            rows.append(np.hstack((np.zeros(half), w, np.zeros(half))))
        lens = np.array([r.shape[0] for r in rows], dtype=np.int64)
        table = np.zeros((len(rows), int(lens.max())), dtype=np.complex128)
        # plan-time precomputation of F short tables (sfreq * real_wave_length samples each), done on the host like
        # the reference does (base.py:253) so that it works for every table length
        for i, r in enumerate(rows):
            spec = np.fft.fft(r)
            table[i, :lens[i]] = np.abs(spec.real) + 1j * np.abs(spec.imag)
        return table, lens

    def _build_plan(self, freqs, n: int) -> "_be.Plan":
        freqs_arr = np.asarray(list(freqs) if not isinstance(freqs, np.ndarray) else freqs, dtype=np.float64)
        if np.any(freqs_arr == 0):
            raise ZeroDivisionError  # reference base.py:234-235
        common = dict(device=self._device_index(), dtype=self.dtype, interpolate=self.interpolate, n=n,
                      sfreq=self.sfreq, freqs=freqs_arr, prune_eps=self.prune_eps, resample=self.resample,
                      resample_tol=self.resample_tol)
        analytic = self.mode in (WaveletMode.Reverse, WaveletMode.Both)
        if self._uses_native():
            return _be.Plan(**common, **self._native(freqs_arr))
        if analytic:
            # user-supplied numpy trans_formula: tabulate it on the reference's grid (base.py:239-246)
            L = n / self.sfreq
            rows = []
            for f in freqs_arr:
                if self.interpolate:
                    grid = self._setup_trans_shape(L, L / 2)
                    rows.append(np.hstack((np.asarray(self.trans_formula(grid, f)), np.zeros(len(grid)))))
                else:
                    grid = self._setup_trans_shape(L, L)
                    rows.append(np.asarray(self.trans_formula(grid, f)))
            lens = np.array([r.shape[0] for r in rows], dtype=np.int64)
            table = np.zeros((len(rows), int(lens.max())), dtype=np.complex128)
            for i, r in enumerate(rows):
                table[i, :lens[i]] = r
            return _be.Plan(**common, family=_be.TABLE, table=table, table_lens=lens)
        table, lens = self._normal_mode_tables(freqs_arr)
        return _be.Plan(**common, family=_be.TABLE, table=table, table_lens=lens)

    # ---- the reference's cache API ------------------------------------------------------
    def make_fft_wavelet(self, freq: float, real_length: float = 1.) -> np.ndarray:
        """One spectrum (reference base.py:221-256), evaluated by the device."""
        self._require_cuda()
        if freq == 0:
            raise ZeroDivisionError
        n = int(round(real_length * self.sfreq))
        plan = self._build_plan([freq], n)
        bank = plan.spectrum_bank_device().cpu().numpy().astype(np.complex128)[0]
        plan.close()
        return bank.real.copy() if self._uses_native() else bank   # tabulated (user / Normal-mode) spectra may be complex

    def make_fft_wavelets(self, freqs: Numbers, real_wave_length: float = 1.):
        """Build the plan that replaces the spectrum cache (reference base.py:258-279)."""
        self._require_cuda()
        self.freq_dist = freqs[1] - freqs[0]   # IndexError / TypeError like the reference (base.py:272)
        n = int(round(real_wave_length * self.sfreq))
        self._plan = self._build_plan(freqs, n)
        self.fft_wavelets = _LazyBank(self._plan, self._uses_native())
        return self.fft_wavelets

    def _plan_for(self, n: int, freqs, reuse: bool) -> "_be.Plan":
        if (not reuse) or (not hasattr(self, 'fft_wavelets')):
            self.freq_dist = freqs[1] - freqs[0]
            self._plan = self._build_plan(freqs, n)
            self.fft_wavelets = _LazyBank(self._plan, self._uses_native())
        elif self._plan.n != n:
            # reference quirk (base.py:396-397): a cached bank built for another length is
            # pad_to'ed (truncated / centre padded) onto the new signal
            old = np.asarray(self.fft_wavelets._materialise(), dtype=np.complex128)
            self._plan = _be.Plan(device=self._device_index(), dtype=self.dtype, interpolate=self.interpolate,
                                  n=n, sfreq=self.sfreq, freqs=self._plan.freqs, family=_be.TABLE, table=old,
                                  prune_eps=self.prune_eps, resample=self.resample, resample_tol=self.resample_tol)
        return self._plan

    # ---- transforms (reference base.py:378-443) -------------------------------------------
    def _transform(self, wave, freqs, reuse, output, baseline=None):
        self._require_cuda()
        on_device = _is_torch(wave)
        if on_device:
            sig = wave if wave.dim() == 2 else wave.reshape(1, -1)
            single = wave.dim() == 1
        else:
            arr = np.asarray(wave)
            single = arr.ndim == 1
            sig = arr.reshape(1, -1) if single else arr
        if len(sig.shape) != 2:
            raise ValueError("wave must be 1-D (reference) or 2-D [signals, samples]")
        n = int(sig.shape[-1])
        plan = self._plan_for(n, freqs, reuse)
        bl_mode, lo, hi = 0, 0, 0
        if baseline is not None:
            mode, start, stop = baseline
            bl_mode = _be.BASELINE_MODES[mode]
            lo, hi = _window(n, self.sfreq, start, stop)
        if on_device:
            out = plan.transform_device(sig, output, bl_mode, lo, hi)
        else:
            out = plan.transform_host(sig, output, bl_mode, lo, hi)
        return out[0] if single else out

    def cwt(self, wave, freqs: Union[Numbers, None], reuse: bool = True):
        """Complex CWT, (F, N) for a 1-D wave (reference base.py:378-407)."""
        return self._transform(wave, freqs, reuse, _be.OUT_CWT)

    def abs(self, wave, freqs: Union[Numbers, None] = None, reuse: bool = True):
        """|cwt| (reference base.py:427-443)."""
        return self._transform(wave, freqs, reuse, _be.OUT_ABS)

    def power(self, wave, freqs: Union[Numbers, None] = None, reuse: bool = True, *, baseline=None):
        """|cwt|**2 (reference base.py:409-425).  `baseline=(mode, start, stop)` additionally applies
        `Baseline(row, sfreq, start, stop).<mode>()` (base.py:46-68) to every output row in the kernel."""
        return self._transform(wave, freqs, reuse, _be.OUT_POWER, baseline)

    def plot(self, freq: float, show: bool = True):
        return plot_wavelet(self, freq, show)


def plot_wavelet(wavelet_obj: WaveletBase, freq: float, show: bool = True):
    """Display helper (reference base.py:449-489); matplotlib is imported lazily."""
    import matplotlib.pyplot as plt
    w = np.asarray(wavelet_obj.make_wavelets(np.array([freq]))[0])
    fig = plt.figure(figsize=(6, 8))
    ax = fig.add_subplot(2, 1, 1)
    ax.plot(np.arange(w.shape[0]), w.real)
    ax.plot(np.arange(w.shape[0]), w.imag)
    ax3 = fig.add_subplot(2, 1, 2, projection='3d')
    ax3.scatter3D(w.real, np.arange(w.shape[0]), w.imag)
    if show:
        plt.show()
    return fig


def plot_tf(data: np.ndarray, sfreq: float = 1000, frange=None, trange=None, vmin=None, vmax=None,
            cmap: str = 'RdBu_r', show: bool = True):
    """Time-frequency image (reference base.py:492-520); matplotlib is imported lazily."""
    import matplotlib.pyplot as plt
    from mpl_toolkits.axes_grid1 import make_axes_locatable
    fig = plt.figure()
    ax = fig.add_subplot(1, 1, 1)
    if frange is not None:
        step = frange[2] / (frange[1] - frange[0]) * data.shape[0]
        plt.yticks(np.arange(0, data.shape[0], step), np.arange(*frange))
    if trange is not None:
        plt.xticks(np.arange(0, data.shape[1], sfreq * trange[2]), np.arange(*trange))
    image = ax.imshow(data, vmin=vmin, vmax=vmax, cmap=cmap)
    ax.invert_yaxis()
    ax.set_aspect('auto')
    cax = make_axes_locatable(ax).new_horizontal(size="2%", pad=0.05)
    fig.add_axes(cax)
    plt.colorbar(image, cax=cax)
    if show:
        plt.show()
    return ax
