"""`EpochsWavelet` of the reference (`ninwavelets/mneutils.py`) on the device path.

The reference loops over epochs in Python (mneutils.py:37-40); here all epochs of
the channel go to the device as one `[E, T]` batch, and the epoch reductions
(mean power, inter-trial coherence; mneutils.py:53-55, 68-71) run there too.
`epochs` is duck-typed like upstream: `.info['sfreq']`, `.ch_names`, `.get_data()`.
"""
import numpy as np

from . import _backend as _be
from .base import WaveletBase, Numbers


class EpochsWavelet:
    def __init__(self, epochs, wavelet: WaveletBase) -> None:
        self.epochs = epochs
        self.wavelet = wavelet
        wavelet.sfreq = self.epochs.info['sfreq']   # reference mneutils.py:24

    def _waves(self, ch_name: str):
        idx = self.epochs.ch_names.index(ch_name)
        return np.asarray(self.epochs.get_data())[:, idx, :]

    def _device_batch(self, ch_name, freqs, output, baseline=None):
        import torch
        w = self.wavelet
        w._require_cuda()
        waves = self._waves(ch_name)
        plan = w._plan_for(int(waves.shape[-1]), freqs, True)
        x = torch.as_tensor(np.ascontiguousarray(waves, dtype=plan.real_dtype), device="cuda:%d" % plan.device)
        bl = (0, 0, 0)
        if baseline is not None:
            from .base import _window
            lo, hi = _window(int(waves.shape[-1]), w.sfreq, baseline[1], baseline[2])
            bl = (_be.BASELINE_MODES[baseline[0]], lo, hi)
        return plan, plan.transform_device(x, output, *bl)

    def cwt(self, ch_name: str, freqs: Numbers) -> np.ndarray:
        """(E, F, T) complex (reference mneutils.py:26-40)."""
        _, z = self._device_batch(ch_name, freqs, _be.OUT_CWT)
        return z.cpu().numpy()

    def power(self, ch_name: str, freqs: Numbers, *, baseline=None) -> np.ndarray:
        """Mean over epochs of |cwt|**2, (F, T) (reference mneutils.py:42-55)."""
        plan, p = self._device_batch(ch_name, freqs, _be.OUT_POWER, baseline)
        return plan.reduce_epochs_device(p, 0).cpu().numpy()

    def itc(self, ch_name: str, freqs: Numbers) -> np.ndarray:
        """Inter-trial coherence |mean(cwt/|cwt|)|, (F, T) (reference mneutils.py:57-71)."""
        plan, z = self._device_batch(ch_name, freqs, _be.OUT_CWT)
        return plan.reduce_epochs_device(z, 1).cpu().numpy()
