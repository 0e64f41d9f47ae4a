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

    def _fused(self, ch_name, freqs, kind):
        """Mean power (kind 0) / inter-trial coherence (kind 1) with the epoch reduction inside the transform kernel;
        None when the rows are too long for the fused kernel (the caller then reduces the materialised rows)."""
        import torch
        w = self.wavelet
        w._require_cuda()
        waves = self._waves(ch_name)
        plan = w._plan_for(int(waves.shape[-1]), freqs, True)
        if plan.info()["path"] != "short_packed":
            return None
        x = torch.as_tensor(np.ascontiguousarray(waves, dtype=plan.real_dtype), device="cuda:%d" % plan.device)
        return plan.transform_epochs_device(x[None], kind)[0].cpu().numpy()

    def cwt(self, ch_name: str, freqs: Numbers) -> np.ndarray:
        """(E, F, T) complex (reference mneutils.py:26-40)."""
        _, z = self._device_batch(ch_name, freqs, _be.OUT_CWT)
        return z.cpu().numpy()

    def power(self, ch_name: str, freqs: Numbers, *, baseline=None) -> np.ndarray:
        """Mean over epochs of |cwt|**2, (F, T) (reference mneutils.py:42-55)."""
        if baseline is None:
            fused = self._fused(ch_name, freqs, 0)
            if fused is not None:
                return fused
        plan, p = self._device_batch(ch_name, freqs, _be.OUT_POWER, baseline)
        return plan.reduce_epochs_device(p, 0).cpu().numpy()

    def itc(self, ch_name: str, freqs: Numbers) -> np.ndarray:
        """Inter-trial coherence |mean(cwt/|cwt|)|, (F, T) (reference mneutils.py:57-71)."""
        fused = self._fused(ch_name, freqs, 1)
        if fused is not None:
            return fused
        plan, z = self._device_batch(ch_name, freqs, _be.OUT_CWT)
        return plan.reduce_epochs_device(z, 1).cpu().numpy()

    def power_all(self, freqs: Numbers) -> np.ndarray:
        """Extension: mean power of EVERY channel in one call, (C, F, T) - one CTA per (channel, frequency share)."""
        import torch
        w = self.wavelet
        w._require_cuda()
        data = np.asarray(self.epochs.get_data())                    # (E, C, T)
        plan = w._plan_for(int(data.shape[-1]), freqs, True)
        x = torch.as_tensor(np.ascontiguousarray(data.transpose(1, 0, 2), dtype=plan.real_dtype), device="cuda:%d" % plan.device)
        if plan.info()["path"] == "short_packed":
            return plan.transform_epochs_device(x, 0).cpu().numpy()
        return np.stack([plan.reduce_epochs_device(plan.transform_device(x[c], _be.OUT_POWER), 0).cpu().numpy()
                         for c in range(x.shape[0])])
