"""Row-wise Baseline on the device (reference base.py:46-68), used by `Baseline` and `EpochsWavelet`."""
import ctypes as C

import numpy as np

from . import _backend as _be


def baseline_rows(wave, sfreq, start, stop, mode, device=None):
    import torch
    is_torch = type(wave).__module__.split(".")[0] == "torch"
    if is_torch:
        t = wave
    else:
        arr = np.asarray(wave)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        dev = torch.cuda.current_device() if device is None else int(device)
        t = torch.as_tensor(arr, device="cuda:%d" % dev)
    shape = t.shape
    n = int(shape[-1])
    rows = t.reshape(-1, n).contiguous().clone()
    lo, hi, _ = slice(int(start * sfreq), int(stop * sfreq)).indices(n)
    hi = max(lo, hi)
    dt = _be.F64 if rows.dtype == torch.float64 else _be.F32
    L = _be.lib()
    L.nwcwt_baseline_rows.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64,
                                      C.c_int64, C.c_void_p]
    L.nwcwt_baseline_rows.restype = C.c_int
    stream = torch.cuda.current_stream(rows.device).cuda_stream
    _be._check(L.nwcwt_baseline_rows(rows.device.index, dt, rows.data_ptr(), rows.shape[0], n,
                                     _be.BASELINE_MODES[mode], lo, hi, stream))
    out = rows.reshape(shape)
    return out if is_torch else out.cpu().numpy()
