"""ctypes binding of libnwcwt.so (include/nwcwt.h).

This is the whole device boundary of the package: plans, workspace sizing and
the transform entry points.  PyTorch is used only as the allocator / stream
provider for device and pinned host memory.  There is no CPU implementation
behind any of these calls: if the shared library is missing, or a compute call
is made without a CUDA device, an exception is raised.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NWCWT_LIB") or os.path.join(_HERE, "libnwcwt.so")   # NWCWT_LIB: experiment builds

F32, F64 = 0, 1
MORSE, MORLET, SHANNON, TABLE = 0, 1, 2, 3
OUT_CWT, OUT_ABS, OUT_POWER = 0, 1, 2
BASELINE_MODES = {None: 0, "none": 0, "mean": 1, "ratio": 2, "percent": 3, "log": 4, "zscore": 5, "zlog": 6}

ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE, ERR_ZERO_FREQ = -1, -2, -3, -4, -5


class PlanDesc(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("dtype", C.c_int32), ("family", C.c_int32), ("interpolate", C.c_int32),
        ("n", C.c_int64), ("n_freqs", C.c_int32), ("resample", C.c_int32),
        ("sfreq", C.c_double), ("freqs", C.POINTER(C.c_double)),
        ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
        ("aux", C.POINTER(C.c_double)),
        ("table", C.POINTER(C.c_double)), ("table_len", C.c_int64), ("table_lens", C.POINTER(C.c_int64)),
        ("prune_eps", C.c_double), ("resample_tol", C.c_double),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("n_freqs", C.c_int32), ("path", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32),
        ("batch", C.c_int32), ("n_stages", C.c_int32 * 2), ("radices", (C.c_int32 * 16) * 2),
        ("band_bins", C.c_int64), ("smem_bytes", C.c_int64),
        ("threads", C.c_int32 * 2), ("rows_per_launch", C.c_int32), ("n_groups", C.c_int32),
        ("group_D", C.c_int32 * 32), ("group_K", C.c_int32 * 32), ("group_rows", C.c_int32 * 32),
        ("group_n1", C.c_int32 * 32), ("group_n2", C.c_int32 * 32), ("group_err", C.c_double * 32),
    ]


class BackendError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nwcwt error %d: %s" % (code, msg))
        self.code = code


_lib = None
_lock = threading.Lock()

# every symbol include/nwcwt.h declares
SYMBOLS = (
    "nwcwt_version", "nwcwt_last_error", "nwcwt_plan_create", "nwcwt_plan_destroy", "nwcwt_plan_get_info",
    "nwcwt_plan_get_bands", "nwcwt_workspace_bytes", "nwcwt_spectrum_bank", "nwcwt_reduce_epochs",
    "nwcwt_baseline_rows", "nwcwt_launch_count", "nwcwt_profile_enable", "nwcwt_profile_read",
    "nwcwt_forward", "nwcwt_transform", "nwcwt_transform_host", "nwcwt_debug_force_generic",
    "nwcwt_debug_force_exact", "nwcwt_fma_peak", "nwcwt_transform_epochs",
)


def lib():
    """Load libnwcwt.so (built in-tree by `__graft_entry__.build()`); fail loudly if absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise ImportError(
                    "ninwavelets_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` (nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
            L = C.CDLL(LIB_PATH)
            vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
            L.nwcwt_version.restype = C.c_int
            L.nwcwt_last_error.restype = C.c_char_p
            L.nwcwt_plan_create.argtypes = [C.POINTER(vp), C.POINTER(PlanDesc)]
            L.nwcwt_plan_destroy.argtypes = [vp]
            L.nwcwt_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
            L.nwcwt_plan_get_bands.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
            L.nwcwt_workspace_bytes.argtypes = [vp, i64, C.POINTER(sz)]
            L.nwcwt_spectrum_bank.argtypes = [vp, vp, vp]
            L.nwcwt_reduce_epochs.argtypes = [vp, vp, vp, i64, i64, i32, vp]
            L.nwcwt_forward.argtypes = [vp, vp, vp, i64, vp, sz, vp]
            L.nwcwt_transform_epochs.argtypes = [vp, vp, vp, i64, i64, i32, vp, sz, vp]
            L.nwcwt_transform.argtypes = [vp, vp, vp, i64, i32, i32, i64, i64, vp, sz, vp]
            L.nwcwt_transform_host.argtypes = [vp, vp, vp, i64, i32, i32, i64, i64]
            L.nwcwt_launch_count.restype = C.c_int64
            L.nwcwt_profile_enable.argtypes = [i32]
            L.nwcwt_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
            for name in SYMBOLS:
                if name not in ("nwcwt_version", "nwcwt_last_error", "nwcwt_launch_count"):
                    getattr(L, name).restype = C.c_int
            _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().nwcwt_last_error().decode("utf-8", "replace")
        if rc == ERR_ZERO_FREQ:
            raise ZeroDivisionError(msg)  # reference base.py:234-235
        raise BackendError(rc, msg)


def launch_count():
    return int(lib().nwcwt_launch_count())


PROFILE_CLASSES = ("short_fused", "fwd_passA", "fwd_passB", "inv_passA", "inv_passB", "baseline_rows", "resample",
                   "reserved")


def force_exact(on):
    """Test hook: plans with resampled rows run the exact length-n transform for every row."""
    _check(lib().nwcwt_debug_force_exact(C.c_int32(1 if on else 0)))


def force_generic(on):
    """Test hook: run the generic kernels even where the plan has the packed fast path."""
    _check(lib().nwcwt_debug_force_generic(C.c_int32(1 if on else 0)))


def fma_peak(device, f32=True):
    """Measured FP32 (FFMA2) / FP64 (DFMA) pipe rate of the device in lane multiply-adds per second."""
    v = C.c_double()
    L = lib()
    L.nwcwt_fma_peak.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    _check(L.nwcwt_fma_peak(int(device), F32 if f32 else F64, C.byref(v)))
    return float(v.value)


def profile_enable(on):
    _check(lib().nwcwt_profile_enable(int(bool(on))))


def profile_read():
    ms = (C.c_double * 8)()
    n = (C.c_int64 * 8)()
    _check(lib().nwcwt_profile_read(ms, n))
    return {k: dict(ms=ms[i], launches=int(n[i])) for i, k in enumerate(PROFILE_CLASSES)}


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Plan:
    """Device-side counterpart of `WaveletBase.fft_wavelets` for one (family, freqs, N, dtype).

    A plan owns one workspace and one set of auxiliary streams: use it from one thread and one CUDA stream
    at a time (build one plan per concurrent stream)."""

    def __init__(self, *, device, dtype, family, interpolate, n, sfreq, freqs, p0=0.0, p1=0.0, p2=0.0,
                 aux=None, table=None, table_lens=None, prune_eps=-1.0, resample=None, resample_tol=0.0):
        self._h = C.c_void_p()
        self.dtype = F64 if np.dtype(dtype) == np.float64 else F32
        self.real_dtype = np.dtype(np.float64 if self.dtype == F64 else np.float32)
        self.cplx_dtype = np.dtype(np.complex128 if self.dtype == F64 else np.complex64)
        self.n = int(n)
        self.device = int(device)
        self.freqs = np.ascontiguousarray(np.asarray(freqs, dtype=np.float64).reshape(-1))
        self.n_freqs = int(self.freqs.shape[0])
        d = PlanDesc()
        d.device, d.dtype, d.family, d.interpolate = self.device, self.dtype, int(family), int(bool(interpolate))
        d.n, d.n_freqs, d.sfreq = self.n, self.n_freqs, float(sfreq)
        d.freqs = _dptr(self.freqs)
        d.p0, d.p1, d.p2 = float(p0), float(p1), float(p2)
        keep = [self.freqs]
        if aux is not None:
            aux = np.ascontiguousarray(np.asarray(aux, dtype=np.float64).reshape(-1))
            d.aux = _dptr(aux)
            keep.append(aux)
        if table is not None:
            table = np.ascontiguousarray(np.asarray(table, dtype=np.complex128))
            if table.ndim != 2 or table.shape[0] != self.n_freqs:
                raise ValueError("table must be [n_freqs, table_len]")
            d.table = table.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
            d.table_len = table.shape[1]
            keep.append(table)
            if table_lens is not None:
                table_lens = np.ascontiguousarray(np.asarray(table_lens, dtype=np.int64))
                d.table_lens = table_lens.ctypes.data_as(C.POINTER(C.c_int64))
                keep.append(table_lens)
        d.prune_eps = -1.0 if prune_eps is None else float(prune_eps)
        d.resample = 0 if resample is None else (1 if resample else -1)
        d.resample_tol = float(resample_tol or 0.0)
        _check(lib().nwcwt_plan_create(C.byref(self._h), C.byref(d)))
        del keep
        self._ws = None

    # -- introspection (host only, no CUDA) ---------------------------------
    def info(self):
        i = PlanInfo()
        _check(lib().nwcwt_plan_get_info(self._h, C.byref(i)))
        out = dict(n=i.n, n_freqs=i.n_freqs, path=("short", "long", "long_packed", "short_packed", "chirp_z")[i.path], n1=i.n1, n2=i.n2,
                   batch=i.batch, band_bins=i.band_bins, smem_bytes=i.smem_bytes, threads=list(i.threads),
                   rows_per_launch=i.rows_per_launch)
        out["radices"] = [list(i.radices[k][: i.n_stages[k]]) for k in range(2)]
        out["groups"] = [dict(D=i.group_D[g], K=i.group_K[g], rows=i.group_rows[g], n1=i.group_n1[g], n2=i.group_n2[g],
                              err=i.group_err[g]) for g in range(i.n_groups)]
        return out

    def bands(self):
        lo = np.zeros(self.n_freqs, dtype=np.int32)
        hi = np.zeros(self.n_freqs, dtype=np.int32)
        _check(lib().nwcwt_plan_get_bands(self._h, lo.ctypes.data_as(C.POINTER(C.c_int32)),
                                          hi.ctypes.data_as(C.POINTER(C.c_int32))))
        return lo, hi

    def workspace_bytes(self, n_signals=1):
        b = C.c_size_t()
        _check(lib().nwcwt_workspace_bytes(self._h, int(n_signals), C.byref(b)))
        return int(b.value)

    # -- device calls ---------------------------------------------------------
    def _torch(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ninwavelets_b200 needs a CUDA device (B200); there is no CPU path")
        return torch

    def _workspace(self, torch, n_signals):
        need = self.workspace_bytes(n_signals)
        if need == 0:
            return None, 0
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device="cuda:%d" % self.device)
        return self._ws, need

    def _check_device_tensor(self, t, name):
        """A plan works on its own device only (its tables and workspace live there)."""
        if not t.is_cuda or t.device.index != self.device:
            raise ValueError("%s must be a CUDA tensor on the plan's device cuda:%d, got %s" % (name, self.device, t.device))

    def _tdtype(self, torch, complex_=False):
        if self.dtype == F64:
            return torch.complex128 if complex_ else torch.float64
        return torch.complex64 if complex_ else torch.float32

    def transform_device(self, signals, output=OUT_POWER, baseline=0, base_lo=0, base_hi=0, out=None):
        """signals: torch CUDA tensor [S, N] of the plan's real dtype -> torch tensor [S, F, N]."""
        torch = self._torch()
        self._check_device_tensor(signals, "signals")
        if signals.dim() != 2 or signals.shape[1] != self.n:
            raise ValueError("signals must be [n_signals, %d], got %s" % (self.n, tuple(signals.shape)))
        signals = signals.contiguous()
        if signals.dtype != self._tdtype(torch):
            signals = signals.to(self._tdtype(torch))
        S = signals.shape[0]
        odt = self._tdtype(torch, output == OUT_CWT)
        if out is None:
            out = torch.empty((S, self.n_freqs, self.n), dtype=odt, device=signals.device)
        else:
            self._check_device_tensor(out, "out")
            if tuple(out.shape) != (S, self.n_freqs, self.n) or out.dtype != odt or not out.is_contiguous():
                raise ValueError("out must be a contiguous %s tensor of shape %s" % (odt, (S, self.n_freqs, self.n)))
        ws, need = self._workspace(torch, S)
        stream = torch.cuda.current_stream(signals.device).cuda_stream
        _check(lib().nwcwt_transform(self._h, signals.data_ptr(), out.data_ptr(), S, int(output), int(baseline),
                                     int(base_lo), int(base_hi), ws.data_ptr() if ws is not None else None, need,
                                     stream))
        return out

    def forward_device(self, signals):
        torch = self._torch()
        self._check_device_tensor(signals, "signals")
        if signals.dim() != 2 or signals.shape[1] != self.n:
            raise ValueError("signals must be [n_signals, %d], got %s" % (self.n, tuple(signals.shape)))
        signals = signals.contiguous().to(self._tdtype(torch))
        S = signals.shape[0]
        spec = torch.empty((S, self.n), dtype=self._tdtype(torch, True), device=signals.device)
        ws, need = self._workspace(torch, S)
        stream = torch.cuda.current_stream(signals.device).cuda_stream
        _check(lib().nwcwt_forward(self._h, signals.data_ptr(), spec.data_ptr(), S,
                                   ws.data_ptr() if ws is not None else None, need, stream))
        return spec

    def spectrum_bank_device(self):
        torch = self._torch()
        bank = torch.empty((self.n_freqs, self.n), dtype=self._tdtype(torch, True), device="cuda:%d" % self.device)
        stream = torch.cuda.current_stream(bank.device).cuda_stream
        _check(lib().nwcwt_spectrum_bank(self._h, bank.data_ptr(), stream))
        return bank

    def reduce_epochs_device(self, x, kind):
        """x: [E, ...] real (kind 0: mean) or complex (kind 1: itc) CUDA tensor -> [...] real."""
        torch = self._torch()
        x = x.contiguous()
        E = x.shape[0]
        out = torch.empty(x.shape[1:], dtype=self._tdtype(torch), device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(lib().nwcwt_reduce_epochs(self._h, x.data_ptr(), out.data_ptr(), E, out.numel(), int(kind), stream))
        return out

    def transform_epochs_device(self, signals, kind):
        """signals: torch CUDA tensor [C, E, N] -> [C, F, N] real: mean power over epochs (kind 0) or inter-trial
        coherence (kind 1), reduced inside the transform kernel.  Raises BackendError(ERR_UNSUPPORTED) for rows that
        do not fit one CTA (callers fall back to transform_device + reduce_epochs_device)."""
        torch = self._torch()
        self._check_device_tensor(signals, "signals")
        if signals.dim() != 3 or signals.shape[2] != self.n:
            raise ValueError("signals must be [n_channels, n_epochs, %d], got %s" % (self.n, tuple(signals.shape)))
        signals = signals.contiguous().to(self._tdtype(torch))
        Cn, E = int(signals.shape[0]), int(signals.shape[1])
        out = torch.empty((Cn, self.n_freqs, self.n), dtype=self._tdtype(torch), device=signals.device)
        ws = None
        if kind == 1:
            ws = torch.empty((Cn, self.n_freqs, self.n), dtype=self._tdtype(torch, True), device=signals.device)
        stream = torch.cuda.current_stream(signals.device).cuda_stream
        _check(lib().nwcwt_transform_epochs(self._h, signals.data_ptr(), out.data_ptr(), Cn, E, int(kind),
                                            ws.data_ptr() if ws is not None else None,
                                            ws.numel() * ws.element_size() if ws is not None else 0, stream))
        return out

    def transform_host(self, signals, output=OUT_POWER, baseline=0, base_lo=0, base_hi=0, out=None, pinned=True):
        """signals: numpy [S, N] -> numpy [S, F, N]; copies are inside the call (pinned result buffer)."""
        torch = self._torch()
        signals = np.ascontiguousarray(signals, dtype=self.real_dtype)
        if signals.ndim != 2 or signals.shape[1] != self.n:
            raise ValueError("signals must be [n_signals, %d], got %s" % (self.n, signals.shape))
        S = signals.shape[0]
        odt = self.cplx_dtype if output == OUT_CWT else self.real_dtype
        if out is None:
            if pinned:
                tdt = self._tdtype(torch, output == OUT_CWT)
                holder = torch.empty((S, self.n_freqs, self.n), dtype=tdt, pin_memory=True)
                out = holder.numpy()
            else:
                out = np.empty((S, self.n_freqs, self.n), dtype=odt)
        if out.dtype != odt or not out.flags.c_contiguous or out.shape != (S, self.n_freqs, self.n):
            raise ValueError("out must be a C-contiguous %s array of shape %s" % (odt, (S, self.n_freqs, self.n)))
        _check(lib().nwcwt_transform_host(self._h, signals.ctypes.data, out.ctypes.data, S, int(output),
                                          int(baseline), int(base_lo), int(base_hi)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().nwcwt_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
