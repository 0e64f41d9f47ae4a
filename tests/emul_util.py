"""Loads tests/emul/libnwemul.so - the kernel bodies of ninwavelets_b200/csrc compiled for the
HOST and stepped block-by-block (test infrastructure; see tests/emul/emul.cpp).  Lets the CPU
suite check the radix plans, index algebra, bands and epilogues of the CUDA code against the
oracle without a GPU.  The product never loads this library."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "emul.cpp")
LIB = os.path.join(HERE, "emul", "libnwemul.so")
CSRC = os.path.join(ROOT, "ninwavelets_b200", "csrc")


def _build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "nwcwt.h")]
    if os.path.isfile(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB, SRC])


_lib = None


def lib():
    global _lib
    if _lib is None:
        _build()
        _lib = C.CDLL(LIB)
    return _lib


def emul_transform(fam_desc, signals, output, baseline=0, lo=0, hi=0, force_long=False):
    """fam_desc: dict for ninwavelets_b200._backend.PlanDesc fields (numpy arrays for pointers)."""
    from ninwavelets_b200._backend import PlanDesc
    d = PlanDesc()
    keep = []
    dtype = fam_desc["dtype"]
    rdt = np.float64 if dtype == 1 else np.float32
    signals = np.ascontiguousarray(signals, dtype=rdt)
    S, N = signals.shape
    freqs = np.ascontiguousarray(fam_desc["freqs"], dtype=np.float64)
    d.device, d.dtype, d.family, d.interpolate = 0, dtype, fam_desc["family"], int(fam_desc.get("interpolate", 0))
    d.n, d.n_freqs, d.sfreq = N, len(freqs), float(fam_desc["sfreq"])
    d.freqs = freqs.ctypes.data_as(C.POINTER(C.c_double))
    d.p0, d.p1, d.p2 = [float(fam_desc.get(k, 0.0)) for k in ("p0", "p1", "p2")]
    if fam_desc.get("aux") is not None:
        aux = np.ascontiguousarray(fam_desc["aux"], dtype=np.float64)
        keep.append(aux)
        d.aux = aux.ctypes.data_as(C.POINTER(C.c_double))
    if fam_desc.get("table") is not None:
        table = np.ascontiguousarray(fam_desc["table"], dtype=np.complex128)
        keep.append(table)
        d.table = table.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
        d.table_len = table.shape[1]
        if fam_desc.get("table_lens") is not None:
            tl = np.ascontiguousarray(fam_desc["table_lens"], dtype=np.int64)
            keep.append(tl)
            d.table_lens = tl.ctypes.data_as(C.POINTER(C.c_int64))
    d.prune_eps = float(fam_desc.get("prune_eps", -1.0))
    d.resample = int(fam_desc.get("resample", 0))
    d.resample_tol = float(fam_desc.get("resample_tol", 0.0))
    F = len(freqs)
    odt = (np.complex128 if dtype == 1 else np.complex64) if output == 0 else rdt
    out = np.zeros((S, F, N), dtype=odt)
    err = C.create_string_buffer(256)
    fn = lib().emul_transform
    fn.argtypes = [C.POINTER(PlanDesc), C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_longlong,
                   C.c_longlong, C.c_int, C.c_char_p, C.c_int]
    rc = fn(C.byref(d), signals.ctypes.data, out.ctypes.data, S, output, baseline, lo, hi, int(force_long), err, 256)
    if rc != 0:
        raise RuntimeError("emul: %d %s" % (rc, err.value.decode()))
    return out


def desc_from_oracle_family(fam, freqs, n, dtype, prune_eps=-1.0):
    """Build the plan description the Python host layer would build for an oracle Family.
    Normal-mode tables are built with the oracle's make_fft_wavelet (host numpy) here; the
    product builds them with the device forward FFT instead."""
    import cwt_oracle as orc
    freqs = np.asarray(freqs, dtype=np.float64)
    d = dict(dtype=dtype, sfreq=fam.sfreq, freqs=freqs, interpolate=fam.interpolate, prune_eps=prune_eps)
    if fam.kind == "morse":
        d.update(family=0, p0=fam.b, p1=fam.r)
    elif fam.kind == "morlet":
        d.update(family=1, p0=fam.sigma, p1=fam.c * np.float_power(np.pi, -1 / 4), p2=fam.k,
                 aux=np.array([orc.peak_freq(fam, f) for f in freqs]))
    elif fam.kind == "shannon":
        d.update(family=2)
    else:
        rows = [orc.make_fft_wavelet(fam, f, n / fam.sfreq) for f in freqs]
        lens = np.array([len(r) for r in rows], dtype=np.int64)
        table = np.zeros((len(rows), lens.max()), dtype=np.complex128)
        for i, r in enumerate(rows):
            table[i, :lens[i]] = r
        d.update(family=3, table=table, table_lens=lens)
    return d
