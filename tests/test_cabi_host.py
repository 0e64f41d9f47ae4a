"""CPU tests of the C-ABI library: it loads, exports every symbol of include/nwcwt.h, and its
host-side planner (no CUDA calls) makes the decisions DESIGN.md describes."""
import os
import re

import numpy as np
import pytest

from ninwavelets_b200 import _backend as be

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nwcwt.h")).read()
    declared = set(re.findall(r"\b(nwcwt_[a-z_]+)\s*\(", hdr))
    assert declared == set(be.SYMBOLS), declared ^ set(be.SYMBOLS)
    L = be.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.nwcwt_version() == 100


def plan(n, freqs, family=be.MORSE, dtype=np.float32, **kw):
    kw.setdefault("p0", 17.5)
    kw.setdefault("p1", 3.0)
    return be.Plan(device=0, dtype=dtype, family=family, interpolate=kw.pop("interpolate", False), n=n,
                   sfreq=kw.pop("sfreq", 1000.0), freqs=freqs, **kw)


def test_short_and_long_shapes():
    i = plan(300, np.arange(1, 100.0)).info()
    assert i["path"] == "short_packed" and i["batch"] == 8 and np.prod(i["radices"][0]) == 300
    i = plan(1500, np.arange(1, 101.0)).info()
    assert i["path"] == "short_packed" and np.prod(i["radices"][0]) == 1500 and i["smem_bytes"] <= 75 * 1024
    i = plan(2 * 7 * 11 * 13, np.arange(1, 101.0)).info()   # a factor the packed engine does not have: generic kernel
    assert i["path"] == "short" and np.prod(i["radices"][0]) == 2002
    i = plan(600000, np.arange(1, 101.0)).info()
    assert i["path"] == "long_packed" and i["n1"] * i["n2"] == 600000
    assert np.prod(i["radices"][0]) == i["n1"] and np.prod(i["radices"][1]) == i["n2"]
    assert i["smem_bytes"] <= 227 * 1024
    i = plan(1 << 20, np.arange(1, 129.0), dtype=np.float64).info()
    assert i["path"] == "long_packed" and i["n1"] * i["n2"] == 1 << 20 and i["smem_bytes"] <= 227 * 1024
    for e in (16, 18, 22):
        assert plan(1 << e, [1.0, 2.0]).info()["n1"] * plan(1 << e, [1.0, 2.0]).info()["n2"] == 1 << e


def test_any_length_plans_and_bad_args():
    # a prime factor > 64: no radix plan, the chirp-z (Bluestein) path on a smooth length >= 2 n - 1 (host planning only here)
    for n in (2 * 10007, 100003, 1234, 599999, 600001):
        i = plan(n, [1.0, 2.0]).info()
        assert i["path"] == "chirp_z" and i["n1"] >= 2 * n - 1 and i["n1"] < 4 * n, (n, i)
        m = i["n1"]
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        assert m == 1
        assert plan(n, [1.0, 2.0]).workspace_bytes(4) > 0
    with pytest.raises(be.BackendError) as ei:
        plan((1 << 31) + 1, [1.0, 2.0])
    assert ei.value.code == be.ERR_UNSUPPORTED
    with pytest.raises(ZeroDivisionError):      # reference base.py:234-235
        plan(300, [1.0, 0.0])
    with pytest.raises(be.BackendError):
        plan(1, [1.0, 2.0])


def test_bands_contain_everything_above_eps():
    """The pruned band must contain every bin whose reference spectrum exceeds eps * peak."""
    import cwt_oracle as orc
    for kind, fam_kw, native in (
            ("morse", dict(b=17.5, r=3.0), dict(family=be.MORSE, p0=17.5, p1=3.0)),
            ("morse", dict(b=3.0, r=3.0), dict(family=be.MORSE, p0=3.0, p1=3.0)),
            ("morlet", dict(sigma=7.0), None), ("morlet", dict(sigma=5.0, gabor=True), None),
            ("shannon", {}, dict(family=be.SHANNON))):
        for n in (300, 1500, 4096):
            for interp in (False, True):
                fam = orc.Family(kind, sfreq=1000.0, interpolate=interp, **fam_kw)
                freqs = np.array([0.5, 1.0, 7.0, 33.0, 100.0, 480.0, 900.0])
                if native is None:
                    nat = dict(family=be.MORLET, p0=fam.sigma, p1=fam.c * np.float_power(np.pi, -0.25), p2=fam.k,
                               aux=np.array([orc.peak_freq(fam, f) for f in freqs]))
                else:
                    nat = native
                for eps in (1e-12, 1e-24):
                    p = be.Plan(device=0, dtype=np.float64, interpolate=interp, n=n, sfreq=1000.0, freqs=freqs,
                                prune_eps=eps, **nat)
                    lo, hi = p.bands()
                    bank = [orc.pad_to(w, n) for w in orc.make_fft_wavelets(fam, freqs, n / 1000.0)]
                    for i, w in enumerate(bank):
                        big = np.nonzero(np.abs(w) > eps * max(np.abs(w).max(), 1e-300))[0]
                        if kind == "shannon":
                            big = np.nonzero(w != 0)[0]
                            assert (lo[i], hi[i]) == ((big.min(), big.max() + 1) if big.size else (lo[i], lo[i]))
                        elif big.size:
                            assert lo[i] <= big.min() and big.max() < hi[i], (kind, n, interp, freqs[i], eps)
                # eps = 0 keeps every evaluated bin
                p0 = be.Plan(device=0, dtype=np.float64, interpolate=interp, n=n, sfreq=1000.0, freqs=freqs,
                             prune_eps=0.0, **nat)
                lo, hi = p0.bands()
                cut = n // 2 if interp else n
                if kind != "shannon":
                    assert (hi == cut).all() and (lo <= 1).all()


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ninwavelets_b200 as nw
    x = np.zeros(300)
    with pytest.raises(RuntimeError):
        nw.Morse(1000, cuda=True).power(x, range(1, 10))
    with pytest.raises(RuntimeError):
        nw.Morse(1000, cuda=False).power(x, range(1, 10))   # no numpy fallback either
    p = plan(300, [1.0, 2.0])
    with pytest.raises(RuntimeError):
        p.transform_host(np.zeros((1, 300), dtype=np.float32))


def test_reference_error_behaviour_of_host_mirror():
    import ninwavelets_b200 as nw
    m = nw.Morse(1000, cuda=True)
    with pytest.raises((IndexError, RuntimeError)):
        m.make_fft_wavelets([5.0])
    with pytest.raises(TypeError):
        m.make_fft_wavelets(None)
    assert nw.Morse().interpolate is False and nw.WaveletBase().interpolate is True   # reference defaults
    ml = nw.Morlet(1000, 7.0)
    assert abs(ml.c - 1.0) < 1e-15 and abs(ml.k - 2.289734845645553e-11) < 1e-25
    assert abs(ml.peak_freq(1.0) - 7.006389) < 1e-5
    assert nw.WaveletMode.Reverse.value == 2 and nw.Morse().mode is nw.WaveletMode.Reverse
