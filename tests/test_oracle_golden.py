"""Pins oracle/cwt_oracle.py to the reference: every fixture in tests/golden was
produced by the unmodified reference (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

import cwt_oracle as orc
from golden_util import case_wave, peak_rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_spectra_bit_exact():
    z = np.load(os.path.join(GOLDEN, "spectra.npz"))
    fam_kw = {
        "morse": ("morse", dict(sfreq=1000, b=17.5, r=3)),
        "morlet": ("morlet", dict(sfreq=1000, sigma=7.)),
        "gabor": ("morlet", dict(sfreq=1000, sigma=7., gabor=True)),
        "shannon": ("shannon", dict(sfreq=1000)),
        "mexicanhat": ("mexicanhat", dict(sfreq=1000)),
        "haar": ("haar", dict(sfreq=1000)),
    }
    fr = [1., 2.5, 10., 99., 400.]
    checked = 0
    for key in z.files:
        name, interp, i = key.rsplit("_", 2)
        kind, kw = fam_kw[name]
        fam = orc.Family(kind, interpolate=bool(int(interp)), **kw)
        mine = orc.make_fft_wavelets(fam, fr, 1.5)[int(i)]
        ref = z[key]
        assert mine.shape == ref.shape, key
        assert np.array_equal(mine, ref), key  # same ufunc sequence -> identical bits
        checked += 1
    assert checked == 55


def test_transforms_match_reference(golden_transforms):
    worst = 0.0
    for name, c in golden_transforms.items():
        fam = orc.Family(c["kind"], **c["kw"])
        wave = case_wave(c)
        z = orc.cwt(fam, wave, c["freqs"])
        ref = c["cwt"]
        if "cols" in c:
            p = np.abs(z) ** 2
            np.testing.assert_allclose(p.sum(axis=1), c["row_power_sum"], rtol=1e-12, err_msg=name)
            np.testing.assert_allclose(p.max(axis=1), c["row_power_max"], rtol=1e-12, err_msg=name)
            z = z[:, c["cols"]]
        assert z.shape == ref.shape and z.dtype == ref.dtype, name
        err = peak_rel_err(z, ref).max()
        worst = max(worst, err)
        assert err <= 1e-15, (name, err)
        if "power" in c:
            assert peak_rel_err(orc.power(fam, wave, c["freqs"]), c["power"]).max() <= 1e-15, name
    assert len(golden_transforms) >= 60


def test_readme_known_answers(golden_transforms):
    """cfg1 (README / test.py:30-35): unit sinusoid at the analysis frequency
    has Morse power 1 (peak of the spectrum is exactly 2)."""
    c = golden_transforms["readme_morse"]
    p = orc.power(orc.Family("morse", **c["kw"]), c["wave"], c["freqs"])
    assert p.shape == (99, 300)
    assert np.unravel_index(p.argmax(), p.shape)[0] == 59
    assert abs(p.max() - 1.0) < 1e-14
    assert abs(p[59, 150] - 1.0) < 1e-14
    fam = orc.Family("morlet", sigma=7.0)
    assert abs(fam.c - 1.0) < 1e-15 and abs(fam.k - 2.289734845645553e-11) < 1e-25


def test_baseline_modes():
    z = np.load(os.path.join(GOLDEN, "baseline.npz"))
    w = z["wave"]
    for mode in orc.BASELINE_MODES:
        assert np.array_equal(orc.baseline(w, 1000, 0.0, 0.2, mode), z[mode]), mode
    assert np.array_equal(orc.baseline(w, 1000, 0.1, 0.35, "zscore"), z["zscore_100_350"])


def test_epochs():
    z = np.load(os.path.join(GOLDEN, "epochs.npz"))
    fam = orc.Family("morlet", sfreq=1000.0, sigma=7.0)
    x = z["data"][:, 1, :]
    assert peak_rel_err(orc.epochs_cwt(fam, x, z["freqs"]), z["cwt"]).max() <= 1e-15
    assert peak_rel_err(orc.epochs_power(fam, x, z["freqs"]), z["power"]).max() <= 1e-15
    assert peak_rel_err(orc.epochs_itc(fam, x, z["freqs"]), z["itc"]).max() <= 1e-15
    p = np.abs(z["cwt"]) ** 2
    zs = orc.baseline_rows(p, 1000.0, 0.0, 0.2, "zscore")
    assert peak_rel_err(zs, z["zscore_power"]).max() <= 1e-15


def test_error_behaviour():
    fam = orc.Family("morse")
    with pytest.raises(ZeroDivisionError):  # base.py:234-235
        orc.make_fft_wavelet(fam, 0)
    with pytest.raises(IndexError):  # base.py:272 with a single frequency
        orc.make_fft_wavelets(fam, [5.0])
    with pytest.raises(TypeError):
        orc.make_fft_wavelets(fam, None)


@pytest.mark.skipif(not os.path.isdir("/root/reference/ninwavelets"), reason="reference tree not mounted")
def test_live_reference_random_cases():
    """Where the reference is mounted, fuzz the oracle against it directly."""
    import refload
    ref = refload.load()
    rng = np.random.default_rng(0)
    for trial in range(12):
        n = int(rng.integers(50, 3000))
        x = rng.standard_normal(n)
        fr = np.sort(rng.uniform(0.5, 400, size=5))
        sf = float(rng.choice([250.0, 1000.0, 512.0]))
        interp = bool(trial % 2)
        pairs = [
            (ref.wavelets.Morse(sf, 9.0, 2.5, interpolate=interp), orc.Family("morse", sfreq=sf, b=9.0, r=2.5, interpolate=interp)),
            (ref.wavelets.Morlet(sf, 6.0, interpolate=interp), orc.Family("morlet", sfreq=sf, sigma=6.0, interpolate=interp)),
            (ref.wavelets.MexicanHat(sf, interpolate=interp), orc.Family("mexicanhat", sfreq=sf, interpolate=interp)),
            (ref.wavelets.Shannon(sf, interpolate=interp), orc.Family("shannon", sfreq=sf, interpolate=interp)),
        ]
        for robj, fam in pairs:
            a = robj.cwt(x, fr)
            b = orc.cwt(fam, x, fr)
            assert a.shape == b.shape
            assert peak_rel_err(b, a).max() <= 1e-15, (trial, fam.kind, n)
