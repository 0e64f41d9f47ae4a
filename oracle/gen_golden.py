"""Generate tests/golden/*.npz from the UNMODIFIED reference (authoring container only).

TEST INFRASTRUCTURE.  Run as `python oracle/gen_golden.py` where
`/root/reference` is mounted.  The reference is imported through
`oracle/refload.py` (stubbed cupy/matplotlib, nothing else changed) and its
own public API is called: `Morse/Morlet/Shannon/MexicanHat/Haar(...).cwt()`,
`.power()`, `.make_fft_wavelets()`, `Baseline(...).<mode>()`,
`EpochsWavelet(...).cwt/power/itc()`.  Inputs are regenerated from seeds by
the tests (and stored, for safety); outputs are stored in float64/complex128.
Large cases store a strided sample of output columns to keep the fixtures
small.  The fixtures are what pins `oracle/cwt_oracle.py` (and through it the
CUDA path) to the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def family(ref, kind, **kw):
    w = ref.wavelets
    ctor = {"morse": w.Morse, "morlet": w.Morlet, "shannon": w.Shannon,
            "mexicanhat": w.MexicanHat, "haar": w.Haar}[kind]
    return ctor(**kw)


class FakeEpochs:
    """Duck-typed stand-in for mne.Epochs (mneutils.py only touches these)."""

    def __init__(self, data, sfreq, ch_names):
        self._d = data
        self.info = {"sfreq": sfreq}
        self.ch_names = ch_names

    def get_data(self):
        return self._d


def broadband(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 1000.0
    return rng.standard_normal(n) + np.sin(2 * np.pi * 10 * t) + 0.5 * np.sin(2 * np.pi * 60 * t + 1.0)


def main():
    ref = refload.load()
    os.makedirs(OUT, exist_ok=True)
    cases = {}

    # ---- cfg1: README example (test.py:30-35 form) --------------------------
    t = np.arange(0, 0.3, 0.001)
    sine = np.sin(t * 60 * 2 * np.pi)
    freqs = np.arange(1, 100).astype(float)
    m = family(ref, "morse", sfreq=1000, b=17.5, r=3)
    cases["readme_morse"] = dict(kind="morse", kw=dict(sfreq=1000, b=17.5, r=3), wave=sine, freqs=freqs,
                                 cwt=m.cwt(sine, range(1, 100)), power=m.power(sine))
    ml = family(ref, "morlet", sfreq=1000, sigma=7.)
    cases["readme_morlet"] = dict(kind="morlet", kw=dict(sfreq=1000, sigma=7.), wave=sine, freqs=freqs,
                                  cwt=ml.cwt(sine, range(1, 100)), power=ml.power(sine))

    # ---- every family on a broadband signal, several lengths ---------------
    fam_kw = {
        "morse": dict(sfreq=1000, b=17.5, r=3),
        "morse_b3": dict(sfreq=1000, b=3.0, r=3.0),
        "morlet": dict(sfreq=1000, sigma=7.),
        "gabor": dict(sfreq=1000, sigma=7., gabor=True),
        "morlet_s5": dict(sfreq=1000, sigma=5.),
        "shannon": dict(sfreq=1000),
        "mexicanhat": dict(sfreq=1000),
        "haar": dict(sfreq=1000),
    }
    kind_of = {"morse_b3": "morse", "gabor": "morlet", "morlet_s5": "morlet"}
    for n in (300, 301, 1000, 1500, 4096):
        x = broadband(n, seed=100 + n)
        fr = np.array([1., 2., 3.5, 7., 10., 20., 40., 60., 100., 150., 333.])
        for name, kw in fam_kw.items():
            kind = kind_of.get(name, name)
            if n > 1500 and name not in ("morse", "morlet", "mexicanhat"):
                continue
            if n > 1500:
                fr = np.array([2., 10., 60., 333.])
            obj = family(ref, kind, **kw)
            z = obj.cwt(x, fr)
            cases["%s_n%d" % (name, n)] = dict(kind=kind, kw=kw, wave=x, freqs=fr, cwt=z)
    # interpolate=True
    for n in (300, 301, 1500):
        x = broadband(n, seed=200 + n)
        fr = np.array([1., 5., 10., 50., 100., 250., 480., 700.])
        for name in ("morse", "morlet", "shannon", "mexicanhat"):
            kw = dict(fam_kw[name], interpolate=True)
            z = family(ref, name, **kw).cwt(x, fr)
            cases["%s_interp_n%d" % (name, n)] = dict(kind=name, kw=kw, wave=x, freqs=fr, cwt=z)
    # fractional frequencies (test.py:181) and beyond-Nyquist (test.py:105-106)
    x = broadband(2000, seed=7)
    fr = np.arange(0.1, 5.0, 0.3)
    cases["morse_fractional"] = dict(kind="morse", kw=fam_kw["morse"], wave=x, freqs=fr,
                                     cwt=family(ref, "morse", **fam_kw["morse"]).cwt(x, fr))
    fr = np.arange(1, 1000, 83).astype(float)
    for name in ("morse", "morlet"):
        cases["%s_beyond_nyquist" % name] = dict(kind=name, kw=fam_kw[name], wave=x, freqs=fr,
                                                 cwt=family(ref, name, **fam_kw[name]).cwt(x, fr))
    # non-default sfreq and real_wave_length (the latter only matters for Normal mode)
    x = broadband(1024, seed=9)
    fr = np.array([2., 4., 8., 16., 32., 64.])
    kw = dict(sfreq=256, b=10.0, r=2.0)
    cases["morse_sfreq256"] = dict(kind="morse", kw=kw, wave=x, freqs=fr, cwt=family(ref, "morse", **kw).cwt(x, fr))
    kw = dict(sfreq=500, sigma=7, real_wave_length=2.0)
    cases["mexicanhat_rwl2"] = dict(kind="mexicanhat", kw=kw, wave=x, freqs=fr,
                                    cwt=family(ref, "mexicanhat", **kw).cwt(x, fr))

    # ---- spectra alone (bit-exact target) ----------------------------------
    spectra = {}
    for name in ("morse", "morlet", "gabor", "shannon", "mexicanhat", "haar"):
        kind = kind_of.get(name, name)
        for interp in (False, True):
            if kind == "haar" and interp:
                continue
            kw = dict(fam_kw[name], interpolate=interp)
            obj = family(ref, kind, **kw)
            fr = [1., 2.5, 10., 99., 400.]
            bank = obj.make_fft_wavelets(fr, 1.5)
            for i, s in enumerate(bank):
                spectra["%s_%d_%d" % (name, int(interp), i)] = np.asarray(s)
    np.savez_compressed(os.path.join(OUT, "spectra.npz"), **spectra)

    # ---- larger rows: strided column sample ---------------------------------
    for n, fr in ((65536, np.array([1., 8., 30., 100.])),
                  (600000, np.array([1., 10., 100.])),
                  (1 << 20, np.array([2., 64.]))):
        x = broadband(n, seed=300 + (n % 1000))
        cols = np.arange(0, n, 997)
        for name in ("morse", "morlet"):
            z = family(ref, name, **fam_kw[name]).cwt(x, fr)
            cases["%s_long_n%d" % (name, n)] = dict(kind=name, kw=fam_kw[name], seed=300 + (n % 1000), n=n,
                                                    freqs=fr, cols=cols, cwt=z[:, cols],
                                                    row_power_sum=(np.abs(z) ** 2).sum(axis=1),
                                                    row_power_max=(np.abs(z) ** 2).max(axis=1))
        if n == 65536:
            for name in ("shannon", "mexicanhat"):
                z = family(ref, name, **fam_kw[name]).cwt(x, fr)
                cases["%s_long_n%d" % (name, n)] = dict(kind=name, kw=fam_kw[name], seed=300 + (n % 1000), n=n,
                                                        freqs=fr, cols=cols, cwt=z[:, cols],
                                                        row_power_sum=(np.abs(z) ** 2).sum(axis=1),
                                                        row_power_max=(np.abs(z) ** 2).max(axis=1))

    # ---- Baseline (base.py:23-68) -------------------------------------------
    rng = np.random.default_rng(11)
    w = rng.uniform(0.5, 2.0, size=1500)
    bl = {"wave": w}
    b = ref.base.Baseline(w, 1000, 0.0, 0.2)
    for mode in ("mean", "ratio", "percent", "log", "zscore", "zlog"):
        bl[mode] = getattr(b, mode)()
    b2 = ref.base.Baseline(w, 1000, 0.1, 0.35)
    bl["zscore_100_350"] = b2.zscore()
    np.savez_compressed(os.path.join(OUT, "baseline.npz"), **bl)

    # ---- cfg3 slice: epochs (mneutils.py) + zscore on each power row --------
    rng = np.random.default_rng(3)
    E, C, T = 4, 2, 1500
    tt = np.arange(T) / 1000.0
    data = rng.standard_normal((E, C, T)) + 2.0 * (((tt >= 0.5) & (tt < 1.0)) * np.sin(2 * np.pi * 10 * tt))
    ep = FakeEpochs(data, 1000.0, ["MEG0", "MEG1"])
    fr = np.arange(1, 101, 18).astype(float)
    ew = ref.mneutils.EpochsWavelet(ep, family(ref, "morlet", sfreq=1000, sigma=7.))
    ecwt = ew.cwt("MEG1", fr)
    ew2 = ref.mneutils.EpochsWavelet(ep, family(ref, "morlet", sfreq=1000, sigma=7.))
    epow = ew2.power("MEG1", fr)
    ew3 = ref.mneutils.EpochsWavelet(ep, family(ref, "morlet", sfreq=1000, sigma=7.))
    eitc = ew3.itc("MEG1", fr)
    p = np.abs(ecwt) ** 2
    z = np.array([[ref.base.Baseline(row, 1000.0, 0.0, 0.2).zscore() for row in e] for e in p])
    np.savez_compressed(os.path.join(OUT, "epochs.npz"), data=data, freqs=fr, cwt=ecwt, power=epow, itc=eitc,
                        zscore_power=z)

    # ---- write transform cases ----------------------------------------------
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            if k == "kw":
                flat["%s/kw" % name] = np.array(repr(v))
            elif k == "kind":
                flat["%s/kind" % name] = np.array(v)
            else:
                flat["%s/%s" % (name, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "transforms.npz"), **flat)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote %d cases, %.2f MB in %s" % (len(cases), tot / 1e6, OUT))


if __name__ == "__main__":
    main()
