"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the public
host mirror (ninwavelets_b200.Morse/...) and therefore through the C ABI of libnwcwt.so; the oracle
and the golden fixtures (produced by the unmodified reference) are only the checker.

Tolerances (BASELINE.json): fp64 max|out-ref|/max|ref| per row <= 1e-12; fp32 per-row relative L2
<= 1e-5 against the fp64 reference evaluated on the float32-rounded input (metrics and their
noise-floor clause: tests/golden_util.py)."""
import os

import numpy as np
import pytest

import cwt_oracle as orc
from golden_util import case_floor, case_wave, l2_rel_err, peak_rel_err

pytestmark = pytest.mark.gpu

F64_TOL = 1e-12
F32_TOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def nw():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ninwavelets_b200 as pkg
    return pkg


def make(nw, kind, kw, **extra):
    kw = dict(kw)
    kw.update(extra)
    ctor = {"morse": nw.Morse, "morlet": nw.Morlet, "shannon": nw.Shannon, "mexicanhat": nw.MexicanHat,
            "haar": nw.Haar}[kind]
    return ctor(cuda=True, **kw)


def test_all_small_golden_cases_fp64(nw, golden_transforms):
    worst = 0.0
    n = 0
    for name, c in golden_transforms.items():
        if "cols" in c:
            continue
        x = case_wave(c)
        obj = make(nw, c["kind"], c["kw"])
        z = obj.cwt(x, c["freqs"])
        assert z.shape == c["cwt"].shape and z.dtype == np.complex128, name
        fl = case_floor(name)                  # zero except for the pure-sine README cases (golden_util.py)
        e = peak_rel_err(z, c["cwt"], fl).max()
        worst = max(worst, e)
        assert e <= F64_TOL, (name, e)
        p = obj.power(x)                       # cached plan, freqs=None like the reference allows
        assert peak_rel_err(p, np.abs(c["cwt"]) ** 2, fl).max() <= F64_TOL, name
        a = obj.abs(x)
        assert peak_rel_err(a, np.abs(c["cwt"]), fl).max() <= F64_TOL, name
        n += 1
    assert n >= 50
    print("worst fp64 error over %d golden cases (rows relative to themselves; floor only for readme_*): %.3e" % (n, worst))


def test_all_small_golden_cases_fp32(nw, golden_transforms):
    worst = 0.0
    for name, c in golden_transforms.items():
        if "cols" in c:
            continue
        x32 = case_wave(c).astype(np.float32)
        ref = orc.cwt(orc.Family(c["kind"], **c["kw"]), x32.astype(np.float64), c["freqs"])
        z = make(nw, c["kind"], c["kw"], dtype="float32").cwt(x32, c["freqs"])
        assert z.dtype == np.complex64
        e = l2_rel_err(z.astype(np.complex128), ref, case_floor(name, f32=True)).max()
        worst = max(worst, e)
        assert e <= F32_TOL, (name, e)
    print("worst fp32 error (rows relative to themselves; floor only for readme_*): %.3e" % worst)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_long_rows_vs_golden_samples(nw, golden_transforms, dtype):
    for name, c in golden_transforms.items():
        if "cols" not in c:
            continue
        x = case_wave(c)
        obj = make(nw, c["kind"], c["kw"], dtype=dtype)
        if dtype == "float64":
            z = obj.cwt(x, c["freqs"])
            assert peak_rel_err(z[:, c["cols"]], c["cwt"]).max() <= F64_TOL, name
            p = obj.power(x)
            np.testing.assert_allclose(p.sum(axis=1), c["row_power_sum"], rtol=1e-11, err_msg=name)
            np.testing.assert_allclose(p.max(axis=1), c["row_power_max"], rtol=1e-11, err_msg=name)
        else:
            # fp32: the golden is for the fp64 input; rounding the input adds ~6e-8 relative, far inside 1e-5
            z = obj.cwt(x.astype(np.float32), c["freqs"]).astype(np.complex128)
            assert l2_rel_err(z[:, c["cols"]], c["cwt"]).max() <= F32_TOL, name


def test_spectrum_generator_vs_reference_spectra(nw):
    """The in-register spectrum generator alone (nwcwt_spectrum_bank) against make_fft_wavelets of the
    reference (golden spectra.npz, N = 1500)."""
    z = np.load(os.path.join(GOLDEN, "spectra.npz"))
    kws = {"morse": ("morse", dict(sfreq=1000, b=17.5, r=3)), "morlet": ("morlet", dict(sfreq=1000, sigma=7.)),
           "gabor": ("morlet", dict(sfreq=1000, sigma=7., gabor=True)), "shannon": ("shannon", dict(sfreq=1000)),
           "mexicanhat": ("mexicanhat", dict(sfreq=1000)), "haar": ("haar", dict(sfreq=1000))}
    fr = [1., 2.5, 10., 99., 400.]
    for name, (kind, kw) in kws.items():
        for interp in (0, 1):
            if kind == "haar" and interp:
                continue
            obj = make(nw, kind, kw, interpolate=bool(interp), prune_eps=0.0)
            bank = obj.make_fft_wavelets(fr, 1.5)
            for i in range(len(fr)):
                ref = orc.pad_to(z["%s_%d_%d" % (name, interp, i)], 1500)
                if interp:
                    ref = orc.interpolate_alias(ref)
                got = np.asarray(bank[i])
                scale = max(np.abs(ref).max(), 1e-300)
                assert np.abs(got - ref).max() / scale <= 1e-13, (name, interp, i)


def test_baseline_modes_and_epochs(nw):
    z = np.load(os.path.join(GOLDEN, "baseline.npz"))
    w = z["wave"]
    for mode in orc.BASELINE_MODES:
        got = getattr(nw.Baseline(w, 1000, 0.0, 0.2), mode)()
        np.testing.assert_allclose(got, z[mode], rtol=1e-13, atol=1e-14, err_msg=mode)
    np.testing.assert_allclose(nw.Baseline(w, 1000, 0.1, 0.35).zscore(), z["zscore_100_350"], rtol=1e-13, atol=1e-14)

    e = np.load(os.path.join(GOLDEN, "epochs.npz"))

    class FakeEpochs:
        info = {"sfreq": 1000.0}
        ch_names = ["MEG0", "MEG1"]

        def get_data(self):
            return e["data"]

    ew = nw.EpochsWavelet(FakeEpochs(), nw.Morlet(500, 7., cuda=True))   # sfreq is overwritten (mneutils.py:24)
    assert ew.wavelet.sfreq == 1000.0
    assert peak_rel_err(ew.cwt("MEG1", e["freqs"]), e["cwt"]).max() <= F64_TOL
    assert peak_rel_err(ew.power("MEG1", e["freqs"]), e["power"]).max() <= F64_TOL
    assert np.abs(ew.itc("MEG1", e["freqs"]) - e["itc"]).max() <= 1e-9
    # cfg3: z-scored power rows, fused epilogue
    m = nw.Morlet(1000, 7., cuda=True)
    zs = m.power(e["data"][:, 1, :], e["freqs"], baseline=("zscore", 0.0, 0.2))
    assert peak_rel_err(zs, e["zscore_power"]).max() <= 1e-10
    zs32 = nw.Morlet(1000, 7., cuda=True, dtype="float32").power(e["data"][:, 1, :].astype(np.float32), e["freqs"],
                                                                 baseline=("zscore", 0.0, 0.2))
    ez = l2_rel_err(zs32, e["zscore_power"]).max()
    assert ez <= F32_TOL, ez


def test_api_quirks_and_extensions(nw):
    import torch
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5, 1000))
    m = nw.Morse(1000, cuda=True)
    batch = m.power(x, np.arange(1, 41.0))
    assert batch.shape == (5, 40, 1000)
    for i in range(5):   # batched [S, N] input is bit-identical to the reference-style 1-D calls
        assert np.array_equal(nw.Morse(1000, cuda=True).power(x[i], np.arange(1, 41.0)), batch[i])
    # reuse=True ignores new freqs (reference base.py:394-395)
    again = m.power(x[0], np.arange(50, 60.0))
    assert np.array_equal(again, batch[0])
    fresh = m.power(x[0], np.arange(50, 60.0), reuse=False)
    assert fresh.shape == (10, 1000)
    # a cached bank of another length is pad_to'ed onto the new signal (base.py:396-397)
    fam = orc.Family("morse")
    bank = orc.make_fft_wavelets(fam, np.arange(50, 60.0), 1.0)
    ref = orc.cwt(fam, x[1][:700], None, bank=bank)
    got = m.cwt(x[1][:700], None)
    assert peak_rel_err(got, ref).max() <= F64_TOL
    # torch CUDA tensor in -> tensor out, same numbers as the host path
    xt = torch.as_tensor(x, device="cuda")
    m2 = nw.Morse(1000, cuda=True)
    pt = m2.power(xt, np.arange(1, 41.0))
    assert pt.is_cuda and np.array_equal(pt.cpu().numpy(), batch)
    # pruning off vs default
    full = nw.Morse(1000, cuda=True, prune_eps=0.0).power(x, np.arange(1, 41.0))
    assert peak_rel_err(full, batch).max() <= 1e-14
    with pytest.raises(ZeroDivisionError):
        nw.Morse(1000, cuda=True).cwt(x[0], [0.0, 1.0])
    with pytest.raises(RuntimeError):
        nw.Morse(1000).cwt(x[0], [1.0, 2.0])     # cuda=False: no CPU path in this package


def test_user_subclass_plugin(nw):
    """README.md:342-355 plugin API: a numpy trans_formula in a user subclass is tabulated and uploaded."""
    class Bump(nw.WaveletBase):
        def __init__(self, sfreq=1000, **kw):
            super().__init__(sfreq, 1., False, True, **kw)
            self.mode = nw.WaveletMode.Reverse

        def trans_formula(self, freqs, freq=1.):
            return np.exp(-np.square((freqs - freq) / (0.2 * freq)))

    rng = np.random.default_rng(2)
    x = rng.standard_normal(1200)
    fr = np.array([5.0, 20.0, 80.0])
    grid = np.arange(0, 1000 / 1.2 * 1.2, 1 / 1.2)
    bank = np.array([orc.pad_to(np.exp(-np.square((grid - f) / (0.2 * f))), 1200) for f in fr])
    from scipy.fftpack import fft, ifft
    ref = ifft(bank * fft(x))
    assert peak_rel_err(Bump().cwt(x, fr), ref).max() <= F64_TOL


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-5)])
def test_properties_at_full_size(nw, dtype, tol):
    """Size-independent checks at a BASELINE.json config-2 row size (N = 600 000)."""
    n = 600000
    t = np.arange(n) / 1000.0
    fr = np.array([10.0, 40.0, 100.0])
    m = make(nw, "morse", dict(sfreq=1000), dtype=dtype)
    # (1) unit sinusoid at the analysis frequency has Morse power 1 (peak of the spectrum is exactly 2)
    p = m.power(np.sin(2 * np.pi * 40.0 * t), fr)
    assert abs(np.median(p[1]) - 1.0) <= 10 * tol and p[0].max() < 1e-6, abs(np.median(p[1]) - 1.0)
    # (2) linearity in the signal
    rng = np.random.default_rng(4)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    za, zb, zab = m.cwt(a, None), m.cwt(b, None), m.cwt(a + 2 * b, None)
    # three transforms, each within tol of the exact one: their combination is within 4 tol (|a| + 2 |b| + |a + 2b|)
    elin = l2_rel_err(zab.astype(np.complex128), (za + 2 * zb).astype(np.complex128)).max()
    assert elin <= 4 * tol, elin
    # (3) Parseval: sum_n |z|^2 = (1/N) sum_k |W X|^2
    X = np.fft.fft(a)
    k = np.arange(n) * (1 / (n / 1000.0))
    for i, f in enumerate(fr):
        W = orc.analytic_spectrum(orc.Family("morse"), k, f)
        lhs = (np.abs(za[i].astype(np.complex128)) ** 2).sum()
        rhs = (np.abs(W * X) ** 2).sum() / n
        assert abs(lhs - rhs) / rhs <= 2 * tol, abs(lhs - rhs) / rhs   # energy: twice the amplitude error


def test_forward_fft_entry_point(nw):
    import torch
    from ninwavelets_b200 import _backend as be
    rng = np.random.default_rng(0)
    for n in (300, 1001, 4096, 65536, 600000):
        x = rng.standard_normal((2, n))
        plan = be.Plan(device=0, dtype=np.float64, family=be.SHANNON, interpolate=False, n=n, sfreq=1000.0, freqs=[])
        X = plan.forward_device(torch.as_tensor(x, device="cuda")).cpu().numpy()
        ref = np.fft.fft(x, axis=1)
        assert np.abs(X - ref).max() / np.abs(ref).max() <= 1e-14, n


@pytest.mark.parametrize("kind,kw", [("morse", dict(sfreq=1000)), ("morlet", dict(sfreq=1000, sigma=7.)),
                                     ("morlet", dict(sfreq=1000, sigma=7., gabor=True)), ("shannon", dict(sfreq=1000)),
                                     ("mexicanhat", dict(sfreq=1000))])
def test_family_sweep_at_2_20(nw, kind, kw):
    """BASELINE.json config 4: every family at N = 2^20 against the oracle (fp32 tolerance, fp64 on one family)."""
    n = 1 << 20
    rng = np.random.default_rng(4)
    x = rng.standard_normal(n)
    fr = np.array([1.0, 9.0, 60.0, 128.0])
    fam = orc.Family(kind, **kw)
    x32 = x.astype(np.float32)
    ref = orc.power(fam, x32.astype(np.float64), fr)
    p = make(nw, kind, kw, dtype="float32").power(x32, fr)
    assert l2_rel_err(p.astype(np.float64), ref).max() <= F32_TOL, kind
    if kind == "morse":
        z = make(nw, kind, kw, dtype="float64").cwt(x, fr)
        assert peak_rel_err(z, orc.cwt(fam, x, fr)).max() <= F64_TOL


@pytest.mark.parametrize("e", [16, 18, 22, 24])
def test_long_signal_sweep_vs_oracle(nw, e):
    """BASELINE.json config 5 sizes that the oracle can still hold: two frequencies, fp32 and (up to 2^22) fp64."""
    n = 1 << e
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    fr = np.array([3.0, 200.0])
    fam = orc.Family("morse", sfreq=1000)
    x32 = x.astype(np.float32)
    ref = orc.power(fam, x32.astype(np.float64), fr)
    p = make(nw, "morse", dict(sfreq=1000), dtype="float32").power(x32, fr)
    assert l2_rel_err(p.astype(np.float64), ref).max() <= F32_TOL, e
    del p, ref
    if e <= 24:
        z = make(nw, "morse", dict(sfreq=1000), dtype="float64").cwt(x, fr)
        assert peak_rel_err(z, orc.cwt(fam, x, fr)).max() <= F64_TOL, e


@pytest.mark.parametrize("ring_rows", [1, 3])
def test_mixed_band_widths_narrow_and_general_pass_a(nw, ring_rows, monkeypatch):
    """Frequencies whose bands are narrow enough for the no-load first pass of pass A next to ones that are not:
    launch groups of 1 or 3 rows make the library pick the kernel group by group (nw_plan.h: group_narrow)."""
    monkeypatch.setenv("NWCWT_RING_ROWS", str(ring_rows))
    n = 1 << 18
    rng = np.random.default_rng(29)
    x = rng.standard_normal((2, n))
    fr = np.array([1.0, 4.0, 20.0, 90.0, 180.0, 300.0, 450.0])
    fam = orc.Family("morse", sfreq=1000)
    x32 = x.astype(np.float32)
    p = make(nw, "morse", dict(sfreq=1000), dtype="float32").power(x32, fr)
    for i in range(2):
        assert l2_rel_err(p[i].astype(np.float64), orc.power(fam, x32[i].astype(np.float64), fr)).max() <= F32_TOL
    z = make(nw, "morse", dict(sfreq=1000), dtype="float64").cwt(x, fr)
    for i in range(2):
        assert peak_rel_err(z[i], orc.cwt(fam, x[i], fr)).max() <= F64_TOL


def test_2_26_properties(nw):
    """N = 2^26 (config 5's largest row; only the packed kernels have a plan): Parseval and linearity in fp32."""
    import torch
    n = 1 << 26
    rng = np.random.default_rng(6)
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    fr = np.array([5.0, 120.0])
    m = make(nw, "morse", dict(sfreq=1000), dtype="float32")
    ta, tb = torch.as_tensor(a, device="cuda")[None], torch.as_tensor(b, device="cuda")[None]
    za = m.cwt(ta, fr)[0]
    zb = m.cwt(tb, None)[0]
    zab = m.cwt(ta + 2 * tb, None)[0]
    d = (zab - (za + 2 * zb))
    num = torch.sqrt((d.real.double() ** 2 + d.imag.double() ** 2).sum(dim=1))
    den = torch.sqrt((zab.real.double() ** 2 + zab.imag.double() ** 2).sum(dim=1))
    # three results, each within F32_TOL of the truth: the residual of  z(a + 2b) - z(a) - 2 z(b)  is bounded by the sum
    # of their errors, (1 + 1 + 2) F32_TOL relative to rows of comparable size (the direct 2^26 comparison with the oracle at
    # 1e-5 is test_2_26_against_oracle)
    assert float((num / den).max()) <= 4 * F32_TOL, float((num / den).max())
    X = torch.fft.fft(ta[0].double())
    k = np.arange(n) * (1 / (n / 1000.0))
    for i, f in enumerate(fr):
        W = torch.as_tensor(orc.analytic_spectrum(orc.Family("morse"), k, f), device="cuda")
        lhs = float((za[i].real.double() ** 2 + za[i].imag.double() ** 2).sum())
        rhs = float(((W * X).abs() ** 2).sum()) / n
        assert abs(lhs - rhs) / rhs <= 2 * F32_TOL, abs(lhs - rhs) / rhs   # energy: twice the relative error of the amplitude


@pytest.mark.parametrize("n", [50625, 56250, 57344, 98304, 30375, 20000, 16384 + 8192])
def test_awkward_long_lengths(nw, n):
    """Odd N1 (unaligned output pairs), partial row tiles, a factor 7 (generic kernels), mixed 2-3-5 lengths: every long-row
    variant against the oracle in fp64 and fp32, cwt and power with a Baseline epilogue."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n))
    fr = np.array([2.0, 17.0, 140.0])
    fam = orc.Family("morlet", sfreq=1000, sigma=7.)
    ref = np.stack([orc.cwt(fam, xi, fr) for xi in x])
    z = make(nw, "morlet", dict(sfreq=1000, sigma=7.), dtype="float64").cwt(x, fr)
    assert peak_rel_err(z.reshape(-1, n), ref.reshape(-1, n)).max() <= F64_TOL, n
    x32 = x.astype(np.float32)
    refp = np.stack([orc.baseline_rows(orc.power(fam, xi.astype(np.float64), fr), 1000., 0.1, 0.9, "zscore") for xi in x32])
    p = make(nw, "morlet", dict(sfreq=1000, sigma=7.), dtype="float32").power(x32, fr, baseline=("zscore", 0.1, 0.9))
    ep = l2_rel_err(p.reshape(-1, n).astype(np.float64), refp.reshape(-1, n)).max()
    assert ep <= F32_TOL, (n, ep)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_resampled_rows_at_cfg2_length(nw, dtype):
    """Resampled rows at BASELINE.json config-2 row size: power and abs of every row against the oracle,
    each row relative to itself, and against the same family with resampling switched off."""
    n = 600000
    rng = np.random.default_rng(31)
    t = np.arange(n) / 1000.0
    x = rng.standard_normal((2, n)) + np.sin(2 * np.pi * 10 * t) + 0.5 * np.sin(2 * np.pi * 60 * t + 1.0)
    fr = np.array([1.0, 3.0, 7.0, 18.0, 30.0, 55.0, 100.0])
    fam = orc.Family("morse", sfreq=1000)
    xin = x.astype(np.float32) if dtype == "float32" else x
    m = make(nw, "morse", dict(sfreq=1000), dtype=dtype)
    p = m.power(xin, fr)
    groups = m._plan.info()["groups"]
    assert groups and all(g["D"] >= 2 for g in groups), groups          # every row of this plan is resampled
    ref = np.stack([orc.power(fam, xi.astype(np.float64), fr) for xi in xin])
    if dtype == "float32":
        e = l2_rel_err(p.reshape(-1, n).astype(np.float64), ref.reshape(-1, n))
        assert e.max() <= F32_TOL, e
        a = m.abs(xin, None)
        assert l2_rel_err(a.reshape(-1, n).astype(np.float64), np.sqrt(ref).reshape(-1, n)).max() <= F32_TOL
    else:
        e = peak_rel_err(p.reshape(-1, n), ref.reshape(-1, n))
        assert e.max() <= F64_TOL, e
    exact = make(nw, "morse", dict(sfreq=1000), dtype=dtype, resample=False).power(xin, fr)
    d = l2_rel_err(p.reshape(-1, n).astype(np.float64), exact.reshape(-1, n).astype(np.float64)).max()
    assert d <= (F32_TOL if dtype == "float32" else 1e-12), d
    print("resampled %s: worst row error %.3e, vs exact rows %.3e, groups %s" % (dtype, e.max(), d, [(g["D"], g["K"]) for g in groups]))


def test_2_26_against_oracle(nw):
    """N = 2^26 (config 5's largest row) against the oracle on two frequencies, fp32, rows relative to themselves."""
    n = 1 << 26
    rng = np.random.default_rng(8)
    x32 = rng.standard_normal(n).astype(np.float32)
    fr = np.array([5.0, 120.0])
    p = make(nw, "morse", dict(sfreq=1000), dtype="float32").power(x32, fr)
    ref = orc.power(orc.Family("morse", sfreq=1000), x32.astype(np.float64), fr)
    e = l2_rel_err(p.astype(np.float64), ref)
    assert e.max() <= F32_TOL, e


def test_interpolate_true_on_long_rows(nw):
    """`interpolate=True` (the WaveletBase default, base.py:239-242, 276) at N = 600 000: half-grid spectrum, aliased half
    zeroed; fp64 complex transform and fp32 power against the oracle."""
    n = 600000
    rng = np.random.default_rng(9)
    x = rng.standard_normal(n)
    fr = np.array([2.0, 40.0, 230.0])
    fam = orc.Family("morse", sfreq=1000, interpolate=True)
    z = make(nw, "morse", dict(sfreq=1000, interpolate=True), dtype="float64").cwt(x, fr)
    assert peak_rel_err(z, orc.cwt(fam, x, fr)).max() <= F64_TOL
    x32 = x.astype(np.float32)
    p = make(nw, "morse", dict(sfreq=1000, interpolate=True), dtype="float32").power(x32, fr)
    assert l2_rel_err(p.astype(np.float64), orc.power(fam, x32.astype(np.float64), fr)).max() <= F32_TOL


@pytest.mark.parametrize("mode", ["mean", "ratio", "percent", "log", "zlog"])
def test_fused_baseline_epilogues(nw, mode):
    """The five Baseline modes other than zscore (base.py:52-68) as epilogues of the fused short-row kernel
    (`power(..., baseline=)`, N = 1500) and of the long-row path (N = 60 000)."""
    fam = orc.Family("morlet", sfreq=1000, sigma=7.)
    fr = np.arange(2, 42.0, 3.0)
    for n in (1500, 60000):
        x = orc.meg_epochs_like(4, n) if n == 1500 else np.random.default_rng(3).standard_normal((2, n))
        ref = np.stack([orc.baseline_rows(orc.power(fam, xi, fr), 1000., 0.1, 0.3, mode) for xi in x])
        got = make(nw, "morlet", dict(sfreq=1000, sigma=7.), dtype="float64").power(x, fr, baseline=(mode, 0.1, 0.3))
        assert peak_rel_err(got.reshape(-1, n), ref.reshape(-1, n)).max() <= 1e-11, (mode, n)
        x32 = x.astype(np.float32)
        ref32 = np.stack([orc.baseline_rows(orc.power(fam, xi.astype(np.float64), fr), 1000., 0.1, 0.3, mode) for xi in x32])
        got32 = make(nw, "morlet", dict(sfreq=1000, sigma=7.), dtype="float32").power(x32, fr, baseline=(mode, 0.1, 0.3))
        e = l2_rel_err(got32.reshape(-1, n).astype(np.float64), ref32.reshape(-1, n)).max()
        assert e <= F32_TOL, (mode, n, e)


@pytest.mark.parametrize("n", [65536, 600000])
def test_generic_kernels_forced(nw, n):
    """nwcwt_debug_force_generic: the generic (any factor <= 64) two-pass kernels on lengths the packed kernels normally
    take, against the oracle and against the packed kernels' own result."""
    from ninwavelets_b200 import _backend as be
    rng = np.random.default_rng(12)
    x = rng.standard_normal(n)
    fr = np.array([4.0, 33.0, 150.0])
    fam = orc.Family("morse", sfreq=1000)
    ref = orc.cwt(fam, x, fr)
    fast = make(nw, "morse", dict(sfreq=1000), dtype="float64").cwt(x, fr)
    be.force_generic(True)
    try:
        slow = make(nw, "morse", dict(sfreq=1000), dtype="float64").cwt(x, fr)
        p32 = make(nw, "morse", dict(sfreq=1000), dtype="float32").power(x.astype(np.float32), fr)
    finally:
        be.force_generic(False)
    assert peak_rel_err(slow, ref).max() <= F64_TOL and peak_rel_err(fast, ref).max() <= F64_TOL
    assert peak_rel_err(slow, fast).max() <= F64_TOL
    ref32 = orc.power(fam, x.astype(np.float32).astype(np.float64), fr)
    assert l2_rel_err(p32.astype(np.float64), ref32).max() <= F32_TOL


@pytest.mark.parametrize("n_ep", [200, 5])
def test_fused_epoch_reductions(nw, n_ep):
    """mneutils.py:53-55 / :68-71 with the reduction over epochs inside the transform kernel (nwcwt_transform_epochs): the
    (E, F, T) rows are never written.  200 epochs (config 3's count) and an odd count, two channels, against the oracle's
    per-epoch transforms reduced in fp64."""
    import torch
    from ninwavelets_b200 import _backend as be
    fam = orc.Family("morlet", sfreq=1000, sigma=7.)
    fr = np.arange(1, 41.0)
    x = np.stack([orc.meg_epochs_like(n_ep, 1500, seed=3), orc.meg_epochs_like(n_ep, 1500, seed=4)])     # (C, E, T)
    zref = np.stack([np.stack([orc.cwt(fam, xe, fr) for xe in xc]) for xc in x])                          # (C, E, F, T)
    pref = (np.abs(zref) ** 2).mean(axis=1)
    iref = np.abs((zref / np.abs(zref)).mean(axis=1))
    m = nw.Morlet(1000, 7., cuda=True)
    m.make_fft_wavelets(fr, 1.5)
    xt = torch.as_tensor(x, device="cuda")
    l0 = be.launch_count()
    p = m._plan.transform_epochs_device(xt, 0).cpu().numpy()
    assert be.launch_count() - l0 == 1                       # one kernel: nothing is materialised and re-read
    assert peak_rel_err(p.reshape(-1, 1500), pref.reshape(-1, 1500)).max() <= F64_TOL
    itc = m._plan.transform_epochs_device(xt, 1).cpu().numpy()
    assert np.abs(itc - iref).max() <= 1e-12
    m32 = nw.Morlet(1000, 7., cuda=True, dtype="float32")
    m32.make_fft_wavelets(fr, 1.5)
    x32 = x.astype(np.float32)
    z32 = np.stack([np.stack([orc.cwt(fam, xe.astype(np.float64), fr) for xe in xc]) for xc in x32])
    p32 = m32._plan.transform_epochs_device(torch.as_tensor(x32, device="cuda"), 0).cpu().numpy()
    e = l2_rel_err(p32.reshape(-1, 1500).astype(np.float64), (np.abs(z32) ** 2).mean(axis=1).reshape(-1, 1500)).max()
    assert e <= F32_TOL, e
    i32 = m32._plan.transform_epochs_device(torch.as_tensor(x32, device="cuda"), 1).cpu().numpy()
    # the unit phasor z / |z| of a sample whose |z| is far below the row's norm carries that sample's RELATIVE rounding error
    # (inherent to the definition): the fp32 bar is the per-row relative L2 error, like every other fp32 comparison
    iref32 = np.abs((z32 / np.abs(z32)).mean(axis=1))
    ei = l2_rel_err(i32.reshape(-1, 1500).astype(np.float64), iref32.reshape(-1, 1500)).max()
    assert ei <= F32_TOL, ei


@pytest.mark.parametrize("n", [100003, 2 * 10007, 1234, 599999])
def test_any_length_chirp_z(nw, n):
    """Lengths with a prime factor > 64 (the reference's scipy.fftpack takes any N, base.py:399, 406): Bluestein's
    chirp-z algorithm on the packed engine, exact circular semantics; 100 003 is prime."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n))
    fr = np.array([2.0, 17.0, 140.0])
    fam = orc.Family("morse", sfreq=1000)
    m = make(nw, "morse", dict(sfreq=1000), dtype="float64")
    z = m.cwt(x, fr)
    assert m._plan.info()["path"] == "chirp_z"
    ref = np.stack([orc.cwt(fam, xi, fr) for xi in x])
    assert peak_rel_err(z.reshape(-1, n), ref.reshape(-1, n)).max() <= F64_TOL, n
    x32 = x.astype(np.float32)
    refp = np.stack([orc.baseline_rows(orc.power(fam, xi.astype(np.float64), fr), 1000., 0.1, 0.9, "zscore") for xi in x32])
    p = make(nw, "morse", dict(sfreq=1000), dtype="float32").power(x32, fr, baseline=("zscore", 0.1, 0.9))
    e = l2_rel_err(p.reshape(-1, n).astype(np.float64), refp.reshape(-1, n)).max()
    assert e <= F32_TOL, (n, e)
    # a Normal-mode family (tabulated spectrum) and the forward entry point at an awkward length
    mh = make(nw, "mexicanhat", dict(sfreq=1000), dtype="float64").cwt(x[0], fr)
    assert peak_rel_err(mh, orc.cwt(orc.Family("mexicanhat", sfreq=1000), x[0], fr)).max() <= F64_TOL


def test_too_long_raises(nw):
    from ninwavelets_b200 import _backend as be
    with pytest.raises(be.BackendError):
        be.Plan(device=0, dtype=np.float32, family=be.MORSE, interpolate=False, n=(1 << 31) + 7, sfreq=1000.0,
                freqs=[1.0, 2.0], p0=17.5, p1=3.0)


@pytest.mark.parametrize("mode", [None, "zscore", "mean", "ratio", "percent", "log", "zlog"])
def test_resampled_short_rows_cfg3_slice(nw, mode, monkeypatch):
    """Resampled SHORT rows (nw_kernels4.cuh) at BASELINE.json config-3 row size: Morlet(7) power at 1..100 Hz on an odd
    number of 1500-sample epochs, fp32, plain and with every Baseline mode as the kernel's epilogue, against the oracle
    (every row relative to itself) and against the exact short-row kernel (`resample=False`); abs output as well.
    NWCWT_SHORT3=0 (read when the plan is created) switches the path off."""
    n = 1500
    fam = orc.Family("morlet", sfreq=1000, sigma=7.)
    fr = np.arange(1, 101.0)
    x32 = orc.meg_epochs_like(7, n).astype(np.float32)
    m = make(nw, "morlet", dict(sfreq=1000, sigma=7.), dtype="float32")
    bl = None if mode is None else (mode, 0.0, 0.2)
    p = m.power(x32, fr, baseline=bl)
    groups = m._plan.info()["groups"]
    assert groups and sum(g["rows"] for g in groups) == 100 and any(g["D"] > 1 for g in groups), groups
    ref = np.stack([orc.power(fam, xi.astype(np.float64), fr) for xi in x32])
    refb = ref if mode is None else np.stack([orc.baseline_rows(r, 1000., 0.0, 0.2, mode) for r in ref])
    e = l2_rel_err(p.reshape(-1, n).astype(np.float64), refb.reshape(-1, n))
    assert e.max() <= F32_TOL, (mode, e.max())
    exact = make(nw, "morlet", dict(sfreq=1000, sigma=7.), dtype="float32", resample=False)
    pe = exact.power(x32, fr, baseline=bl)
    assert exact._plan.info()["groups"] == []
    d = l2_rel_err(p.reshape(-1, n).astype(np.float64), pe.reshape(-1, n).astype(np.float64)).max()
    assert d <= F32_TOL, (mode, d)
    if mode is None:
        a = m.abs(x32, None)
        assert l2_rel_err(a.reshape(-1, n).astype(np.float64), np.sqrt(ref).reshape(-1, n)).max() <= F32_TOL
    print("short resampled rows, baseline %s: worst row %.3e, vs exact kernel %.3e, groups %s" % (
        mode, e.max(), d, [(g["D"], g["K"], g["rows"]) for g in groups]))


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_graph_replay_matches_direct_launches(nw, dtype):
    """nwcwt_transform on the fast long path: the first call with an argument set launches directly, the second records a CUDA
    graph, later ones replay it.  Replays must give the direct launches' bits, count the same launches, and read the
    buffers' CURRENT contents (nothing of the data is baked into the graph)."""
    import torch
    from ninwavelets_b200 import _backend as be
    n, fr = 60000, np.arange(2, 42.0)
    obj = make(nw, "morse", dict(sfreq=1000), dtype=dtype)
    obj.make_fft_wavelets(fr, n / 1000.0)
    plan = obj._plan
    tdt = torch.float32 if dtype == "float32" else torch.float64
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((3, n), device="cuda", dtype=tdt, generator=g)
    out = torch.empty((3, len(fr), n), device="cuda", dtype=tdt)
    l0 = be.launch_count()
    plan.transform_device(x, be.OUT_POWER, out=out)          # direct launches
    n_direct = be.launch_count() - l0
    direct = out.clone()
    counts = []
    for _ in range(3):                                        # records, then replays
        out.zero_()
        l0 = be.launch_count()
        plan.transform_device(x, be.OUT_POWER, out=out)
        counts.append(be.launch_count() - l0)
        assert torch.equal(out, direct)
    assert counts == [n_direct] * 3 and n_direct > 3
    x.copy_(torch.randn((3, n), device="cuda", dtype=tdt, generator=g))
    plan.transform_device(x, be.OUT_POWER, out=out)          # replay on new data in the same buffers
    fresh = plan.transform_device(x.clone(), be.OUT_POWER)   # new argument set: direct launches
    assert torch.equal(out, fresh)
    xs = x[0].double().cpu().numpy()
    ref = orc.power(orc.Family("morse", sfreq=1000), xs, fr)
    got = out[0].double().cpu().numpy()
    if dtype == "float32":
        assert l2_rel_err(got, ref).max() <= F32_TOL
    else:
        assert peak_rel_err(got, ref).max() <= F64_TOL


def test_plan_close_leaves_no_cuda_error_behind(nw):
    """A plan whose graph cache holds argument sets that were only seen once (no recorded graph) is destroyed; the next
    launches - the host-buffer entry point of another plan, whose kernels check cudaGetLastError - must not see a stale
    CUDA error (r02: cudaGraphExecDestroy(nullptr) left 'invalid argument' for bench.py's end-to-end leg)."""
    import torch
    from ninwavelets_b200 import _backend as be
    n, fr = 60000, np.array([3.0, 17.0, 40.0])
    a = make(nw, "morse", dict(sfreq=1000), dtype="float32")
    a.make_fft_wavelets(fr, n / 1000.0)
    b = make(nw, "morse", dict(sfreq=1000), dtype="float32")
    b.make_fft_wavelets(fr, n / 1000.0)
    x = torch.randn((2, n), device="cuda", dtype=torch.float32)
    a._plan.transform_device(x, be.OUT_POWER)        # one-off argument sets: remembered, never recorded
    a._plan.transform_device(x[:1], be.OUT_POWER)
    torch.cuda.synchronize()
    a._plan.close()
    hx = np.random.default_rng(4).standard_normal((2, n)).astype(np.float32)
    got = b._plan.transform_host(hx, be.OUT_POWER)
    ref = orc.power(orc.Family("morse", sfreq=1000), hx[1].astype(np.float64), fr)
    assert l2_rel_err(got[1].astype(np.float64), ref).max() <= F32_TOL
