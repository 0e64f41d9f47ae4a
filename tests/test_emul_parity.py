"""CPU check of the CUDA kernel bodies (host emulation, tests/emul) against the golden fixtures.

This does NOT replace the GPU parity tests: it validates the arithmetic and index algebra of
the exact code the kernels run (same headers), sequentially, so that GPU time is not spent on
logic errors.  Tolerances are the BASELINE.json ones: fp64 max|out-ref|/max|ref| per row <=
1e-12, fp32 per-row relative L2 <= 1e-5 against the fp64 reference on the float32-rounded input.
"""
import numpy as np
import pytest

import cwt_oracle as orc
from emul_util import desc_from_oracle_family, emul_transform
from golden_util import case_floor, case_wave, l2_rel_err, peak_rel_err

F64_TOL = 1e-12
F32_TOL = 1e-5

SMALL = ["readme_morse", "readme_morlet", "morse_n300", "morse_n301", "morlet_n1000", "gabor_n1500", "morse_b3_n1500",
         "morlet_s5_n300", "shannon_n1500", "mexicanhat_n1500", "mexicanhat_n301", "haar_n1000", "morse_n4096",
         "morse_interp_n301", "morlet_interp_n1500", "shannon_interp_n300", "mexicanhat_interp_n300",
         "morse_fractional", "morse_beyond_nyquist", "morlet_beyond_nyquist", "morse_sfreq256", "mexicanhat_rwl2"]


@pytest.mark.parametrize("name", SMALL)
def test_emul_fp64_matches_reference(golden_transforms, name):
    c = golden_transforms[name]
    fam = orc.Family(c["kind"], **c["kw"])
    x = case_wave(c)
    d = desc_from_oracle_family(fam, c["freqs"], len(x), dtype=1)
    z = emul_transform(d, x[None, :], output=0)[0]
    fl = case_floor(name)
    assert peak_rel_err(z, c["cwt"], fl).max() <= F64_TOL, name
    p = emul_transform(d, x[None, :], output=2)[0]
    assert peak_rel_err(p, np.abs(c["cwt"]) ** 2, fl).max() <= F64_TOL, name


@pytest.mark.parametrize("name", ["readme_morse", "morse_n1500", "morlet_n1500", "shannon_n1000", "mexicanhat_n1500",
                                  "morse_interp_n300", "morse_n4096"])
def test_emul_fp32_matches_reference(golden_transforms, name):
    c = golden_transforms[name]
    fam = orc.Family(c["kind"], **c["kw"])
    x32 = case_wave(c).astype(np.float32)
    ref = orc.cwt(fam, x32.astype(np.float64), c["freqs"])
    d = desc_from_oracle_family(fam, c["freqs"], len(x32), dtype=0)
    z = emul_transform(d, x32[None, :], output=0)[0]
    err = l2_rel_err(z.astype(np.complex128), ref, case_floor(name, f32=True))
    assert (err <= F32_TOL).all(), (name, err)


@pytest.mark.parametrize("name,dtype", [("morse_n1500", 1), ("morlet_n1000", 1), ("mexicanhat_n1500", 1),
                                        ("morse_n4096", 0), ("shannon_n1500", 1), ("morse_interp_n301", 1)])
def test_emul_long_path_forced(golden_transforms, name, dtype):
    """The two-pass (four-step) kernels on sizes small enough to emulate."""
    c = golden_transforms[name]
    fam = orc.Family(c["kind"], **c["kw"])
    x = case_wave(c)
    if dtype == 0:
        x = x.astype(np.float32)
    ref = orc.cwt(fam, x.astype(np.float64), c["freqs"])
    d = desc_from_oracle_family(fam, c["freqs"], len(x), dtype=dtype)
    z = emul_transform(d, x[None, :], output=0, force_long=True)[0]
    if dtype == 1:
        assert peak_rel_err(z, ref).max() <= F64_TOL
    else:
        assert l2_rel_err(z.astype(np.complex128), ref).max() <= F32_TOL


@pytest.mark.parametrize("flags", [1, 1 | 16, 1 | 8 | 16, 1 | 2, 1 | 4, 1 | 32, 1 | 16 | 64])
def test_emul_long_path_variants_agree(flags):
    """Same input through every variant of the long path: pruned pass A (default), unpruned compile-time or
    run-time plans (16, 8|16), generic kernels (2), one stepping thread instead of fibers (4), generic forward
    transform under the packed inverse (32), no narrow-band first pass (64; 16 alone runs it where the plan allows)."""
    rng = np.random.default_rng(11)
    for kind, kw, n, freqs in (("morse", {}, 6000, np.array([1., 2.5, 9., 30., 77., 210., 499.])),
                                ("morlet", dict(sigma=7.0), 3600, np.arange(1., 60., 7.)),
                                ("mexicanhat", {}, 3000, np.array([2., 8., 40.])),
                                ("shannon", {}, 2400, np.array([1., 2.])),
                                ("morse", dict(interpolate=True), 4500, np.array([3., 50., 400.]))):
        fam = orc.Family(kind, sfreq=1000.0, **kw)
        x = rng.standard_normal((2, n))
        ref = np.stack([orc.cwt(fam, xi, freqs) for xi in x])
        d = desc_from_oracle_family(fam, freqs, n, dtype=1)
        z = emul_transform(d, x, output=0, force_long=flags)
        assert peak_rel_err(z.reshape(-1, n), ref.reshape(-1, n)).max() <= F64_TOL, (kind, flags)
        d32 = desc_from_oracle_family(fam, freqs, n, dtype=0)
        x32 = x.astype(np.float32)
        ref32 = np.stack([orc.power(fam, xi.astype(np.float64), freqs) for xi in x32])
        p = emul_transform(d32, x32, output=2, force_long=flags)
        assert l2_rel_err(p.reshape(-1, n), ref32.reshape(-1, n)).max() <= F32_TOL, (kind, flags)


@pytest.mark.parametrize("ring_rows", [1, 2, 3])
def test_emul_narrow_pass_per_launch_group(ring_rows, monkeypatch):
    """Bands of mixed width: the planner allows the narrow-band first pass of pass A only for launch groups whose
    frequencies all qualify (nw_plan.h: group_narrow); small groups make both kernels run in one transform."""
    monkeypatch.setenv("NWCWT_RING_ROWS", str(ring_rows))
    rng = np.random.default_rng(23)
    fam = orc.Family("morse", sfreq=1000.0)
    for n, freqs in ((6000, np.array([1., 2.5, 9., 30., 77., 210., 499.])), (16384, np.array([2., 11., 60., 200., 450.]))):
        x = rng.standard_normal((2, n))
        ref = np.stack([orc.cwt(fam, xi, freqs) for xi in x])
        d = desc_from_oracle_family(fam, freqs, n, dtype=1)
        z = emul_transform(d, x, output=0, force_long=1 | 16)
        assert peak_rel_err(z.reshape(-1, n), ref.reshape(-1, n)).max() <= F64_TOL
        x32 = x.astype(np.float32)
        ref32 = np.stack([orc.power(fam, xi.astype(np.float64), freqs) for xi in x32])
        p = emul_transform(desc_from_oracle_family(fam, freqs, n, dtype=0), x32, output=2, force_long=1 | 16)
        assert l2_rel_err(p.reshape(-1, n), ref32.reshape(-1, n)).max() <= F32_TOL


def test_emul_long_65536(golden_transforms):
    c = golden_transforms["morse_long_n65536"]
    fam = orc.Family(c["kind"], **c["kw"])
    x = case_wave(c)
    d = desc_from_oracle_family(fam, c["freqs"], len(x), dtype=1)
    z = emul_transform(d, x[None, :], output=0)[0]
    assert peak_rel_err(z[:, c["cols"]], c["cwt"]).max() <= F64_TOL
    p = np.abs(z) ** 2
    np.testing.assert_allclose(p.sum(axis=1), c["row_power_sum"], rtol=1e-11)


def test_emul_baseline_and_batch():
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "epochs.npz"))
    fam = orc.Family("morlet", sfreq=1000.0, sigma=7.0)
    x = z["data"][:, 1, :]
    d = desc_from_oracle_family(fam, z["freqs"], x.shape[1], dtype=1)
    out = emul_transform(d, x, output=2, baseline=5, lo=0, hi=200)
    assert peak_rel_err(out, z["zscore_power"]).max() <= 1e-11
    zc = emul_transform(d, x, output=0)
    assert peak_rel_err(zc, z["cwt"]).max() <= F64_TOL
    for mode, code in (("mean", 1), ("ratio", 2), ("percent", 3), ("log", 4), ("zlog", 6)):
        ref = orc.baseline_rows(np.abs(z["cwt"][:2]) ** 2, 1000.0, 0.1, 0.3, mode)
        got = emul_transform(d, x[:2], output=2, baseline=code, lo=100, hi=300)
        assert peak_rel_err(got, ref).max() <= 1e-11, mode


def test_emul_odd_lengths_and_generic_radix():
    rng = np.random.default_rng(5)
    for n in (2, 3, 7, 77, 91, 143, 1001, 1331, 2 * 3 * 5 * 7 * 11):
        x = rng.standard_normal(n)
        fr = np.array([3.0, 20.0, 110.0])
        fam = orc.Family("morse", sfreq=1000.0)
        ref = orc.cwt(fam, x, fr)
        d = desc_from_oracle_family(fam, fr, n, dtype=1)
        z = emul_transform(d, x[None, :], output=0)[0]
        assert peak_rel_err(z, ref).max() <= F64_TOL, n


@pytest.mark.parametrize("kind,kw,n,freqs", [
    ("morse", {}, 24000, np.array([1., 2., 4., 7., 12., 20., 45., 100.])),
    ("morlet", dict(sigma=7.0), 30000, np.array([2., 5., 11., 30., 80.])),
    ("morse", dict(interpolate=True), 36000, np.array([3., 9., 50.])),
    ("mexicanhat", {}, 24000, np.array([2., 8., 40.])),
])
def test_emul_resampled_rows(kind, kw, n, freqs, monkeypatch):
    """Resampled rows: band-limited rows are transformed at a decimated length N / D and interpolated back
    (nw_plan.h: plan_multirate; nw_resample.cuh).  The planner's shortest decimated length is lowered so that lengths
    the emulation can afford take the path; every row is checked against the reference RELATIVE TO ITSELF (no floor),
    in fp32 (power, abs) and fp64 (power), and against the same plan forced to exact length-N transforms (flag 128)."""
    monkeypatch.setenv("NWCWT_RESAMPLE_MMIN", "600")
    rng = np.random.default_rng(17)
    fam = orc.Family(kind, sfreq=1000.0, **kw)
    x = rng.standard_normal((2, n))
    x32 = x.astype(np.float32)
    ref32 = np.stack([orc.power(fam, xi.astype(np.float64), freqs) for xi in x32])
    d32 = desc_from_oracle_family(fam, freqs, n, dtype=0)
    p = emul_transform(d32, x32, output=2, force_long=1)
    e = l2_rel_err(p.reshape(-1, n), ref32.reshape(-1, n))
    assert e.max() <= F32_TOL, (kind, e)
    a = emul_transform(d32, x32, output=1, force_long=1)
    assert l2_rel_err(a.reshape(-1, n), np.sqrt(ref32).reshape(-1, n)).max() <= F32_TOL, kind
    pe = emul_transform(d32, x32, output=2, force_long=1 | 128)          # exact transforms, same plan
    assert l2_rel_err(p.reshape(-1, n), pe.astype(np.float64).reshape(-1, n)).max() <= F32_TOL, kind
    ref = np.stack([orc.power(fam, xi, freqs) for xi in x])
    p64 = emul_transform(desc_from_oracle_family(fam, freqs, n, dtype=1), x, output=2, force_long=1)
    assert peak_rel_err(p64.reshape(-1, n), ref.reshape(-1, n)).max() <= F64_TOL, kind
    # the complex transform of the same plan never resamples (the modulation matters there): exact rows
    z = emul_transform(desc_from_oracle_family(fam, freqs, n, dtype=1), x[:1], output=0, force_long=1)
    assert peak_rel_err(z[0], orc.cwt(fam, x[0], freqs)).max() <= F64_TOL, kind


def test_emul_resampling_is_planned_and_bounded(monkeypatch):
    """Planner facts through the C ABI (host only): cfg2's plan resamples every frequency, reports the decimation, the
    tap count and the error bound of every group, keeps the bound below the default tolerance, and `resample=False` /
    a complex-output transform keep the exact path available."""
    from ninwavelets_b200 import _backend as be
    plan = be.Plan(device=0, dtype=np.float32, family=be.MORSE, interpolate=False, n=600000, sfreq=1000.0,
                   freqs=np.arange(1, 101.0), p0=17.5, p1=3.0)
    g = plan.info()["groups"]
    assert g and sum(x["rows"] for x in g) == 100 and all(x["D"] >= 2 for x in g)
    assert all(0 < x["err"] <= 1e-6 and x["K"] % 2 == 0 and 4 <= x["K"] <= 16 for x in g)
    assert all(x["n1"] * x["n2"] * x["D"] == 600000 for x in g)
    pts = sum(x["rows"] * 600000 // x["D"] for x in g)
    assert pts < 0.12 * 100 * 600000          # the decimated transforms hold < 12 % of the exact rows' points
    off = be.Plan(device=0, dtype=np.float32, family=be.MORSE, interpolate=False, n=600000, sfreq=1000.0,
                  freqs=np.arange(1, 101.0), p0=17.5, p1=3.0, resample=False)
    assert off.info()["groups"] == []
    tight = be.Plan(device=0, dtype=np.float64, family=be.MORSE, interpolate=False, n=600000, sfreq=1000.0,
                    freqs=np.arange(1, 101.0), p0=17.5, p1=3.0)
    assert all(x["err"] <= 5e-14 for x in tight.info()["groups"])


def test_emul_resampled_rows_at_cfg2_length():
    """BASELINE.json config-2 row length (N = 600 000) through the emulation: the decimated lengths 600 000 / D and their
    compile-time plans (100, 120, 150 x 125 ... 1000; nw_kernels2.cuh StaticPlan 20-26), the vector interpolation kernel
    in its persistent form, fp32 power against the oracle, every row relative to itself."""
    rng = np.random.default_rng(5)
    n = 600000
    fam = orc.Family("morse", sfreq=1000.0)
    freqs = np.array([3., 14., 33., 52., 71., 100.])
    x = rng.standard_normal((1, n)).astype(np.float32)
    ref = orc.power(fam, x[0].astype(np.float64), freqs)
    p = emul_transform(desc_from_oracle_family(fam, freqs, n, dtype=0), x, output=2)
    e = l2_rel_err(p[0], ref)
    assert e.max() <= F32_TOL, e


@pytest.mark.parametrize("kind,kw,n,freqs,bl", [
    ("morlet", dict(sigma=7.0), 1500, np.array([1., 3., 8., 13., 21., 34., 55., 89., 100.]), ("zscore", 0.0, 0.2)),
    ("morse", {}, 1500, np.array([2., 5., 17., 60., 99.]), None),
    ("morse", {}, 300, np.array([1., 7., 30., 80.]), None),
    ("shannon", {}, 1200, np.array([3., 10., 40., 90.]), None),
    ("morse", dict(interpolate=True), 1000, np.array([4., 25., 70.]), ("ratio", 0.1, 0.4)),
])
def test_emul_resampled_short_rows(kind, kw, n, freqs, bl, monkeypatch):
    """Resampled SHORT rows (nw_kernels4.cuh; nw_plan.h: plan_multirate(short)): rows that fit one CTA are transformed at
    N / D points and interpolated inside the fused kernel.  fp32 power and abs (with a Baseline epilogue where given)
    against the oracle, every row relative to itself, and against the same plan forced to exact transforms (flag 128:
    nw_kernels3.cuh); three signals (an odd count: the last CTA owns one signal)."""
    rng = np.random.default_rng(23)
    fam = orc.Family(kind, sfreq=1000.0, **kw)
    x32 = rng.standard_normal((3, n)).astype(np.float32)
    d32 = desc_from_oracle_family(fam, freqs, n, dtype=0)
    modes = {"mean": 1, "ratio": 2, "percent": 3, "log": 4, "zscore": 5, "zlog": 6}
    args = {}
    if bl is not None:
        lo, hi = int(bl[1] * 1000), int(bl[2] * 1000)
        args = dict(baseline=modes[bl[0]], lo=lo, hi=hi)
    ref = np.stack([orc.power(fam, xi.astype(np.float64), freqs) for xi in x32])
    refb = ref if bl is None else np.stack([orc.baseline_rows(r, 1000.0, bl[1], bl[2], bl[0]) for r in ref])
    p = emul_transform(d32, x32, output=2, **args)
    e = l2_rel_err(p.reshape(-1, n), refb.reshape(-1, n))
    assert e.max() <= F32_TOL, (kind, e)
    pe = emul_transform(d32, x32, output=2, force_long=128, **args)          # exact transforms (short2 kernel), same plan
    assert l2_rel_err(p.reshape(-1, n), pe.astype(np.float64).reshape(-1, n)).max() <= F32_TOL, kind
    a = emul_transform(d32, x32, output=1)
    assert l2_rel_err(a.reshape(-1, n), np.sqrt(ref).reshape(-1, n)).max() <= F32_TOL, kind


def test_emul_short_resampling_is_planned(monkeypatch):
    """Planner facts (host only): config 3's plan (Morlet, N = 1500, 1..100 Hz, fp32) resamples, its groups cover every
    frequency once with decimated lengths that divide N, the bound of every group is below the default tolerance; fp64 and
    `resample=False` (or NWCWT_SHORT3=0) keep the exact short-row kernel."""
    from ninwavelets_b200 import _backend as be
    fr = np.arange(1, 101.0)
    fam = orc.Family("morlet", sfreq=1000.0, sigma=7.0)
    aux = np.array([orc.peak_freq(fam, f) for f in fr])
    kw = dict(device=0, family=be.MORLET, interpolate=False, n=1500, sfreq=1000.0, freqs=fr, p0=fam.sigma,
              p1=fam.c * np.float_power(np.pi, -1 / 4), p2=fam.k, aux=aux)
    monkeypatch.setenv("NWCWT_SHORT3", "0")
    assert be.Plan(dtype=np.float32, **kw).info()["groups"] == []
    monkeypatch.delenv("NWCWT_SHORT3")
    g = be.Plan(dtype=np.float32, **kw).info()["groups"]
    assert g and sum(x["rows"] for x in g) == 100
    assert all(1500 % x["D"] == 0 and x["n1"] * x["D"] == 1500 for x in g)
    assert all(x["D"] == 1 or (0 < x["err"] <= 1e-6 and x["K"] % 2 == 0 and 4 <= x["K"] <= 12) for x in g)
    assert sum(x["rows"] * 1500 // x["D"] for x in g) < 0.4 * 100 * 1500
    assert be.Plan(dtype=np.float64, **kw).info()["groups"] == []
    assert be.Plan(dtype=np.float32, resample=False, **kw).info()["groups"] == []
