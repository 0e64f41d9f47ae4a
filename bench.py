#!/usr/bin/env python
"""bench.py - CWT power output points/s (T x F x ch) on B200, BASELINE.json's metric.

Workload (config.workload): BASELINE.json configs[1] - 64-channel EEG, 600 000 samples per
channel at 1 kHz, Morse(17.5, 3) power at 1..100 Hz, fp32 (the configuration the metric is
quoted on; 15.4 GB of fp32 output per GPU, fits one B200).  One "step" = one pass of the hot
path over the 64 channels of this rank (forward FFTs, spectrum generation, inverse FFTs, |z|^2).
Multi-GPU: every rank owns its own 64 channels (weak scaling, no collective on the data path).

Prints ONE JSON line on rank 0 (see the task contract): value = device-resident throughput,
e2e = the same through the C ABI's host-buffer call (H2D and D2H inside the timed region),
roofline = algorithmic HBM bytes / measured duration of the path's kernels vs MEASURED_PEAKS.json,
cpu_baseline = the numpy oracle (port of the reference) timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]
                  [--workload cfg2|cfg3|cfg4] [--dtype f32|f64]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "cwt_power_output_points_per_sec"
UNIT = "points/s"

WORKLOADS = {
    # name: (family, kwargs, n_signals, n, freqs, baseline)
    "cfg2": dict(desc="cfg2: 64 ch x 600000 samples @1 kHz, Morse(17.5,3) power, freqs 1-100",
                 kind="morse", S=64, N=600000, freqs=np.arange(1, 101.0), baseline=None),
    "cfg3": dict(desc="cfg3: 306 ch x 200 epochs x 1500 samples @1 kHz, Morlet(7) power + zscore[0,0.2s], freqs 1-100",
                 kind="morlet", S=306 * 200, N=1500, freqs=np.arange(1, 101.0), baseline=("zscore", 0.0, 0.2)),
    # cfg3, epoch-mean variant (mneutils.py:53-55): per-epoch power, then the mean over the 200 epochs of each channel
    "cfg3_mean": dict(desc="cfg3 epoch-mean: 306 ch x 200 epochs x 1500 samples @1 kHz, Morlet(7) power, mean over epochs fused into the transform kernel, freqs 1-100",
                      kind="morlet", S=306 * 200, N=1500, freqs=np.arange(1, 101.0), baseline=None, epochs=200),
    "cfg4": dict(desc="cfg4: 32 ch x 2^20 samples, Morse power, freqs 1-128",
                 kind="morse", S=32, N=1 << 20, freqs=np.arange(1, 129.0), baseline=None),
}
# cfg5: long-signal sweep, 256 ch x 2^k samples x 256 freqs.  The full output (256 x 256 x N reals) exceeds device memory
# from 2^20 on, so one step processes as many of the 256 channels as fit a 70 GB output buffer and the buffer is
# recycled step after step ("streamed"); points/s is per processed point either way.
for _k in (16, 18, 20, 21, 22, 24, 26):
    _S = max(1, min(256, int(70e9 // (256 * (1 << _k) * 4))))
    WORKLOADS["cfg5_%d" % _k] = dict(
        desc="cfg5: 2^%d samples x 256 freqs, Morse power, %d of 256 ch per step (output buffer recycled)" % (_k, _S),
        kind="morse", S=_S, S_total=256, N=1 << _k, freqs=np.arange(1, 257.0), baseline=None)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.15:
                continue
            parts = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass


# ---------------------------------------------------------------------------------------
# CPU legs (oracle = port of the reference; the ONLY place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------
def _oracle_family(wl):
    import cwt_oracle as orc
    if wl["kind"] == "morse":
        return orc.Family("morse", sfreq=1000.0, b=17.5, r=3.0)
    return orc.Family("morlet", sfreq=1000.0, sigma=7.0)


def _oracle_one(args):
    """One worker's share: `per` signals through the reference algorithm.  The spectrum bank is built once per worker
    (outside the timed calls of a warm sample) exactly like the reference's default `reuse=True` keeps
    `self.fft_wavelets` after the first call (base.py:394-395); `cold` rebuilds it for every signal (reuse=False)."""
    import cwt_oracle as orc
    kind, n, freqs, seed, baseline, per, cold = args
    fam = orc.Family(kind, sfreq=1000.0) if kind == "morse" else orc.Family("morlet", sfreq=1000.0, sigma=7.0)
    rng = np.random.default_rng(seed)
    xs = [rng.standard_normal(n) for _ in range(per)]
    bank = None if cold else orc.make_fft_wavelets(fam, freqs, n / 1000.0)
    t0 = time.perf_counter()
    acc = 0.0
    for x in xs:
        p = orc.power(fam, x, freqs) if cold else orc.power(fam, x, None, bank=bank)
        if baseline is not None:
            p = orc.baseline_rows(p, 1000.0, baseline[1], baseline[2], baseline[0])
        acc += float(p[0, 0])
    return time.perf_counter() - t0, acc


def cpu_sample(wl, workers, signals_per_worker, freqs, cold=False):
    """Time `workers * signals_per_worker` signals of the workload on host cores; returns (points/s, seconds).
    Warm samples time the transforms only (max over workers), as the reference's steady state does."""
    from multiprocessing import get_context
    jobs = [(wl["kind"], wl["N"], freqs, 1000 + i, wl["baseline"], signals_per_worker, cold) for i in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        res = [_oracle_one(jobs[0])]
    else:
        with get_context("fork").Pool(workers) as pool:
            res = pool.map(_oracle_one, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    dt = wall if cold else max(r[0] for r in res)
    pts = workers * signals_per_worker * len(freqs) * wl["N"]
    return pts / dt, dt


def sample_shape(wl):
    """Bounded CPU sample: ~1e8 output points per worker-step (a few seconds each)."""
    n, F = wl["N"], len(wl["freqs"])
    if n * F > 2.5e7:       # long rows: one signal, a quarter of the frequencies
        nfr = max(2, min(25, int(1.5e7 // n)))           # keeps the (F, N) complex128 block of a cold call below ~1 GB
        fr = wl["freqs"][:: max(1, F // nfr)][:nfr]
        return 1, fr
    per = max(1, int(2e7 // (n * F)))
    return per, wl["freqs"]


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = max(1, min(os.cpu_count() or 1, 16))
    per, fr = sample_shape(wl)
    per = max(per, 2) if wl["N"] > 100000 else per
    for _ in range(args.warmup):
        cpu_sample(wl, workers, per, fr)
    pts, total = 0, 0.0
    for _ in range(args.steps):
        v, dt = cpu_sample(wl, workers, per, fr)
        pts += workers * per * len(fr) * wl["N"]
        total += dt
    value = pts / total
    cold_v, _ = cpu_sample(wl, workers, 1, fr, cold=True)
    sample = ("%d signals x %d of %d freqs x %d samples per step (%s), %d worker processes, warm calls: spectrum bank "
              "built once per worker outside the timed region (the reference's reuse=True steady state, base.py:394-395)" % (
                  workers * per, len(fr), len(wl["freqs"]), wl["N"], wl["desc"].split(":")[0], workers))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "cold_call_value": cold_v},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def _executed_lane_ops(info, S, N, F):
    """FP32 / FP64 lane operations one step executes, from the plan: the packed two-pass engine spends about 77 lane-ops
    per point of a transform (3.5 per radix-2 level and point, spectrum, twiddles, |z|^2), a resampled
    row additionally 2 K + 2 per OUTPUT sample for its K-tap interpolation (nw_resample.cuh)."""
    groups = info.get("groups") or []
    if not groups:
        return float(S) * (F + 1) * N * 77.0
    ops = float(S) * N * 77.0                       # forward transforms
    for g in groups:
        M = N // g["D"]
        ops += float(S) * g["rows"] * (M * 77.0 + (N * (2.0 * g["K"] + 2.0) if g["D"] > 1 else 0.0))
    return ops


def run_graft(args, wl):
    import torch
    import torch.distributed as dist
    import ninwavelets_b200 as nw
    from ninwavelets_b200 import _backend as be
    from ninwavelets_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    f32 = args.dtype == "f32"
    tdt = torch.float32 if f32 else torch.float64
    N, freqs = wl["N"], wl["freqs"]
    F = len(freqs)
    real_b = 4 if f32 else 8
    strong = args.scaling == "strong"
    # weak scaling: every rank owns its own wl["S"] signals.  strong scaling: the workload's wl["S_total"] (cfg5: 256)
    # or wl["S"] signals are ONE job, sharded over the ranks by the product's driver (ninwavelets_b200/sharding.py).
    S_job = wl.get("S_total", wl["S"]) if strong else wl["S"]
    if strong:
        lo_sig, hi_sig = sharding.shard_range(S_job, rank, world)
        S = hi_sig - lo_sig
    else:
        S = S_job
    fit = max(1, int(70e9 // (F * N * real_b)))          # signals whose output fits a 70 GB buffer (recycled per chunk)
    S_buf = min(S, fit) if S > 0 else 1

    # synthetic signals, generated on the host and made resident before timing (value),
    # kept on the host (pinned) for the end-to-end leg
    rng = np.random.default_rng(2 + rank)
    host_x = torch.empty((max(S, 1), N), dtype=tdt).pin_memory()
    hx = host_x.numpy()
    t = np.arange(N) / 1000.0
    for sidx in range(max(S, 1)):
        hx[sidx] = rng.standard_normal(N)
        if wl["kind"] == "morse":
            for f0 in (10.0, 40.0, 60.0):
                hx[sidx] += np.sin(2 * np.pi * f0 * t + rng.uniform(0, 2 * np.pi))
    x = host_x.to(dev)

    ctor = nw.Morse if wl["kind"] == "morse" else nw.Morlet

    def make_plan(dtype_name):
        obj = ctor(1000, cuda=True, dtype=dtype_name, device=local)
        obj.make_fft_wavelets(freqs, N / 1000.0)
        return obj, obj._plan

    obj, plan = make_plan("float32" if f32 else "float64")
    info = plan.info()
    bl = (0, 0, 0)
    if wl["baseline"] is not None:
        from ninwavelets_b200.base import _window
        lo, hi = _window(N, 1000.0, wl["baseline"][1], wl["baseline"][2])
        bl = (be.BASELINE_MODES[wl["baseline"][0]], lo, hi)
    n_ep = wl.get("epochs", 0)
    if n_ep:
        S_buf = min(S_buf, n_ep)
    out = torch.empty((S_buf, F, N), dtype=tdt, device=dev)

    def local_transform(xs, fr):
        """The single-GPU call handed to the product's multi-GPU driver: device-resident signals -> [s, F, N] in `out`."""
        res = None
        for c0 in range(0, xs.shape[0], S_buf):
            c1 = min(xs.shape[0], c0 + S_buf)
            res = plan.transform_device(xs[c0:c1], be.OUT_POWER, *bl, out=out[: c1 - c0])
        return res

    def step():
        if strong:
            # the product driver decides this rank's shard (signals, or frequencies when there are fewer signals than
            # ranks) and runs the local call on it; gather=False: results stay sharded on the ranks (no collective)
            class _Dev:   # signals are already resident: hand the driver the local block of the job
                shape = (S_job, N)
                def __getitem__(self, sl):
                    a, b2, _ = sl.indices(S_job)
                    return x[a - lo_sig: b2 - lo_sig]
            sharding.distributed_transform(local_transform, _Dev(), freqs, gather=False, rank=rank, world=world)
            return None
        if n_ep:   # signals are channel-major (c * E + e): mean power over epochs, reduced inside the transform kernel
            return plan.transform_epochs_device(x.view(S // n_ep, n_ep, N), 0)
        local_transform(x, freqs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = be.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = be.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    points_per_step = (S_job if strong else S * world) * F * N
    value = points_per_step * args.steps / (ms * 1e-3)
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    step_ms = ms / args.steps

    # ---- roofline of the path's kernels --------------------------------------------------------
    # `achieved` = algorithmic bytes of one step / device time of one step's kernels, measured with CUDA
    # events on the launching stream around the timed region above (the library forks onto its own
    # auxiliary streams and joins back, so the events bracket all of its kernels).  A separate profiling
    # pass with events around every launch gives each kernel class's share; with several row groups in flight
    # on different streams those per-launch times overlap, so a second plan restricted to ONE stream gives the
    # true per-class kernel times.
    be.profile_enable(True)
    step()
    torch.cuda.synchronize()
    prof = be.profile_read()
    be.profile_enable(False)
    alg_bytes_step = S * N * (F + 1) * real_b              # write one real per point + read each sample once
    if n_ep:                                               # fused epoch mean: one (F, N) result per channel is written
        alg_bytes_step = S * N * real_b + (S // n_ep) * F * N * real_b
    kern = {k: v for k, v in prof.items() if v["launches"]}
    tot_ms = sum(v["ms"] for v in kern.values())
    dominant = max(kern, key=lambda k: kern[k]["ms"])
    peak, peak_src = peaks()
    achieved = alg_bytes_step / (step_ms * 1e-3) / 1e9
    n_launch = sum(v["launches"] for v in kern.values())
    single = None
    if info["path"].startswith("long") and not args.tuning:
        os.environ["NWCWT_STREAMS"] = "1"
        try:
            _, plan1 = make_plan("float32" if f32 else "float64")
            for _ in range(2):
                plan1.transform_device(x[:S_buf], be.OUT_POWER, *bl, out=out)
            torch.cuda.synchronize()
            be.profile_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan1.transform_device(x[:S_buf], be.OUT_POWER, *bl, out=out)
            e1.record()
            torch.cuda.synchronize()
            p1 = be.profile_read()
            be.profile_enable(False)
            single = {"ms_per_step": e0.elapsed_time(e1) * (S / float(S_buf)),
                      "classes_ms": {k: round(v["ms"] * (S / float(S_buf)), 4) for k, v in p1.items() if v["launches"]},
                      "note": "the same step on ONE stream: per-class sums are kernel durations (no overlap)"}
            plan1.close()
        finally:
            del os.environ["NWCWT_STREAMS"]
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r02", "traffic_%s.json" % args.workload)
    if os.path.isfile(tpath) and f32 and not strong:
        try:
            tj = json.load(open(tpath))
            traffic = float(tj["dram_bytes_per_step"]) * (S / float(tj["signals"]))
            traffic_note = tj["note"]
        except Exception:
            traffic = None
    # FLOP side: the FP32 (FP64) pipe peak is measured in this run by the library's FFMA2 / DFMA chain
    # kernel; `nominal` counts the reference algorithm's 5 N log2 N per transform (what a full-length FFT-based CWT would
    # execute: one forward and F inverse transforms per signal), `executed` what this implementation actually runs
    # (decimated transforms + interpolation taps, _executed_lane_ops).
    try:
        pipe_peak = be.fma_peak(local, f32)              # lane-ops / s
    except Exception:
        pipe_peak = None
    nominal_flops = 5.0 * N * np.log2(N) * (F + 1) * S
    executed_ops = _executed_lane_ops(info, S, N, F)
    flop = {"nominal_tflops": nominal_flops / (step_ms * 1e-3) / 1e12,
            "nominal_convention": "5 N log2 N per length-N transform, (F + 1) transforms per signal",
            "executed_lane_ops_per_step": executed_ops,
            "executed_tlaneops": executed_ops / (step_ms * 1e-3) / 1e12,
            "peak_tlaneops": None if pipe_peak is None else pipe_peak / 1e12,
            "peak_source": "in-run %s chain microbenchmark of the library (nwcwt_fma_peak)" % ("FFMA2" if f32 else "DFMA"),
            "frac": None if not pipe_peak else executed_ops / (step_ms * 1e-3) / pipe_peak}
    t_hbm = alg_bytes_step / (peak * 1e9)
    t_pipe = executed_ops / pipe_peak if pipe_peak else None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_note": traffic_note,
        "peak_source": peak_src,
        "kernel": "all kernels of one step (%s); dominant class %s" % (
            "one fused kernel" if info["path"].startswith("short") else
            "%d launches: per launch group of rows an inverse passA + passB pair (+ the interpolation kernel for resampled "
            "rows), forward transforms once per signal" % n_launch, dominant),
        "algorithmic_bytes_per_step": alg_bytes_step,
        "kernel_ms_per_step": step_ms,
        "avg_launch_ms": step_ms / max(n_launch, 1),
        "flop": flop,
        "binding": None if t_pipe is None else {
            "hbm_floor_ms": 1e3 * t_hbm, "pipe_floor_ms": 1e3 * t_pipe,
            "binds": "fp%d pipe" % (32 if f32 else 64) if t_pipe > t_hbm else "hbm",
            "note": "floors of one step at 100 % of the measured HBM bandwidth / of the measured pipe rate"},
        "classes": {k: {"ms_sum_of_launches": round(v["ms"], 4), "launches": v["launches"], "share": v["ms"] / tot_ms,
                        "avg_launch_ms": v["ms"] / v["launches"]} for k, v in kern.items()},
        "single_stream": single,
    }

    # ---- parity spot check of the timed output against the oracle (not timed) -------------------
    def spot_check(out_t, hx_row, f32_):
        import cwt_oracle as orc
        fam = _oracle_family(wl)
        sub = freqs[:: max(1, F // 4)][:4] if N > 100000 else freqs
        xs = hx_row.astype(np.float32).astype(np.float64) if f32_ else hx_row.astype(np.float64)
        ref = orc.power(fam, xs, sub)
        if wl["baseline"] is not None:
            ref = orc.baseline_rows(ref, 1000.0, wl["baseline"][1], wl["baseline"][2], wl["baseline"][0])
        idx = [int(np.nonzero(freqs == f)[0][0]) for f in sub]
        got = out_t[0, idx].double().cpu().numpy()
        if f32_:
            num = np.sqrt(((got - ref) ** 2).sum(axis=1))
            den = np.sqrt((ref ** 2).sum(axis=1))
            return {"rows_checked": len(idx), "max_row_rel_l2": float((num / den).max()), "metric": "per-row relative L2, no floor"}
        num = np.abs(got - ref).max(axis=1)
        den = np.abs(ref).max(axis=1)
        return {"rows_checked": len(idx), "max_row_rel_peak": float((num / den).max()), "metric": "per-row max|diff| / max|ref|, no floor"}

    parity = None
    if rank == 0 and S > 0:
        plan.transform_device(x[:1], be.OUT_POWER, *bl, out=out[:1])
        torch.cuda.synchronize()
        parity = spot_check(out, hx[0], f32)
        if n_ep:   # the fused epoch mean of channel 0 against a float64 mean of its materialised per-epoch rows
            m = step()[0].double()
            rows0 = plan.transform_device(x[:n_ep], be.OUT_POWER, 0, 0, 0)
            mref = rows0.double().mean(dim=0)
            parity["epoch_mean_max_rel"] = float(((m - mref).abs().max() / mref.abs().max()).item())
            del rows0

    # ---- fp64 sibling (the reference's own arithmetic) on the same workload ------------------------
    fp64 = None
    try_fp64 = f32 and not args.tuning and not strong and world == 1 and not n_ep and S_buf * F * N * 8 <= 75e9
    def fp64_leg():
        _, plan64 = make_plan("float64")
        x64 = x.double()
        out64 = torch.empty((S_buf, F, N), dtype=torch.float64, device=dev)

        def step64():
            for c0 in range(0, S, S_buf):
                c1 = min(S, c0 + S_buf)
                plan64.transform_device(x64[c0:c1], be.OUT_POWER, *bl, out=out64[: c1 - c0])
        for _ in range(3):
            step64()
        torch.cuda.synchronize()
        k64 = max(3, min(args.steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k64):
            step64()
        e1.record()
        torch.cuda.synchronize()
        ms64 = e0.elapsed_time(e1) / k64
        plan64.transform_device(x64[:1], be.OUT_POWER, *bl, out=out64[:1])
        torch.cuda.synchronize()
        res = {"value": S * F * N / (ms64 * 1e-3), "unit": UNIT, "ms_per_step": ms64, "steps": k64,
               "roofline_frac_hbm": S * N * (F + 1) * 8 / (ms64 * 1e-3) / 1e9 / peak,
               "parity_spot_check": spot_check(out64, hx[0], False),
               "groups": [(g["D"], g["K"]) for g in plan64.info().get("groups", [])]}
        del out64, x64
        plan64.close()
        return res

    if try_fp64:
        del out
        torch.cuda.empty_cache()
        try:   # a failure of a secondary leg is reported in the line, it does not take the headline measurement with it
            fp64 = fp64_leg()
        except Exception as e:
            fp64 = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
        torch.cuda.empty_cache()
        out = torch.empty((S_buf, F, N), dtype=tdt, device=dev)

    # ---- end to end through the C ABI's host-buffer entry point ---------------------------------
    if args.tuning or strong:
        e2e = None
    elif F * N * real_b > (8 << 30):
        e2e = {"value": None, "unit": UNIT, "note": "one signal's output exceeds 8 GiB; host leg not run for this sweep size"}
    else:
      try:
        e2e_S = min(S, max(1, int((16 << 30) // (F * N * real_b))))       # the whole job when its output fits 16 GiB of pinned memory
        host_out = torch.empty((e2e_S, F, N), dtype=tdt, pin_memory=True)   # allocated pinned (no pageable copy first)
        hin = hx[:e2e_S]
        hout = host_out.numpy()
        plan.transform_host(hin, be.OUT_POWER, *bl, out=hout)        # warm-up (allocates staging)
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.transform_host(hin, be.OUT_POWER, *bl, out=hout)
        torch.cuda.synchronize()
        e_dt = time.perf_counter() - t0
        if world > 1:
            tms = torch.tensor([e_dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            e_dt = float(tms.item())
        e2e = {"value": e2e_S * F * N * world * e2e_steps / e_dt, "unit": UNIT,
               "h2d_bytes_per_step": int(e2e_S * N * real_b), "d2h_bytes_per_step": int(e2e_S * F * N * real_b),
               "steps": e2e_steps, "signals_per_step": e2e_S,
               "api": "nwcwt_transform_host (pinned host buffers; chunked H2D -> kernels -> D2H on two streams)",
               "note": "bound by the device-to-host copy of 4 B per output point over PCIe"}
      except Exception as e:
        if world > 1:   # the other ranks are waiting in the all-reduce: fail loudly rather than hang
            raise
        e2e = {"value": None, "unit": UNIT, "error": repr(e)[:300]}

    cpu = None
    if rank == 0 and not args.tuning:
        # bounded CPU sample of the same workload on this box's host cores (single process = as shipped), warm calls
        per, fr = sample_shape(wl)
        reps, dt_sum, pts_sum = 0, 0.0, 0.0
        while dt_sum < 10.0 and reps < 64:          # ~10-15 s of single-core work
            v, dt = cpu_sample(wl, 1, per, fr)
            reps, dt_sum, pts_sum = reps + 1, dt_sum + dt, pts_sum + v * dt
        cold_v, _ = cpu_sample(wl, 1, 1, fr, cold=True)
        cpu = {"value": pts_sum / dt_sum, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d x (%d signal(s) x %d of %d freqs x %d samples), warm calls (spectrum bank cached like the "
                         "reference's reuse=True), %.1f s" % (reps, per, len(fr), F, N, dt_sum),
               "cold_call_value": cold_v, "host_cpus": os.cpu_count()}
    if sampler:
        sampler.stop()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": wl["desc"], "signals_per_gpu": S, "signals_total": S_job if strong else S * world,
                       "n": N, "n_freqs": F, "path": info["path"],
                       "sharding": ("ninwavelets_b200.sharding.distributed_transform: the job's %d signals are split over "
                                    "the ranks, no collective on the data path" % S_job) if strong else
                                   "every rank owns its own signals (no collective on the data path)",
                       "split": [info["n1"], info["n2"]], "radices": info["radices"], "batch": info["batch"],
                       "threads": info["threads"], "rows_per_launch": info["rows_per_launch"],
                       "resampled_groups": [{"D": g["D"], "K": g["K"], "rows": g["rows"], "n1": g["n1"], "n2": g["n2"],
                                             "err_bound": g["err"]} for g in info.get("groups", [])],
                       "l2": "inputs+outputs per step (%.1f GB) far exceed the 126 MB L2" % (
                           (S * N * (F + 1) * real_b) / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "fp64": fp64, "gpu_launches": int(launches),
            "clocks": clocks, "parity_spot_check": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns its own signals; strong: ONE job (cfg2: 64 channels, cfg5: 256) sharded over "
                         "the ranks by ninwavelets_b200.sharding.distributed_transform")
    ap.add_argument("--tuning", action="store_true",
                    help="kernel tuning runs: skip the host-buffer leg and the CPU baseline (not a valid bench line)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_graft(args, wl)


if __name__ == "__main__":
    main()
