"""Small driver for ncu captures: one plan, a couple of transforms of a reduced workload.
usage: python profiles/prof_run.py {cfg2|cfg3|cfg4} [f32|f64] [n_signals] [n_freqs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ninwavelets_b200 as nw
from ninwavelets_b200 import _backend as be

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dt = sys.argv[2] if len(sys.argv) > 2 else "f32"
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
F = int(sys.argv[4]) if len(sys.argv) > 4 else 20
N = {"cfg2": 600000, "cfg3": 1500, "cfg4": 1 << 20}[wl]
freqs = np.arange(1, 101.0)[:: max(1, 100 // F)][:F]
ctor = nw.Morlet if wl == "cfg3" else nw.Morse
obj = ctor(1000, cuda=True, dtype="float32" if dt == "f32" else "float64")
obj.make_fft_wavelets(freqs, N / 1000.0)
plan = obj._plan
print(plan.info())
x = torch.randn((S, N), device="cuda", dtype=torch.float32 if dt == "f32" else torch.float64)
bl = (5, 0, 200) if wl == "cfg3" else (0, 0, 0)
for _ in range(3):
    out = plan.transform_device(x, be.OUT_POWER, *bl)
torch.cuda.synchronize()
print("ok", float(out[0, 0, :10].sum()))
