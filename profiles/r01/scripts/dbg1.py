import sys
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/oracle']
import numpy as np, torch
import cwt_oracle as orc
from ninwavelets_b200 import _backend as be
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
dt = np.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else np.float32
freqs = np.array([2., 10., 40.])
pl = be.Plan(device=0, dtype=dt, family=0, interpolate=False, n=N, sfreq=1000.0, freqs=freqs, p0=17.5, p1=3.)
print(pl.info())
rng = np.random.default_rng(0)
x = rng.standard_normal((1, N)).astype(dt)
xt = torch.from_numpy(x).cuda()
ref = orc.cwt(orc.Family("morse"), x[0].astype(np.float64), freqs)
for force in (1, 0):
    be.force_generic(force)
    z = pl.transform_device(xt, be.OUT_CWT)
    torch.cuda.synchronize()
    zc = z[0].cpu().numpy()
    err = np.abs(zc - ref).max(axis=1) / np.abs(ref).max(axis=1)
    print("force_generic", force, "max|z|", np.abs(zc).max(), "ref max", np.abs(ref).max(), "err", err)
    ws = pl._ws
    cdt = torch.complex128 if dt == np.float64 else torch.complex64
    X = ws[: N * (16 if dt == np.float64 else 8)].view(cdt).cpu().numpy()
    Xref = np.fft.fft(x[0].astype(np.float64))
    print("  X err", np.abs(X - Xref).max() / np.abs(Xref).max())
    i = pl.info()
    n1, n2 = i["n1"], i["n2"]
    xb = (pl.workspace_bytes(1) - 0)
    print("  ws bytes", xb, "launches", be.launch_count())
