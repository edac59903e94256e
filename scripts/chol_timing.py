"""Phase breakdown of the one-launch bordered Cholesky (needs a library built with -DOAK_CHOL_TIMING)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _cabi, _device
lib = _cabi.load()
for n, ident in ((1024, True), (1024, False)):
    rng = np.random.default_rng(0)
    Q = rng.standard_normal((n, n)); A = Q @ Q.T / n + np.eye(n)
    nb = n if ident else 1
    buf = np.zeros((n, 2 * n + 8)); buf[:, :n] = A
    if ident: buf[:, n:2 * n] = np.eye(n)
    else: buf[:, n] = 1.0
    d0 = torch.as_tensor(buf).cuda()
    for _ in range(3):
        d = d0.clone(); _device.chol(d, n, n + nb, border_identity=ident)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ds = [d0.clone() for _ in range(10)]
    e0.record()
    for d in ds: _device.chol(d, n, n + nb, border_identity=ident)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} identity border={ident}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per factorisation")
    if hasattr(lib, "oak_debug_chol_timing"):
        t = (C.c_longlong * (512 + 1024))()
        lib.oak_debug_chol_timing(t)
        t2 = np.array(t[512:]).reshape(256, 4)
        t = np.array(t[:512]).reshape(64, 8)[: n // 64]
        d = np.diff(t, axis=1) / 1.965e3  # us at 1965 MHz
        names = ["load", "potf2", "inverse", "trsm", "sync1", "update", "sync2"]
        print("  panel " + " ".join(f"{x:>8s}" for x in names))
        for p in range(len(t)):
            print(f"  {p:5d} " + " ".join(f"{x:8.2f}" for x in d[p]))
        print("  total " + " ".join(f"{x:8.1f}" for x in np.where(d > 0, d, 0).sum(0)))
        print("  panel 1, per CTA (us): phase 1 [potf2+trsm]   ", np.round(t2[:148:8, 0] / 1.965e3, 1))
        print("  panel 1, per CTA (us): phase 2 [update]       ", np.round(t2[:148:8, 1] / 1.965e3, 1))
