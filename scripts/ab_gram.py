"""A/B timing of the fused Gram kernel (development aid, not the bench).  Variants are chosen by
environment (OAK_B200_LIB, OAK_GRAM_VARIANT, OAK_GRAM_NOFAST) -> one process per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oak_b200 import _device
from oak_b200.workloads import config_B, config_C, build_kernel

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

import subprocess, threading
class Clk:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._rd, daemon=True).start()
    def _rd(self):
        for l in self.p.stdout:
            try:
                a, b = l.split(","); self.rows.append((float(a), float(b)))
            except Exception: pass
    def stop(self):
        self.p.terminate()
        r = sorted(self.rows)
        if not r: return "clk n/a"
        return f"clk med {r[len(r)//2][0]:.0f} min {r[0][0]:.0f} MHz, power max {max(x[1] for x in r):.0f} W ({len(r)} samples)"
clk = Clk()
tag = " ".join(f"{k}={os.environ[k].split('/')[-1]}" for k in ("OAK_B200_LIB", "OAK_GRAM_VARIANT", "OAK_GRAM_NOFAST", "OAK_GRAM_NOTMA") if k in os.environ) or "default"
peak = _device.measure_fp64_peak(0.5)
n = int(os.environ.get("AB_N", 32768))
out_line = [f"[{tag}] peak {2*peak/1e12:.2f} TF"]
for algo in (0, 1):
    cfg = config_B(n)
    k = build_kernel(cfg); k.esp_algorithm = algo
    spec = k._make_spec()
    Xd = _device.to_device(cfg["X"])
    px = _device.Points(spec, Xd)
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    t = timeit(lambda: _device.gram(spec, px, out=out))
    uniq = n * (n + 1) / 2
    out_line.append(f"B sym algo{algo} n={n}: {t:.3f} ms frac {uniq*306/(t*1e-3)/peak:.3f}")
    if algo == 0:
        px2 = _device.Points(spec, Xd)
        t = timeit(lambda: _device.gram(spec, px, px2, out=out))
        out_line.append(f"B cross: {t:.3f} ms frac {n*n*306/(t*1e-3)/peak:.3f}")
    spec.close()
cfg = config_C(131072)
k = build_kernel(cfg); spec = k._make_spec()
Xd, Zd = _device.to_device(cfg["X"]), _device.to_device(cfg["Z"])
pz, px = _device.Points(spec, Zd), _device.Points(spec, Xd)
out = torch.empty((1024, 131072), dtype=torch.float64, device="cuda")
t = timeit(lambda: _device.gram(spec, pz, px, out=out))
out_line.append(f"C Kuf 1024x131072: {t:.3f} ms frac(352) {1024*131072*352/(t*1e-3)/peak:.3f}")
from oak_b200.workloads import config_D, config_A
for algo in (0, 1):
    cfg = config_D(32768)
    k = build_kernel(cfg); k.esp_algorithm = algo; spec = k._make_spec()
    Xd = _device.to_device(cfg["X"])
    px = _device.Points(spec, Xd)
    out = torch.empty((32768, 32768), dtype=torch.float64, device="cuda")
    t = timeit(lambda: _device.gram(spec, px, out=out))
    out_line.append(f"D sym algo{algo} n=32768 (8 cont + 4 discrete, P=2): {t:.3f} ms frac(135) {32768*32769/2*135/(t*1e-3)/peak:.3f}")
    spec.close()
    cfg = config_A(8192)
    k = build_kernel(cfg); k.esp_algorithm = algo; spec = k._make_spec()
    Xd = _device.to_device(cfg["X"])
    px = _device.Points(spec, Xd)
    out = torch.empty((8192, 8192), dtype=torch.float64, device="cuda")
    t = timeit(lambda: _device.gram(spec, px, out=out))
    out_line.append(f"A-shaped sym algo{algo} n=8192 (D=8, P=8): {t:.3f} ms frac(244) {8192*8193/2*244/(t*1e-3)/peak:.3f}")
    spec.close()
out_line.append(clk.stop())
print(" | ".join(out_line), flush=True)
