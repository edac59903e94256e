#!/usr/bin/env python
"""bench.py -- OAK Gram entries/sec (FP64) and SGPR ELBO evals/sec on B200, vs roofline.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (under torchrun for N > 1) prints ONE
JSON line on rank 0.  A *step* is one pass of the hot path over the workload:

  headline   config B of BASELINE.json: K(X, X), N=65536, D=16, max_interaction_depth=4, FP64,
             prepare + fused Gram kernel, inputs resident in HBM; the FULL symmetric matrix is written on
             every N: one GPU evaluates lower-triangle tiles + mirrored stores; N>1 ranks own folded row
             strips and write the lower trapezoid of each strip and its mirror image (same product, no
             collective) -> strong scaling of a fixed N.  value = unique entries N(N+1)/2 / time.
             config_B_sweep: the same at N = 8k, 16k, 32k, 64k.
  e2e        the same metric through the host-buffer C-ABI call (oak_gram_host_lower_f64): pinned NumPy
             X in, H2D, prepare, row-blocked lower trapezoids, D2H of every unique entry into pinned host memory.
  sgpr_elbo  config C: SGPR ELBO evals/s (N=1M, D=20, M=1024, depth 3), N axis sharded over ranks, L = chol(Kuu)
             first, route chosen on the device, one all-reduce of M^2+M+2 doubles; statistics / factor front /
             tail / all-reduce timed separately; the training step (ELBO + gradient through the backward tiles).
  extra      one evaluation of configs A, D, E and an A-shaped Gram; the widening rows (SVGP, flow objective).
  roofline   FP64-pipe bound: algorithmic flop (306 slots x 2 per unique entry) / Gram-kernel time
             (CUDA events on the launch stream) / measured DFMA peak of the same run.
  cpu_baseline  the reference op sequence (oracle/cpu_baseline.py, torch-CPU FP64, all host cores)
             on a bounded sample, rank 0 at N=1 only.

``--impl reference`` times that CPU restatement alone (the reference itself needs TensorFlow /
gpflow, which are not installable in this image -- see DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SLOTS_B = 306.0  # BASELINE.md work model: FP64 issue slots per Gram entry at D=16, P=4
SLOTS_C = 865.5  # per Kuf entry at D=20, P=3, M=1024 (tile 352 + SYRK 512.5 + Kuf y 1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=65536, help="Gram size (config B: 8k..64k)")
    ap.add_argument("--n-e2e", type=int, default=0, help="Gram size of the host-buffer leg (0 = auto)")
    ap.add_argument("--n-cpu", type=int, default=8192, help="sample size of the CPU baseline (BASELINE.md section 4)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the N = 8k / 16k / 32k points of config B")
    ap.add_argument("--elbo-n", type=int, default=1_000_000)
    ap.add_argument("--elbo-m", type=int, default=1024)
    ap.add_argument("--no-elbo", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--algo", type=int, default=1, help="0 Newton-Girard (the reference's scheme), 1 direct recurrence (package default)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled from a thread
    (sub-millisecond period, so even the 20 ms timed region of an 8-GPU run gets samples);
    falls back to `nvidia-smi -lms` when pynvml is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.stop_flag, self.samples = None, None, False, []
        self.max_mhz, self.reason_bits = None, 0
        self.smi_id = self._physical_id(index)
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = (pynvml.nvmlDeviceGetHandleByUUID(self.smi_id) if self.smi_id.startswith("GPU-")
                           else pynvml.nvmlDeviceGetHandleByIndex(int(self.smi_id)))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_id(local):
        """NVML / nvidia-smi identifier of CUDA device `local`: CUDA_VISIBLE_DEVICES may renumber the
        devices, NVML never does -- go through the UUID (or the visible-devices list)."""
        try:
            import torch

            u = str(torch.cuda.get_device_properties(local).uuid)
            if len(u) >= 32:
                return u if u.startswith("GPU-") else "GPU-" + u
        except Exception:
            pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").strip()
        if vis:
            toks = [t.strip() for t in vis.split(",") if t.strip()]
            if local < len(toks):
                return toks[local]
        return str(local)

    def _poll(self):
        nv = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    self.reason_bits |= int(get_reasons(self.handle))
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.smi_id, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(self.samples)
            if not sm:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            # NVML clocks-event-reason bits (nvml.h): 0x8 hw_slowdown, 0x40 hw_thermal_slowdown,
            # 0x20 sw_thermal_slowdown, 0x4 sw_power_cap
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            reasons = sorted(v for k, v in names.items() if self.reason_bits & k)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


# ---------------------------------------------------------------------------------------------
def _host_threads():
    """All host cores, whatever torchrun put into OMP_NUM_THREADS (it exports 1 to every rank)."""
    import torch

    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


def cpu_gram_sample(n_cpu, reps, warm=0, keep=False):
    """Reference op sequence on the host cores: K(X, X) of config B at N=n_cpu. Returns
    (unique entries/s, seconds per evaluation, threads, K of the last evaluation or None)."""
    import torch

    from oak_b200.workloads import config_B
    from oracle import cpu_baseline

    cpu_baseline.tune_allocator()
    threads = _host_threads()
    cfg = config_B(n_cpu)
    X = torch.as_tensor(cfg["X"])
    cpu_baseline.gram(cfg, X[:512])  # warm the thread pool / allocator
    for _ in range(warm):
        cpu_baseline.gram(cfg, X)
    ts, K = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        K = cpu_baseline.gram(cfg, X)
        ts.append(time.perf_counter() - t0)
        if not keep:
            K = None
    ts.sort()
    t = ts[len(ts) // 2]
    return n_cpu * (n_cpu + 1) / 2 / t, t, threads, K


def cpu_elbo_sample(n_slice, m, reps=1):
    """gpflow SGPR.elbo with the unfused OAK kernel on a slice of config C (N=20000, M=1024)."""
    import torch

    from oak_b200.workloads import config_C
    from oracle import cpu_baseline

    _host_threads()
    cfg = config_C(n_slice, 20, m, 3)
    X, Y, Z = (torch.as_tensor(cfg[k]) for k in ("X", "y", "Z"))
    ts, val = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        val = cpu_baseline.sgpr_elbo(cfg, X, Y, Z, cfg["noise"])
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return cfg, val, ts[len(ts) // 2]


def run_reference(args):
    """The reference's CPU implementation of the path (restated op sequence: TensorFlow / gpflow are not
    installable here) on ALL host cores; under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n_cpu
    val, t_eval, threads, _ = cpu_gram_sample(n, max(args.steps, 1), warm=max(args.warmup, 1))
    _, _, t_elbo = cpu_elbo_sample(20000, args.elbo_m, reps=1)
    line = {
        "impl": "reference",
        "metric": "OAK Gram entries/sec (FP64)", "value": val, "unit": "unique entries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_eval * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config B: OAK Gram K(X,X) D=16 depth=4 FP64; reference CPU op sequence on a "
                               f"bounded sample N={n} (BASELINE.md section 4: the unfused path needs "
                               f"~(D+P+2) N^2 x 8 B, N=16k+ does not fit)"},
        "cpu_baseline": {"value": val, "unit": "unique entries/s", "cores": threads, "kind": "port",
                         "sample": f"K(X,X) N={n}, D=16, depth=4, median of {args.steps} after {args.warmup} warm-ups; "
                                   "TensorFlow/gpflow not installable -> oracle/cpu_baseline.py (torch-CPU FP64)",
                         "elbo_evals_per_s_scaled_to_n1e6": 1.0 / (t_elbo * 1e6 / 20000),
                         "elbo_sample": f"config C slice N=20000, M={args.elbo_m}: {t_elbo:.2f} s per evaluation, "
                                        "scaled linearly in N"},
        "e2e": {"value": val, "unit": "unique entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs NVML reports as local to its GPU, so that pinned host buffers are first-touched
    on the GPU's NUMA node (8 ranks writing 1 GB each over PCIe into one node's memory is what held the 8-GPU
    host-buffer leg at 11 GB/s per rank in round 1)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByUUID(ClockSampler._physical_id(local)) \
            if ClockSampler._physical_id(local).startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from oak_b200 import _cabi, _device, parallel
    from oak_b200.models import SGPR
    from oak_b200.workloads import build_kernel, config_B, config_C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _cabi.require_device()
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def by_rank(v):
        """The per-rank values of a timing (diagnostic: which rank is the straggler of a max-over-ranks number)."""
        if world == 1:
            return [v]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = v
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [round(float(x), 4) for x in t.tolist()]

    # ---- measured FP64 peak (roofline denominator) -------------------------------------------
    peak_slots = _device.measure_fp64_peak(1.0)
    peak_tflops = 2 * peak_slots / 1e12
    ncu_file = {}
    try:
        ncu_file = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    hbm_peak = 6521.4
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    # ---- config B: symmetric Gram, the same product on every N -----------------------------------
    # N = 1: lower-triangle tiles + mirrored stores, the full matrix on one GPU.  N > 1: folded row strips; every
    # rank evaluates the lower trapezoid of its strips AND writes their mirror image (oak_gram_lower_mirror_f64),
    # so the ranks hold the full matrix between them: the same product, no collective.
    def gram_leg(n, steps, sample_clocks=False):
        cfg = config_B(n)
        kern = build_kernel(cfg)
        kern.esp_algorithm = args.algo
        spec = kern._make_spec()
        Xd = _device.to_device(cfg["X"])
        if world == 1:
            strips = [(0, n)]
            outs = [torch.empty((n, n), dtype=torch.float64, device="cuda")]
        else:
            strips = [s for s in parallel.balanced_symmetric_rows(n, world)[rank] if s[1] > s[0]]
            outs = [(torch.empty((e - b, e), dtype=torch.float64, device="cuda"),
                     torch.empty((e, e - b), dtype=torch.float64, device="cuda")) for b, e in strips]
        kev = []

        def step(record=False):
            px = _device.Points(spec, Xd)  # per-point prologue (O(N D))
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if world == 1:
                _device.gram(spec, px, out=outs[0])
            else:
                for (b, e), (o, ot) in zip(strips, outs):
                    _device.gram_lower_mirror(spec, px, b, e, out=o, out_t=ot)
            if record:
                e1.record()
                kev.append((e0, e1))

        for _ in range(warm):
            step()
        sampler = ClockSampler(local) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        launches0 = _cabi.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            step(record=True)
        ev1.record()
        torch.cuda.synchronize()
        launches = _cabi.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        barrier()
        ms_step = max_over_ranks(ev0.elapsed_time(ev1) / steps)
        kern_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in kev) / len(kev))
        unique = n * (n + 1) / 2
        my_unique = max_over_ranks(sum((e - b) * (b + e + 1) / 2 for b, e in strips))
        achieved = my_unique * SLOTS_B * 2 / (kern_ms * 1e-3) / 1e12
        out_bytes = max_over_ranks(float(n) * n * 8.0 if world == 1 else 16.0 * sum((e - b) * (b + e) / 2 for b, e in strips))
        spec.close()
        del outs
        torch.cuda.empty_cache()
        return {"n": n, "ms_per_step": ms_step, "kernel_ms": kern_ms, "value": unique / (ms_step * 1e-3),
                "roofline_frac": achieved / peak_tflops, "achieved_tflops": achieved, "out_bytes": out_bytes,
                "launches": int(launches), "clocks": clocks}

    sweep = []
    if not args.no_sweep:
        for n_s in (8192, 16384, 32768):
            if n_s < args.n:
                r = gram_leg(n_s, 3)
                sweep.append({k: r[k] for k in ("n", "ms_per_step", "value", "roofline_frac")})
    head = gram_leg(args.n, args.steps, sample_clocks=True)
    sweep.append({k: head[k] for k in ("n", "ms_per_step", "value", "roofline_frac")})
    n = args.n
    value, ms_step, kern_ms, clocks, launches = head["value"], head["ms_per_step"], head["kernel_ms"], head["clocks"], head["launches"]
    roofline = {
        "bound": "fp64",
        "bound_note": "FP64 (DFMA) pipe: neither of the template's roofs binds -- 8 B of HBM traffic per entry "
                      "against >= 213 FP64 instructions, no tensor-core work in this kernel; the hbm sub-object "
                      "carries the HBM line",
        "achieved": head["achieved_tflops"], "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": head["roofline_frac"], "traffic": ncu_file.get("gram_kernel_dram_bytes_per_launch"),
        "kernel": "oak::gram_kernel<4,4,4,%s>" % ("NG" if args.algo == 0 else "direct"), "kernel_ms": kern_ms,
        "fp64_pipe_active_pct_ncu": ncu_file.get("gram_kernel_fp64_pipe_active_pct" if args.algo == 1
                                                 else "gram_kernel_fp64_pipe_active_pct_newton_girard"),
        "fp64_pipe_note": "frac is the ALGORITHMIC roofline: the work model charges 306 slots per entry (the reference's "
                          "power sums + Newton-Girard); the kernel issues ~%d FP64 instructions per entry (%s), so frac "
                          "can pass 1.0; the pipe utilisation behind it is ncu's sm__pipe_fp64_cycles_active of the "
                          "committed capture (profiles/)" % ((242, "Newton-Girard tiles") if args.algo == 0
                                                              else (213, "direct-recurrence tiles, the package default")),
        "peak_source": "measured in this run: oak_measure_fp64_peak (register-resident DFMA chains, burst); "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "work_model": f"{SLOTS_B:.0f} FP64 issue slots (x2 flop) per unique entry (BASELINE.md section 3)",
        "hbm": {"achieved": head["out_bytes"] / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": head["out_bytes"] / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                "note": "8 B per written entry (slowest rank); the kernel is FP64-pipe bound, not HBM bound"},
    }

    # ---- e2e: the symmetric Gram through HOST buffers (oak_gram_host_lower_f64) ---------------------
    e2e = None
    if not args.no_e2e:
        ncpu_bound = bind_to_gpu_numa_node(local)
        n_e = args.n_e2e
        if n_e == 0:
            # pinned result buffers on this host: the lower trapezoids of all ranks' strips, ~N^2/2 * 8 B in total
            # (a full N x N matrix, 34.4 GB at 65536, on a single-GPU run: the reference-shaped output)
            need_gb = 8.0 * args.n * args.n / 1e9 * (1.0 if world == 1 else 0.6) + 12.0
            n_e = args.n if mem_available_gb() > need_gb else 32768
        cfg_e = config_B(n_e)
        kern_e = build_kernel(cfg_e)
        kern_e.esp_algorithm = args.algo
        spec_e = kern_e._make_spec()
        strips_e = [(0, n_e)] if world == 1 else [s for s in parallel.balanced_symmetric_rows(n_e, world)[rank] if s[1] > s[0]]
        Xh = torch.as_tensor(np.ascontiguousarray(cfg_e["X"])).pin_memory()
        block = 2048
        Kh = [torch.empty((e - b, e), dtype=torch.float64).pin_memory() for b, e in strips_e]
        for k_ in Kh:
            k_.zero_()  # first touch on this rank's NUMA node
        wb = max(lib.oak_gram_host_lower_work_bytes(spec_e.handle, n_e, 16, e, block) for b, e in strips_e)
        work = torch.empty(wb // 8 + 1, dtype=torch.float64, device="cuda")

        def e2e_step():
            for (b, e), k_ in zip(strips_e, Kh):
                rc = lib.oak_gram_host_lower_f64(spec_e.handle, Xh.data_ptr(), n_e, 16, b, e, k_.data_ptr(), e, block, 0,
                                                 work.data_ptr(), C.c_void_p(_device.stream_ptr()))
                assert rc == 0, _cabi.last_error()

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            e2e_step()
        b_.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / e2e_steps
        ms_e = max_over_ranks(max(a.elapsed_time(b_) / e2e_steps, wall * 1e3))
        barrier()
        d2h = 0
        for b, e in strips_e:
            for r0 in range(b, e, block):
                r1 = min(r0 + block, e)
                d2h += (r1 - r0) * r1 * 8
        e2e = {
            "value": n_e * (n_e + 1) / 2 / (ms_e * 1e-3), "unit": "unique entries/s",
            "h2d_bytes_per_step": int(Xh.numel() * 8 * len(strips_e)), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": ms_e, "d2h_gb_per_s_this_rank": d2h / (ms_e * 1e-3) / 1e9,
            "cpus_bound_to_gpu_numa_node": ncpu_bound,
            "workload": f"oak_gram_host_lower_f64: K(X,X) N={n_e} D=16 depth=4, pinned host X in; the lower trapezoid of "
                        f"every {block}-row block (all unique entries) copied back to pinned host memory, overlapped "
                        "with compute; the upper triangle is the mirror image (mirror=1 fills it on the host, not timed)"
                        + (f"; folded row strips over {world} ranks" if world > 1 else ""),
        }
        del Kh, work, Xh
        spec_e.close()

    # ---- SGPR ELBO evals/s (config C): the second half of BASELINE.json's metric ---------------------
    elbo = None
    if not args.no_elbo:
        cfg_c = config_C(args.elbo_n, 20, args.elbo_m, 3)
        b, e = parallel.partition_rows(args.elbo_n, world)[rank]
        kc = build_kernel(cfg_c)
        kc.esp_algorithm = args.algo
        model = SGPR((cfg_c["X"][b:e], cfg_c["y"][b:e]), kernel=kc, inducing_variable=cfg_c["Z"], chunk=262144,
                     distributed=(world > 1))
        model.likelihood.variance.assign(cfg_c["noise"])
        model._device_data()
        for _ in range(2):
            model.elbo()
        barrier()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_e = max(3, min(args.steps, 5))
        a.record()
        vals = [model.elbo() for _ in range(k_e)]
        b2.record()
        torch.cuda.synchronize()
        ms_elbo_rank = by_rank(a.elapsed_time(b2) / k_e)
        ms_elbo = max(ms_elbo_rank)
        # phases on their own: statistics (tiles + DMMA contraction), the factor-first front, the tail
        sp = kc._make_spec()
        Xs, Ys = model._device_data()
        pz = _device.Points(sp, model._Z_device())
        pxs = _device.Points(sp, Xs)
        fac = _device.sgpr_factor(sp, pz, 1e-6)
        _device.sgpr_stats2(sp, pz, pxs, Ys, fac)
        torch.cuda.synchronize()

        def timed_dev(fn):
            a.record()
            for _ in range(k_e):
                r = fn()
            b2.record()
            torch.cuda.synchronize()
            return max_over_ranks(a.elapsed_time(b2) / k_e), r

        ms_stats, st = timed_dev(lambda: _device.sgpr_stats2(sp, pz, pxs, Ys, fac))
        ms_factor, _ = timed_dev(lambda: _device.sgpr_factor(sp, pz, 1e-6, buf=fac.buf))
        ms_finish, _ = timed_dev(lambda: _device.sgpr_finish2(fac, st, args.elbo_n, cfg_c["noise"], want_alpha=False))
        # what the model calls: factorisation on a side stream next to the first chunk's Kuf tiles
        ms_fused, _ = timed_dev(lambda: _device.sgpr_factor_stats(sp, pz, pxs, Ys, 1e-6, buf=fac.buf))
        ms_allreduce = None
        if world > 1:
            ms_allreduce, _ = timed_dev(lambda: parallel.allreduce_sum_(st))
        fac_w = _device.sgpr_factor(sp, pz, 1e-6, route=_device.ROUTE_WHITENED)
        _device.sgpr_stats2(sp, pz, pxs, Ys, fac_w)
        ms_stats_w, _ = timed_dev(lambda: _device.sgpr_stats2(sp, pz, pxs, Ys, fac_w))
        sp.close()
        del fac_w
        kuf_entries = float(args.elbo_m) * (e - b)
        elbo = {
            "metric": "SGPR ELBO evals/sec", "value": 1e3 / ms_elbo, "unit": "evals/s", "ms_per_eval": ms_elbo,
            "ms_stats_phase": ms_stats, "ms_tail_and_collective": ms_elbo - ms_stats, "elbo": vals[-1],
            "ms_factor_front": ms_factor, "ms_finish_tail": ms_finish, "ms_allreduce": ms_allreduce,
            "ms_factor_and_stats_overlapped": ms_fused, "ms_per_eval_by_rank": ms_elbo_rank,
            "overlap_note": "oak_sgpr_factor_stats_f64: chol(Kuu) + condition estimate run on a side stream on 4-8 CTAs "
                            "while the first chunk's Kuf tiles leave them as many of the 148 SMs; ms_stats_phase and "
                            "ms_factor_front are the two pieces timed serially on their own",
            "route": model.last_route, "cond_estimate_kuu": model.last_cond_estimate,
            "ms_stats_phase_whitened_route": ms_stats_w,
            "workload": f"config C: N={args.elbo_n}, D=20, M={args.elbo_m}, depth 3; N axis sharded over {world} "
                        f"rank(s); all-reduce of {args.elbo_m ** 2 + args.elbo_m + 2} doubles; L = chol(Kuu) first, route "
                        "chosen on the device (0 = Phi statistics, 1 = gpflow's whitened order)",
            "roofline_stats_phase": {
                "bound": "fp64", "unit": "TFLOP/s", "peak": peak_tflops,
                "achieved": kuf_entries * SLOTS_C * 2 / (ms_stats * 1e-3) / 1e12,
                "frac": kuf_entries * SLOTS_C * 2 / (ms_stats * 1e-3) / 1e12 / peak_tflops,
                "work_model": f"{SLOTS_C} slots per Kuf entry (tile 352 + DSYRK 512.5 + Kuf y 1)"},
        }

        # one training step of the same model: ELBO + gradient w.r.t. lengthscales, order variances and
        # noise (backward tiles) -- what each BFGS iteration of oak_model.fit costs
        try:
            from oak_b200.training import freeze_unsupported, sgpr_elbo_and_grad

            model.inducing_variable.Z.trainable = False  # zfixed=True, the reference's default
            freeze_unsupported(model)
            for _ in range(2):  # the first calls size the kept-Kuf store and the allocator's pools
                sgpr_elbo_and_grad(model)
            barrier()
            a.record()
            for _ in range(2):
                g_out = sgpr_elbo_and_grad(model)
            b2.record()
            torch.cuda.synchronize()
            ms_grad = max_over_ranks(a.elapsed_time(b2) / 2)
            elbo["training_step"] = {"metric": "SGPR ELBO + gradient evals/sec", "value": 1e3 / ms_grad,
                                     "ms_per_eval": ms_grad, "grad_norm_lengthscales": float(np.abs(g_out[1]).max())}
            # the same step with trainable inducing points (zfixed=False): adds d/dZ through Kuf and Kuu
            model.inducing_variable.Z.trainable = True
            sgpr_elbo_and_grad(model)
            barrier()
            a.record()
            for _ in range(2):
                sgpr_elbo_and_grad(model)
            b2.record()
            torch.cuda.synchronize()
            ms_gz = max_over_ranks(a.elapsed_time(b2) / 2)
            elbo["training_step"]["with_inducing_point_gradients"] = {
                "ms_per_eval": ms_gz, "grad_norm_Z": float(np.abs(model._inducing_grad).max())}
        except Exception as exc:  # the headline must not depend on the widening row
            elbo["training_step"] = {"error": repr(exc)}
        del model
        torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations (A, D, E) and an A-shaped Gram, N=1 only ---------------
    others = None
    if world == 1 and not args.no_elbo:
        try:
            from oak_b200.models import GPR
            from oak_b200.utils import compute_sobol_oak
            from oak_b200.workloads import config_A, config_D, config_E

            def timed(fn, reps=3):
                fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    r = fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / reps * 1e3, r

            others = {}
            ca = config_A()
            ga = GPR((ca["X"], ca["y"]), kernel=build_kernel(ca))
            ga.likelihood.variance.assign(ca["noise"])
            ms, lml = timed(ga.log_marginal_likelihood)
            others["A_gpr_lml_n1030_d8_depth8"] = {"ms": ms, "lml": lml}
            ms, _ = timed(lambda: compute_sobol_oak(ga, 1.0, 0.0), reps=1)
            others["A_sobol_255_components"] = {"ms": ms}
            # config A's kernel (D=8, full depth 8) at a size where the tile kernel is the whole cost
            na = 32768
            cal = config_A(na)
            sa = build_kernel(cal)._make_spec()
            pxa = _device.Points(sa, _device.to_device(cal["X"]))
            outa = torch.empty((na, na), dtype=torch.float64, device="cuda")
            ms, _ = timed(lambda: _device.gram(sa, pxa, out=outa))
            others["A_shaped_gram_n32768_d8_depth8"] = {
                "ms": ms, "unique_entries_per_s": na * (na + 1) / 2 / (ms * 1e-3),
                "roofline_frac_244_slots": na * (na + 1) / 2 * 244 / (ms * 1e-3) / peak_slots}
            sa.close()
            del outa
            cd = config_D()
            kd_ = build_kernel(cd)
            sd = kd_._make_spec()
            pxd = _device.Points(sd, _device.to_device(cd["X"]))
            nd_ = cd["X"].shape[0]
            outd = torch.empty((nd_, nd_), dtype=torch.float64, device="cuda")
            ms, _ = timed(lambda: _device.gram(sd, pxd, out=outd))
            others["D_gram_mixed_n50000_d12_depth2"] = {
                "ms": ms, "unique_entries_per_s": nd_ * (nd_ + 1) / 2 / (ms * 1e-3),
                "roofline_frac_135_slots": nd_ * (nd_ + 1) / 2 * 135 / (ms * 1e-3) / peak_slots}
            sd.close()
            del outd
            ce = config_E()
            me = SGPR((ce["X"], ce["y"]), kernel=build_kernel(ce), inducing_variable=ce["Z"])
            me.likelihood.variance.assign(ce["noise"])
            ms, _ = timed(lambda: compute_sobol_oak(me, 1.0, 0.0), reps=1)
            others["E_sobol_d50_n200000_m512_1275_components"] = {"ms": ms}
        except Exception as exc:
            others = {"error": repr(exc)}

    # ---- widening rows: one timed evaluation each, N=1 only -------------------------
    widening = None
    if world == 1 and not args.no_elbo:
        try:
            from oak_b200._gpflow_shim import Bernoulli, inv_logit
            from oak_b200.models import SVGP
            from oak_b200.training import svgp_elbo_and_grad

            widening = {}
            rng_w = np.random.default_rng(5)
            nw, dw, mw = 100_000, 10, 200
            Xw = rng_w.standard_normal((nw, dw))
            yw = (rng_w.random((nw, 1)) < 1.0 / (1.0 + np.exp(-2.0 * np.sin(Xw[:, :1]) - Xw[:, 1:2]))).astype(np.float64)
            cw = {"dims": [{"type": "rbf", "lengthscale": 1.0, "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)}] * dw,
                  "depth": 4, "variances": [1.0] * 5, "share_var": True}
            sv = SVGP(kernel=build_kernel(cw), likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Xw[:mw].copy(),
                      whiten=True, q_diag=True)
            sv.inducing_variable.Z.trainable = False
            dataw = (_device.to_device(Xw), _device.to_device(yw))
            ms, _ = timed(lambda: svgp_elbo_and_grad(sv, dataw), reps=3)
            widening["svgp_bernoulli_elbo_and_gradient_n100000_d10_m200_depth4"] = {"ms": ms}
            ms, _ = timed(lambda: sv.elbo(dataw), reps=3)
            widening["svgp_bernoulli_elbo_n100000_d10_m200_depth4"] = {"ms": ms}
            # normalising-flow objective: one L-BFGS-B iteration = one pass over a resident column
            xcol = _device.to_device(np.exp(0.5 * rng_w.standard_normal(4_000_000)) + 2.0, ndim=1)
            ms, _ = timed(lambda: _device.flow_objective(xcol, 0.5, True, [0.1, -0.5, 0.05, 0.0]), reps=5)
            widening["flow_objective_pass_n4000000"] = {"ms": ms, "points_per_s": 4e6 / (ms * 1e-3)}
            del xcol, dataw
            # k-means inducing points at config C's scale: scikit-learn's algorithm on the device (kmeans.KMeans)
            from oak_b200.kmeans import KMeans

            Xk = _device.to_device(rng_w.standard_normal((1_000_000, 20)) + 2.0 * rng_w.standard_normal((64, 20))[
                rng_w.integers(0, 64, 1_000_000)])
            tk = {}
            for iters in (1, 11):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                KMeans(n_clusters=1024, random_state=0, max_iter=iters).fit(Xk)
                torch.cuda.synchronize()
                tk[iters] = time.perf_counter() - t0
            widening["kmeans_n1000000_d20_k1024"] = {"kmeanspp_seeding_plus_one_iteration_s": tk[1],
                                                     "ms_per_lloyd_iteration": (tk[11] - tk[1]) * 100.0}
            del Xk
        except Exception as exc:
            widening = {"error": repr(exc)}

    # ---- CPU baseline beside it (rank 0, N=1 only): timing AND parity on the same inputs ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, t_eval, threads, Kc = cpu_gram_sample(args.n_cpu, 3, warm=1, keep=True)
        cfg_p = config_B(args.n_cpu)
        Kg = build_kernel(cfg_p).K(_device.to_device(cfg_p["X"])).cpu()
        gram_err = float(((Kg - Kc).abs() / Kc.abs().clamp_min(1e-12)).max())
        del Kg, Kc
        cfg_s, elbo_cpu, t_elbo = cpu_elbo_sample(20000, args.elbo_m)
        ms_ = SGPR((cfg_s["X"], cfg_s["y"]), kernel=build_kernel(cfg_s), inducing_variable=cfg_s["Z"])
        ms_.likelihood.variance.assign(cfg_s["noise"])
        elbo_gpu = ms_.elbo()
        # per-configuration parity on samples of the named shapes (element-wise for Gram entries)
        per_cfg = {}
        try:
            from oak_b200.workloads import config_A, config_D, config_E
            from oracle import cpu_baseline as cb

            rng_p = np.random.default_rng(123)
            for name, cfg_x in (("A", config_A()), ("D", config_D()), ("E", config_E())):
                rows = rng_p.choice(cfg_x["X"].shape[0], 384, replace=False)
                cols = rng_p.choice(cfg_x["X"].shape[0], 320, replace=False)
                Xa, Xb = cfg_x["X"][rows], cfg_x["X"][cols]
                Kg = np.asarray(build_kernel(cfg_x).K(Xa, Xb))
                Kr = cb.gram(cfg_x, torch.as_tensor(Xa), torch.as_tensor(Xb)).numpy()
                dg = np.asarray(build_kernel(cfg_x).K_diag(Xa))
                dr = cb.gram_diag(cfg_x, torch.as_tensor(Xa)).numpy().reshape(-1)
                per_cfg[name] = {"gram_max_elementwise_rel_err": float(np.max(np.abs(Kg - Kr) / np.maximum(np.abs(Kr), 1e-12))),
                                 "gram_max_err_over_max": float(np.max(np.abs(Kg - Kr)) / np.max(np.abs(Kr))),
                                 "k_diag_max_rel_err": float(np.max(np.abs(dg.reshape(-1) - dr) / np.abs(dr))),
                                 "sample": "384 x 320 random rows / columns of the configuration's inputs"}
        except Exception as exc:
            per_cfg = {"error": repr(exc)}
        cpu = {"value": v, "unit": "unique entries/s", "cores": threads, "kind": "port",
               "parity_other_configs": per_cfg,
               "sample": f"K(X,X) N={args.n_cpu}, D=16, depth=4 (reference op sequence, oracle/cpu_baseline.py, "
                         f"torch-CPU FP64), median of 3: {t_eval:.2f} s per evaluation",
               "parity_max_elementwise_rel_err_gram": gram_err,
               "elbo": {"evals_per_s_scaled_to_n1e6": 1.0 / (t_elbo * args.elbo_n / 20000),
                        "sample": f"config C slice N=20000, M={args.elbo_m}: {t_elbo:.2f} s per evaluation, scaled "
                                  f"linearly to N={args.elbo_n}",
                        "parity_rel_err_elbo": abs(elbo_gpu - elbo_cpu) / abs(elbo_cpu)}}

    if rank == 0:
        line = {
            "metric": "OAK Gram entries/sec (FP64)", "value": value, "unit": "unique entries/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"config B: OAK Gram K(X,X), N={n}, D=16, max_interaction_depth=4, Gaussian measure, "
                            "FP64; step = prepare + fused Gram kernel, X resident in HBM; the full symmetric N x N matrix "
                            "is written: " + ("lower-triangle tiles + mirrored stores on one GPU" if world == 1 else
                                              f"folded row strips over {world} ranks, each rank its lower trapezoids and "
                                              "their mirror images (same product as N=1), no collective"),
                "l2": f"output of {head['out_bytes'] / 1e9:.1f} GB per step and rank >> 126 MB L2 (streaming stores); inputs 16 MB",
                "esp": "newton_girard" if args.algo == 0 else "direct_recurrence",
            },
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu,
            "elbo_evals_per_s": None if elbo is None else elbo["value"],
            "sgpr_elbo": elbo,
            "config_B_sweep": sweep,
            "extra": {"other_configs": others, "widening": widening, "fp64_peak_slots_per_s": peak_slots},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
