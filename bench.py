#!/usr/bin/env python
"""Benchmark of the DESMO fused train step (BASELINE.json metric: train iters/s + HBM GB/s of the fused residual+grad pass).

  python bench.py --gpus N --steps K --warmup W            # this framework, N GPUs of one node (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/torch_port.py) on the host cores

Workload (config.workload): "aneurysm-scale" -- n = 3 * 2^20 mesh points per GPU x m = 1000 snapshots, r = 4, polyorder = 2
(K = 27 library terms), fp32, synthetic pulsatile data generated on the device (12.6 GB per GPU, >> L2), POD init on the
device.  Weak scaling: every rank owns an equally sized slab of points; only the (K*m + 29)-float `red` buffer is all-reduced.
One step = build_w + fused residual/grad pass + partial reduction + (all-reduce) + regulariser/Adamax update.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("DESMO_KERNEL_EVENTS", "1")  # lets the library time its dominant kernel with CUDA events

WORKLOADS = {
    # name: (points per GPU, m, r, polyorder, nF)
    "aneurysm-scale": (3 * 2 ** 20, 1000, 4, 2, 0),
    "aneurysm-script": (27000, 1000, 4, 2, 0),
    "cylinder-script": (3961, 1001, 4, 3, 0),
    "cylinder-8modes": (3961, 1001, 8, 2, 0),   # BASELINE.json configs[0] ("8 modes"): K = 69, FFMA path
    "cylinder-fourier": (3961, 1001, 2, 2, 10),
    "channel-script": (16384, 1000, 4, 2, 0),
}
METRIC = "train_iters_per_s"  # x slabs: one unit = one train iteration over one GPU-slab (weak scaling: N slabs per step at N GPUs)
UNIT = "it/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 2 ms (nvidia-smi as a fallback, ~10 Hz)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index: int):
        self.index, self.samples, self._stop = index, [], threading.Event()  # samples: (sm_mhz, sm_max_mhz, {reasons})
        self.th = threading.Thread(target=self._run, daemon=True)
        self.source = "nvml"

    def _run_nvml(self):
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            mask = int(get_reasons(h))
            self.samples.append((int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(mx), {n for n, b in self.BITS if mask & b}))
            self._stop.wait(0.002)

    def _run_smi(self):
        self.source = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                c = [v.strip() for v in out.split(",")]
                if len(c) >= 6 and c[0].isdigit():
                    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                    self.samples.append((int(c[0]), int(c[1]), {n for n, v in zip(names, c[2:6]) if v.lower().startswith("active")}))
            except Exception:
                pass
            self._stop.wait(0.05)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        reasons = set().union(*[s[2] for s in self.samples]) if self.samples else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(s[1] for s in self.samples) if self.samples else None, "reasons": sorted(reasons),
                "samples": len(self.samples), "source": self.source}


def synth_on_device(torch, n, m, dev, seed, x_offset=0, n_global=None):
    """Pulsatile aneurysm-like data, generated on the device straight into the padded time-major layout:
    8 smooth spatial fields x 5 temporal harmonics + noise, temporal mean removed (CYL:136-149), scaled by 1/sqrt(m) (ANEU:143)."""
    ld = (n + 255) // 256 * 256
    n_global = n_global or n
    g = torch.Generator(device="cpu").manual_seed(seed)
    U = torch.zeros(m, ld, dtype=torch.float32, device=dev)
    x = (torch.arange(n, device=dev, dtype=torch.float64) + x_offset) / max(n_global - 1, 1)
    t = torch.arange(m, device=dev, dtype=torch.float64)
    for q in range(8):
        c = torch.randn(4, generator=g, dtype=torch.float64)
        ph = torch.rand(4, generator=g, dtype=torch.float64) * 3.141592653589793
        gq = sum(c[j] * torch.sin(3.141592653589793 * (j + 1) * (q + 1) * x + ph[j]) / (j + 1) for j in range(4))
        a = torch.randn(6, generator=g, dtype=torch.float64)
        psi = torch.rand(6, generator=g, dtype=torch.float64) * 6.283185307179586
        aq = a[0] * 0 + sum(a[h] / h * torch.cos(6.283185307179586 * h * t / m + psi[h]) for h in range(1, 6))
        U[:, :n].add_((aq[:, None] * gq[None, :]).to(torch.float32) / (q + 1))
    gen = torch.Generator(device=dev).manual_seed(seed + 17 + x_offset % 9973)
    chunk = 64
    for t0 in range(0, m, chunk):
        U[t0:t0 + chunk, :n].add_(0.02 * torch.randn(min(chunk, m - t0), n, device=dev, generator=gen))
    U[:, :n].sub_(U[:, :n].mean(dim=0, keepdim=True))
    U.mul_(1.0 / m ** 0.5)
    return U


def cpu_reference_leg(n_full, m, r, p, nF, steps, warmup, sample_points=None):
    """The reference's CPU implementation of the step (oracle/torch_port.py) on a bounded slab of the same workload."""
    import numpy as np
    import torch

    from oracle import desmo_oracle as orc
    from oracle.torch_port import time_steps

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ns = min(n_full, sample_points or 2 ** 14)
    X = orc.synthetic_snapshots("aneurysm", ns, m, seed=2)
    modes, _, _, _ = orc.pod_analysis(X, r)
    prm = orc.init_params(ns, m, p, r, nF=nF or None)
    sec, _ = time_steps(prm, modes, np.ascontiguousarray(X.T), steps, warmup)
    its_sample = 1.0 / sec
    # per-step cost is linear in the number of points (every op is O(n*m*K)); scale the slab rate to the full workload
    return {"value": its_sample * ns / n_full, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ns} of {n_full} points x {m} snapshots, {steps} steps after {warmup} warm-up, torch CPU fp32 "
                      f"{torch.get_num_threads()} threads, incl. the per-epoch fp64->fp32 re-collation (CYL:707-708); "
                      f"{its_sample:.3f} it/s on the slab, scaled by {ns}/{n_full}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="desmo_b200", choices=["desmo_b200", "reference"])
    ap.add_argument("--workload", default="aneurysm-scale", choices=sorted(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override points per GPU")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 fp32 FFMA, 2 tcgen05")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pod", action="store_true", help="profiling runs: random orthonormal-scale modes instead of the POD init")
    ap.add_argument("--cpu-sample", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    n, m, r, p, nF = WORKLOADS[args.workload]
    if args.points:
        n = args.points
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"workload": f"{args.workload}: {n} points/GPU x {m} snapshots, r={r}, polyorder={p}" + (f", nF={nF}" if nF else "") +
           ", fp32, point-sharded", "points_per_gpu": n, "snapshots": m, "r": r, "polyorder": p,
           "cache": "inputs larger than L2 (12.6 GB/GPU streamed per step)" if n * m * 4 > 4e8 else "L2 flushed between steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        # unit of `value` (both arms): slab-iterations per second, a slab being one GPU's share (n points x m snapshots); the host
        # cores process slabs at the same rate whatever N is, the GPUs process N of them per step (weak scaling)
        leg = cpu_reference_leg(n, m, r, p, nF, max(args.steps // 4, 3), 1, args.cpu_sample or None)
        line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000.0 / leg["value"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": leg,
                "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (desmo_b200 has no CPU fallback)")
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from desmo_b200 import DESMO, DesmoTrainer, _lib

    n_global = n * world
    import contextlib

    with contextlib.redirect_stdout(sys.stderr):  # the module prints the reference's banner (CYL:510); keep stdout = one JSON line
        model = DESMO(n, m, p, r, 10000, device=dev, n_global=n_global, path=args.path)
    e = model.engine
    e.U = synth_on_device(torch, n, m, dev, seed=2, x_offset=rank * n, n_global=n_global)
    if args.no_pod:
        gen = torch.Generator(device=dev).manual_seed(5)
        e.P[:, :n] = torch.randn(r, n, device=dev, generator=gen) / n ** 0.5
        sigma = torch.zeros(r)
        pod_info = None
    else:
        sigma = e.pod_from_snapshot()  # POD init on the device (method of snapshots)
        pod_info = dict(e.pod_timing)
        # Gram = 2 n m^2 useful flop; on the tensor pipe it is executed as 6 bf16 passes over 128-padded tiles (upper triangle)
        nt = (m + 127) // 128
        pod_info["gram_tflops_fp32_equiv"] = 2.0 * n * m * m / (pod_info["gram_ms"] * 1e-3) / 1e12
        pod_info["gram_tensor_tflops_bf16"] = 6 * 2.0 * n * (nt * (nt + 1) // 2) * 128 * 128 / (pod_info["gram_ms"] * 1e-3) / 1e12
        try:
            pod_info["gram_tensor_pipe_util"] = pod_info["gram_tensor_tflops_bf16"] / float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
    torch.cuda.synchronize()
    trainer = DesmoTrainer(model, sched_every=10 ** 9, use_cuda_graph=True)  # no host sync inside the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l2_flush = None if n * m * 4 > 4e8 else torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        trainer.step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        if l2_flush is None:
            ev0.record()
            for _ in range(args.steps):
                trainer.step()
            ev1.record()
            barrier()
            ms_total = ev0.elapsed_time(ev1)
        else:
            ms_total = 0.0
            for _ in range(args.steps):
                l2_flush.fill_(1)
                ev0.record()
                trainer.step()
                ev1.record()
                torch.cuda.synchronize()
                ms_total += ev0.elapsed_time(ev1)
            barrier()
        # dominant kernel alone: average launch duration of the fused residual+grad pass on its launching stream
        e.build_w(False)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = 0.0      # the fused call (dominant kernel + chain rule + partial reduction), torch events on the current stream
        kms_dom = 0.0  # the dominant kernel alone, CUDA events recorded inside the library on the launching stream
        dom = ctypes.c_float(0.0)
        for _ in range(args.steps):
            if l2_flush is not None:
                l2_flush.fill_(1)
            k0.record()
            e.fused_residual_grad()
            k1.record()
            torch.cuda.synchronize()
            kms += k0.elapsed_time(k1)
            _lib.check(e.lib.desmo_last_fused_kernel_ms(ctypes.byref(dom)), "desmo_last_fused_kernel_ms")
            kms_dom += dom.value
        kms /= args.steps
        kms_dom /= args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    losses = [float(v) for v in e.losses.tolist()]

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D of the batch + step + D2H of the losses per step) ----
    e2e = None
    if not args.no_e2e:
        try:
            lib = _lib.load()
            import psutil

            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            if 4.0 * n * m * local_world > 0.6 * psutil.virtual_memory().available:  # never drive the box out of memory
                raise RuntimeError(f"host memory too small for {local_world} pinned batches of {4.0 * n * m / 1e9:.1f} GB")
            host = torch.empty(m, n, dtype=torch.float32, pin_memory=True)
            host.copy_(e.U[:, :n])
            ss = ctypes.c_void_p()
            _lib.check(lib.desmo_session_create(n, m, r, p, nF, args.path, ctypes.byref(ss)), "session_create")
            import numpy as np

            pod = np.ascontiguousarray(e.P[:, :n].t().double().cpu().numpy())
            K = e.K
            phi0 = np.ones((r, n), np.float32); gates0 = np.ones(K, np.float32); rows0 = np.ones((K, m), np.float32)
            om0 = np.full(3 * r, 1e4, np.float32); lrs = np.array([1e-2, 1e-3, 1e-2, 1e3, 1e-2], np.float32)
            cp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
            _lib.check(lib.desmo_session_set_pod_host(ss, cp(pod)), "set_pod")
            _lib.check(lib.desmo_session_set_params_host(ss, cp(phi0), cp(gates0), cp(rows0), None, cp(om0)), "set_params")
            _lib.check(lib.desmo_session_set_hyper(ss, cp(lrs), 1e-3, 1e-4), "set_hyper")
            lo = np.zeros(4, np.float32)
            e_steps = max(2, min(args.steps, 5))
            _lib.check(lib.desmo_session_step_host(ss, ctypes.c_void_p(host.data_ptr()), cp(lo)), "step_host")  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                _lib.check(lib.desmo_session_step_host(ss, ctypes.c_void_p(host.data_ptr()), cp(lo)), "step_host")
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / e_steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            lib.desmo_session_destroy(ss)
            e2e = {"value": world / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(n * m * 4), "d2h_bytes_per_step": 16,
                   "steps": e_steps, "call": "desmo_session_step_host (pinned host batch -> device, fused step, losses -> host)"}
            del host
        except Exception as ex:  # keep the device-resident number even if the host leg cannot run (e.g. pinned alloc)
            e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    if rank == 0:
        peak, peak_src = load_peaks()
        alg_bytes = 4.0 * n * m + 12.0 * n * r + 8.0 * e.K * m  # U once; phi, P in, dphi out; W in, E out
        achieved = alg_bytes / (kms_dom * 1e-3) / 1e9
        line = {"metric": METRIC, "value": world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg,
                "global_iters_per_s": 1.0 / (ms_step * 1e-3),
                "snapshot_gbs": world * 4.0 * n * m / (ms_step * 1e-3) / 1e9,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": "fused_tc_kernel" if e.uses_tensor_cores() else "fused_fp32_kernel", "kernel_ms": kms_dom,
                             "fused_call_ms": kms, "peak_source": peak_src,
                             "algorithmic_bytes": alg_bytes},
                "clocks": clk.summary(), "e2e": e2e,
                "gpu_launches": (5 if e.uses_tensor_cores() else 4) * args.steps + (1 if world > 1 else 0) * args.steps,
                "losses_last_step": losses, "pod_sigma": [float(v) for v in sigma.tolist()], "pod_init": pod_info,
                "path": "tcgen05 (bf16x3 split)" if e.uses_tensor_cores() else "fp32 ffma"}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as fh:
                tr = json.load(fh)
            if tr.get("points_per_gpu") == n and tr.get("path") == line["path"]:
                line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        if not args.no_cpu and world == 1:
            try:
                line["cpu_baseline"] = cpu_reference_leg(n, m, r, p, nF, 5, 1, args.cpu_sample or None)
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down of a process group whose collectives were captured in a CUDA graph can block; everything is measured and
        # printed, so leave without the destructor dance (exit code 0 for torchrun)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
