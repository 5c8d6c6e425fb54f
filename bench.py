#!/usr/bin/env python
"""bench.py — headline benchmark of the WeatherConverter hot path on B200 (contract: see the task README).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1]

Workload (N = 1 and every N, weak scaling): BASELINE.json configs[1] — UNet DDPM sampling at 128x256, batch 16
per GPU, 1000-step linear schedule, reference UNet (config.yaml, im_size 128, 110.6 M params), random-init
weights, synthetic noise.  A "step" is ONE reverse-diffusion step over the batch: Unet.forward(x_t, t) followed by
the fused posterior update x_{t-1} = mean + sigma_t z.  Per-step cost does not depend on t, so
    images/s = (N * batch) / (T_schedule * seconds_per_step).
`value` times the steps with inputs resident in HBM; `e2e` times the same steps through the public Python API with
the per-step noise coming from pinned HOST memory (the reference draws z on the CPU every step,
linear_noise_scheduler.py:110) and x_{t-1} read back to the host, copies inside the timed region.
`--impl reference` times the reference algorithm's CPU implementation (the oracle port; the reference itself is a
Python package that is not shipped to the GPU box) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "guided translation images/sec (256x512, T steps) at 1/2/4/8 B200 vs host CPU"
UNIT = "images/s"

WORKLOADS = {
    # name: (im_size, batch/GPU, H, W, schedule steps, description)
    "c2": (128, 16, 128, 256, 1000, "C2: UNet DDPM sampling 128x256, batch 16/GPU, 1000-step schedule (BASELINE.json configs[1])"),
    "c1": (64, 4, 64, 64, 50, "C1: UNet DDPM sampling 64x64, batch 4, 50 steps (BASELINE.json configs[0])"),
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU via NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_model(im_size, dev):
    import torch
    from weatherconverter_b200.diffusion_model.models.unet_base import Unet
    from weatherconverter_b200.diffusion_model.config.models import ModelConfig
    torch.manual_seed(3455)   # config.yaml:32
    cfg = ModelConfig(im_size=im_size)
    model = Unet(cfg).to(dev).eval()
    return model, cfg


def cpu_reference_rate(im_size, H, W, T, steps, warmup, batch=1):
    """Oracle port (plain PyTorch fp32, all host threads) on a bounded sample: `batch` image(s), `steps` timed
    reverse steps.  Returns (images/s extrapolated to the T-step schedule, seconds/step, cores)."""
    import torch
    from oracle.scheduler import OracleScheduler
    from oracle.unet import DEFAULT_MODEL_CONFIG, unet_forward, unet_param_spec
    from oracle.weights import synth_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = im_size
    sd = synth_state_dict(unet_param_spec(cfg), 3455)
    sched = OracleScheduler(T, 1e-4, 0.02)
    g = torch.Generator().manual_seed(1234)
    xt = torch.randn(batch, 3, H, W, generator=g)
    times = []
    with torch.no_grad():
        for k in range(warmup + steps):
            i = T - 1 - k
            t0 = time.perf_counter()
            eps = unet_forward(sd, cfg, xt, torch.tensor([i]))
            z = torch.randn(xt.shape, generator=g)
            mean, sz, _ = sched.sample_prev_timestep(xt, eps, i, z=z)
            xt = mean + sz
            dt = time.perf_counter() - t0
            if k >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return batch / (T * sec), sec, cores


def run_reference(args):
    """Reference arm: rank 0 only; CPU implementation of the same step on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    im_size, B, H, W, T, desc = WORKLOADS[args.workload]
    value, sec, cores = cpu_reference_rate(im_size, H, W, T, args.steps, args.warmup, batch=1)
    sample = f"1 image of the batch, {args.steps} timed reverse steps after {args.warmup} warm-up, extrapolated linearly to T={T}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "batch_per_gpu": B, "height": H, "width": W, "schedule_steps": T,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from weatherconverter_b200 import _lib, ops
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)
    K = args.steps
    im_size, B, H, Wd, T, desc = WORKLOADS[args.workload]
    peaks = read_peaks()

    model, cfg = build_model(im_size, dev)
    sched = LinearNoiseScheduler(T, 1e-4, 0.02)
    # per-image RNG keyed by the GLOBAL image index -> results independent of the number of GPUs
    xs = []
    for b in range(B):
        g = torch.Generator().manual_seed(1234 + rank * B + b)
        xs.append(torch.randn(3, H, Wd, generator=g))
    xt0 = torch.stack(xs).to(dev)
    gz = torch.Generator(device=dev).manual_seed(99 + rank)
    eps = torch.empty_like(xt0)
    nsteps_total = W_ + K
    t_devs = [torch.tensor([T - 1 - (k % T)], device=dev) for k in range(nsteps_total)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: z pre-generated in HBM ----------------
    zs = [torch.randn(xt0.shape, device=dev, generator=gz) for _ in range(min(nsteps_total, 8))]
    xt = xt0.clone()
    for k in range(W_):
        model(xt, t_devs[k], out=eps)
        xt = sched.step(xt, eps, T - 1 - (k % T), z=zs[k % len(zs)])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W_, W_ + K):
        model(xt, t_devs[k], out=eps)
        xt = sched.step(xt, eps, T - 1 - (k % T), z=zs[k % len(zs)])
    e1.record()
    barrier()
    clocks = sampler.finish()
    launches = ops.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(xt).all())

    # ---------------- end-to-end leg: z from pinned host memory, x_{t-1} read back each step ----------------
    z_host = [torch.randn(xt0.shape).pin_memory() for _ in range(4)]
    x_host = torch.empty(xt0.shape).pin_memory()
    z_dev = torch.empty_like(xt0)
    xt = xt0.clone()
    for k in range(W_):
        z_dev.copy_(z_host[k % 4], non_blocking=True)
        model(xt, t_devs[k], out=eps)
        xt = sched.step(xt, eps, T - 1 - (k % T), z=z_dev)
        x_host.copy_(xt, non_blocking=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for k in range(W_, W_ + K):
        z_dev.copy_(z_host[k % 4], non_blocking=True)
        model(xt, t_devs[k], out=eps)
        xt = sched.step(xt, eps, T - 1 - (k % T), z=z_dev)
        x_host.copy_(xt, non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    bytes_io = xt0.numel() * 4

    # ---------------- per-kernel-class timing (CUDA events around every launch), roofline of the dominant kernel
    import ctypes as C
    lib = _lib.lib()
    lib.wc_profile_begin()
    xt = xt0.clone()
    for k in range(K):
        model(xt, t_devs[k], out=eps)
        xt = sched.step(xt, eps, T - 1 - (k % T), z=zs[k % len(zs)])
    ms_c, cnt_c, work_c = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_double * 8)()
    _lib.check(lib.wc_profile_end(ms_c, cnt_c, work_c))
    names = ["igemm_tcgen05", "flash_attention_tcgen05", "groupnorm_silu", "boundary_conv", "ddpm_step", "other"]
    classes = {}
    for i, n in enumerate(names):
        if cnt_c[i]:
            classes[n] = {"ms_per_step": ms_c[i] / K, "launches_per_step": cnt_c[i] / K,
                          "work_per_step": work_c[i] / K}
    prof_total = sum(v["ms_per_step"] for v in classes.values())
    dom = "igemm_tcgen05"
    achieved = classes[dom]["work_per_step"] / (classes[dom]["ms_per_step"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tf_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"], "traffic": None,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "share_of_step": classes[dom]["ms_per_step"] / prof_total,
                "flops_per_launch": classes[dom]["work_per_step"] / classes[dom]["launches_per_step"],
                "avg_launch_ms": classes[dom]["ms_per_step"] / classes[dom]["launches_per_step"]}
    for n, v in classes.items():
        if n in ("igemm_tcgen05", "flash_attention_tcgen05"):
            v["tflops"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12
        else:
            v["gbs"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9

    # ---------------- reduce over ranks (max time), assemble the line ----------------
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0]), float(tt[1])
    ms_step = ms_total / K
    value = world * B / (T * ms_step * 1e-3)
    e2e_value = world * B / (T * (ms_e2e / K) * 1e-3)
    flops_step = model.flops_per_forward()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_rate(im_size, H, Wd, T, args.cpu_steps, 1, batch=1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sec_per_step_per_image": sec,
                        "sample": f"oracle port (PyTorch fp32, {cores} threads): 1 image, {args.cpu_steps} timed reverse "
                                  f"steps after 1 warm-up, extrapolated linearly to T={T}"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "height": H, "width": Wd, "schedule_steps": T,
                       "step": "one reverse-diffusion step (Unet.forward + fused posterior update) over the batch",
                       "l2": "per-step working set (activations ~GBs) exceeds the 126 MB L2; no explicit flush",
                       "unet_params_m": sum(p.numel() for p in model.parameters()) / 1e6,
                       "guidance": "not in this workload (SGG path = configs[2], see DESIGN.md)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches),
            "launches_per_step": launches / K,
            "clocks": clocks,
            "roofline": roofline,
            "kernel_classes": classes,
            "step_tflops": flops_step / (ms_step * 1e-3) / 1e12,
            "step_gflop_per_image": flops_step / B / 1e9,
            "step_frac_of_tensor_peak": flops_step / (ms_step * 1e-3) / 1e12 / peaks["tf_sustained"],
            "cpu_baseline": cpu_baseline,
            "finite": finite,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
