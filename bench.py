#!/usr/bin/env python
"""bench.py — headline benchmark of the WeatherConverter hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c5]

Default workload (every N, weak scaling) = the configuration BASELINE.json's metric is quoted on:
  c3  SGG-guided translation at 256x512 (BASELINE.json configs[2]/[3]), batch 32 per GPU, N = 500 reverse steps,
      geometry A of SURVEY.md 8d (reference-faithful): latent 64x128 (UNet config.yaml with im_size 64) -> SRGAN x4 ->
      DeepLabV3+-ResNet-50 (os16, 19 classes) forward + CE + input gradient at 256x512 -> avg-pool 4 -> guidance
      update (lambda = 60), GSG on every step (the repaired driver of SURVEY.md 8c).
  c2  UNet DDPM sampling at 128x256, batch 16 per GPU, 1000-step schedule (BASELINE.json configs[1]).
  c5  denoising-loss TRAINING step at 128x256, batch 64 per GPU, NCCL gradient all-reduce at N > 1 (BASELINE.json
      configs[4]; metric = trained images/s; step = add_noise + forward + MSE + backward + all-reduce + Adam).
A "step" is ONE reverse-diffusion step over the batch; per-step cost does not depend on t, so
    images/s = (N_gpus * batch) / (schedule_steps * seconds_per_step).
`value` times the steps with inputs resident in HBM; `e2e` times the same steps through the public Python API with
the per-step noise coming from pinned HOST memory (the reference draws z on the CPU every step,
linear_noise_scheduler.py:110) and x_{t-1} read back to the host, copies inside the timed region.
`--impl reference` times the reference algorithm's CPU implementation (the oracle port — the reference is a Python
source tree that cannot travel to the GPU box) on all host cores, on a bounded sample of the same workload.
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
    "c3": dict(kind="sgg", im_size=64, batch=32, h=64, w=128, T=500,
               desc="C3/C4: SGG-guided translation 256x512 (latent 64x128, SRGAN x4, DeepLabV3+-R50 gradient guidance every "
                    "step, lambda 60), batch 32/GPU, 500 reverse steps (BASELINE.json configs[2]/[3], geometry A)"),
    "c2": dict(kind="ddpm", im_size=128, batch=16, h=128, w=256, T=1000,
               desc="C2: UNet DDPM sampling 128x256, batch 16/GPU, 1000-step schedule (BASELINE.json configs[1])"),
    "c5": dict(kind="train", im_size=128, batch=64, h=128, w=256, T=1000,
               desc="C5: denoising-loss training step 128x256, batch 64/GPU, Adam, NCCL gradient all-reduce across GPUs "
                    "(BASELINE.json configs[4])"),
    "c1": dict(kind="ddpm", im_size=64, batch=4, h=64, w=64, T=50,
               desc="C1: UNet DDPM sampling 64x64, batch 4, 50 steps (BASELINE.json configs[0])"),
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
            self._halt.wait(0.05)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def block_labels(gen, B, H, W, blk=8):
    """Cityscapes-shaped synthetic masks: 8x8 blocks of uniform class in 0..18, 5 % of the blocks = 255."""
    import torch
    lab = torch.randint(0, 19, (B, H // blk, W // blk), generator=gen)
    lab[torch.rand(B, H // blk, W // blk, generator=gen) < 0.05] = 255
    return lab.repeat_interleave(blk, 1).repeat_interleave(blk, 2)


# ------------------------------------------------------------------------------------------------ GPU workloads
class GpuWorkload:
    def __init__(self, spec, dev, rank):
        import torch
        from weatherconverter_b200.diffusion_model.config.models import ModelConfig
        from weatherconverter_b200.diffusion_model.models.unet_base import Unet
        from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
        from weatherconverter_b200.sharding import initial_noise
        self.spec, self.dev = spec, dev
        B, h, w = spec["batch"], spec["h"], spec["w"]
        torch.manual_seed(3455)   # diffusion config.yaml:32
        self.unet = Unet(ModelConfig(im_size=spec["im_size"])).to(dev).eval()
        self.sched = LinearNoiseScheduler(spec["T"] if spec["kind"] == "ddpm" else 1000, 1e-4, 0.02)
        self.x0 = initial_noise((3, h, w), rank * B, (rank + 1) * B).to(dev)   # RNG keyed by the global image index
        self.eps = torch.empty_like(self.x0)
        self.T = spec["T"]
        if spec["kind"] == "sgg":
            from weatherconverter_b200.seg_model.network import modeling
            from weatherconverter_b200.srgan_model.models import Generator
            torch.manual_seed(42)   # seg config.yaml:3
            self.seg = modeling.deeplabv3plus_resnet50(num_classes=19, output_stride=16, pretrained_backbone=False)
            g = torch.Generator().manual_seed(7)
            for n, b in self.seg.named_buffers():   # non-degenerate eval BatchNorm statistics
                if n.endswith("running_mean"):
                    b.copy_(0.1 * torch.randn(b.shape, generator=g))
                elif n.endswith("running_var"):
                    b.copy_(0.5 + torch.rand(b.shape, generator=g))
            self.seg = self.seg.to(dev).eval()
            torch.manual_seed(0)
            self.srgan = Generator(upscale_factor=4)
            for n, b in self.srgan.named_buffers():
                if n.endswith("running_mean"):
                    b.copy_(0.1 * torch.randn(b.shape, generator=g))
                elif n.endswith("running_var"):
                    b.copy_(0.5 + torch.rand(b.shape, generator=g))
            self.srgan = self.srgan.to(dev).eval()
            self.gt = block_labels(torch.Generator().manual_seed(1234 + rank), B, 4 * h, 4 * w).to(dev)
        self.t_dev = {}

    def t_tensor(self, i):
        import torch
        if i not in self.t_dev:
            self.t_dev[i] = torch.tensor([i], device=self.dev)
        return self.t_dev[i]

    def step(self, xt, k, z):
        """One reverse step at timestep i = T-1-(k mod (T-1)) (never 0: the timed steps all inject noise)."""
        i = self.T - 1 - (k % (self.T - 1))
        self.unet(xt, self.t_tensor(i), out=self.eps)
        if self.spec["kind"] == "ddpm":
            return self.sched.step(xt, self.eps, i, z=z)
        from weatherconverter_b200.sgg.sgg import apply_gsg_batch
        mu, sigma, _ = self.sched.sample_prev_timestep(xt, self.eps, i, z=z)
        sr = self.srgan(xt)
        return apply_gsg_batch(self.seg, mu, sigma, sr, self.gt, 60.0)

    def flops_per_step(self):
        f = self.unet.flops_per_forward()
        parts = {"unet": f}
        if self.spec["kind"] == "sgg":
            sf, sb = self.seg.flops()
            parts.update(srgan=self.srgan.flops(), seg_fwd=sf, seg_dgrad=sb)
        return sum(parts.values()), parts


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_rate(spec, steps, warmup):
    """Oracle port (plain PyTorch fp32, all host threads) on a bounded sample: ONE image, `steps` timed reverse steps.
    Returns (images/s extrapolated linearly to the schedule, seconds/step/image, cores)."""
    import torch
    from oracle import deeplab, sgg, srgan
    from oracle.scheduler import OracleScheduler
    from oracle.unet import DEFAULT_MODEL_CONFIG, unet_forward, unet_param_spec
    from oracle.weights import synth_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = spec["im_size"]
    T, h, w = spec["T"], spec["h"], spec["w"]
    sd = synth_state_dict(unet_param_spec(cfg), 3455)
    sched = OracleScheduler(T if spec["kind"] == "ddpm" else 1000, 1e-4, 0.02)
    g = torch.Generator().manual_seed(1234)
    xt = torch.randn(1, 3, h, w, generator=g)
    if spec["kind"] == "sgg":
        seg_sd = synth_state_dict(deeplab.deeplab_param_spec("resnet50"), 42)
        gan_sd = synth_state_dict(srgan.srgan_param_spec(), 0)
        gt = block_labels(g, 1, 4 * h, 4 * w)
    times = []
    with torch.no_grad():
        for k in range(warmup + steps):
            i = T - 1 - k
            t0 = time.perf_counter()
            eps = unet_forward(sd, cfg, xt, torch.tensor([i]))
            z = torch.randn(xt.shape, generator=g)
            mean, sz, _ = sched.sample_prev_timestep(xt, eps, i, z=z)
            if spec["kind"] == "sgg":
                sr = srgan.generator_forward(gan_sd, xt)
                with torch.enable_grad():
                    xt, _, _ = sgg.apply_gsg(seg_sd, mean, sz, sr, gt, 60.0)
            else:
                xt = mean + sz
            dt = time.perf_counter() - t0
            if k >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return 1.0 / (T * sec), sec, cores


TRAIN_METRIC = "denoising-loss training images/sec (128x256, batch 64/GPU, Adam, NCCL gradient all-reduce)"


def cpu_train_rate(spec, steps, warmup):
    """Oracle port of the training step (PyTorch fp32 autograd + Adam, all host threads) on ONE image."""
    import torch
    from oracle.scheduler import OracleScheduler
    from oracle.train import train_step
    from oracle.unet import DEFAULT_MODEL_CONFIG, unet_param_spec
    from oracle.weights import synth_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = spec["im_size"]
    sd = synth_state_dict(unet_param_spec(cfg), 3455)
    sched = OracleScheduler(1000, 1e-4, 0.02)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(1, 3, spec["h"], spec["w"], generator=g) * 2 - 1
    times = []
    for k in range(warmup + steps):
        noise = torch.randn(x.shape, generator=g)
        t = torch.randint(0, 1000, (1,), generator=g)
        t0 = time.perf_counter()
        _, _, sd = train_step(sd, cfg, x, noise, t, sched, step=k + 1)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return 1.0 / sec, sec, cores


def run_train(args):
    """C5: training step.  value = images trained per second over all ranks (inputs resident in HBM); e2e = the same
    step with the image batch coming from pinned host memory and the loss read back every step."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from weatherconverter_b200 import _lib, ops
    from weatherconverter_b200.diffusion_model.config.models import ModelConfig
    from weatherconverter_b200.diffusion_model.models.unet_base import Unet
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
    from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE JSON line on stdout: NCCL's banner / debug output (NCCL_DEBUG from the environment or from
        # /etc/nccl.conf) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    spec = dict(WORKLOADS["c5"])
    if args.batch:
        spec["batch"] = args.batch
    B, h, w = spec["batch"], spec["h"], spec["w"]
    W_, K = max(args.warmup, 3), args.steps
    peaks = read_peaks()
    torch.manual_seed(3455)                       # same initial replica on every rank (config.yaml:32)
    model = Unet(ModelConfig(im_size=spec["im_size"])).to(dev)
    sched = LinearNoiseScheduler(1000, 1e-4, 0.02)
    trainer = DenoisingTrainer(model, sched, lr=1e-4)
    g = torch.Generator().manual_seed(1234 + rank)   # every rank trains on its own shard of the global batch
    imgs_host = [(torch.rand(B, 3, h, w, generator=g) * 2 - 1).pin_memory() for _ in range(2)]
    imgs = [x.to(dev) for x in imgs_host]
    ts = [torch.randint(0, 1000, (B,), generator=g) for _ in range(4)]
    gz = torch.Generator(device=dev).manual_seed(99 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(k, x):
        noise = torch.randn(x.shape, device=dev, generator=gz)
        return trainer.step(x, noise=noise, t=ts[k % 4])

    for k in range(W_):
        one(k, imgs[k % 2])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W_, W_ + K):
        loss = one(k, imgs[k % 2])
    e1.record()
    barrier()
    clocks = sampler.finish()
    launches = ops.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    loss_val = float(loss)
    # e2e: batch from pinned host memory, loss read back each step (as the reference's loss.item(), train_ddpm.py:116)
    x_dev = torch.empty_like(imgs[0])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for k in range(K):
        x_dev.copy_(imgs_host[k % 2], non_blocking=True)
        loss_val = float(one(k, x_dev))
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    classes, roofline = {}, None
    flops_step = trainer.flops_per_step
    if not args.no_profile and not args.resident_only:
        lib = _lib.lib()
        KP = 1
        lib.wc_profile_begin()
        one(0, imgs[0])
        ms_c, cnt_c, work_c = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_double * 8)()
        _lib.check(lib.wc_profile_end(ms_c, cnt_c, work_c))
        names = ["igemm_wgrad_tcgen05", "flash_attention_fwd_bwd_tcgen05", "groupnorm_fwd_bwd", "boundary_conv", "elementwise_adam_mse", "other"]
        for i, n in enumerate(names):
            if cnt_c[i]:
                classes[n] = {"ms_per_step": ms_c[i] / KP, "launches_per_step": cnt_c[i] / KP, "work_per_step": work_c[i] / KP}
        prof_total = sum(v["ms_per_step"] for v in classes.values())
        for n, v in classes.items():
            if "tcgen05" in n:
                v["tflops"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12
            elif v["work_per_step"] > 0:
                v["gbs"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9
            v["share_of_step"] = v["ms_per_step"] / prof_total
        dom = max(("igemm_wgrad_tcgen05", "flash_attention_fwd_bwd_tcgen05"), key=lambda n: classes.get(n, {}).get("ms_per_step", 0))
        achieved = classes[dom]["tflops"]
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tf_sustained"], "traffic": None,
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "share_of_step": classes[dom]["share_of_step"],
                    "flops_per_launch": classes[dom]["work_per_step"] / classes[dom]["launches_per_step"],
                    "avg_launch_ms": classes[dom]["ms_per_step"] / classes[dom]["launches_per_step"],
                    "launches_per_step": classes[dom]["launches_per_step"]}
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0]), float(tt[1])
    ms_step = ms_total / K
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / K * 1e-3)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_train_rate(spec, 1, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sec_per_step_per_image": sec,
                        "sample": f"oracle port (PyTorch fp32 autograd + Adam, {cores} threads): batch of 1 image, 1 timed step after 1 warm-up"}
    if rank == 0:
        line = {
            "metric": TRAIN_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": spec["desc"], "batch_per_gpu": B, "image": [h, w],
                       "step": "add_noise + UNet forward + MSE + full backward (data + weight gradients) + bucketed NCCL all-reduce + fused Adam",
                       "l2": "per-step working set (tens of GB of saved activations) exceeds the 126 MB L2; no explicit flush",
                       "weights": "random init (seed 3455), fp32 master weights, bf16 activations / GEMM operands, fp32 accumulation"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * h * w * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches), "launches_per_step": launches / K, "clocks": clocks, "roofline": roofline,
            "kernel_classes": classes, "step_gflop_per_image": flops_step / B / 1e9,
            "step_tflops": flops_step / (ms_step * 1e-3) / 1e12,
            "step_frac_of_tensor_peak": flops_step / (ms_step * 1e-3) / 1e12 / peaks["tf_sustained"],
            "cpu_baseline": cpu_baseline, "final_loss": loss_val, "finite": loss_val == loss_val,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: rank 0 only; CPU implementation of the same step on a bounded sample."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    spec = WORKLOADS[args.workload]
    if spec["kind"] == "train":
        steps, warm = min(args.steps, 2), 1
        value, sec, cores = cpu_train_rate(spec, steps, warm)
        sample = (f"oracle port (PyTorch fp32 autograd + Adam, {cores} threads): batch of 1 image, {steps} timed training steps "
                  f"after {warm} warm-up")
        print(json.dumps({
            "impl": "reference", "metric": TRAIN_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": spec["desc"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    steps = min(args.steps, 5)
    warm = min(args.warmup, 1)
    value, sec, cores = cpu_reference_rate(spec, steps, warm)
    sample = (f"oracle port (PyTorch fp32, {cores} threads): 1 image of the batch, {steps} timed reverse steps after {warm} "
              f"warm-up, extrapolated linearly to {spec['T']} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": spec["desc"], "batch_per_gpu": spec["batch"], "latent": [spec["h"], spec["w"]],
                   "schedule_steps": spec["T"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _dram_traffic_per_launch(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per igemm launch (bytes), averaged over the launches of one step, from the
    committed ncu capture of the same command (profiles/r1_dram_traffic_<workload>.json); None when no capture exists."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", f"r1_dram_traffic_{workload}.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="skip the e2e and profiling legs (ncu launch lists)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if WORKLOADS[args.workload]["kind"] == "train":
        return run_train(args)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from weatherconverter_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE JSON line on stdout: NCCL's banner / debug output (NCCL_DEBUG from the environment or from
        # /etc/nccl.conf) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(args.warmup, 3)
    K = args.steps
    spec = dict(WORKLOADS[args.workload])
    if args.batch:
        spec["batch"] = args.batch
    B, T = spec["batch"], spec["T"]
    peaks = read_peaks()
    wl = GpuWorkload(spec, dev, rank)
    x0 = wl.x0
    gz = torch.Generator(device=dev).manual_seed(99 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: z pre-generated in HBM ----------------
    zs = [torch.randn(x0.shape, device=dev, generator=gz) for _ in range(4)]
    xt = x0.clone()
    for k in range(W_):
        xt = wl.step(xt, k, zs[k % 4])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W_, W_ + K):
        xt = wl.step(xt, k, zs[k % 4])
    e1.record()
    barrier()
    clocks = sampler.finish()
    launches = ops.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(xt).all())

    # ---------------- end-to-end leg: z from pinned host memory, x_{t-1} read back each step ----------------
    if args.resident_only:
        args.no_profile = True
    z_host = [torch.randn(x0.shape).pin_memory() for _ in range(4)]
    x_host = torch.empty(x0.shape).pin_memory()
    z_dev = torch.empty_like(x0)
    xt = x0.clone()
    for k in range(0 if args.resident_only else W_):
        z_dev.copy_(z_host[k % 4], non_blocking=True)
        xt = wl.step(xt, k, z_dev)
        x_host.copy_(xt, non_blocking=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for k in range(W_, W_ + (1 if args.resident_only else K)):
        z_dev.copy_(z_host[k % 4], non_blocking=True)
        xt = wl.step(xt, k, z_dev)
        x_host.copy_(xt, non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    bytes_io = x0.numel() * 4

    # ---------------- per-kernel-class timing (CUDA events around every launch), roofline of the dominant kernel
    classes, roofline = {}, None
    flops_step, flop_parts = wl.flops_per_step()
    if not args.no_profile:
        lib = _lib.lib()
        KP = min(K, 3)
        lib.wc_profile_begin()
        xt = x0.clone()
        for k in range(KP):
            xt = wl.step(xt, k, zs[k % 4])
        ms_c, cnt_c, work_c = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_double * 8)()
        _lib.check(lib.wc_profile_end(ms_c, cnt_c, work_c))
        names = ["igemm_tcgen05", "flash_attention_tcgen05", "groupnorm_silu", "boundary_conv", "scheduler", "other"]
        for i, n in enumerate(names):
            if cnt_c[i]:
                classes[n] = {"ms_per_step": ms_c[i] / KP, "launches_per_step": cnt_c[i] / KP, "work_per_step": work_c[i] / KP}
        prof_total = sum(v["ms_per_step"] for v in classes.values())
        for n, v in classes.items():
            if n in ("igemm_tcgen05", "flash_attention_tcgen05"):
                v["tflops"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12
            elif v["work_per_step"] > 0:
                v["gbs"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9
            v["share_of_step"] = v["ms_per_step"] / prof_total
        dom = "igemm_tcgen05"
        achieved = classes[dom]["tflops"]
        roofline = {"bound": "tensor", "kernel": "igemm_kernel (csrc/igemm.cu)", "achieved": achieved,
                    "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                    "traffic": _dram_traffic_per_launch(args.workload),
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "share_of_step": classes[dom]["share_of_step"],
                    "flops_per_launch": classes[dom]["work_per_step"] / classes[dom]["launches_per_step"],
                    "avg_launch_ms": classes[dom]["ms_per_step"] / classes[dom]["launches_per_step"],
                    "launches_per_step": classes[dom]["launches_per_step"]}

    # ---------------- reduce over ranks (max time), assemble the line ----------------
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0]), float(tt[1])
    ms_step = ms_total / K
    value = world * B / (T * ms_step * 1e-3)
    e2e_value = world * B / (T * (ms_e2e / K) * 1e-3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_rate(WORKLOADS[args.workload], args.cpu_steps, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sec_per_step_per_image": sec,
                        "sample": f"oracle port (PyTorch fp32, {cores} threads): 1 image, {args.cpu_steps} timed reverse steps "
                                  f"after 1 warm-up, extrapolated linearly to {T} steps"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": spec["desc"], "batch_per_gpu": B, "latent": [spec["h"], spec["w"]], "schedule_steps": T,
                       "step": "one reverse-diffusion step over the batch (UNet forward + posterior"
                               + (" + SRGAN x4 + DeepLabV3+ forward/CE/input-gradient + guidance update)" if spec["kind"] == "sgg" else " update)"),
                       "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush",
                       "weights": "random init (UNet seed 3455, seg 42, SRGAN 0), BatchNorm statistics randomised"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches), "launches_per_step": launches / K,
            "clocks": clocks,
            "roofline": roofline,
            "kernel_classes": classes,
            "step_gflop_per_image": flops_step / B / 1e9,
            "step_gflop_parts_per_image": {k: v / B / 1e9 for k, v in flop_parts.items()},
            "step_tflops": flops_step / (ms_step * 1e-3) / 1e12,
            "step_frac_of_tensor_peak": flops_step / (ms_step * 1e-3) / 1e12 / peaks["tf_sustained"],
            "cpu_baseline": cpu_baseline,
            "finite": finite,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
