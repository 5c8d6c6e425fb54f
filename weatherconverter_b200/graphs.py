"""CUDA-graph capture of one reverse-diffusion step (SURVEY.md section 7 step 6 / 8b "capturable").

A reverse step is a fixed sequence of ~440 kernel launches whose parameters do not depend on the step index once the
timestep is read from device memory (scheduler.sample_prev_timestep_indexed; the UNet already reads t on the device).
``StepGraph`` captures that sequence ONCE on static buffers and replays it with a single cudaGraphLaunch per step:
the host then issues 3 operations per step (copy t, copy / draw z, launch the graph) instead of ~440, which is what the
launch-bound regimes need (batch 1 at the reference's own geometry, 8 ranks driven by 8 host processes).

Nothing here computes: the captured nodes are the library's own kernels, recorded by torch.cuda.graph on the capture
stream exactly as the eager path launches them, so a replay is bit-identical to the eager step.
"""
import torch


class StepGraph:
    """step_fn(xt, t_dev, z) -> x_{t-1}, all CUDA tensors; xt / z fp32 [B,3,h,w], t_dev int64 [1].

    Static buffers: ``xt`` (input AND output: the captured region ends with xt <- step_fn(...), so consecutive replays
    chain without host work), ``z``, ``t``.  ``warmup`` eager calls (default 2, on the capture stream) let every plan bind
    its workspace, pack its weights and set its function attributes before the capture starts."""

    def __init__(self, step_fn, shape, device, warmup=2, t_warm=1):
        self.xt = torch.zeros(shape, device=device, dtype=torch.float32)
        self.z = torch.zeros(shape, device=device, dtype=torch.float32)
        self.t = torch.full((1,), int(t_warm), device=device, dtype=torch.int64)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step_fn(self.xt, self.t, self.z)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        with torch.cuda.graph(self.graph):
            out = step_fn(self.xt, self.t, self.z)
            self.xt.copy_(out)
        self.xt.zero_()

    def load(self, xt):
        self.xt.copy_(xt)

    def replay(self, t_src, z=None):
        """One reverse step in place on ``self.xt``.  t_src: CUDA int64 [1] tensor holding the timestep (copied into the
        static scalar on the stream - no host synchronisation); z: noise for this step (None: keep what is in self.z)."""
        self.t.copy_(t_src, non_blocking=True)
        if z is not None:
            self.z.copy_(z, non_blocking=True)
        self.graph.replay()
        return self.xt
