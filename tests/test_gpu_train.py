"""GPU parity of the full training step (forward + MSE + backward + Adam) against the CPU oracle and against the golden
made by the reference's own step (diffusion_model/train_ddpm.py:95-114; tests/golden/make_golden.py:g_train)."""
import zlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda")


def _build(cfg, seed, dev):
    from oracle.weights import synth_state_dict
    from weatherconverter_b200.diffusion_model.models.unet_base import Unet, param_spec
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
    sd = synth_state_dict({k: (v, torch.float32) for k, v in param_spec(cfg).items()}, seed)
    model = Unet(cfg).to(dev)
    model.load_state_dict(sd)
    return sd, model, LinearNoiseScheduler(1000, 1e-4, 0.02)


def test_train_step_vs_reference_golden_and_oracle(golden):
    from oracle.scheduler import OracleScheduler
    from oracle.train import train_step
    from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer
    dev = _dev()
    d = golden("train_step.pt")
    sd, model, sched = _build(d["cfg"], d["seed"], dev)
    trainer = DenoisingTrainer(model, sched, lr=d["lr"])
    loss = trainer.step(d["images"].to(dev), noise=d["noise"].to(dev), t=d["t"])
    torch.cuda.synchronize()
    loss_ref = float(d["loss"])
    print(f"loss {float(loss):.6f} vs reference {loss_ref:.6f}")
    assert abs(float(loss) - loss_ref) < 2e-3 * abs(loss_ref)          # bf16 activations, fp32 accumulation
    # full gradients vs the fp32 oracle (itself pinned to the reference in tests/test_oracle_golden.py)
    _, grads, new_sd = train_step(sd, d["cfg"], d["images"], d["noise"], d["t"], OracleScheduler(1000, 1e-4, 0.02), lr=d["lr"])
    params = dict(model.named_parameters())
    worst, worst_name, cos_min = 0.0, "", 1.0
    tot_num, tot_den = 0.0, 0.0
    for k, g_ref in grads.items():
        g = params[k].grad.detach().cpu()
        assert g.shape == g_ref.shape and torch.isfinite(g).all(), k
        num, den = float((g - g_ref).norm()), float(g_ref.norm())
        tot_num += num * num
        tot_den += den * den
        rel = num / (den + 1e-12)
        cos = float((g * g_ref).sum() / (g.norm() * g_ref.norm() + 1e-20))
        if rel > worst:
            worst, worst_name = rel, k
        cos_min = min(cos_min, cos)
        # golden (reference) norm of this gradient
        gn = float(d["params"][k]["grad_norm"])
        assert abs(float(g.norm()) - gn) < 6e-2 * gn + 1e-6, (k, float(g.norm()), gn)
    total_rel = (tot_num / tot_den) ** 0.5
    print(f"gradients: global rms-rel {total_rel:.3e}, worst tensor {worst_name} {worst:.3e}, min cosine {cos_min:.5f}")
    assert total_rel < 3e-2, total_rel
    assert worst < 1e-1, (worst_name, worst)
    assert cos_min > 0.995, cos_min
    # Adam: the first step moves every weight by ~lr*sign(grad); compare the sampled post-step values with the reference
    bad = 0
    n_s = 0
    for k, ref in d["params"].items():
        p = params[k].detach().cpu().flatten()
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()) & 0x7FFFFFFF)
        idx = torch.randint(0, p.numel(), (min(32, p.numel()),), generator=g)
        diff = (p[idx] - ref["new_samples"]).abs()
        bad += int((diff > 5e-5).sum())
        n_s += idx.numel()
    print(f"post-Adam sampled weights: {bad} of {n_s} differ by more than 0.5*lr")
    # Adam's first update is lr*g/(|g|+eps): entries whose gradient is within bf16 noise of zero may move the other way
    assert bad <= 0.05 * n_s


def test_two_steps_deterministic_and_loss_decreases():
    """Two trainers on the same data give bit-identical weights (no atomics anywhere), and a few steps on a fixed batch
    reduce the loss."""
    from oracle.unet import DEFAULT_MODEL_CONFIG
    from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer
    dev = _dev()
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = 32
    g = torch.Generator().manual_seed(3)
    images = (torch.rand(2, 3, 32, 64, generator=g) * 2 - 1).to(dev)
    noise = torch.randn(2, 3, 32, 64, generator=g).to(dev)
    t = torch.tensor([100, 600])
    finals, losses = [], []
    for rep in range(2):
        if rep == 1:
            # the second trainer gets recycled device memory full of NaNs: nothing may depend on uninitialised buffers
            junk = [torch.full((n,), float("nan"), device=dev) for n in (1 << 28, 1 << 26, 1 << 20, 65536, 4096, 256) for _ in range(3)]
            del junk
        _, model, sched = _build(cfg, 7, dev)
        tr = DenoisingTrainer(model, sched, lr=1e-4)
        ls = [float(tr.step(images, noise=noise, t=t)) for _ in range(4)]
        finals.append(tr.flat_params.clone())
        losses.append(ls)
    assert torch.equal(finals[0], finals[1])
    assert losses[0] == losses[1]
    assert losses[0][-1] < losses[0][0], losses[0]
    # the inference module sees the trained weights
    with torch.no_grad():
        y = model(images, t.to(dev))
    assert torch.isfinite(y).all()


def test_train_step_ragged_shapes_vs_oracle():
    """Odd batch (3), non-square 32x64 input with im_size 64: token counts 2048 / 512 / 128 / 32 (attention tiles larger than
    the sequence, pixel tiles spanning several samples in the weight-gradient GEMMs).  Full gradients vs the fp32 oracle."""
    from oracle.scheduler import OracleScheduler
    from oracle.train import train_step
    from oracle.unet import DEFAULT_MODEL_CONFIG
    from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer
    dev = _dev()
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = 64
    sd, model, sched = _build(cfg, 99, dev)
    g = torch.Generator().manual_seed(12)
    images = torch.rand(3, 3, 32, 64, generator=g) * 2 - 1
    noise = torch.randn(3, 3, 32, 64, generator=g)
    t = torch.tensor([5, 420, 999])
    trainer = DenoisingTrainer(model, sched, lr=1e-4)
    loss = trainer.forward_backward(sched.add_noise(images.to(dev), noise.to(dev), t.to(dev)), t, noise.to(dev))
    torch.cuda.synchronize()
    loss_ref, grads, _ = train_step(sd, cfg, images, noise, t, OracleScheduler(1000, 1e-4, 0.02))
    assert abs(float(loss) - float(loss_ref)) < 3e-3 * abs(float(loss_ref)), (float(loss), float(loss_ref))
    params = dict(model.named_parameters())
    num = den = 0.0
    worst, worst_name = 0.0, ""
    for k, g_ref in grads.items():
        gg = params[k].grad.detach().cpu()
        assert torch.isfinite(gg).all(), k
        a, b = float((gg - g_ref).norm()), float(g_ref.norm())
        num += a * a
        den += b * b
        if a / (b + 1e-12) > worst:
            worst, worst_name = a / (b + 1e-12), k
    print(f"ragged step: loss {float(loss):.5f} vs {float(loss_ref):.5f}, gradient rms-rel {(num / den) ** 0.5:.3e}, worst {worst_name} {worst:.3e}")
    assert (num / den) ** 0.5 < 3e-2
    assert worst < 1.5e-1, (worst_name, worst)


def test_resume_from_checkpoint_continues_adam():
    """The reference's resume flow (train_ddpm.py:63-68,81-84): save model + optimizer state dicts, load them into a fresh model
    and a fresh torch.optim.Adam, keep training.  The resumed run must continue the moments and the bias-correction step count:
    two steps + save / load + one step == three uninterrupted steps, bit for bit; and train() (the drop-in loop) imports the
    optimizer state by itself and exports a loadable one after every epoch."""
    import io
    from oracle.unet import DEFAULT_MODEL_CONFIG
    from weatherconverter_b200.diffusion_model.train_ddpm import DenoisingTrainer, train
    dev = _dev()
    cfg = dict(DEFAULT_MODEL_CONFIG); cfg["im_size"] = 32
    g = torch.Generator().manual_seed(5)
    images = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).to(dev)
    noises = [torch.randn(2, 3, 32, 32, generator=g).to(dev) for _ in range(3)]
    ts = [torch.tensor([100 + 7 * k, 600 - 11 * k]) for k in range(3)]

    _, model_a, sched = _build(cfg, 7, dev)
    tr_a = DenoisingTrainer(model_a, sched, lr=1e-4)
    for k in range(3):
        tr_a.step(images, noise=noises[k], t=ts[k])

    _, model_b, _ = _build(cfg, 7, dev)
    opt_b = torch.optim.Adam(model_b.parameters(), lr=1e-4)
    tr_b = DenoisingTrainer(model_b, sched, lr=1e-4)
    for k in range(2):
        tr_b.step(images, noise=noises[k], t=ts[k])
    tr_b.export_optimizer_state(opt_b)
    buf = io.BytesIO()      # the reference's save_checkpoint dictionary (train_ddpm.py:56-61)
    torch.save({"epoch": 1, "model_state_dict": model_b.state_dict(), "optimizer_state_dict": opt_b.state_dict()}, buf)
    buf.seek(0)
    ckpt = torch.load(buf, map_location=dev)

    _, model_c, _ = _build(cfg, 99, dev)          # different initial weights: everything must come from the checkpoint
    model_c.load_state_dict(ckpt["model_state_dict"])
    opt_c = torch.optim.Adam(model_c.parameters(), lr=1e-4)
    opt_c.load_state_dict(ckpt["optimizer_state_dict"])
    tr_c = DenoisingTrainer(model_c, sched, lr=1e-4)
    assert tr_c.import_optimizer_state(opt_c) == len(tr_c.slices) and tr_c.step_count == 2
    tr_c.step(images, noise=noises[2], t=ts[2])
    torch.cuda.synchronize()
    for (na, pa), (nc, pc) in zip(model_a.named_parameters(), model_c.named_parameters()):
        assert na == nc and torch.equal(pa, pc), na
    # a trainer that ignores the saved moments (fresh Adam) does NOT reproduce the uninterrupted run
    _, model_d, _ = _build(cfg, 99, dev)
    model_d.load_state_dict(ckpt["model_state_dict"])
    tr_d = DenoisingTrainer(model_d, sched, lr=1e-4)
    tr_d.step(images, noise=noises[2], t=ts[2])
    assert not torch.equal(tr_d.flat_params, tr_a.flat_params)
    # train(): imports the state of the optimizer it is given, exports after every epoch
    _, model_e, _ = _build(cfg, 99, dev)
    model_e.load_state_dict(ckpt["model_state_dict"])
    opt_e = torch.optim.Adam(model_e.parameters(), lr=1e-4)
    opt_e.load_state_dict(ckpt["optimizer_state_dict"])
    seen = []
    train([images.cpu()], model_e, opt_e, torch.nn.MSELoss(), sched, epochs=2, seed=11,
          on_epoch_end=lambda ep: seen.append(int(float(next(iter(opt_e.state.values()))["step"]))))
    assert seen == [3, 4]       # continued from step 2: one batch per epoch
