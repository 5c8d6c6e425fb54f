"""CUDA-graph replay of the reverse step == the eager loop, bit for bit (weatherconverter_b200/graphs.py): the captured nodes
are the same kernels with the timestep and the noise fed through static device buffers."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda")


def _unet(cfg, seed, dev):
    from oracle.weights import synth_state_dict
    from weatherconverter_b200.diffusion_model.models.unet_base import Unet, param_spec
    m = Unet(cfg).to(dev).eval()
    m.load_state_dict(synth_state_dict({k: (v, torch.float32) for k, v in param_spec(cfg).items()}, seed))
    return m


def test_indexed_step_equals_scalar_step(golden):
    """wc_ddpm_step_indexed (timestep read on the device, coefficients from tables) == wc_ddpm_step == the reference golden."""
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
    dev = _dev()
    g = golden("scheduler.pt")
    s = LinearNoiseScheduler(1000, 1e-4, 0.02)
    x0, eps = g["x0"].to(dev), g["eps"].to(dev)
    for ti, d in g["sample_prev_timestep"].items():
        z = d["z"].to(dev)
        t_dev = torch.tensor([ti], device=dev)
        out = torch.empty_like(x0)
        mean, sigz, _ = s.sample_prev_timestep_indexed(x0, eps, t_dev, z, out=out)
        assert torch.equal(mean.cpu(), d["mean"]), ti
        if ti != 0:
            assert torch.equal(sigz.cpu(), d["sigma_z"]) and torch.equal(out.cpu(), d["mean"] + d["sigma_z"]), ti
            assert torch.equal(s.step_indexed(x0, eps, t_dev, z), s.step(x0, eps, ti, z=z))
        else:
            assert torch.equal(out.cpu(), d["mean"]) and float(sigz.abs().max()) == 0.0


def test_graphed_sampling_equals_eager(golden):
    from weatherconverter_b200.diffusion_model.sample_ddpm import sample_tensor
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
    dev = _dev()
    d = golden("sample_traj.pt")
    m = _unet(d["cfg"], d["seed"], dev)
    s = LinearNoiseScheduler(d["T"], 1e-4, 0.02)
    rec_e, rec_g = [], []
    a = sample_tensor(m, s, xT=d["xT"], noise=d["zs"], num_timesteps=d["T"], record=rec_e)
    b = sample_tensor(m, s, xT=d["xT"], noise=d["zs"], num_timesteps=d["T"], record=rec_g, use_graph=True)
    assert len(rec_e) == len(rec_g) == d["T"]
    for k in range(d["T"]):
        assert torch.equal(rec_e[k], rec_g[k]), k
    assert torch.equal(a, b)
    # the eager path still works after the capture (plans stay bound), and a second graphed run reproduces the first
    assert torch.equal(a, sample_tensor(m, s, xT=d["xT"], noise=d["zs"], num_timesteps=d["T"]))
    assert torch.equal(b, sample_tensor(m, s, xT=d["xT"], noise=d["zs"], num_timesteps=d["T"], use_graph=True))


@pytest.mark.parametrize("mode", ["gsg", "alternate"])
def test_graphed_translation_equals_eager(golden, mode):
    from oracle.weights import synth_state_dict
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
    from weatherconverter_b200.seg_model.network import modeling
    from weatherconverter_b200.srgan_model.models import Generator
    from weatherconverter_b200.translation import sample_with_sgg
    dev = _dev()
    d = golden("sgg.pt")["driver_alternate"]
    unet = _unet(d["cfg"], d["unet_seed"], dev)
    seg = modeling.deeplabv3plus_resnet50(19, 16, False)
    seg.load_state_dict(synth_state_dict(seg.state_dict(), d["seg_seed"]))
    seg = seg.to(dev).eval()
    G = Generator(upscale_factor=4)
    G.load_state_dict(synth_state_dict(G.state_dict(), d["srgan_seed"]))
    G = G.to(dev).eval()
    sched = LinearNoiseScheduler(1000, 1e-4, 0.02)
    B = 2
    x0 = torch.cat([d["x0"], d["x0"].flip(-1)])
    gt = torch.cat([d["gt"], d["gt"].flip(-1)])
    gen = torch.Generator().manual_seed(3)
    noise = torch.randn(B, *d["x0"].shape[1:], generator=gen)
    zs = torch.randn(d["N"], B, *d["x0"].shape[1:], generator=gen)
    t_fwd = torch.tensor([d["N"] - 1] * B)
    kw = dict(n_steps=d["N"], noise=noise, t_forward=t_fwd, step_noise=zs, mode=mode)
    rec_e, rec_g = [], []
    a = sample_with_sgg(x0, unet, sched, seg, gt, G, record=rec_e, **kw)
    b = sample_with_sgg(x0, unet, sched, seg, gt, G, record=rec_g, use_graph=True, **kw)
    for k in range(d["N"]):
        assert torch.equal(rec_e[k], rec_g[k]), (mode, k)
    assert torch.equal(a, b)
