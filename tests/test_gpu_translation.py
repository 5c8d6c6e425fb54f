"""GPU parity of the SRGAN generator and of the repaired SGG translation driver against golden vectors produced by
the reference's own modules (tests/golden/srgan.pt, sgg.pt)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda")


def _psnr(y, ref, peak=None):
    peak = float(ref.abs().max()) if peak is None else peak
    return 10 * math.log10(peak ** 2 / float(((y - ref) ** 2).mean()))


def _srgan(seed, dev):
    from weatherconverter_b200.srgan_model.models import Generator
    from oracle.weights import synth_state_dict
    G = Generator(upscale_factor=4)
    G.load_state_dict(synth_state_dict(G.state_dict(), seed))
    return G.to(dev).eval()


def test_srgan_vs_golden(golden):
    dev = _dev()
    d = golden("srgan.pt")
    G = _srgan(d["seed"], dev)
    y = G(d["x"].to(dev)).cpu()
    rel = float((y - d["y"]).norm() / d["y"].norm())
    print(f"srgan: rms-rel {rel:.3e} psnr {_psnr(y, d['y'], 1.0):.1f} dB max-abs {float((y - d['y']).abs().max()):.3e}")
    assert y.shape == d["y"].shape
    # bf16 storage alone (fp32 arithmetic) gives rms-rel 2.05e-2 / PSNR 38.4 dB on this fixture (35 stacked convolutions)
    assert rel < 3e-2 and _psnr(y, d["y"], 1.0) > 35


def _models(d, dev):
    from oracle.weights import synth_state_dict
    from weatherconverter_b200.diffusion_model.models.unet_base import Unet, param_spec
    from weatherconverter_b200.diffusion_model.scheduler.linear_noise_scheduler import LinearNoiseScheduler
    from weatherconverter_b200.seg_model.network import modeling
    unet = Unet(d["cfg"]).to(dev).eval()
    unet.load_state_dict(synth_state_dict({k: (v, torch.float32) for k, v in param_spec(d["cfg"]).items()}, d["unet_seed"]))
    seg = modeling.deeplabv3plus_resnet50(19, 16, False)
    seg.load_state_dict(synth_state_dict(seg.state_dict(), d["seg_seed"]))
    seg = seg.to(dev).eval()
    return unet, seg, _srgan(d["srgan_seed"], dev), LinearNoiseScheduler(1000, 1e-4, 0.02)


def _check_driver(d, rec, base, out, seg, dev, tag):
    """Per-step x_t (PSNR, measured 77-85 dB on B200; asserted at 1.25x the measured error = -2 dB), the GUIDANCE TERM
    ALONE x_t - (mu + sigma) against the reference's (x_t itself cannot see a guidance error: the term is ~1e-5 of x_t),
    the final super-resolved image, and the segmentation argmax of the final image (north-star wording)."""
    N = d["N"]
    for k in range(N):
        ref = d["traj"][k]
        psnr = _psnr(rec[k].cpu(), ref)
        line = f"{tag} step {k} (i = {N - 1 - k}): psnr {psnr:.1f} dB max-abs {float((rec[k].cpu() - ref).abs().max()):.3e}"
        assert psnr > 72, (k, psnr)
        if k < N - 1:   # i > 0: a guided step
            delta_ref = (d["traj"][k].double() - d["base"][k].double())
            delta = (rec[k].double() - base[k].double()).cpu()
            rel = float((delta - delta_ref).norm() / delta_ref.norm())
            cos = float(torch.nn.functional.cosine_similarity(delta.flatten(), delta_ref.flatten(), dim=0))
            line += f" | guidance term: rms-rel {rel:.3f} cosine {cos:.4f} (|term| max {float(delta_ref.abs().max()):.2e})"
            # bf16 segmentor gradient vs the fp32 reference on this random-init fixture: DESIGN.md section 4 (measured
            # rms-rel 0.11-0.2); a deleted / mis-wired guidance call gives rms-rel 1.0, a sign error 2.0
            assert rel < 0.30 and cos > 0.95, (k, rel, cos)
        else:
            assert torch.equal(rec[k], base[k])      # i == 0: x_0 = mu, no guidance
        print(line)
    psnr = _psnr(out, d["sr_x0"], 1.0)
    pred = seg.infer(out.to(dev), d["gt"].to(dev), want_grad=False)["pred"].cpu()
    agree = float((pred[0].to(torch.uint8) == d["final_pred"][0]).float().mean())
    print(f"{tag} sr_x0: psnr {psnr:.1f} dB; segmentation argmax of the final image agrees on {100 * agree:.2f} % of the pixels")
    assert psnr > 36.5
    assert agree > 0.98


def test_repaired_driver_vs_golden(golden):
    from weatherconverter_b200.translation import sample_with_sgg
    dev = _dev()
    d = golden("sgg.pt")["driver"]
    unet, seg, G, sched = _models(d, dev)
    rec, base = [], []
    N = d["N"]
    out = sample_with_sgg(d["x0"], unet, sched, seg, d["gt"], G, n_steps=N, noise=d["noise"], t_forward=d["t_fwd"],
                          step_noise=d["zs"], record=rec, record_base=base).cpu()
    _check_driver(d, rec, base, out, seg, dev, "driver")
    # D1 switch: guidance computed but discarded == unguided chain
    a = sample_with_sgg(d["x0"], unet, sched, seg, d["gt"], G, n_steps=N, noise=d["noise"], t_forward=d["t_fwd"],
                        step_noise=d["zs"], reference_quirks=True)
    b = sample_with_sgg(d["x0"], unet, sched, seg, d["gt"], G, n_steps=N, noise=d["noise"], t_forward=d["t_fwd"],
                        step_noise=d["zs"], guidance=False)
    assert torch.equal(a, b)
    # ... and the guided chain is NOT the unguided one (the guidance term is small but present)
    assert not torch.equal(a, sample_with_sgg(d["x0"], unet, sched, seg, d["gt"], G, n_steps=N, noise=d["noise"],
                                              t_forward=d["t_fwd"], step_noise=d["zs"]))


def test_alternating_lcg_gsg_driver_vs_golden(golden):
    """mode="alternate": the shipped schedule of translation.py:84-87 (LCG on even steps, GSG on odd steps, i != 0), golden
    made with the reference's own apply_gsg / infer / compute_gradient_magnitude (LCG's final sum repaired)."""
    from weatherconverter_b200.translation import sample_with_sgg
    dev = _dev()
    d = golden("sgg.pt")["driver_alternate"]
    unet, seg, G, sched = _models(d, dev)
    rec, base = [], []
    out = sample_with_sgg(d["x0"], unet, sched, seg, d["gt"], G, n_steps=d["N"], noise=d["noise"], t_forward=d["t_fwd"],
                          step_noise=d["zs"], mode="alternate", record=rec, record_base=base).cpu()
    _check_driver(d, rec, base, out, seg, dev, "alternate")
    # the schedule really alternates: forcing GSG everywhere changes the even steps only from the first LCG step on
    rec_g = []
    sample_with_sgg(d["x0"], unet, sched, seg, d["gt"], G, n_steps=d["N"], noise=d["noise"], t_forward=d["t_fwd"],
                    step_noise=d["zs"], mode="gsg", record=rec_g)
    assert torch.equal(rec[0], rec_g[0]) == ((d["N"] - 1) % 2 == 1)
    assert not torch.equal(rec[1], rec_g[1])
