"""GPU parity of DeepLabV3+ forward / loss head / input gradient and of the guidance update against golden vectors
from the reference (tests/golden/seg_infer.pt, sgg.pt).

Stated tolerances (bf16 storage + fp32 accumulation vs the fp32 reference):
  * logits rms-rel <= 3e-2, argmax pixel agreement >= 99 %, per-image loss within 2 %;
  * input gradient: this random-init 50-layer ReLU network is chaotic at bf16 resolution.  Measured with the oracle
    alone (fp32 arithmetic, EMULATE="bf16" storage): rounding moves the reference's own gradient to cosine ~0.98 /
    rms-rel ~0.20, and perturbing just 0.05 % of the stored activations by ONE bf16 ulp moves it by rms-rel 0.18
    (DESIGN.md section 4).  Hence: vs the fp32 golden cosine >= 0.96 / rms-rel <= 0.30; vs the bf16-storage oracle
    cosine >= 0.975 / rms-rel <= 0.25; and a wiring check that does not depend on the noise floor: the CUDA gradient
    must be closer to the full oracle gradient than to the oracle gradient with any one branch cut (every branch
    carries >= 20 % of the gradient).  Each backward building block is checked tightly in test_gpu_seg_ops.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda")


def _seg(backbone, seed, dev):
    from weatherconverter_b200.seg_model.network import modeling
    from oracle.weights import synth_state_dict
    m = getattr(modeling, "deeplabv3plus_" + backbone)(num_classes=19, output_stride=16, pretrained_backbone=False)
    m.load_state_dict(synth_state_dict(m.state_dict(), seed))
    return m.to(dev).eval()


def test_seg_infer_vs_golden(golden):
    from oracle import deeplab
    from oracle.weights import synth_state_dict
    dev = _dev()
    g = golden("seg_infer.pt")
    for tag, d in g.items():
        bb = tag.split("_")[0]
        m = _seg(bb, d["seed"], dev)
        out = m.infer(d["x"].to(dev), d["gt"].to(dev), want_grad=True, want_logits=True)
        pred, grad = out["pred"].cpu(), out["grad"].cpu()
        agree = float((pred[0].to(torch.uint8) == d["pred"]).float().mean())
        cos = float(torch.nn.functional.cosine_similarity(grad.flatten(), d["grad"].flatten(), dim=0))
        rel = float((grad - d["grad"]).norm() / d["grad"].norm())
        sd = synth_state_dict(deeplab.deeplab_param_spec(bb), d["seed"])
        with torch.no_grad():
            ref_logits = deeplab.deeplab_forward(sd, d["x"], bb)
        lrel = float((out["logits"].cpu() - ref_logits).norm() / ref_logits.norm())
        ref_loss = float(torch.nn.functional.cross_entropy(ref_logits, d["gt"], ignore_index=255))
        deeplab.EMULATE = "bf16"
        try:
            _, grad_emu, _ = deeplab.infer(sd, d["x"], d["gt"], bb)
        finally:
            deeplab.EMULATE = None
        cos_e = float(torch.nn.functional.cosine_similarity(grad.flatten(), grad_emu.flatten(), dim=0))
        rel_e = float((grad - grad_emu).norm() / grad_emu.norm())
        print(f"{tag}: argmax agreement {agree:.4f} logits rms-rel {lrel:.3e} loss {float(out['loss'][0]):.5f} (ref {ref_loss:.5f}) "
              f"| grad vs fp32 golden: cosine {cos:.5f} rms-rel {rel:.3e} | vs bf16-storage oracle: cosine {cos_e:.5f} rms-rel {rel_e:.3e}")
        assert lrel < 3e-2, (tag, lrel)
        assert agree > 0.99, (tag, agree)
        assert abs(float(out["loss"][0]) - ref_loss) < 2e-2 * abs(ref_loss) + 1e-3
        assert cos > 0.96 and rel < 0.30, (tag, cos, rel)
        assert cos_e > 0.975 and rel_e < 0.25, (tag, cos_e, rel_e)


def test_seg_output_stride_8_vs_golden(golden):
    """deeplabv3plus_resnet50(output_stride=8) - the reference factories' DEFAULT argument (modeling.py:182-202): layers 3 and 4
    dilated (dilation 2 / 4), ASPP rates 12 / 24 / 36 - against a golden made by the reference itself."""
    from oracle import deeplab
    from oracle.weights import synth_state_dict
    from weatherconverter_b200.seg_model.network import modeling
    dev = _dev()
    for tag, d in golden("seg_infer_os8.pt").items():
        m = modeling.deeplabv3plus_resnet50(num_classes=19, pretrained_backbone=False)      # output_stride defaults to 8
        assert m.output_stride == 8
        m.load_state_dict(synth_state_dict(m.state_dict(), d["seed"]))
        m = m.to(dev).eval()
        out = m.infer(d["x"].to(dev), d["gt"].to(dev), want_grad=True, want_logits=True)
        sd = synth_state_dict(deeplab.deeplab_param_spec("resnet50"), d["seed"])
        with torch.no_grad():
            ref_logits = deeplab.deeplab_forward(sd, d["x"], "resnet50", output_stride=8)
        lrel = float((out["logits"].cpu() - ref_logits).norm() / ref_logits.norm())
        agree = float((out["pred"][0].cpu().to(torch.uint8) == d["pred"]).float().mean())
        grad = out["grad"].cpu()
        cos = float(torch.nn.functional.cosine_similarity(grad.flatten(), d["grad"].flatten(), dim=0))
        rel = float((grad - d["grad"]).norm() / d["grad"].norm())
        print(f"{tag}: argmax agreement {agree:.4f} logits rms-rel {lrel:.3e} | grad vs fp32 golden: cosine {cos:.5f} rms-rel {rel:.3e}")
        assert lrel < 3e-2 and agree > 0.99, (tag, lrel, agree)
        assert cos > 0.96 and rel < 0.30, (tag, cos, rel)


def test_seg_gradient_on_conditioned_fixture(golden):
    """Input gradient on a WELL-CONDITIONED fixture (tests/golden/seg_conditioned.pt, made by the reference's infer): the last
    BatchNorm scale of every bottleneck x 0.2, i.e. residual branches are small corrections of the identity path as in a
    trained ResNet.  There bf16 storage (oracle, EMULATE="bf16") moves the reference's own gradient only to cosine 0.994 /
    rms-rel 0.11, against 0.974 / 0.23 on the plain random-init fixture and 0.12 / 1.27 with BatchNorm statistics calibrated
    on a forward pass (random weights + normalisation = exploding backward gains; numbers in DESIGN.md section 4).  The CUDA
    gradient must reach the same level: cosine >= 0.99, rms-rel <= 0.15, batch of two distinct images."""
    from oracle import deeplab
    from oracle.weights import synth_state_dict
    from weatherconverter_b200.seg_model.network import modeling
    dev = _dev()
    d = golden("seg_conditioned.pt")
    m = modeling.deeplabv3plus_resnet50(num_classes=19, output_stride=16, pretrained_backbone=False)
    sd = synth_state_dict(m.state_dict(), d["seed"])
    for k in sd:
        if k.endswith("bn3.weight"):
            sd[k] = sd[k] * d["bn3_gain"]
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    out = m.infer(d["x"].to(dev), d["gt"].to(dev), want_grad=True)
    for b in range(d["x"].shape[0]):
        grad, ref = out["grad"][b].cpu(), d["grad"][b]
        cos = float(torch.nn.functional.cosine_similarity(grad.flatten(), ref.flatten(), dim=0))
        rel = float((grad - ref).norm() / ref.norm())
        agree = float((out["pred"][b].cpu().to(torch.uint8) == d["pred"][b]).float().mean())
        deeplab.EMULATE = "bf16"
        try:
            _, grad_emu, _ = deeplab.infer(sd, d["x"][b:b + 1], d["gt"][b:b + 1], "resnet50")
        finally:
            deeplab.EMULATE = None
        cos_o = float(torch.nn.functional.cosine_similarity(grad_emu.flatten(), ref.flatten(), dim=0))
        rel_o = float((grad_emu[0] - ref).norm() / ref.norm())
        print(f"conditioned image {b}: CUDA vs fp32 reference cosine {cos:.5f} rms-rel {rel:.3e} argmax {agree:.4f} | "
              f"bf16-storage oracle vs fp32 reference cosine {cos_o:.5f} rms-rel {rel_o:.3e}")
        assert cos > 0.99 and rel < 0.15, (b, cos, rel)
        assert agree > 0.99, (b, agree)


def test_seg_gradient_has_every_branch(golden):
    """Wiring check: cutting any one gradient branch in the oracle moves its gradient further from the CUDA
    gradient than the full oracle gradient is."""
    import torch.nn.functional as F
    from oracle import deeplab
    from oracle.weights import synth_state_dict
    dev = _dev()
    d = golden("seg_infer.pt")["resnet50_64x128"]
    m = _seg("resnet50", d["seed"], dev)
    grad = m.infer(d["x"].to(dev), d["gt"].to(dev))["grad"].cpu()
    sd = synth_state_dict(deeplab.deeplab_param_spec("resnet50"), d["seed"])

    def oracle_grad(cut=None):
        x = d["x"].clone().requires_grad_(True)
        orig = deeplab._cb

        def cb(sd_, wkey, bnkey, xx, **kw):
            return orig(sd_, wkey, bnkey, xx.detach() if (cut and wkey.endswith(cut)) else xx, **kw)
        deeplab._cb = cb
        try:
            loss = F.cross_entropy(deeplab.deeplab_forward(sd, x, "resnet50"), d["gt"], ignore_index=255)
            return torch.autograd.grad(loss, x)[0]
        finally:
            deeplab._cb = orig

    full = float((grad - oracle_grad()).norm())
    for cut in ("classifier.project.0", "aspp.convs.4.1", "aspp.convs.1.0", "aspp.convs.2.0", "aspp.convs.3.0", "aspp.convs.0.0",
                "layer2.0.downsample.0", "layer3.0.downsample.0", "layer1.0.downsample.0", "layer3.0.conv1", "layer4.2.conv1"):
        dist = float((grad - oracle_grad(cut)).norm())
        print(f"cut {cut}: distance {dist / full:.2f} x the distance to the full gradient")
        assert dist > 1.1 * full, cut


def test_fused_pooled_gradient_equals_avg_pool_of_gradient(golden):
    """wc_seg_infer_pooled(grad_pool=4) == F.avg_pool2d(full input gradient, 4) (sgg.py:18) up to fp32 re-association."""
    import torch.nn.functional as F
    dev = _dev()
    d = golden("seg_infer.pt")["resnet50_128x256"]
    m = _seg("resnet50", d["seed"], dev)
    x, gt = d["x"].to(dev).repeat(2, 1, 1, 1), d["gt"].to(dev).repeat(2, 1, 1)
    full = m.infer(x, gt)["grad"]
    for P in (2, 4, 8):
        pooled = m.infer(x, gt, grad_pool=P)["grad"]
        ref = F.avg_pool2d(full, P, P)
        rel = float((pooled - ref).norm() / ref.norm())
        print(f"pool {P}: rms-rel {rel:.2e}")
        assert pooled.shape == ref.shape and rel < 1e-4


def test_sgg_update_kernel_matches_reference_arithmetic(golden):
    """wc_sgg_update == avg_pool2d(4) -> compute_gradient_magnitude (float64) -> mu + lambda*sigma*mag + sigma
    (sgg.py:18-22, inference.py:39-43), bit-exact after the final cast to fp32, for a batch of gradients."""
    import torch.nn.functional as F
    from oracle.sgg import compute_gradient_magnitude
    from weatherconverter_b200._lib import check, lib, ptr, stream_ptr
    dev = _dev()
    d = golden("sgg.pt")["gsg"]
    g = torch.Generator().manual_seed(77)
    B, h, w = 3, d["mu"].shape[2], d["mu"].shape[3]
    grad = 1e-4 * torch.randn(B, 3, 4 * h, 4 * w, generator=g)
    mu = torch.randn(B, 3, h, w, generator=g)
    sig = 0.1 * torch.randn(B, 3, h, w, generator=g)
    ref = []
    for b in range(B):
        mag = compute_gradient_magnitude(F.avg_pool2d(grad[b:b + 1], 4, 4))
        ref.append(((mu[b:b + 1] + 60.0 * sig[b:b + 1] * mag) + sig[b:b + 1]).float())
    ref = torch.cat(ref)
    xt = torch.empty(B, 3, h, w, device=dev)
    mag_out = torch.empty(B, h, w, device=dev)
    grad_d, mu_d, sig_d = grad.to(dev), mu.to(dev), sig.to(dev)   # keep the device tensors alive across the raw-pointer call
    check(lib().wc_sgg_update(ptr(grad_d), ptr(mu_d), ptr(sig_d), ptr(xt), ptr(mag_out), B, h, w, 4, 60.0, stream_ptr()))
    assert torch.equal(xt.cpu(), ref)


def test_apply_gsg_end_to_end(golden):
    from weatherconverter_b200.sgg.sgg import apply_gsg
    dev = _dev()
    d = golden("sgg.pt")["gsg"]
    m = _seg("resnet50", 42, dev)
    xt = apply_gsg(m, d["mu"].to(dev), d["sigma"].to(dev), d["sr_xt"].to(dev), d["gt"].to(dev), d["lam"]).cpu()
    ref = d["xt"].float()
    guid = ref - (d["mu"] + d["sigma"])          # the guidance term alone
    got = xt - (d["mu"] + d["sigma"])
    rel = float((got - guid).norm() / guid.norm())
    cos = float(torch.nn.functional.cosine_similarity(got.flatten(), guid.flatten(), dim=0))
    print(f"apply_gsg: guidance-term rms-rel {rel:.3e} cosine {cos:.4f}, total max-abs {float((xt - ref).abs().max()):.3e}")
    assert rel < 0.25 and cos > 0.97
    assert (xt - ref).abs().max() < 1e-3


def test_apply_lcg_vs_golden(golden):
    """Repaired local class guidance: 19 masked passes batched through the plan vs the reference's per-class loop."""
    from weatherconverter_b200.sgg.sgg import apply_lcg
    dev = _dev()
    d = golden("sgg.pt")["lcg"]
    m = _seg("resnet50", 42, dev)
    xt = apply_lcg(m, d["mu"].to(dev), d["sigma"].to(dev), d["sr_xt"].to(dev), d["gt"].to(dev), d["lam"]).cpu()
    ref = d["xt"]
    base = d["mu"] + d["sigma"]
    rel = float(((xt - base) - (ref - base)).norm() / (ref - base).norm())
    print(f"apply_lcg: guidance-term rms-rel {rel:.3e}, total max-abs {float((xt - ref).abs().max()):.3e}")
    assert rel < 0.25 and (xt - ref).abs().max() < 1e-3
