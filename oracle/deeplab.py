"""Oracle: DeepLabV3+ (ResNet-50/101, output_stride 16) forward + input gradient
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional fp32 restatement driven by a reference-format state_dict:
  seg_model/network/backbone/resnet.py  Bottleneck.forward :98-118, ResNet._make_layer :164-194,
                                        stem :142-153
  seg_model/network/_deeplab.py         DeepLabHeadV3Plus.forward :47-51, ASPP.forward :157-162,
                                        ASPPPooling.forward :128-131
  seg_model/network/utils.py            _SimpleSegmentationModel.forward :13-18
  seg_model/network/modeling.py         _segm_resnet :32-58 (os=16: dilate layer4, ASPP rates 6/12/18)
  seg_model/inference.py                infer :118-152 (CE ignore_index=255, gradient wrt the input)
"""
import torch
import torch.nn.functional as F

LAYERS = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3]}

# Optional storage-precision emulation (tests only): when set to "bf16", every activation that the CUDA path stores
# in bf16 is rounded to bf16 here too (straight-through in autograd) and conv weights are rounded to bf16.  This
# separates "bf16 storage changes ReLU on/off patterns" from genuine implementation differences.
EMULATE = None


class _RoundBoth(torch.autograd.Function):
    """bf16 rounding of an activation in the forward pass AND of its gradient in the backward pass (the CUDA path
    stores both in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def _q(x):
    return _RoundBoth.apply(x) if EMULATE == "bf16" else x


def _qw(w):
    return w.to(torch.bfloat16).float() if EMULATE == "bf16" else w


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _cb(sd, wkey, bnkey, x, **kw):
    """conv + eval BatchNorm.  Reference order (conv, then batch_norm) unless EMULATE: then BN is folded into the
    weights BEFORE the bf16 rounding, exactly like the CUDA path packs them (csrc/seg.cu: conv_bn)."""
    w = sd[wkey + ".weight"]
    if EMULATE == "bf16":
        scale = sd[bnkey + ".weight"] / torch.sqrt(sd[bnkey + ".running_var"] + 1e-5)
        shift = sd[bnkey + ".bias"] - sd[bnkey + ".running_mean"] * scale
        wf = (w * scale[:, None, None, None]).to(torch.bfloat16).float()
        return F.conv2d(x, wf, **kw) + shift[None, :, None, None]
    return _bn(sd, bnkey, F.conv2d(x, w, **kw))


def _bottleneck(sd, p, x, stride, dilation, has_down):
    out = _q(F.relu(_cb(sd, p + ".conv1", p + ".bn1", x)))
    out = _q(F.relu(_cb(sd, p + ".conv2", p + ".bn2", out, stride=stride, padding=dilation, dilation=dilation)))
    out = _cb(sd, p + ".conv3", p + ".bn3", out)
    idt = x
    if has_down:
        idt = _q(_cb(sd, p + ".downsample.0", p + ".downsample.1", x, stride=stride))
    return _q(F.relu(out + idt))


def block_plan(backbone="resnet50", output_stride=16):
    """[(prefix, inplanes, planes, stride, dilation, has_down)]; os=16: replace_stride_with_dilation=[F,F,T], os=8: [F,T,T]
    (modeling.py:34-39)."""
    plan, inplanes, dilation = [], 64, 1
    dilate_flags = [False, False, False, True] if output_stride == 16 else [False, False, True, True]
    for li, (planes, nblocks, stride, dilate) in enumerate(
            zip([64, 128, 256, 512], LAYERS[backbone], [1, 2, 2, 2], dilate_flags)):
        prev = dilation
        if dilate:
            dilation *= stride
            stride = 1
        for b in range(nblocks):
            first = b == 0
            has_down = first and (stride != 1 or inplanes != planes * 4)
            plan.append((f"backbone.layer{li + 1}.{b}", inplanes, planes, stride if first else 1,
                         prev if first else dilation, has_down))
            inplanes = planes * 4
    return plan


def deeplab_forward(sd, x, backbone="resnet50", taps=None, output_stride=16):
    """x [B,3,H,W] -> logits [B,nc,H,W].  output_stride 16 -> ASPP rates (6, 12, 18); 8 -> (12, 24, 36) (modeling.py:34-39)."""
    assert output_stride in (8, 16)
    H, W = x.shape[-2:]
    if EMULATE == "bf16":
        # the CUDA stem runs on the tensor core (csrc/conv.cu: build_conv_stem7s2): image and BN-folded weights in bf16, fp32
        # accumulation; the gradient w.r.t. the image stays fp32 (straight-through on the input rounding)
        xq = x + (x.to(torch.bfloat16).float() - x).detach()
        h = _q(F.relu(_cb(sd, "backbone.conv1", "backbone.bn1", xq, stride=2, padding=3)))
    else:
        h = _q(F.relu(_bn(sd, "backbone.bn1", F.conv2d(x, sd["backbone.conv1.weight"], stride=2, padding=3))))
    h = F.max_pool2d(h, 3, 2, 1)
    low = None
    for (p, _, _, stride, dil, has_down) in block_plan(backbone, output_stride):
        h = _bottleneck(sd, p, h, stride, dil, has_down)
        if p.startswith("backbone.layer1.") and p.endswith(str(LAYERS[backbone][0] - 1)):
            low = h
    if taps is not None:
        taps["low"], taps["out"] = low, h
    c = "classifier"
    ll = _q(F.relu(_cb(sd, c + ".project.0", c + ".project.1", low)))
    res = [_q(F.relu(_cb(sd, c + ".aspp.convs.0.0", c + ".aspp.convs.0.1", h)))]
    for k, r in zip((1, 2, 3), (6, 12, 18) if output_stride == 16 else (12, 24, 36)):
        res.append(_q(F.relu(_cb(sd, f"{c}.aspp.convs.{k}.0", f"{c}.aspp.convs.{k}.1", h, padding=r, dilation=r))))
    g = _q(F.adaptive_avg_pool2d(h, 1))
    g = _q(F.relu(_cb(sd, c + ".aspp.convs.4.1", c + ".aspp.convs.4.2", g)))
    res.append(F.interpolate(g, size=h.shape[-2:], mode="bilinear", align_corners=False))
    a = _q(F.relu(_cb(sd, c + ".aspp.project.0", c + ".aspp.project.1", torch.cat(res, 1))))
    a = _q(F.interpolate(a, size=ll.shape[-2:], mode="bilinear", align_corners=False))
    y = _q(F.relu(_cb(sd, c + ".classifier.0", c + ".classifier.1", torch.cat([ll, a], 1), padding=1)))
    y = F.conv2d(y, _qw(sd[c + ".classifier.3.weight"]), sd[c + ".classifier.3.bias"])
    if taps is not None:
        taps["aspp"], taps["logits_lowres"] = a, y
    return F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False)


def infer(sd, x, labels, backbone="resnet50", output_stride=16):
    """seg_model/inference.py:118-152 for B=1 (looped per image for B>1).
    Returns (pred int64 [B,H,W], input_grad [B,3,H,W], loss [B])."""
    preds, grads, losses = [], [], []
    for b in range(x.shape[0]):
        xb = x[b:b + 1].detach().clone().requires_grad_(True)
        out = deeplab_forward(sd, xb, backbone, output_stride=output_stride)
        preds.append(out.argmax(1))
        loss = F.cross_entropy(out, labels[b:b + 1], ignore_index=255)
        g, = torch.autograd.grad(loss, xb)
        grads.append(g); losses.append(loss.detach())
    return torch.cat(preds), torch.cat(grads), torch.stack(losses)


def deeplab_param_spec(backbone="resnet50", num_classes=19):
    f32, i64 = torch.float32, torch.int64
    spec = {}

    def conv(p, o, i, k):
        spec[p + ".weight"] = ((o, i, k, k), f32)

    def bn(p, c):
        spec[p + ".weight"] = ((c,), f32); spec[p + ".bias"] = ((c,), f32)
        spec[p + ".running_mean"] = ((c,), f32); spec[p + ".running_var"] = ((c,), f32)
        spec[p + ".num_batches_tracked"] = ((), i64)

    conv("backbone.conv1", 64, 3, 7); bn("backbone.bn1", 64)
    for (p, inpl, planes, stride, dil, has_down) in block_plan(backbone):
        conv(p + ".conv1", planes, inpl, 1); bn(p + ".bn1", planes)
        conv(p + ".conv2", planes, planes, 3); bn(p + ".bn2", planes)
        conv(p + ".conv3", planes * 4, planes, 1); bn(p + ".bn3", planes * 4)
        if has_down:
            conv(p + ".downsample.0", planes * 4, inpl, 1); bn(p + ".downsample.1", planes * 4)
    c = "classifier"
    conv(c + ".project.0", 48, 256, 1); bn(c + ".project.1", 48)
    conv(c + ".aspp.convs.0.0", 256, 2048, 1); bn(c + ".aspp.convs.0.1", 256)
    for k in (1, 2, 3):
        conv(f"{c}.aspp.convs.{k}.0", 256, 2048, 3); bn(f"{c}.aspp.convs.{k}.1", 256)
    conv(c + ".aspp.convs.4.1", 256, 2048, 1); bn(c + ".aspp.convs.4.2", 256)
    conv(c + ".aspp.project.0", 256, 1280, 1); bn(c + ".aspp.project.1", 256)
    conv(c + ".classifier.0", 256, 304, 3); bn(c + ".classifier.1", 256)
    spec[c + ".classifier.3.weight"] = ((num_classes, 256, 1, 1), f32)
    spec[c + ".classifier.3.bias"] = ((num_classes,), f32)
    return spec
