"""Layer-by-layer check: feed each conv of the bf16-storage oracle's forward to the CUDA igemm and compare."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import deeplab
from oracle.weights import synth_state_dict
from weatherconverter_b200 import ops
dev = torch.device("cuda")
g = torch.load("tests/golden/seg_infer.pt", weights_only=False)
d = g["resnet50_64x128"]
osd = synth_state_dict(deeplab.deeplab_param_spec("resnet50"), 42)
rec = []
orig_cb = deeplab._cb
def cb(sd, wkey, bnkey, x, **kw):
    out = orig_cb(sd, wkey, bnkey, x, **kw)
    rec.append((wkey, bnkey, x.detach().clone(), out.detach().clone(), dict(kw)))
    return out
deeplab._cb = cb
deeplab.EMULATE = "bf16"
with torch.no_grad():
    deeplab.deeplab_forward(osd, d["x"], "resnet50")
deeplab.EMULATE = None
worst = []
for (wkey, bnkey, x, out, kw) in rec:
    w = osd[wkey + ".weight"]
    scale = osd[bnkey + ".weight"] / torch.sqrt(osd[bnkey + ".running_var"] + 1e-5)
    shift = osd[bnkey + ".bias"] - osd[bnkey + ".running_mean"] * scale
    wf = (w * scale[:, None, None, None])
    K = w.shape[-1]; stride = kw.get("stride", 1); dil = kw.get("dilation", 1)
    pad = kw.get("padding", 0)
    y = ops.conv2d(ops.to_nhwc_bf16(x.to(dev)), wf.to(dev), shift.to(dev), stride=stride, pad=pad, dil=dil)
    y = ops.to_nchw_f32(y).cpu()
    ref = out.to(torch.bfloat16).float()
    diff = (y - ref).abs()
    nbad = int((diff > 0).sum())
    worst.append((float(diff.max() / (ref.abs().max() + 1e-9)), nbad / ref.numel(), wkey, tuple(x.shape), K, stride, dil))
for w_ in sorted(worst, reverse=True)[:12]:
    print("maxrel %.3e mismatch-frac %.4f %s in%s K%d s%d d%d" % w_)
print("layers", len(worst), "with any mismatch", sum(1 for w_ in worst if w_[1] > 0))
