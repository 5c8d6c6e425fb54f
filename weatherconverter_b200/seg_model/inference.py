"""Segmentation inference + input gradient with the reference's entry points (seg_model/inference.py:20-152).

``infer(model, input_tensor, encoded_label_tensor)`` returns ``(pred ndarray [H,W] int64, input_gradients Tensor
[1,3,H,W], gradients ndarray)`` like the reference and keeps its side effects (``input_tensor.requires_grad`` is
set, the returned tensor aliases ``input_tensor.grad``).  The forward, the cross-entropy head and the data-gradient
backward all run in the C plan (csrc/seg.cu); no autograd graph is built and no weight gradients are computed.
``infer_batch`` is the batched, sync-free form used by the translation loop.
"""
import numpy as np
import torch
import yaml

from . import network
from .. import _lib
from .._lib import check, lib, ptr, stream_ptr

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def load_config(config_path: str):
    with open(config_path, 'r') as file:
        return yaml.safe_load(file)


def load_model(model_path, model_config) -> torch.nn.Module:
    g = (lambda k: model_config[k]) if isinstance(model_config, dict) else (lambda k: getattr(model_config, k))
    model = network.modeling.__dict__[g("name")](num_classes=g("num_classes"), output_stride=g("output_stride"),
                                                pretrained_backbone=False)
    model.load_state_dict(torch.load(model_path, map_location="cpu")['model_state_dict'])
    model.to(device)
    model.eval()
    return model


def preprocess(ori_img_path: str, img_path: str, gt_label_ids_path: str, gt_color_path: str, verbose=False):
    """Reference :60-116 with the tensor work on the GPU (byte-exact, weatherconverter_b200/image_io.py): PNG decoding stays
    on the host (PIL), then ExtToTensor + ExtNormalize of the image, NEAREST resize (540, 960) + centre crop 512 + labelIds ->
    trainIds of the label run as kernels.  Returns (original_image PIL 512, input_tensor [1,3,H,W], encoded_label_tensor
    [1,512,512] int64, lbl_colored_img PIL 512) like the reference."""
    from PIL import Image
    import torchvision.transforms as T
    from .. import image_io
    img = Image.open(img_path).convert("RGB")
    ori_img = Image.open(ori_img_path).convert("RGB")
    label = Image.open(gt_label_ids_path)
    label_colored = Image.open(gt_color_path)
    input_tensor = image_io.normalize_image(torch.from_numpy(np.array(img)).to(device))
    encoded_label_tensor = image_io.encode_label(torch.from_numpy(np.array(label, dtype=np.uint8)).to(device))
    host_tf = T.Compose([T.Resize(size=(1080 // 2, 1920 // 2), interpolation=Image.BILINEAR), T.CenterCrop(size=(512, 512))])
    return host_tf(ori_img), input_tensor, encoded_label_tensor, host_tf(label_colored)   # the two PIL outputs are display-only


def infer_batch(model, input_tensor, encoded_label_tensor, want_grad=True, grad_pool=1):
    """Batched, stream-ordered: returns dict(pred int64 [B,H,W], grad [B,3,H,W], loss [B]); image b's loss is the CE
    mean over its own valid pixels, i.e. a vmap of the reference's B = 1 call."""
    labels = encoded_label_tensor
    if labels.dim() == 4:
        labels = labels.squeeze(1)
    return model.infer(input_tensor, labels, want_grad=want_grad, grad_pool=grad_pool)


def infer(model, input_tensor, encoded_label_tensor, verbose=False):
    """Reference contract (batch must be 1, inference.py:123)."""
    if input_tensor.shape[0] != 1:
        raise RuntimeError("infer() follows the reference's B = 1 contract; use infer_batch for batches")
    out = infer_batch(model, input_tensor, encoded_label_tensor)
    input_tensor.requires_grad = True                      # reference :132
    input_tensor.grad = out["grad"]                        # reference: loss.backward() fills input.grad (:141-143)
    pred = out["pred"].squeeze(0).cpu().numpy()            # reference :137 (D2H sync)
    input_gradients = input_tensor.grad
    gradients_np = input_gradients.detach().cpu().squeeze(0).numpy()
    return pred, input_gradients, gradients_np


def compute_gradient_magnitude(input_gradients, denormalize=True, norm=False):
    """Reference :36-53 — float64 [h,w] magnitude of the (de-normalised) gradient, B = 1.  Kept as a host-side
    compatibility helper (numpy, like the reference); the hot loop uses the fused wc_sgg_update kernel instead."""
    g = input_gradients.squeeze(0).cpu().numpy()
    if denormalize:
        g = g * np.array([0.229, 0.224, 0.225])[:, None, None]
    mag = np.sqrt(np.sum(g ** 2, axis=0))
    if norm:
        mag = (mag - mag.min()) / (mag.max() - mag.min())
    return torch.from_numpy(mag).to(input_gradients.device)
