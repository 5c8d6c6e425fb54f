"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

SURVEY.md section 8c: stub matplotlib + seg_model.utils.visualizer, chdir to the reference root
(train_ddpm.py:26 loads its YAML by relative path at import time), disable wandb, and make
Tensor.cuda the identity (unet_base.py:461 hard-codes .cuda()).  Never used on the GPU box.
"""
import os
import sys
import types

REF = os.environ.get("WC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "diffusion_model"))


def install():
    import torch
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    viz = types.ModuleType("seg_model.utils.visualizer")
    viz.Visualizer = object
    sys.modules.setdefault("seg_model.utils.visualizer", viz)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.chdir(REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    return REF
