"""SRGAN inference entry points with the reference's names (srgan_model/inference.py:9-39)."""
import torch

from .models import Generator

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def load_model(model_path: str) -> torch.nn.Module:
    netG = Generator(upscale_factor=4).to(device)
    checkpoint = torch.load(model_path, map_location=device)
    netG.load_state_dict(checkpoint['model'])       # reference checkpoints keep the weights under 'model' (:12-13)
    netG.eval()
    return netG


def inference(netG: torch.nn.Module, lr_image: torch.Tensor) -> torch.Tensor:
    with torch.no_grad():
        return netG(lr_image)
