"""Config boundary of the diffusion package (mirrors the field names of the reference's pydantic models,
diffusion_model/config/models.py:5-66, so the same config.yaml loads unchanged)."""
from typing import List, Optional

import yaml
from pydantic import BaseModel


class DataConfig(BaseModel):
    root_dir: str = "data"
    acdc_dir: str = "ACDC"
    acdc_labels: str = "ACDC/gt"
    acdc_images: str = "ACDC/rgb_anon"
    bdd_dir: str = "BDD"
    dawn_dir: str = "DAWN"
    weather: List[str] = ["fog", "rain"]
    image_size: List[int] = [128, 128]


class DiffusionConfig(BaseModel):
    num_timesteps: int = 1000
    beta_start: float = 0.0001
    beta_end: float = 0.02


class ModelConfig(BaseModel):
    name: str = "ddpm"
    im_channels: int = 3
    im_size: int = 128
    down_channels: List[int] = [64, 128, 256, 512, 768]
    mid_channels: List[int] = [768, 768, 512]
    down_sample: List[bool] = [True, True, True, False]
    time_emb_dim: int = 128
    num_down_layers: int = 2
    num_mid_layers: int = 2
    num_up_layers: int = 2
    num_heads: int = 4
    attn_resolutions: List[int] = [8, 16, 32, 64]


class FolderConfig(BaseModel):
    output: str = "diffusion_model/outputs"
    weights: str = "diffusion_model/weights"
    logs: str = "diffusion_model/logs"
    checkpoints: str = "diffusion_model/outputs/checkpoints"
    samples: str = "diffusion_model/outputs/samples"


class TrainingConfig(BaseModel):
    device: str = "cuda"
    random_seed: int = 3455
    epochs: int = 200
    batch_size: int = 4
    num_workers: int = 0
    lr: float = 0.0001
    log_interval: int = 10
    save_interval: int = 10
    sample_interval: int = 1000000000
    resume_training: bool = False
    resume_checkpoint: Optional[str] = None
    sample_size: int = 8
    num_grid_rows: int = 4


class Config(BaseModel):
    training: TrainingConfig = TrainingConfig()
    diffusion: DiffusionConfig = DiffusionConfig()
    data: DataConfig = DataConfig()
    model: ModelConfig = ModelConfig()
    folders: FolderConfig = FolderConfig()


def load_config(config_path: str) -> Config:
    with open(config_path, "r") as f:
        return Config(**yaml.safe_load(f))
