"""Model registry — same surface as the reference's ``src/models/models.py:6-18``."""
from typing import Dict


def get_model(model_name: str, config: Dict, device: str):
    if model_name == "lcnn":
        from .lcnn import LCNN

        return LCNN(device=device, **config)
    elif model_name == "specrnet":
        from .specrnet import SpecRNet, get_config

        return SpecRNet(get_config(config.get("input_channels", 1)), device=device, **config)
    elif model_name == "rawnet3":
        from .rawnet3 import prepare_model

        return prepare_model()
    else:
        raise ValueError(f"Model '{model_name}' not supported")
