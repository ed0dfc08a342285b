"""SpecRNet parameter holder.

Same constructor, parameter names and shapes as the reference's ``src/models/specrnet.py:23-214``
(``SpecRNet(get_config(input_channels), device=..., frontend_algorithm=[...])``), so reference checkpoints load with
``load_state_dict``.  The sub-modules are real ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.GRU`` / ``nn.Linear`` objects used
only as storage (``block2.0.bn1`` / ``block4.0.bn1`` exist in the checkpoint although the reference's forward discards
their output, specrnet.py:76-81); ``forward`` hands the waveform to the CUDA engine.
"""
from typing import Dict

import torch
from torch import nn

from .. import frontends


def get_config(input_channels: int) -> Dict:
    return {
        "filts": [input_channels, [input_channels, 20], [20, 64], [64, 64]],
        "nb_fc_node": 64,
        "gru_node": 64,
        "nb_gru_layer": 2,
        "nb_classes": 1,
    }


class Residual_block2D(nn.Module):
    def __init__(self, nb_filts, first=False):
        super().__init__()
        self.first = first
        if not self.first:
            self.bn1 = nn.BatchNorm2d(num_features=nb_filts[0])
        self.conv1 = nn.Conv2d(nb_filts[0], nb_filts[1], kernel_size=3, padding=1, stride=1)
        self.bn2 = nn.BatchNorm2d(num_features=nb_filts[1])
        self.conv2 = nn.Conv2d(nb_filts[1], nb_filts[1], kernel_size=3, padding=1, stride=1)
        self.downsample = nb_filts[0] != nb_filts[1]
        if self.downsample:
            self.conv_downsample = nn.Conv2d(nb_filts[0], nb_filts[1], kernel_size=1, padding=0, stride=1)


class SpecRNet(nn.Module):
    def __init__(self, d_args, **kwargs):
        super().__init__()
        if d_args["filts"][0] != 1:
            raise ValueError("advb200 SpecRNet supports input_channels=1 (lfcc / mfcc frontends)")
        self.device = kwargs.get("device", "cuda")
        f = d_args["filts"]
        self.first_bn = nn.BatchNorm2d(num_features=f[0])
        self.block0 = nn.Sequential(Residual_block2D(nb_filts=f[1], first=True))
        self.block2 = nn.Sequential(Residual_block2D(nb_filts=f[2]))
        self.block4 = nn.Sequential(Residual_block2D(nb_filts=[f[2][1], f[2][1]]))
        self.fc_attention0 = nn.Sequential(nn.Linear(f[1][-1], f[1][-1]))
        self.fc_attention2 = nn.Sequential(nn.Linear(f[2][-1], f[2][-1]))
        self.fc_attention4 = nn.Sequential(nn.Linear(f[2][-1], f[2][-1]))
        self.bn_before_gru = nn.BatchNorm2d(num_features=f[2][-1])
        self.gru = nn.GRU(input_size=f[2][-1], hidden_size=d_args["gru_node"], num_layers=d_args["nb_gru_layer"],
                          batch_first=True, bidirectional=True)
        self.fc1_gru = nn.Linear(d_args["gru_node"] * 2, d_args["nb_fc_node"] * 2)
        self.fc2_gru = nn.Linear(d_args["nb_fc_node"] * 2, d_args["nb_classes"], bias=True)
        self.frontend = frontends.get_frontend(kwargs.get("frontend_algorithm", []))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .. import engine

        return engine.model_forward(self, x)
