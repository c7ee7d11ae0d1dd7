"""Geo math of the path, restated from reference preprocessing/geo_utils.py and preprocessing/utils.py."""
from __future__ import annotations

import torch

RAD_M = torch.tensor(6378137.0, dtype=torch.float64)  # geo_utils.py:10  (Earth radius in metres)
LABEL_SMOOTHING_CONSTANT = 65                          # config.py:55


def haversine(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """geo_utils.py:40-55 — (lng, lat) degrees, [n,2] x [n,2] -> km [n]. dtype promotion left to torch, as there."""
    x_rad, y_rad = torch.deg2rad(x), torch.deg2rad(y)
    delta = y_rad - x_rad
    a = torch.sin(delta[:, 1] / 2) ** 2 + torch.cos(x_rad[:, 1]) * torch.cos(y_rad[:, 1]) * torch.sin(delta[:, 0] / 2) ** 2
    c = 2 * torch.arcsin(torch.sqrt(a))
    return (RAD_M * c) / 1000


def haversine_matrix(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """geo_utils.py:58-74 — x [N,2], y [2,M] -> km [N,M]."""
    x_rad, y_rad = torch.deg2rad(x), torch.deg2rad(y)
    delta = x_rad.unsqueeze(2) - y_rad
    p = torch.cos(x_rad[:, 1]).unsqueeze(1) * torch.cos(y_rad[1, :]).unsqueeze(0)
    a = torch.sin(delta[:, 1, :] / 2) ** 2 + p * torch.sin(delta[:, 0, :] / 2) ** 2
    c = 2 * torch.arcsin(torch.sqrt(a))
    return (RAD_M * c) / 1000


def smooth_labels(distances: torch.Tensor) -> torch.Tensor:
    """preprocessing/utils.py:7-19."""
    adj = distances - distances.min(dim=-1, keepdim=True)[0]
    sm = torch.exp(-adj / LABEL_SMOOTHING_CONSTANT)
    return torch.nan_to_num(sm, nan=0.0, posinf=0.0, neginf=0.0)
