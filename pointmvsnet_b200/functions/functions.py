"""Helpers kept for API compatibility (reference functions/functions.py:128-175).
The fused PointFlow kernel computes the pixel grid in registers and never calls ``get_pixel_grids``;
``get_propability_map`` belongs to the coarse stage (model.py:127) and is here so that an unchanged ``model.py``
finds every name it imports."""
import torch


def get_pixel_grids(height, width):
    """[3, H*W]: rows (x + 0.5, y + 0.5, 1), row-major over (y, x)."""
    with torch.no_grad():
        xs = torch.linspace(0.5, width - 0.5, width).view(1, width).expand(height, width)
        ys = torch.linspace(0.5, height - 0.5, height).view(height, 1).expand(height, width)
        return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width)], dim=0)


def get_propability_map(cv, depth_map, depth_start, depth_interval):
    """functions/functions.py:141-175: probability of the two depth planes that bracket the regressed depth.
    cv [B,D,H,W] (soft-max over D), depth_map [B,1,H,W], depth_start / depth_interval [B] -> [B,1,H,W]."""
    with torch.no_grad():
        D = cv.size(1)
        d = ((depth_map - depth_start.view(-1, 1, 1, 1)) / depth_interval.view(-1, 1, 1, 1)).detach()
        lo = d.floor().clamp(0, D - 1).long()
        hi = d.ceil().clamp(0, D - 1).long()
        return cv.gather(1, lo) + cv.gather(1, hi)
