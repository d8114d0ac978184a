"""Grid helper kept for API compatibility (reference functions/functions.py:128-138).
The fused PointFlow kernel computes the pixel grid in registers and never calls this."""
import torch


def get_pixel_grids(height, width):
    """[3, H*W]: rows (x + 0.5, y + 0.5, 1), row-major over (y, x)."""
    with torch.no_grad():
        xs = torch.linspace(0.5, width - 0.5, width).view(1, width).expand(height, width)
        ys = torch.linspace(0.5, height - 0.5, height).view(height, 1).expand(height, width)
        return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width)], dim=0)
