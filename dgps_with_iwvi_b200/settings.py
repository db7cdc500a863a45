"""Numerics settings of the path (the reference takes them from gpflow.settings: float64, jitter 1e-6)."""
import torch

float_type = torch.float64
jitter = 1e-6   # gpflow.settings.numerics.jitter_level, used at reference temp_workaround.py:39


def device():
    """Parameters live on the GPU; on a machine without one only host-side logic (shapes, flattening, sharding)
    can run and every compute entry point raises."""
    return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
