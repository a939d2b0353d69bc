from .seed import set_seed
from .metrics import compute_metrics, compute_spatial_metrics, print_metrics
from .ema import ModelEMA

__all__ = ["set_seed", "compute_metrics", "compute_spatial_metrics", "ModelEMA", "print_metrics"]
