"""Data I/O.  Only `load_kaust_csv_single` is on the ST-DADK training path (upstream
scripts/train_st_interp.py:33, :2187); the legacy forecasting-window API of the upstream package is not part
of this hot-path drop-in."""
from .kaust_loader import load_kaust_csv_single, ObservationTable

__all__ = ["load_kaust_csv_single", "ObservationTable"]
