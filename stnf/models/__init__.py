from .st_interp import STInterpMLP, create_model

__all__ = ["STInterpMLP", "create_model"]
