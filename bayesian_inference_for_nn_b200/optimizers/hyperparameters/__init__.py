from .HyperParameters import HyperParameters

__all__ = ["HyperParameters"]
