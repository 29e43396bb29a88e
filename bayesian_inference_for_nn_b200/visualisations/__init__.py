from .Metrics import Metrics
from .Plotter import Plotter

__all__ = ["Metrics", "Plotter"]
