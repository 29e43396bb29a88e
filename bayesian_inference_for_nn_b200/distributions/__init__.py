from .Distribution import Distribution
from .GaussianPrior import GaussianPrior
from .Sampled import Sampled

__all__ = ["Distribution", "GaussianPrior", "Sampled"]
