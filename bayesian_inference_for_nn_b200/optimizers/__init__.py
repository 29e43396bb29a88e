from .Optimizer import Optimizer
from .HMC import HMC
from .SVGD import SVGD, SVGDResult

__all__ = ["Optimizer", "HMC", "SVGD", "SVGDResult"]
