from .Optimizer import Optimizer
from .HMC import HMC
from .SVGD import SVGD, SVGDResult
from .SGLD import SGLD
from .SWAG import SWAG

__all__ = ["Optimizer", "HMC", "SVGD", "SVGDResult", "SGLD", "SWAG"]
