from .Distribution import Distribution
from .GaussianPrior import GaussianPrior
from .Sampled import Sampled
from .Normal import Normal, TensorflowProbabilityDistribution
from .MultivariateNormalDiagPlusLowRank import MultivariateNormalDiagPlusLowRank
from .Mixture import Mixture

__all__ = ["Distribution", "GaussianPrior", "Sampled", "Normal", "TensorflowProbabilityDistribution",
           "MultivariateNormalDiagPlusLowRank", "Mixture"]
