from .BayesianModel import BayesianModel, ParticleModel

__all__ = ["BayesianModel", "ParticleModel"]
