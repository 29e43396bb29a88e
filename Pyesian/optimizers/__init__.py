from bayesian_inference_for_nn_b200.optimizers import *  # noqa: F401,F403
from bayesian_inference_for_nn_b200.optimizers import Optimizer, HMC, SVGD, SGLD, SWAG
