from bayesian_inference_for_nn_b200.distributions import (Distribution, GaussianPrior, Sampled, Normal,  # noqa: F401
                                                          TensorflowProbabilityDistribution,
                                                          MultivariateNormalDiagPlusLowRank, Mixture)
