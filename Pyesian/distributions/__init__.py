from bayesian_inference_for_nn_b200.distributions import Distribution, GaussianPrior, Sampled  # noqa: F401
