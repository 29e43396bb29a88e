from bayesian_inference_for_nn_b200.optimizers.hyperparameters import HyperParameters  # noqa: F401
