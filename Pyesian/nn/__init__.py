from bayesian_inference_for_nn_b200.nn import BayesianModel, ParticleModel  # noqa: F401
