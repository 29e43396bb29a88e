from bayesian_inference_for_nn_b200.visualisations import Metrics, Plotter  # noqa: F401
