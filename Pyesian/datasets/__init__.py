from bayesian_inference_for_nn_b200.datasets import Dataset, ArrayDataset  # noqa: F401
