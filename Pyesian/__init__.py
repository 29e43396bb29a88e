"""Import-path alias: `from Pyesian.optimizers import HMC` keeps working against the B200 build
(the reference's scripts import exactly these paths, e.g. HMC_classification.py:3-8)."""
