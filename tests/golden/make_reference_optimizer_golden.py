#!/usr/bin/env python
"""Observable behaviour of the reference's Optimizer base class (Pyesian/optimizers/Optimizer.py: compile-once guard,
train's argument checks, loss-file reset, model folder handling and periodic result().store, progress line), recorded by
driving THE REFERENCE's class with a trivial subclass in the build container (tensorflow / wandb stubbed for import only).

    python -B tests/golden/make_reference_optimizer_golden.py      # writes tests/golden/reference_optimizer.json
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import warnings
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))

SCENARIOS = [dict(n=5, freq=None, path=False, loss_file=False), dict(n=5, freq=2, path=True, loss_file=True),
             dict(n=4, freq=1, path=True, loss_file=False), dict(n=3, freq=None, path=True, loss_file=False),
             dict(n=3, freq=2, path=False, loss_file=False), dict(n=7, freq=3, path=True, loss_file=True),
             dict(n=0, freq=1, path=True, loss_file=False)]


def drive(Optimizer, sc, verbose):
    """shared by the generator (reference class) and the test (this repo's class)"""
    class Store:
        def __init__(self, tag):
            self.tag = tag

        def store(self, path):
            with open(os.path.join(path, "stored.txt"), "w") as f:
                f.write(self.tag)

    class Toy(Optimizer):
        def __init__(self):
            super().__init__()
            self.calls = 0

        def step(self, save_document_path=None):
            self.calls += 1
            if save_document_path is not None:
                with open(save_document_path, "a") as f:
                    f.write("%d\n" % self.calls)
            return 1.0 / self.calls

        def compile_extra_components(self, **kwargs):
            self.kw = sorted(kwargs)

        def update_parameters_step(self):
            pass

        def result(self):
            return Store("after %d" % self.calls)

    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        opt = Toy()
        opt.compile("hyper", "{}", "dataset", verbose=verbose, alpha=1, beta=2)
        rec["kwargs_seen"] = opt.kw
        try:
            opt.compile("hyper", "{}", "dataset")
            rec["second_compile"] = None
        except Exception as e:
            rec["second_compile"] = [type(e).__name__, str(e)]
        model_dir = os.path.join(tmp, "models")
        os.makedirs(os.path.join(model_dir, "stale", "deep"))
        open(os.path.join(model_dir, "stale.txt"), "w").close()
        loss_file = os.path.join(tmp, "loss.txt")
        with open(loss_file, "w") as f:
            f.write("old\n")
        out = io.StringIO()
        try:
            with contextlib.redirect_stdout(out):
                opt.train(sc["n"], loss_file if sc["loss_file"] else None, sc["freq"], model_dir if sc["path"] else None)
            rec["error"] = None
        except Exception as e:
            rec["error"] = [type(e).__name__, str(e)]
        rec["calls"] = opt.calls
        rec["models"] = {d: open(os.path.join(model_dir, d, "stored.txt")).read()
                         for d in sorted(os.listdir(model_dir)) if os.path.exists(os.path.join(model_dir, d, "stored.txt"))}
        rec["leftovers"] = sorted(d for d in os.listdir(model_dir) if d.startswith("stale"))
        rec["loss_file"] = open(loss_file).read()
        rec["stdout"] = out.getvalue()
    return rec


def main():
    sys.dont_write_bytecode = True
    warnings.simplefilter("ignore")
    for name in ["tensorflow", "tensorflow_probability", "wandb", "wandb.integration", "wandb.integration.keras",
                 "tensorflow_datasets", "ucimlrepo", "matplotlib", "matplotlib.pyplot", "scikitplot"]:
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, "/root/reference")
    import Pyesian.optimizers  # noqa: F401
    Optimizer = sys.modules["Pyesian.optimizers.Optimizer"].Optimizer
    out = [dict(scenario=sc, verbose=v, observed=drive(Optimizer, sc, v)) for sc in SCENARIOS for v in (False, True)]
    with open(os.path.join(HERE, "reference_optimizer.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(out), "runs;", sum(o["observed"]["error"] is not None for o in out), "raised")


if __name__ == "__main__":
    main()
