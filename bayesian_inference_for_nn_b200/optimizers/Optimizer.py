"""Optimizer — abstract driver: compile once, train loop, progress line.

Mirrors Pyesian/optimizers/Optimizer.py:14-164 for the part the hot path uses: the compile-once
guard with the same ``Exception("Model Already compiled")`` (:54-55), the generic ``train`` loop with
its argument-consistency errors, loss-file reset and periodic ``result().store`` (:94-137), and the
``\\r`` progress bar (:149-159).  W&B logging is optional here (imported only when asked for) —
the reference hard-imports wandb (:10-11).
"""
import math
import os
import shutil
from abc import ABC, abstractmethod


class Optimizer(ABC):
    def __init__(self):
        self._model_config = None
        self._hyperparameters = None
        self.__compiled = False
        self._dataset = None
        self._verbose = True

    @abstractmethod
    def step(self, save_document_path=None):
        pass

    def compile(self, hyperparameters, model_config: str, dataset, verbose=True, **kwargs):
        if self.__compiled:
            raise Exception("Model Already compiled")
        self.__compiled = True
        self._hyperparameters = hyperparameters
        self._model_config = model_config
        self._dataset = dataset
        self._verbose = verbose
        self.compile_extra_components(**kwargs)

    @abstractmethod
    def compile_extra_components(self, **kwargs):
        pass

    def update_parameters_step(self):
        pass

    def _hp(self, name, default=None):
        """optional hyper-parameter (new knobs must not break existing scripts)."""
        try:
            return getattr(self._hyperparameters, name)
        except AttributeError:
            return default

    def _empty_folder(self, path):
        for name in os.listdir(path):
            p = os.path.join(path, name)
            try:
                if os.path.isfile(p) or os.path.islink(p):
                    os.unlink(p)
                elif os.path.isdir(p):
                    shutil.rmtree(p)
            except Exception as e:
                print("Failed to delete %s. Reason: %s" % (p, e))

    def train_with_weights_and_biases(self, nb_iterations, project_name, weights_and_biases_config):
        import wandb
        wandb.login()
        wandb.init(project=project_name, config=weights_and_biases_config)
        self.train(nb_iterations, weights_and_biases_log=True)

    def train(self, nb_iterations: int, loss_save_document_path: str = None, model_save_frequency: int = None,
              model_save_path: str = None, weights_and_biases_log=False):
        if model_save_frequency is None and model_save_path is not None:
            raise Exception("Error: save path precised and save frequency is None, please provide a savong frequency")
        if model_save_frequency is not None and model_save_path is None:
            raise Exception("Error: save frequency precised and save path is None, please provide a saving path")
        if loss_save_document_path is not None and os.path.exists(loss_save_document_path):
            os.remove(loss_save_document_path)
        if model_save_path is not None:
            self._empty_folder(model_save_path)
        saved = 0
        for i in range(nb_iterations):
            loss = self.step(loss_save_document_path)
            self._print_progress(i / nb_iterations, loss=loss)
            if weights_and_biases_log:
                import wandb
                wandb.log({"loss": loss})
            if model_save_frequency is not None and i % model_save_frequency == 0:
                target = os.path.join(model_save_path, "model" + str(saved))
                if os.path.exists(target):
                    shutil.rmtree(target)
                os.makedirs(target)
                res = self.result()
                getattr(res, "bayesian_model", res).store(target)
                saved += 1
        if self._verbose:
            print()

    @abstractmethod
    def result(self):
        pass

    def _print_progress(self, progress: float, bar_length=10, suffix="Training", **kwargs):
        if not self._verbose:
            return
        filled = math.ceil(progress * bar_length)
        bar = "[" + filled * "=" + (">" if filled < bar_length else "") + "]"
        infos = " ".join("{}: {}".format(k, v) for k, v in kwargs.items())
        print("\r" + suffix + " " + str(math.ceil(progress * 100)) + " % " + bar + " " + infos, end="")

    def _new_progress_line(self):
        if self._verbose:
            print()
