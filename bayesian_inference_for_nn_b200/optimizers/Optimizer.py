"""Optimizer — the driver every inference method shares: compile once, iterate ``step``, report progress.

Behavioural mirror of Pyesian/optimizers/Optimizer.py:14-164 for the part the hot path uses.  What callers can observe
is kept: ``compile`` works once (``Exception("Model Already compiled")``, :54-55) and hands the keyword arguments to
``compile_extra_components``; ``train`` rejects a save path without a save frequency and vice versa with the reference's
messages (:109-112), starts the loss file afresh (:114-115), empties the model folder (:117-118), stores
``result()`` under ``model<k>`` every ``model_save_frequency`` iterations (:127-132) and prints the ``\\r`` progress line
(:149-159).  Weights & Biases logging is imported only when it is asked for (the reference imports wandb at module
load, :10-11).
"""
import math
import os
import shutil
from abc import ABC, abstractmethod

_MSG_NO_FREQUENCY = "Error: save path precised and save frequency is None, please provide a savong frequency"
_MSG_NO_PATH = "Error: save frequency precised and save path is None, please provide a saving path"


def _clear_directory(path):
    """remove everything inside ``path`` (files, links, sub-folders); failures are reported and skipped"""
    for entry in os.scandir(path):
        try:
            if entry.is_dir(follow_symlinks=False):
                shutil.rmtree(entry.path)
            else:
                os.unlink(entry.path)
        except Exception as e:
            print("Failed to delete %s. Reason: %s" % (entry.path, e))


class _Checkpointer:
    """``model<k>`` folders under ``root``, one per ``every`` iterations (disabled when ``every`` is None)"""

    def __init__(self, root, every):
        if root is not None and every is None:
            raise Exception(_MSG_NO_FREQUENCY)
        if every is not None and root is None:
            raise Exception(_MSG_NO_PATH)
        self.root, self.every, self.count = root, every, 0

    def prepare(self):
        if self.root is not None:
            _clear_directory(self.root)

    def maybe_store(self, iteration, make_result):
        if self.every is None or iteration % self.every:
            return
        target = os.path.join(self.root, "model%d" % self.count)
        shutil.rmtree(target, ignore_errors=True)
        os.makedirs(target)
        res = make_result()
        getattr(res, "bayesian_model", res).store(target)
        self.count += 1


class Optimizer(ABC):
    def __init__(self):
        self._hyperparameters = self._model_config = self._dataset = None
        self._verbose = True
        self.__compiled = False

    # ---- to be provided by an inference method ---------------------------------------------------------------
    @abstractmethod
    def step(self, save_document_path=None):
        """one iteration; returns its loss"""

    @abstractmethod
    def compile_extra_components(self, **kwargs):
        """method-specific part of ``compile``"""

    @abstractmethod
    def result(self):
        """the posterior at this stage of the training"""

    def update_parameters_step(self):
        pass

    # ---- shared driver -------------------------------------------------------------------------------------------
    def compile(self, hyperparameters, model_config: str, dataset, verbose=True, **kwargs):
        if self.__compiled:
            raise Exception("Model Already compiled")
        self.__compiled = True
        self._hyperparameters, self._model_config, self._dataset, self._verbose = (
            hyperparameters, model_config, dataset, verbose)
        self.compile_extra_components(**kwargs)

    def _hp(self, name, default=None):
        """a hyper-parameter that may be absent (knobs this build adds must not break existing scripts)"""
        try:
            return getattr(self._hyperparameters, name)
        except AttributeError:
            return default

    def _empty_folder(self, path):
        _clear_directory(path)

    def train(self, nb_iterations: int, loss_save_document_path: str = None, model_save_frequency: int = None,
              model_save_path: str = None, weights_and_biases_log=False):
        saver = _Checkpointer(model_save_path, model_save_frequency)
        if loss_save_document_path is not None and os.path.exists(loss_save_document_path):
            os.remove(loss_save_document_path)
        saver.prepare()
        log = None
        if weights_and_biases_log:
            import wandb
            log = wandb.log
        for it in range(nb_iterations):
            loss = self.step(loss_save_document_path)
            self._print_progress(it / nb_iterations, loss=loss)
            if log is not None:
                log({"loss": loss})
            saver.maybe_store(it, self.result)
        self._new_progress_line()

    def train_with_weights_and_biases(self, nb_iterations, project_name, weights_and_biases_config):
        import wandb
        wandb.login()
        wandb.init(project=project_name, config=weights_and_biases_config)
        self.train(nb_iterations, weights_and_biases_log=True)

    # ---- progress line: "\\rTraining 37 % [====>     ] loss: 0.42" ----------------------------------------------
    def _print_progress(self, progress: float, bar_length=10, suffix="Training", **kwargs):
        if not self._verbose:
            return
        done = math.ceil(progress * bar_length)
        arrow = ">" if done < bar_length else ""
        fields = " ".join("%s: %s" % kv for kv in kwargs.items())
        print("\r%s %d %% [%s%s] %s" % (suffix, math.ceil(progress * 100), "=" * done, arrow, fields), end="")

    def _new_progress_line(self):
        if self._verbose:
            print()
