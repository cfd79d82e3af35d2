"""The launcher loop that feeds the learner: ``store_episode`` then one ``update`` per stored transition
(machin/auto/launcher.py:61-66 for the local frameworks, :152-171 for the distributed ones).

The reference implements both launchers as pytorch-lightning modules and leaves the loop itself, early stopping,
checkpointing and the process world to ``pl.Trainer`` and a patched DDP plugin (launcher.py:13, :105; pl_plugin.py:34-69).
pytorch-lightning is not part of this stack: :class:`Trainer` below is that loop with the knobs ``launch`` sets
(``limit_train_batches``, ``max_steps``, callbacks), and the world is the ``torchrun`` job the process was started in.

What differs for the distributed launcher, and why: every member of a sharded framework is an SPMD rank and
its ``update`` is a collective (a non-learner serves the learners' draws inside it), so all ranks must call ``update``
the same number of times.  The reference lets each learner run ``len(episode)`` updates while samplers only store
(launcher.py:157-160); here the ranks agree on that count first -- the largest count any LEARNER asks for -- and every
rank runs it."""
import json
import os
from typing import Any, Callable, Dict, List

import torch as t
import torch.distributed as dist
from torch.utils.data.dataloader import DataLoader

from .config import init_algorithm_from_config, is_algorithm_distributed
from .dataset import RLDataset


def _world():
    on = dist.is_available() and dist.is_initialized()
    return (dist.get_rank(), dist.get_world_size()) if on else (0, 1)


def _as_is(batch):
    return batch


class _EpisodeModule:
    """What both launchers share: the datasets, the log table and the handler protocol of the log entries."""

    def __init__(self, config: Dict[str, Any], env_dataset_creator: Callable[[Any, Dict[str, Any]], RLDataset]):
        self.config, self.env_dataset_creator = config, env_dataset_creator
        self.frame, self.trainer = None, None
        self.current_epoch = 0
        self.logged: Dict[str, float] = {}            # newest value of every scalar log key
        self._batch_num = 1
        self._train_esm = self._test_esm = None       # early-stopping monitor keys of the two datasets

    # ---- datasets
    def _loader(self, which):
        dataset = self.env_dataset_creator(self.frame, self.config[which])
        return dataset, DataLoader(dataset=dataset, batch_size=self._batch_num, collate_fn=_as_is)

    def train_dataloader(self):
        dataset, loader = self._loader("train_env_config")
        self._train_esm = getattr(dataset, "early_stopping_monitor", None)
        return loader

    def test_dataloader(self):
        dataset, loader = self._loader("test_env_config")
        self._test_esm = getattr(dataset, "early_stopping_monitor", None)
        return loader

    # ---- logging
    def log(self, key: str, value, prog_bar: bool = False, sync_dist: bool = False):
        """Record a scalar; with ``sync_dist`` its mean over the ranks (collective)."""
        value = float(value.item() if t.is_tensor(value) else value)
        rank, world = _world()
        if sync_dist and world > 1:
            cell = t.tensor([value], dtype=t.float64, device=self._collective_device())
            dist.all_reduce(cell)
            value = cell.item() / world
        self.logged[key] = value
        if self.trainer is not None:
            self.trainer.record(key, value)

    def _collective_device(self):
        return t.device("cuda", t.cuda.current_device()) if dist.get_backend() == "nccl" else t.device("cpu")

    def _log(self, logs: List[Dict[str, Any]], sync_keys=()):
        in_job = dist.is_available() and dist.is_initialized()
        for entry in logs:
            for key, value in entry.items():
                if isinstance(value, tuple) and len(value) == 2 and callable(value[1]):
                    value[1](self, key, value[0])
                else:
                    # only the monitored key is reduced: other keys may be missing on other ranks (launcher.py:227-239)
                    self.log(key, value, prog_bar=True, sync_dist=bool(key) and key in sync_keys and in_job)

    def _updates_for(self, sample) -> int:
        """One update per stored transition (launcher.py:63-64).  An on-policy framework empties its buffer inside
        ``update``, so it gets one update per stored episode (the reference's loop asks it for a second update on an empty
        buffer, which fails at a2c.py:349 unpacking ``None``)."""
        return 1 if getattr(self.frame, "on_policy", False) else len(sample.observations)

    def test_step(self, batch, _batch_idx):
        for sample in batch:
            self._log(sample.logs)

    def on_test_start(self):
        self.on_train_start()


class Launcher(_EpisodeModule):
    """Launcher of the single-process frameworks (launcher.py:13-102)."""

    def __init__(self, config, env_dataset_creator, model_device="cuda"):
        super().__init__(config, env_dataset_creator)
        assert not is_algorithm_distributed(config)
        self.frame = init_algorithm_from_config(config, model_device=model_device)
        # config["cuda_graph"]: capture the framework's update as one CUDA graph once the first episode is stored (the
        # frameworks that offer enable_cuda_graph() without arguments: the DQN and DDPG families)
        self._graph_pending = bool(config.get("cuda_graph", False))

    def on_train_start(self):
        if _world()[1] > 1:
            raise RuntimeError(f"Current framework: {self.config['frame']} is not a distributed framework, you should "
                               "launch it as a single process.")

    def configure_optimizers(self):
        return self.frame.optimizers, self.frame.lr_schedulers

    def training_step(self, batch, _batch_idx):
        for sample in batch:
            self.frame.store_episode(sample.observations)
            if self._graph_pending:
                self.frame.enable_cuda_graph()
                self._graph_pending = False
            for _ in range(self._updates_for(sample)):
                self.frame.update()
            self._log(sample.logs)


class DistributedLauncher(_EpisodeModule):
    """Launcher of the sharded frameworks (launcher.py:105-239): one instance per rank of the ``torchrun`` job.  The
    framework is built lazily, by :meth:`init_frame`, because building it is a collective over the process group."""

    def __init__(self, config, env_dataset_creator, model_device="cuda"):
        super().__init__(config, env_dataset_creator)
        assert is_algorithm_distributed(config)
        self._model_device = model_device

    def init_frame(self):
        if self.frame is None:
            self.frame = init_algorithm_from_config(self.config, model_device=self._model_device)
            role = getattr(self.frame, "role", "learner")
            self._batch_num = int(self.config.get("batch_num", {}).get(role, 1))

    def on_train_start(self):
        self.init_frame()

    def train_dataloader(self):
        self.init_frame()
        return super().train_dataloader()

    def test_dataloader(self):
        self.init_frame()
        return super().test_dataloader()

    def configure_optimizers(self):
        return None                                   # the framework keeps its own optimizers (launcher.py:177-181)

    def _agreed_update_count(self, wanted: int) -> int:
        rank, world = _world()
        if world == 1:
            return wanted
        mine = wanted if getattr(self.frame, "role", "learner") == "learner" else 0
        cell = t.tensor([mine], dtype=t.int64, device=self._collective_device())
        dist.all_reduce(cell, op=dist.ReduceOp.MAX)
        return int(cell.item())

    def training_step(self, batch, batch_idx):
        wanted = 0
        for sample in batch:
            self.frame.store_episode(sample.observations)
            wanted += self._updates_for(sample)
        # the first batch only fills the buffers (launcher.py:157-158)
        first = self.current_epoch == 0 and batch_idx == 0
        for _ in range(self._agreed_update_count(0 if first else wanted)):
            self.frame.update()
        if len(batch) > 0:
            # only the first episode's logs: ranks hold different numbers of episodes per batch, and the monitored key is
            # reduced over the ranks (launcher.py:162-166)
            self._log(batch[0].logs, sync_keys=(self._train_esm, self._test_esm))
        if _world()[1] > 1:
            dist.barrier()                            # leave the step together

    def test_step(self, batch, _batch_idx):
        for sample in batch:
            self._log(sample.logs, sync_keys=(self._train_esm, self._test_esm))


# ------------------------------------------------------------------------------------------------ the loop and its callbacks
class EarlyStopping:
    """Stop when ``monitor`` has not improved for ``patience`` epochs in a row."""

    def __init__(self, monitor: str, mode: str = "max", patience: int = 3, min_delta: float = 0.0):
        assert mode in ("max", "min")
        self.monitor, self.sign, self.patience, self.min_delta = monitor, (1.0 if mode == "max" else -1.0), patience, min_delta
        self.best, self.stale = None, 0

    def on_epoch_end(self, trainer, module):
        value = trainer.callback_metrics.get(self.monitor)
        if value is None:
            return
        if self.best is None or self.sign * (value - self.best) > self.min_delta:
            self.best, self.stale = value, 0
        else:
            self.stale += 1
            if self.stale >= self.patience:
                trainer.should_stop = True


class ModelCheckpoint:
    """Keep the restorable models of the best epoch (by ``monitor``) as ``<name>_<epoch>.pt`` under ``dirpath``; written
    by rank 0 only, older bests removed."""

    def __init__(self, dirpath: str, monitor: str, mode: str = "max"):
        assert mode in ("max", "min")
        self.dirpath, self.monitor, self.sign = dirpath, monitor, (1.0 if mode == "max" else -1.0)
        self.best, self.best_version = None, None

    def on_epoch_end(self, trainer, module):
        value = trainer.callback_metrics.get(self.monitor)
        if value is None or (self.best is not None and self.sign * (value - self.best) <= 0.0):
            return
        previous, self.best, self.best_version = self.best_version, value, trainer.current_epoch
        if _world()[0] != 0:
            return
        module.frame.save(self.dirpath, version=trainer.current_epoch)
        if previous is not None:
            for name in module.frame.get_restorable_model_names():
                stale = os.path.join(self.dirpath, module.frame._file_name(name, previous))
                if os.path.exists(stale):
                    os.remove(stale)


class JsonLinesLogger:
    """One JSON object per training step with the scalars logged during it (rank 0 only)."""

    def __init__(self, path: str):
        self.path = path

    def on_step_end(self, trainer, module, metrics: Dict[str, float]):
        if _world()[0] != 0 or not metrics:
            return
        os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
        with open(self.path, "a") as out:
            out.write(json.dumps({"epoch": trainer.current_epoch, "step": trainer.global_step, **metrics}, sort_keys=True) + "\n")


class Trainer:
    """The episode loop.  An epoch is ``limit_train_batches`` batches of the (endless) training dataset; a batch is
    ``batch_num[role]`` episodes (one for the local frameworks); training ends after ``max_steps`` batches, ``max_epochs``
    epochs or when a callback sets ``should_stop``.  Callbacks may define ``on_step_end(trainer, module, metrics)`` and
    ``on_epoch_end(trainer, module)``."""

    def __init__(self, max_steps: int = None, max_epochs: int = None, limit_train_batches: int = None,
                 limit_test_batches: int = 1, callbacks: List[Any] = None):
        if max_steps is None and max_epochs is None:
            raise ValueError("Trainer needs max_steps or max_epochs: the episode datasets never end")
        if max_steps is None and limit_train_batches is None:
            raise ValueError("Trainer needs limit_train_batches with max_epochs: an epoch of an endless dataset never ends")
        self.max_steps, self.max_epochs = max_steps, max_epochs
        self.limit_train_batches, self.limit_test_batches = limit_train_batches, limit_test_batches
        self.callbacks = list(callbacks or [])
        self.current_epoch, self.global_step, self.should_stop = 0, 0, False
        self.callback_metrics: Dict[str, float] = {}
        self._step_metrics: Dict[str, float] = {}

    def record(self, key, value):
        self.callback_metrics[key] = value
        self._step_metrics[key] = value

    def _notify(self, hook, *args):
        for callback in self.callbacks:
            method = getattr(callback, hook, None)
            if callable(method):
                method(self, *args)

    def _done(self):
        return self.should_stop or (self.max_steps is not None and self.global_step >= self.max_steps) or \
            (self.max_epochs is not None and self.current_epoch >= self.max_epochs)

    def fit(self, module):
        module.trainer = self
        module.on_train_start()
        loader = iter(module.train_dataloader())
        while not self._done():
            module.current_epoch = self.current_epoch
            batch_idx = 0
            while self.limit_train_batches is None or batch_idx < self.limit_train_batches:
                try:
                    batch = next(loader)
                except StopIteration:
                    self.should_stop = True
                    break
                self._step_metrics = {}
                module.training_step(batch, batch_idx)
                self.global_step += 1
                batch_idx += 1
                self._notify("on_step_end", module, dict(self._step_metrics))
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    break
            self._notify("on_epoch_end", module)
            self.current_epoch += 1
        return self

    def test(self, module):
        module.trainer = self
        module.on_test_start()
        loader = iter(module.test_dataloader())
        for batch_idx in range(self.limit_test_batches):
            try:
                batch = next(loader)
            except StopIteration:
                break
            module.test_step(batch, batch_idx)
        return dict(self.callback_metrics)
