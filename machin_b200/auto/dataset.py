"""Episode datasets for the launcher (surface of machin/auto/dataset.py:14-108).

A dataset is an endless iterator of :class:`DatasetResult`: one episode's transitions (the dictionaries
``store_episode`` takes) plus the log entries that episode produced.  A log value is a scalar, or a pair
``(value, handler)`` whose handler is called as ``handler(launcher, key, value)`` (dataset.py:79, launcher.py:96-102).
Image / video handlers (dataset.py:38-73) belong to the media utilities, which are outside the learner path."""
from typing import Any, Callable, Dict, Iterator, List, Tuple, Union

import torch as t
from torch.utils.data import IterableDataset

Scalar = Any


def determine_precision(models) -> t.dtype:
    """The one parameter dtype of ``models`` -- observations are cast to it before acting (dataset.py:14-28)."""
    kinds = []
    for model in models:
        for _, param in model.named_parameters():
            if param.dtype not in kinds:
                kinds.append(param.dtype)
    if len(kinds) > 1:
        raise RuntimeError(f"Multiple data types of parameters detected in models: {kinds}, this is currently not supported "
                           "since we need to determine the data type of your model input from your model parameter data type.")
    if not kinds:
        raise RuntimeError("Models have no parameters to take a data type from.")
    return kinds[0]


class DatasetResult:
    def __init__(self, observations: List[Dict[str, Any]] = None,
                 logs: List[Dict[str, Union[Scalar, Tuple[Scalar, Callable]]]] = None):
        self.observations = observations if observations is not None else []
        self.logs = logs if logs is not None else []

    def add_observation(self, obs: Dict[str, Any]):
        self.observations.append(obs)

    def add_log(self, log: Dict[str, Union[Scalar, Tuple[Any, Callable]]]):
        self.logs.append(log)

    def __len__(self):
        return len(self.observations)


class RLDataset(IterableDataset):
    """Base of the episode datasets; ``early_stopping_monitor`` names the log key the trainer watches."""
    early_stopping_monitor = ""

    def __init__(self, **_kwargs):
        super().__init__()

    def __iter__(self) -> Iterator:
        return self

    def __next__(self) -> DatasetResult:
        raise StopIteration()
