"""Server construction helpers: reference ``machin.frame.helpers.servers`` surface (servers.py:7-165).

Collective: every rank of the torch.distributed job enters (outside torch.distributed every group is a group of one).
``members`` are ranks -- or ``"all"`` -- instead of RPC process names; the first member is the primary, which owns the
managed copy of each model and its optimizer (servers.py:88-131).
"""
from typing import Any, Callable, Dict, List, Tuple, Union

import torch as t
from torch.optim import Adam


def _group(name: str, members, device):
    import torch.distributed as dist
    from ...parallel import LocalGroup, ShardGroup
    if not (dist.is_available() and dist.is_initialized()):
        return LocalGroup(name, device=device)
    return ShardGroup(name, ranks=None if members == "all" else [int(m) for m in members], device=device)


def _default_device():
    return t.device("cuda", t.cuda.current_device())


def model_server_helper(model_num: int, group_name: str = "model_server", members: Union[str, List[int]] = "all",
                        device=None) -> Tuple:
    """``model_num`` push/pull weight servers named ``model_server_<i>`` over one group (servers.py:112-165)."""
    from ...parallel.server import PushPullModelServer
    group = _group(group_name, members, device if device is not None else _default_device())
    return tuple(PushPullModelServer(f"model_server_{i}", group, 0) for i in range(model_num))


def grad_server_helper(model_creators: List[Callable], group_name: str = "grad_server",
                       members: Union[str, List[int]] = "all", optimizer: Any = Adam,
                       learning_rate: Union[float, List[float]] = 1e-3, optimizer_kwargs: List[Dict[str, Any]] = None,
                       lr_scheduler: Any = None, lr_scheduler_args: List[Tuple] = None,
                       lr_scheduler_kwargs: List[Dict[str, Any]] = None, device=None) -> Tuple:
    """One gradient-reduction server ``grad_server_<i>`` per model creator (servers.py:7-109).  Only the primary member
    calls the creators: it holds the managed model, its optimizer and, if given, its learning-rate scheduler."""
    from ...parallel.server import PushPullGradServer
    device = device if device is not None else _default_device()
    group = _group(group_name, members, device)
    n = len(model_creators)
    rates = [learning_rate] * n if isinstance(learning_rate, float) else list(learning_rate)
    optimizer_kwargs = optimizer_kwargs or [{}] * n
    lr_scheduler_args = lr_scheduler_args or [()] * n
    lr_scheduler_kwargs = lr_scheduler_kwargs or [{}] * n
    servers = tuple(PushPullGradServer(f"grad_server_{i}", group, 0) for i in range(n))
    if group.index == 0:
        for i, (server, create) in enumerate(zip(servers, model_creators)):
            model = create().to(device)
            opt = optimizer(model.parameters(), lr=rates[i], **optimizer_kwargs[i])
            scheduler = lr_scheduler(opt, *lr_scheduler_args[i], **lr_scheduler_kwargs[i]) if lr_scheduler is not None else None
            server.manage_model(model, opt, scheduler)
    return servers
