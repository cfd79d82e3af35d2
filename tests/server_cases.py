"""Weight / gradient server cases shared by the gloo (CPU) and NCCL (GPU) tests: the reference's goldens
(test/parallel/server/test_param_server.py:30-173) plus the version-chain cases of ordered_server.py:112-128.
Written for any world >= 2; the reference's own tests use 3 processes (server + 2 workers)."""
import torch as t


class Lin(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = t.nn.Linear(1, 1, bias=False)
        self.fc2 = t.nn.Linear(1, 1, bias=False)
        self.fc3 = t.nn.Linear(1, 1, bias=False)
        with t.no_grad():
            self.fc1.weight.fill_(1)
            self.fc2.weight.fill_(2)
            self.fc3.weight.fill_(3)

    def forward(self, x):
        return self.fc3(self.fc2(self.fc1(x)))


def _w(model):
    return [model.fc1.weight.item(), model.fc2.weight.item(), model.fc3.weight.item()]


def model_server_case(rank, world, device="cpu"):
    from machin_b200.parallel.group import ShardGroup
    from machin_b200.parallel.server import PushPullModelServer
    group = ShardGroup("group")
    server = PushPullModelServer("model", group, src_index=0)
    model = Lin().to(device)
    assert server.pull(model) is False            # nothing pushed yet (param_server.py:84-86)
    for i in range(10):                           # test_param_server.py:92-101: weights +1 per push -> 11,12,13
        if rank == 0:
            with t.no_grad():
                for p in model.parameters():
                    p += 1
            assert server.push(model)
        else:
            assert server.pull(model)
    assert _w(model) == [11, 12, 13]
    assert model.pp_version == 10
    # a pull by a member that already holds the newest version loads nothing (param_server.py:87-90), returns True
    if rank != 0:
        with t.no_grad():
            model.fc1.weight.fill_(-7)
    assert server.pull(model) is True
    assert model.fc1.weight.item() == (11 if rank == 0 else -7)
    # a push on a stale chain is rejected and the pusher receives the newest parameters instead
    # (ordered_server.py:116-121, param_server.py:49-63): rank 1 pretends to be at version 3
    if rank == 1:
        model.pp_version = 3
        with t.no_grad():
            model.fc2.weight.fill_(99)
        assert server.push(model) is False
        assert model.pp_version == 10 and _w(model) == [11, 12, 13]
    else:
        assert server.pull(model) is True and model.pp_version == 10
    # two pushers on the same version in one round: the first in member order extends the chain, the other is
    # rejected and pulls what the first one pushed
    first, second = (1, 2) if world > 2 else (0, 1)
    with t.no_grad():
        model.fc3.weight.fill_(100.0 + rank)
        model.fc1.weight.fill_(11.0)
    ok = server.push(model) if rank in (first, second) else server.pull(model)
    assert ok is (rank != second)
    assert model.pp_version == 11 and model.fc3.weight.item() == 100.0 + first
    # pull_on_fail=False leaves the rejected pusher's model alone
    if rank == second:
        model.pp_version = 4
        with t.no_grad():
            model.fc3.weight.fill_(-1.0)
        assert server.push(model, pull_on_fail=False) is False
        assert model.pp_version == 4 and model.fc3.weight.item() == -1.0
    else:
        assert server.pull(model) is True
    return True


def grad_expect(method, world, pushes=2):
    """test_param_server.py:108-173 (world 3): member 0 serves, the other world-1 members push `pushes` gradients
    (6, 3, 2) each; SGD with lr 1.  mean: w - g; sum: w - (world-1) * pushes * g."""
    k = 1 if method == "mean" else (world - 1) * pushes
    return (1 - 6 * k, 2 - 3 * k, 3 - 2 * k)


def grad_server_case(rank, world, method, device="cpu"):
    from machin_b200.parallel.group import ShardGroup
    from machin_b200.parallel.server import PushPullGradServer
    group = ShardGroup("group")
    server = PushPullGradServer("server", group, primary_index=0, reduce_method=method, pushes_per_reduce=2)
    model = Lin().to(device)
    if rank == 0:
        server.manage_model(model, t.optim.SGD(model.parameters(), lr=1.0))
    for i in range(2):
        if rank == 0:
            server.push(None)                 # the server process itself has no gradients
        else:
            model.zero_grad()
            model(t.ones([1, 1], device=device)).backward()  # grads 6, 3, 2
            server.push(model)
    assert server.pull(model) is True
    got = tuple(_w(model))
    assert got == grad_expect(method, world), got
    return True
