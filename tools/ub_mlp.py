import os, sys, json
import torch as t
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200 import _lib
from machin_b200._lib import lib
from machin_b200.model import _Workspace, FusedMLP
from tools.microbench import graphed
dev = t.device("cuda", 0)
ws = _Workspace(dev)
out = {}
for (M, N, K) in ((512, 256, 84), (512, 256, 256), (512, 6, 256)):
    x, w, b = t.randn(M, K, device=dev), t.randn(N, K, device=dev), t.randn(N, device=dev)
    y = t.empty(M, N, device=dev); dy = t.randn(M, N, device=dev); dx = t.empty(M, K, device=dev)
    dw = t.empty(N, K, device=dev); db = t.empty(N, device=dev)
    f = lambda: _lib.check(lib.mb_mlp_linear_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), M, N, K, 1, *ws.args(), _lib.stream_ptr()))
    f0 = lambda: _lib.check(lib.mb_mlp_linear_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), M, N, K, 1, None, 0, None, 0, _lib.stream_ptr()))
    bi = lambda: _lib.check(lib.mb_mlp_linear_backward_input(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(x), _lib.ptr(dx), M, N, K, *ws.args(), _lib.stream_ptr()))
    bw = lambda: _lib.check(lib.mb_mlp_linear_backward_weight(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(db), M, N, K, *ws.args(), _lib.stream_ptr()))
    tl = lambda: t.relu(t.nn.functional.linear(x, w, b))
    out[f"{M}x{N}x{K}"] = {"fwd_us": graphed([f] * 16) * 1e6, "fwd_nosplit_us": graphed([f0] * 16) * 1e6,
                           "bwd_input_us": graphed([bi] * 16) * 1e6, "bwd_weight_us": graphed([bw] * 16) * 1e6,
                           "torch_linear_relu_us": graphed([tl] * 16) * 1e6}
net = FusedMLP([84, 256, 256, 6]).to(dev)
xs = t.randn(512, 84, device=dev)
def fb():
    y = net(xs); y.backward(t.ones_like(y))
out["fused_mlp_fwd_bwd_us"] = graphed([fb] * 4) * 1e6
ref = t.nn.Sequential(t.nn.Linear(84, 256), t.nn.ReLU(), t.nn.Linear(256, 256), t.nn.ReLU(), t.nn.Linear(256, 6)).to(dev)
def fb2():
    y = ref(xs); y.backward(t.ones_like(y))
out["torch_mlp_fwd_bwd_us"] = graphed([fb2] * 4) * 1e6
print(json.dumps(out, indent=1))
