"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (iffiX/machin).

Only the golden-vector generator scripts under ``oracle/`` (run in the build
container, where ``/root/reference`` is mounted read-only) use this module.  Nothing
under ``machin_b200/`` may import it; ``/root/reference`` does not exist on the GPU box.

Recipe follows SURVEY.md Appendix A: (1) the reference's optional dependencies that
are absent from this image are replaced by ``MagicMock`` modules, (2) four
``torch.distributed.*_multigpu`` symbols removed from torch>=2 are stubbed, because
``machin/parallel/distributed/_world.py:546-570`` touches them in decorators at import.
"""
import os
import sys
from unittest.mock import MagicMock

# the mounted reference in the build container; on the GPU box the copy oracle/vendor_ref.py made (oracle/_ref)
_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("MACHIN_REFERENCE") or (
    "/root/reference" if os.path.isdir("/root/reference/machin") else _VENDORED)

_OPTIONAL = [
    "gym", "gym.spaces", "gym.wrappers", "gym.wrappers.time_limit", "GPUtil",
    "colorlog", "matplotlib", "matplotlib.pyplot", "moviepy", "moviepy.editor",
    "tensorboardX", "torchviz", "pytorch_lightning",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "machin"))


def load():
    """Import and return the reference ``machin`` package (raises if not mounted)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for name in _OPTIONAL:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    import torch.distributed as dist

    for name in ("broadcast_multigpu", "all_reduce_multigpu", "reduce_multigpu",
                 "all_gather_multigpu"):
        if not hasattr(dist, name):
            setattr(dist, name, lambda *a, **k: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import machin  # noqa: F401

    return machin
