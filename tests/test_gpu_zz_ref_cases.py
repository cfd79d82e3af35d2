"""GPU: the reference's own algorithm test cases (act, store, update with every flag combination, save / load, lr
scheduler; int32 actions and integer rewards) through the product's classes on the device.  Same case functions as the
CPU dry runs in tests/test_algorithm_ref_cases_cpu.py, here with real buffers and kernels."""
import pytest


from test_algorithm_ref_cases_cpu import ddpg_cases, value_based_cases

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("kind", ["DQNPer", "DQN", "RAINBOW"])
def test_value_based_reference_cases_on_device(kind, tmp_path):
    value_based_cases(kind, tmp_path, "cuda")


def test_ddpg_per_reference_cases_on_device(tmp_path):
    ddpg_cases(tmp_path, "cuda")
