"""CPU: host layer of the group-fairness metrics against goldens from the reference, with the counting kernel replaced
by its torch stand-in (fixture `cpu_kernel_standins`); the same replay runs on the real kernel in
tests/test_zz_group_fairness_gpu.py."""
import pytest
import torch

from metrics_b200._native import NativeLibraryError
from metrics_b200.classification import BinaryFairness, BinaryGroupStatRates
from metrics_b200.functional.classification import binary_fairness, binary_groups_stat_rates
from tests.fairness_cases import replay


def test_replay_reference_goldens(golden_fairness, cpu_kernel_standins):
    assert replay(golden_fairness, "cpu") == 27


def test_docstring_examples_of_the_reference(cpu_kernel_standins):
    target = torch.tensor([0, 1, 0, 1, 0, 1])
    preds = torch.tensor([0.11, 0.84, 0.22, 0.73, 0.33, 0.92])
    groups = torch.tensor([0, 1, 0, 1, 0, 1])
    rates = binary_groups_stat_rates(preds, target, groups, 2)
    assert rates["group_0"].tolist() == [0.0, 0.0, 1.0, 0.0] and rates["group_1"].tolist() == [1.0, 0.0, 0.0, 0.0]
    out = binary_fairness(preds, target, groups)
    assert list(out) == ["DP_0_1", "EO_0_1"] and [float(v) for v in out.values()] == [0.0, 0.0]
    # group ids that are not 0..G-1: the functional works on the ids present, like the reference's sort-and-split
    sparse = binary_groups_stat_rates(torch.tensor([0.9, 0.2, 0.8, 0.1]), torch.tensor([1, 0, 0, 1]), torch.tensor([0, 2, 2, 0]), 3)
    assert sparse["group_0"].tolist() == [0.5, 0.0, 0.0, 0.5] and sparse["group_1"].tolist() == [0.0, 0.5, 0.5, 0.0]


def test_argument_validation(cpu_kernel_standins):
    preds, target, groups = torch.rand(8), torch.randint(2, (8,)), torch.randint(2, (8,))
    with pytest.raises(ValueError, match="Expected argument `task`"):
        binary_fairness(preds, target, groups, task="nope")
    with pytest.raises(ValueError, match="Expected argument `task`"):
        BinaryFairness(2, task="nope")
    with pytest.raises(ValueError, match="num_groups"):
        BinaryGroupStatRates(1)
    with pytest.raises(ValueError, match="dtype of argument groups to be long"):
        binary_groups_stat_rates(preds, target, groups.int(), 2)
    with pytest.raises(ValueError, match="largest number in the groups tensor"):
        binary_groups_stat_rates(preds, target, groups + 5, 2)
    with pytest.raises(ValueError, match="threshold"):
        BinaryFairness(2, threshold=2.0)
    with pytest.warns(UserWarning, match="does not require a target"):
        binary_fairness(preds, target, groups, task="demographic_parity")
    metric = BinaryFairness(2, task="demographic_parity")
    with pytest.warns(UserWarning, match="does not require a target"):
        metric.update(preds, target, groups)
    assert list(metric.compute())[0].startswith("DP_")


def test_there_is_no_cpu_path_without_the_fixture():
    with pytest.raises(NativeLibraryError):
        binary_groups_stat_rates(torch.rand(4), torch.randint(2, (4,)), torch.randint(2, (4,)), 2)
