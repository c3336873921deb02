"""Dry run of the gated GPU tests themselves: the sampler-level test functions of tests/test_gpu_temporal.py are called
here, unchanged, with the emulated sampler standing in for MultiHopSampler and `.cuda()` patched to a no-op -- so that
the first GPU run of the next round spends its minutes on the kernels, not on typos in tests that have never executed.
(The loader-level functions of that file need the WholeMemory-backed FeatureStore and are covered, with the same
expectations, by tests/test_loaders_emulated_cpu.py.)  Test infrastructure only.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import test_gpu_temporal as G  # noqa: E402
from test_emulated_multihop_cpu import emu  # noqa: E402,F401  (fixture)
from test_loaders_emulated_cpu import stack  # noqa: E402,F401  (fixture)


@pytest.fixture()
def fake_env(stack):  # noqa: F811
    _, _, sampler = stack
    cls = type(sampler)
    if not hasattr(cls, "sample_temporal"):
        cls.sample_temporal = lambda self, *a, **k: self.sample_temporal_async(*a, **k).result()
        cls.sample_hetero = lambda self, *a, **k: self.sample_hetero_async(*a, **k).result()
    return None, None, sampler


@pytest.mark.parametrize("comparison,fanout,col_dtype", [("strictly_increasing", [3, 2, 4, 2, 2, 2], np.int32),
                                                         ("monotonically_decreasing", [-1, 3, 0, 2, -1, 1], np.int64)])
def test_dryrun_hetero_bit_exact(fake_env, oracle, comparison, fanout, col_dtype):
    G.test_temporal_hetero_bit_exact_vs_oracle(fake_env, oracle, comparison, fanout, col_dtype)


def test_dryrun_homogeneous(fake_env, oracle):
    G.test_temporal_homogeneous_matches_oracle(fake_env, oracle, [4, 3])


def test_dryrun_biased_one_hop(fake_env, oracle):
    G.test_biased_temporal_one_hop_sets_vs_oracle(fake_env, oracle)


def test_dryrun_argument_checks(fake_env):
    G.test_temporal_rejects_bad_arguments(fake_env)
