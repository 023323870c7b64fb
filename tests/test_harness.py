"""Row f4 pieces that need no GPU: the restated regret computation (scamlgp/benchmarking/plotting.py:21-53) and the
study generators shared by the product arm and the CPU oracle arm of scripts/bo_parity.py."""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
import harness  # noqa: E402


def test_compute_regrets_running_minimum_and_sign():
    vals = [{"loss": 3.0}, {"loss": 5.0}, {"loss": 1.5}, {"loss": 2.0}]
    assert harness.compute_regrets(False, "loss", 1.0, vals) == [2.0, 2.0, 0.5, 0.5]
    acc = [{"acc": 0.5}, {"acc": 0.9}, {"acc": 0.7}]
    np.testing.assert_allclose(harness.compute_regrets(True, "acc", 1.0, acc), [0.5, 0.1, 0.1])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert harness.compute_regrets(False, "loss", 1.0, [{"loss": 0.9}]) == [0.9 - 1.0]
        assert any("negative regret" in str(x.message) for x in w)


def test_studies_are_reproducible_and_share_the_noise_stream():
    a, b = harness.branin_study(3), harness.branin_study(3)
    assert all(np.array_equal(x, y) for x, y in zip(a.meta_X, b.meta_X))
    assert all(np.array_equal(x, y) for x, y in zip(a.meta_y, b.meta_y))
    assert a.optimum == b.optimum and a.rng.normal() == b.rng.normal()  # identical target-noise draws in every arm
    x = np.array([np.pi, 2.275])
    assert a.objective(x) >= a.optimum - 1e-9  # the located optimum is a lower bound on the dense grid's scale
    h = harness.hartmann6_study(1, tasks=2, points=8)
    assert h.bounds.shape == (6, 2) and len(h.meta_X) == 2 and h.optimum < -2.5  # Hartmann-6 family minimum ~ -3.3
