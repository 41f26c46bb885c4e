"""CPU: host-side logic of biceps_b200.convergence that needs no kernel (labels, argument checks, exponential fit,
block splitting) -- the numbers themselves are covered by the GPU tests."""
import numpy as np
import pytest


def test_labels_blocks_and_argument_checks():
    from biceps_b200 import convergence as BC
    C = BC.Convergence.from_arrays(np.zeros((3, 8)), np.zeros((3, 8), np.int32), [np.arange(4.)] * 3,
                                   ["sigma_J", "sigma_noe", "gamma_noe"], outdir=None)
    assert C.labels == ["$\\sigma_{J}$", "$\\sigma_{noe}$", "$\\gamma_{noe}$"]           # convergence.py:242-256
    assert C.outdir and C.freq_save_traj == 1
    with pytest.raises(ValueError):
        BC.Convergence.from_arrays(np.zeros((3, 8)), np.zeros((2, 8), np.int32), [np.arange(4.)] * 3, ["a", "b", "c"])
    with pytest.raises(ValueError):
        BC.Convergence()
    C.tau_c = np.zeros(3)
    with pytest.raises(ValueError):
        C.process(rng="mt")                                   # unknown generator
    b = BC.get_blocks([np.arange(11), np.arange(7)], nblocks=3)
    assert [len(x) for x in b[0]] == [3, 3, 3] and [len(x) for x in b[1]] == [2, 2, 2]   # tails dropped (:128-141)
    with pytest.raises(KeyError):
        C.get_autocorrelation_curves(method="nope")          # rejected before any kernel is launched (:437)


def test_exponential_fit_recovers_tau():
    from biceps_b200 import convergence as BC
    x = np.arange(400)
    y = 0.05 + 0.9 * np.exp(-x / 37.0)
    fit, tau = BC.exponential_fit(y, exp_function="single", v0=[0.0, 1.0, 50.0])
    assert abs(tau - 37.0) < 1e-6 and np.allclose(fit, y, atol=1e-9)
    y2 = 0.7 * np.exp(-x / 80.0) + 0.3 * np.exp(-x / 5.0)
    fit2, tau2 = BC.exponential_fit(y2, exp_function="double", v0=[0.0, 0.8, 0.2, 60.0, 8.0])
    assert abs(tau2 - 80.0) < 1e-4
