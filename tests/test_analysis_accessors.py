"""Host-side accessors of Analysis (Analysis.py:119-133, 240-267) on the reference's stored tutorial
trajectories; no GPU work."""
import numpy as np
import pytest

from oracle import ref_harness as RH

SRC = RH.REFERENCE_ROOT + "/docs/examples/Tutorials/Prep_Rest_Post_Ana/results/"


@pytest.mark.skipif(not RH.reference_available(), reason="reference tree not present")
def test_max_likelihood_parameters_and_traces_match_the_reference():
    ref = RH.import_reference()
    import biceps_b200 as mine
    trajs = [np.load(SRC + f, allow_pickle=True)['arr_0'].item() for f in ('traj_lambda0.00.npz', 'traj_lambda1.00.npz')]
    A = ref.Analysis.__new__(ref.Analysis)
    B = mine.Analysis.__new__(mine.Analysis)
    for X in (A, B):
        X.traj, X.scheme = trajs, trajs[0]['rest_type']
    assert A.get_max_likelihood_parameters().equals(B.get_max_likelihood_parameters())
    assert A.get_max_likelihood_parameters(model=1, sigma_only=True).equals(B.get_max_likelihood_parameters(model=1, sigma_only=True))
    # the reference's own get_traces fails on NumPy >= 1.24 (ragged np.array, Analysis.py:246); check the layout
    df = B.get_traces()
    assert list(df.columns) == ["sigma_J0", "sigma_noe0", "gamma_noe0"] and df.shape == (1000, 3)
    assert np.array_equal(df.to_numpy(), np.array(trajs[-1]["traces"]))
