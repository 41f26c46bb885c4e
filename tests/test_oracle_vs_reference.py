"""CPU, build container only: the oracle against the LIVE reference imported from /root/reference
(skipped where the reference tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import philox as PX
from oracle import ref_harness as H
from oracle import sampler_oracle as O

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference tree not present")

OPTS = [dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
        dict(ref="exponential", sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02))]


@pytest.fixture(scope="module")
def ref():
    biceps = H.import_reference()
    energies, input_data = H.cineromycin_inputs(biceps)
    return biceps, energies, input_data


@pytest.mark.parametrize("lam,seed,burn", [(1.0, 3, 0), (0.5, 21, 250)])
def test_live_reference_replay(ref, lam, seed, burn):
    biceps, energies, input_data = ref
    s = H.build_reference_sampler(biceps, lam, energies, input_data, OPTS, seed)
    with H.quiet():
        s.sample(4000, burn=burn)
    tabs = H.extract_tables(s)
    r = O.run_chain(tabs, PX.MTGlobalRNG(seed), 4000, burn=burn, traj_every=100)
    ot = [(t[0], t[1], t[2], int(t[3][0]), tuple(np.concatenate(t[4]).tolist())) for t in s.traj.trajectory]
    mine = list(zip(r.traj_step, r.traj_E, r.traj_accept, r.traj_state, [tuple(i) for i in r.traj_indices]))
    assert ot == mine
    assert s.traj.state_trace == r.state_trace
    assert np.array_equal(s.traj.state_counts, r.state_counts)
    assert all(np.array_equal(a, b) for a, b in zip(s.traj.sampled_parameters, r.sampled_parameters))
    assert s.accepted == r.accepted and np.array_equal(s.traj.sep_accept[0], r.sep_accept[0])
    assert s.traj.traces == r.traces


def test_live_reference_second_sample_call_continues(ref):
    """sample() twice on one sampler: state/E/counters persist, indices restart at len/2."""
    biceps, energies, input_data = ref
    s = H.build_reference_sampler(biceps, 1.0, energies, input_data, OPTS, 5)
    with H.quiet():
        s.sample(1500)
        s.sample(1000)
    tabs = H.extract_tables(s)
    rng = PX.MTGlobalRNG(5)
    a = O.run_chain(tabs, rng, 1500)
    b = O.run_chain(tabs, rng, 1000, chain=a.chain, reset_indices=True)
    assert b.chain["E"] == s.E and b.chain["state"] == int(s.state[0])
    assert b.accepted == s.accepted and b.total == s.total
    assert np.array_equal(s.traj.sep_accept[2], b.sep_accept[0])


def test_neglogp_matches_reference(ref):
    biceps, energies, input_data = ref
    s = H.build_reference_sampler(biceps, 1.0, energies, input_data, OPTS, 1)
    tabs = H.extract_tables(s)
    rng = np.random.default_rng(0)
    allowed = O.allowed_parameters(tabs)
    for _ in range(200):
        st = int(rng.integers(100))
        idx = [int(rng.integers(len(a))) for a in allowed]
        vals = [[allowed[0][idx[0]]], [allowed[1][idx[1]], allowed[2][idx[2]]]]
        want = s.neglogP([st], vals, [[idx[0]], [idx[1], idx[2]]])
        assert O.neglogP(tabs, st, idx) == want
