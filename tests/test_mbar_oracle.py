"""CPU: pins the numpy MBAR oracle against the reference's own stored BS.dat / populations.dat
(written by its Analysis with the real pymbar) and checks solver-independent properties for K > 2."""
import numpy as np

from helpers import load_npz, tables_from_npz
from oracle import mbar_oracle as M


def tutorial():
    z = load_npz("tutorial_mbar.npz")
    tabs = [tables_from_npz(z, "k%d_" % k) for k in range(2)]
    trajs = [dict(E=z["k%d_traj_E" % k], state=z["k%d_traj_state" % k], indices=z["k%d_traj_indices" % k])
             for k in range(2)]
    return z, tabs, trajs


def test_stored_energies_are_self_consistent():
    """sampler.neglogP(snapshot) == stored E exactly (SURVEY.md section 4 fixture table)."""
    from oracle import sampler_oracle as O
    z, tabs, trajs = tutorial()
    for k in range(2):
        for n in range(0, 1000, 37):
            e = O.neglogP(tabs[k], int(trajs[k]["state"][n]), [int(x) for x in trajs[k]["indices"][n]])
            assert e == trajs[k]["E"][n]


def test_reproduces_reference_BS_and_populations():
    z, tabs, trajs = tutorial()
    r = M.analysis(tabs, trajs, 100)
    assert np.allclose(r["f_df"], z["BS"], rtol=0, atol=5e-9)          # BS.dat is printed with %.18e
    assert abs(r["f_df"][1, 0] - 0.30276934) < 1e-8 and abs(r["f_df"][1, 1] - 0.03992001) < 1e-8
    sp = z["populations"]
    assert np.max(np.abs(sp[:, :2] - r["P_dP"][:, :2])) < 1e-12
    a, b = sp[:, 2:], r["P_dP"][:, 2:]
    assert np.array_equal(np.isnan(a), np.isnan(b))                    # never-sampled states: dP = nan
    ok = ~np.isnan(a)
    assert np.max(np.abs(a[ok] - b[ok]) / a[ok]) < 1e-10


def test_fixed_point_and_gauge_for_many_lambdas():
    """K = 6 harmonic ensembles with known free energies: f_k = -ln Z_k analytically."""
    rng = np.random.default_rng(0)
    K, n = 6, 4000
    kspring = np.linspace(1.0, 3.0, K)
    x = np.concatenate([rng.normal(0, 1 / np.sqrt(k), n) for k in kspring])
    u_kn = 0.5 * kspring[:, None] * x[None, :] ** 2
    f, W, it = M.mbar_solve(u_kn, np.full(K, n))
    exact = 0.5 * np.log(kspring / kspring[0])
    assert np.max(np.abs(f - exact)) < 0.03
    assert np.allclose(W.sum(axis=0), 1.0, atol=1e-10)                 # fixed point: sum_n W_nk = 1
    f2, _, _ = M.mbar_solve(u_kn + 7.0, np.full(K, n))                 # adding a constant to every u changes nothing
    assert np.allclose(f, f2, atol=1e-9)
