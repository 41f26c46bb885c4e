"""CPU: HDX protection factors on the 6-D nuisance grid.
The Python oracle (dense tables, reference summation order) replays the reference's own MT-seeded
trajectory bit for bit; the C oracle evaluates the term on the fly from <N_c>/<N_h> with the
lane-tree summation order the CUDA kernel uses, and must make the same decisions."""
import numpy as np

from helpers import load_npz, tables_from_npz, pf_raw_tables
from oracle import philox as PX
from oracle import sampler_oracle as O


def test_python_oracle_replays_reference_pf_trajectory():
    g = load_npz("pf_grid.npz")
    t = tables_from_npz(g, "t_")
    assert t["restraints"][0]["kind"] == "pfgrid" and list(g["t_rest_type"]) == \
        ["sigma_pf", "c_pf", "h_pf", "0_pf", "xcs_pf", "xhs_pf", "bs_pf"]
    r = O.run_chain(t, PX.MTGlobalRNG(int(g["t_seed"])), int(g["t_nsteps"]), traj_every=100)
    assert np.array_equal(np.array(r.traj_E), g["t_E"])
    assert np.array_equal(np.array(r.traj_indices), g["t_indices"])
    assert np.array_equal(np.array(r.state_trace), g["t_state_trace"])
    assert np.array_equal(np.concatenate(r.sampled_parameters), g["t_sampled_parameters"])
    assert r.accepted == g["t_accepted"] and np.array_equal(r.sep_accept[0], g["t_sep_accept0"])


def test_dense_tables_from_raw_counts_in_reference_order():
    """sse / exponential reference of every grid cell, residues summed left to right
    (Restraint.py:742-749, 843-849) == the reference's dense tables, bit for bit."""
    g = load_npz("pf_grid.npz")
    rt = tables_from_npz(g, "t_")["restraints"][0]
    Nc, Nh, e = g["Nc"], g["Nh"], g["exp"]
    bc, bh, b0 = rt["grids"][0], rt["grids"][1], rt["grids"][2]
    sse = np.zeros(rt["sse"].shape); ref = np.zeros(rt["sse"].shape)
    for j in range(Nc.shape[-1]):
        model = (bc[None, :, None, None, None, None, None] * Nc[:, None, None, None, :, None, :, j]
                 + bh[None, None, :, None, None, None, None] * Nh[:, None, None, None, None, :, :, j]) \
            + b0[None, None, None, :, None, None, None]
        sse += 1 * (model - e[j]) ** 2.0
        ref += 1 * np.maximum(-1.0 * model, 0.0)
    assert np.array_equal(sse, rt["sse"]) and np.array_equal(ref, rt["refsum"])


def test_c_oracle_on_the_fly_matches_dense_oracle(oracle_lib):
    from oracle import cbind
    g = load_npz("pf_grid.npz")
    dense = tables_from_npz(g, "t_")
    raw = pf_raw_tables(g, dense)
    n, every = 3000, 10
    o = cbind.sample(raw, 6, 21, n, burn=30, traj_every=every, record_state_trace=True, chain_group=np.arange(6), n_groups=6)
    for c in range(6):
        r = O.run_chain(dense, PX.PhiloxStepRNG(21, c), n, burn=30, traj_every=every, accept_fn=O.accept_bcp)
        assert np.array_equal(o["traj_state"][:, c], np.array(r.traj_state))
        assert np.array_equal(o["traj_indices"][:, :, c], np.array(r.traj_indices))
        assert np.array_equal(o["state_trace"][:, c], np.array(r.state_trace))
        assert np.allclose(o["traj_energy"][:, c], np.array(r.traj_E), rtol=1e-12, atol=0)
        assert o["accepted"][c] == r.accepted and np.array_equal(o["sep_accepted"][c], r.sep_accepted)
        assert np.array_equal(o["param_counts"][c], np.concatenate(r.sampled_parameters))


def test_python_oracle_replays_reference_on_the_full_default_grid(oracle_lib):
    """The reference's FULL default grid (20 x 26 x 50 x 7 x 8 x 1 cells, the C4 grid): dense sse / exponential
    reference tables rebuilt from the raw counts in the reference's residue order, then the Python oracle on the
    reference's MT19937 stream == the live reference's trajectory (tests/golden/pf_full_grid.npz), bit for bit."""
    g = load_npz("pf_full_grid.npz")
    grids = [g["grid%d" % k] for k in range(6)]
    Nc, Nh, e = g["Nc"], g["Nh"], g["exp"]
    bc, bh, b0 = grids[0], grids[1], grids[2]
    shape = (Nc.shape[0],) + tuple(len(x) for x in grids)
    sse = np.zeros(shape); ref = np.zeros(shape)
    for j in range(Nc.shape[-1]):
        model = (bc[None, :, None, None, None, None, None] * Nc[:, None, None, None, :, None, :, j]
                 + bh[None, None, :, None, None, None, None] * Nh[:, None, None, None, None, :, :, j]) \
            + b0[None, None, None, :, None, None, None]
        sse += 1 * (model - e[j]) ** 2.0
        ref += 1 * np.maximum(-1.0 * model, 0.0)
    rt = dict(kind="pfgrid", name="pf", extension="pf", ref="exponential", ndof=float(g["ndof"]), cnorm=float(g["cnorm"]),
              sigma=g["sigma"], nlsig=g["nlsig"], two_sig2=g["two_sig2"], sse=sse, refsum=ref, grids=grids,
              prior=np.zeros(shape[1:]))
    t = dict(n_states=shape[0], lam=1.0, energy=g["energy"], logZ=float(g["logZ"]), restraints=[rt])
    r = O.run_chain(t, PX.MTGlobalRNG(int(g["t_seed"])), int(g["t_nsteps"]), traj_every=100)
    assert np.array_equal(np.array(r.traj_E), g["t_E"])
    assert np.array_equal(np.array(r.traj_indices), g["t_indices"])
    assert np.array_equal(np.array(r.state_trace), g["t_state_trace"])
    assert np.array_equal(np.concatenate(r.sampled_parameters), g["t_sampled_parameters"])
    assert r.accepted == g["t_accepted"] == 761
    # ... and the C oracle (term evaluated on the fly from the raw counts in the kernel's lane-tree order, Philox) makes
    # the same decisions as the dense reference-order oracle on this grid: every index, state and counter identical,
    # energies to 1e-12 (the summation order is the only difference)
    from oracle import cbind
    raw = dict(n_states=shape[0], energy=g["energy"], logZ=float(g["logZ"]),
               restraints=[dict(rt, sse=None, refsum=None, prior=None, pf=dict(Nc=Nc, Nh=Nh, exp=e, weight=g["weight"]))])
    n = 1500
    o = cbind.sample(raw, 3, 77, n, burn=10, traj_every=10, record_state_trace=True, chain_group=np.arange(3), n_groups=3)
    for c in range(3):
        rc = O.run_chain(t, PX.PhiloxStepRNG(77, c), n, burn=10, traj_every=10, accept_fn=O.accept_bcp)
        assert np.array_equal(o["traj_indices"][:, :, c], np.array(rc.traj_indices))
        assert np.array_equal(o["state_trace"][:, c], np.array(rc.state_trace))
        assert np.allclose(o["traj_energy"][:, c], np.array(rc.traj_E), rtol=1e-12, atol=0)
        assert o["accepted"][c] == rc.accepted and np.array_equal(o["param_counts"][c], np.concatenate(rc.sampled_parameters))
