"""GPU: the reference-facing Python API (Ensemble -> PosteriorSampler.sample -> process_results ->
Analysis) end to end, on per-state DataFrame files written from the golden cineromycin observables."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from helpers import cineromycin_tables, stack_lambdas

pytestmark = pytest.mark.gpu

OPTS = [dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
        dict(ref="exponential", sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02))]


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    """<i>.J and <i>.noe pandas pickles, the reference's per-state input format (Restraint.py:411-416, 521-525)."""
    z, t0, t1 = cineromycin_tables()
    d = tmp_path_factory.mktemp("J_NOE")
    S = z["obs0_model"].shape[0]
    for i in range(S):
        J = z["obs0_model"].shape[1]
        pd.DataFrame({"atom_index1": range(J), "atom_index2": range(J), "atom_index3": range(J), "atom_index4": range(J),
                      "exp": z["obs0_exp"], "model": z["obs0_model"][i],
                      "restraint_index": z["obs0_restraint_index"]}).to_pickle(str(d / ("%d.J" % i)))
        pd.DataFrame({"exp": z["obs1_exp"], "model": z["obs1_model"][i],
                      "restraint_index": z["obs1_restraint_index"]}).to_pickle(str(d / ("%d.noe" % i)))
    return str(d), z, t0, t1


def test_ensemble_tables_match_the_reference(dataset):
    import biceps_b200 as biceps
    d, z, t0, t1 = dataset
    input_data = biceps.toolbox.sort_data(d)
    assert len(input_data) == 100 and input_data[3][0].endswith("3.J") and input_data[3][1].endswith("3.noe")
    ens = biceps.Ensemble(1.0, z["raw_energies"])
    ens.initialize_restraints(input_data, OPTS)
    lst = ens.to_list()
    assert len(lst) == 100 and type(lst[0][1]).__name__ == "Restraint_noe"
    for k in range(2):
        want = t1["restraints"][k]
        got = np.array([s[k].sse for s in lst])
        assert np.max(np.abs(got - want["sse"]) / np.abs(want["sse"])) < 1e-11          # tolerance stated: 1e-6
        assert lst[0][k].Ndof == want["ndof"]
        assert np.array_equal(lst[0][k].allowed_sigma, want["sigma"])
    r = lst[5][1].restraints
    assert len(r) == 33 and set(r[0]) == {"exp", "model", "restraint_index", "weight"}
    assert r[7]["model"] == z["obs1_model"][5, 7] and r[7]["weight"] == z["obs1_weight"][7]
    s = biceps.PosteriorSampler(ens, seed=3)
    assert abs(s.logZ - t1["logZ"]) < 1e-12
    assert np.max(np.abs(lst[0][1].betas - z["obs1_betas"]) / z["obs1_betas"]) < 1e-12
    assert s.traj.rest_type == ["sigma_J", "sigma_noe", "gamma_noe"]
    assert s.traj.trajectory_headers[-1] == "para_index = [151, 116, 81]"
    # energies through the API == oracle on the reference's own tables, to the table tolerance
    from oracle import sampler_oracle as O
    e = s.neglogP([7], [[1.0], [1.0, 1.0]], [[150], [100, 80]])
    assert abs(e - O.neglogP(t1, 7, [150, 100, 80])) < 1e-9 * abs(e)


def test_option_errors_like_the_reference(dataset):
    import biceps_b200 as biceps
    d, z, t0, t1 = dataset
    input_data = biceps.toolbox.sort_data(d)
    with pytest.raises(ValueError):
        biceps.Ensemble(1, z["raw_energies"])                                   # lambda must be float
    with pytest.raises(ValueError):
        biceps.Ensemble(1.0, np.arange(100))                                     # energies must be float
    ens = biceps.Ensemble(1.0, z["raw_energies"])
    with pytest.raises(TypeError):
        ens.initialize_restraints(input_data, [dict(ref="uniform", sgima=(1, 2, 1.1)), dict()])
    ens = biceps.Ensemble(1.0, z["raw_energies"])
    ens.initialize_restraints(input_data, [dict(ref="uniform"), dict(ref="exp")])
    with pytest.raises(ValueError):
        biceps.PosteriorSampler(ens)                                             # 'exp' is rejected (:57-59)
    with pytest.raises(ValueError):
        biceps.Analysis("/nonexistent", nstates=0)


def test_full_workflow_matches_oracle_and_mbar(dataset, tmp_path):
    import biceps_b200 as biceps
    from oracle import cbind, mbar_oracle as M
    d, z, t0, t1 = dataset
    input_data = biceps.toolbox.sort_data(d)
    out = str(tmp_path)
    samplers = []
    for lam in (0.0, 1.0):
        ens = biceps.Ensemble(lam, z["raw_energies"])
        ens.initialize_restraints(input_data, OPTS)
        s = biceps.PosteriorSampler(ens, seed=77)
        s.sample(20000, burn=1000)
        s.traj.process_results(os.path.join(out, "traj_lambda%2.2f.npz" % lam))
        samplers.append(s)
        # the same chain from the oracle, on the tables the API built
        o = cbind.sample(s.tables, 1, 77, 20000, burn=1000, traj_every=100, record_state_trace=True)
        tr = s.traj.trajectory
        assert len(tr) == 200 and tr[1][0] == 100
        assert [t[3][0] for t in tr] == list(o["traj_state"][:, 0])
        assert [float(t[1]) for t in tr] == list(o["traj_energy"][:, 0])
        assert [int(t[2]) for t in tr] == list(o["traj_accept"][:, 0])
        assert [list(np.concatenate(t[4])) for t in tr] == [list(x) for x in o["traj_indices"][:, :, 0]]
        assert s.traj.state_trace == list(o["state_trace"][:, 0])
        assert np.array_equal(s.traj.state_counts, 1.0 + o["state_counts"][0])
        assert np.array_equal(np.concatenate(s.traj.sampled_parameters), o["param_counts"][0].astype(float))
        assert s.accepted == o["accepted"][0] and s.total == 21000
        assert np.allclose(s.traj.sep_accept[0], o["sep_accepted"][0] / 21000 * 100.)
        assert s.traj.traces[5] == [s.traj.allowed_parameters[p][tr[5][4][r][j]] for p, (r, j) in enumerate([(0, 0), (1, 0), (1, 1)])]
    # files: npz dict schema + sampler pickle (PosteriorSampler.py:444-470)
    npz = np.load(os.path.join(out, "traj_lambda1.00.npz"), allow_pickle=True)["arr_0"].item()
    assert sorted(npz) == sorted(['rest_type', 'trajectory_headers', 'trajectory', 'sep_accept', 'allowed_parameters',
                                  'sampled_parameters', 'model', 'ref', 'traces', 'state_trace'])
    assert len(npz["model"][1]) == 33 and len(npz["model"][1][0]) == 100 and npz["ref"][0] == ['Nan']
    with open(os.path.join(out, "traj_lambda1.00.pkl"), "rb") as f:
        assert pickle.load(f).lam == 1.0
    A = biceps.Analysis(out, nstates=100, precheck=True)
    assert A.K == 2 and A.lam == [0.0, 1.0] and list(A.N_k) == [200, 200]
    # oracle analysis on the same trajectories and tables
    tabs = [dict(s.tables, energy=s.tables["energy"][0], logZ=float(s.tables["logZ"][0])) for s in samplers]
    trajs = [dict(E=np.array([t[1] for t in s.traj.trajectory]), state=np.array([t[3][0] for t in s.traj.trajectory]),
                  indices=np.array([np.concatenate(t[4]) for t in s.traj.trajectory])) for s in samplers]
    want = M.analysis(tabs, trajs, 100)
    assert np.array_equal(A.u_kln, want["u_kln"])
    assert np.allclose(A.f_df, want["f_df"], atol=1e-9)
    ok = ~np.isnan(want["P_dP"])
    assert np.array_equal(np.isnan(A.P_dP), ~ok) and np.allclose(A.P_dP[ok], want["P_dP"][ok], rtol=1e-7, atol=1e-13)
    assert np.allclose(np.loadtxt(os.path.join(out, "BS.dat")), A.f_df)
    assert abs(A.P_dP[:, :2].sum(axis=0) - 1.0).max() < 1e-9


def test_many_chains_and_second_sample_call(dataset):
    import biceps_b200 as biceps
    d, z, t0, t1 = dataset
    ens = biceps.Ensemble(1.0, z["raw_energies"])
    ens.initialize_restraints(biceps.toolbox.sort_data(d), OPTS)
    s = biceps.PosteriorSampler(ens, nchains=4096, seed=5)
    s.sample(20000, burn=2000)
    assert s.traj.state_counts.sum() == 100 + 4096 * 20000
    assert s.traj.chains["energy"].shape == (200, 4096)
    pops = s.traj.state_counts / s.traj.state_counts.sum()
    s2 = biceps.PosteriorSampler(ens, nchains=4096, seed=6)
    s2.sample(20000, burn=2000)
    p2 = s2.traj.state_counts / s2.traj.state_counts.sum()
    assert float(np.sum(pops * np.log(pops / p2))) < 1e-3            # KL between independent runs
    # second call on the same sampler: counters persist, indices restart (reference quirk)
    s1 = biceps.PosteriorSampler(ens, seed=9)
    s1.sample(1000)
    s1.sample(500)
    assert s1.total == 1500 and len(s1.traj.trajectory) == 15 and len(s1.traj.sep_accept) == 4


def test_analysis_from_samplers_without_files(dataset, tmp_path):
    """Analysis.from_samplers (no npz / pkl round trip) gives the file route's numbers for one chain per
    lambda, and uses every chain of a many-chain sampler."""
    import biceps_b200 as biceps
    d, z, t0, t1 = dataset
    input_data = biceps.toolbox.sort_data(d)
    out = str(tmp_path)
    samplers = []
    for lam in (0.0, 1.0):
        ens = biceps.Ensemble(lam, z["raw_energies"])
        ens.initialize_restraints(input_data, OPTS)
        s = biceps.PosteriorSampler(ens, seed=5)
        s.sample(10000, burn=500)
        s.traj.process_results(os.path.join(out, "traj_lambda%2.2f.npz" % lam))
        samplers.append(s)
    A = biceps.Analysis(out, nstates=100, precheck=False)
    B = biceps.Analysis.from_samplers(samplers, nstates=100)
    assert B.K == 2 and B.lam == [0.0, 1.0] and list(B.N_k) == list(A.N_k)
    assert np.array_equal(B.u_kn, A.u_kn)
    assert np.allclose(B.f_df, A.f_df, atol=1e-12)
    ok = ~np.isnan(A.P_dP)
    assert np.array_equal(np.isnan(B.P_dP), ~ok) and np.allclose(B.P_dP[ok], A.P_dP[ok], rtol=1e-10, atol=1e-15)
    scores = []
    for seed in (9, 10):                                     # two independent many-chain runs
        many = []
        for lam in (0.0, 1.0):
            ens = biceps.Ensemble(lam, z["raw_energies"])
            ens.initialize_restraints(input_data, OPTS)
            s = biceps.PosteriorSampler(ens, nchains=512, seed=seed)
            s.sample(10000, burn=1000)
            many.append(s)
        Cn = biceps.Analysis.from_samplers(many, nstates=100)
        assert list(Cn.N_k) == [512 * 100, 512 * 100]
        assert Cn.f_df[1, 1] < 0.01 and abs(Cn.P_dP[:, :2].sum(axis=0) - 1.0).max() < 1e-9
        scores.append(Cn.f_df[1, 0])
    assert abs(scores[0] - scores[1]) < 0.05                 # north_star tolerance for independent runs
    assert abs(scores[0] - A.f_df[1, 0]) < 5 * A.f_df[1, 1] + 0.05   # and consistent with the one-chain file route


def test_from_arrays_bulk_constructor(dataset):
    import biceps_b200 as biceps
    d, z, t0, t1 = dataset
    ens = biceps.Ensemble.from_arrays(1.0, z["raw_energies"], [
        dict(cls="J", model=z["obs0_model"], exp=z["obs0_exp"], restraint_index=list(z["obs0_restraint_index"]), options=OPTS[0]),
        dict(cls="noe", model=z["obs1_model"], exp=z["obs1_exp"], restraint_index=list(z["obs1_restraint_index"]), options=OPTS[1])])
    s = biceps.PosteriorSampler(ens, seed=3)
    for k in range(2):
        assert np.max(np.abs(s.tables["restraints"][k]["sse"] - t1["restraints"][k]["sse"]) / t1["restraints"][k]["sse"]) < 1e-11
    assert np.max(np.abs(s.tables["restraints"][1]["refsum"] - t1["restraints"][1]["refsum"])) < 1e-10


def test_pf_grid_through_the_api(tmp_path):
    """Restraint_pf(precomputed=False) from per-state <N_c>/<N_h> .npy files; energies and the chain must
    equal the oracle on the reference-derived golden data."""
    import biceps_b200 as biceps
    from helpers import load_npz, tables_from_npz, pf_raw_tables
    from oracle import cbind
    g = load_npz("pf_grid.npz")
    dense = tables_from_npz(g, "t_")
    raw = pf_raw_tables(g, dense)
    d = str(tmp_path)
    S, J = g["Nc"].shape[0], g["Nc"].shape[-1]
    xcs = np.arange(5.0, 8.5, 0.5); xhs = np.arange(2.0, 2.7, 0.1); bs = np.arange(15.0, 16.0, 1.0)
    files = []
    for i in range(S):
        for o, xc in enumerate(xcs):
            np.save("%s/Nc_x%0.1f_b%d_state%03d.npy" % (d, xc, bs[0], i), g["Nc"][i, o, 0])
        for p_, xh in enumerate(xhs):
            np.save("%s/Nh_x%0.1f_b%d_state%03d.npy" % (d, xh, bs[0], i), g["Nh"][i, p_, 0])
        pd.DataFrame({"atom_index1": np.arange(J), "exp": g["exp"]}).to_pickle("%s/%d.pf" % (d, i))
        files.append(["%s/%d.pf" % (d, i)])
    np.save(d + "/prior.npy", dense["restraints"][0]["prior"])
    opts = dict(ref="exponential", sigma=(0.5, 8.0, 1.2), beta_c=(0.05, 0.15, 0.05), beta_h=(0.0, 0.6, 0.2),
                beta_0=(-1.0, 0.0, 0.5), xcs=(5.0, 8.5, 0.5), xhs=(2.0, 2.7, 0.1), bs=(15.0, 16.0, 1.0),
                Ncs_fi=d, Nhs_fi=d, pf_prior=d + "/prior.npy")
    ens = biceps.Ensemble(1.0, g["raw_energies"])
    ens.initialize_restraints(files, [opts])
    s = biceps.PosteriorSampler(ens, seed=12)
    assert s.traj.rest_type == ["sigma_pf", "c_pf", "h_pf", "0_pf", "xcs_pf", "xhs_pf", "bs_pf"]
    assert np.array_equal(s.tables["restraints"][0]["nlsig"], dense["restraints"][0]["nlsig"])
    s.sample(3000, burn=100)
    o = cbind.sample(raw, 1, 12, 3000, burn=100, traj_every=100, record_state_trace=True)
    assert s.traj.state_trace == list(o["state_trace"][:, 0])
    assert [list(np.concatenate(t[4])) for t in s.traj.trajectory] == [list(x) for x in o["traj_indices"][:, :, 0]]
    assert [t[1] for t in s.traj.trajectory] == list(o["traj_energy"][:, 0])
    with pytest.raises(TypeError):
        e2 = biceps.Ensemble(1.0, g["raw_energies"])
        e2.initialize_restraints(files, [dict(opts, ref="gaussian")])
        biceps.PosteriorSampler(e2)


def test_preparation_columnar_route_equals_the_file_route(tmp_path):
    """Preparation -> per-state files -> initialize_restraints  ==  Preparation(write_as="columnar") ->
    Ensemble.from_columnar: same tables, same chain."""
    import biceps_b200 as biceps
    z, t0, t1 = cineromycin_tables()
    S = z["obs0_model"].shape[0]
    J0, J1 = z["obs0_model"].shape[1], z["obs1_model"].shape[1]
    exp0 = np.column_stack([z["obs0_restraint_index"], z["obs0_exp"]])
    exp1 = np.column_stack([z["obs1_restraint_index"], z["obs1_exp"]])
    ind0 = np.arange(4 * J0).reshape(J0, 4)
    ind1 = np.arange(2 * J1).reshape(J1, 2)
    samplers = []
    for mode in ("pickle", "columnar"):
        d = tmp_path / mode
        d.mkdir()
        p = biceps.Preparation(nstates=S, outdir=str(d))
        p.prepare_J(exp0, list(z["obs0_model"]), ind0, write_as=mode)
        p.prepare_noe(exp1, list(z["obs1_model"]), ind1, write_as=mode)
        if mode == "pickle":
            ens = biceps.Ensemble(1.0, z["raw_energies"])
            ens.initialize_restraints(p.to_sorted_list(), OPTS)
        else:
            ens = biceps.Ensemble.from_columnar(1.0, z["raw_energies"], [str(d / "ensemble.J"), str(d / "ensemble.noe")], OPTS)
        s = biceps.PosteriorSampler(ens, seed=3)
        s.sample(3000)
        samplers.append(s)
    a, b = samplers
    for k in range(2):
        assert np.array_equal(a.tables["restraints"][k]["sse"], b.tables["restraints"][k]["sse"])
        assert np.max(np.abs(a.tables["restraints"][k]["sse"] - t1["restraints"][k]["sse"]) / t1["restraints"][k]["sse"]) < 1e-11
    assert a.traj.trajectory == b.traj.trajectory and a.traj.state_trace == b.traj.state_trace


def test_restraint_compute_neglogp_and_rebuilt_reference_potentials(dataset):
    """R.compute_neglogP (Restraint.py:356-383, 572-592) on the kernels' tables adds up to sampler.neglogP, and
    PosteriorSampler.build_exp_ref / build_gaussian_ref (PosteriorSampler.py:74-150) recompute a reference potential."""
    import biceps_b200 as biceps
    from oracle import precompute_oracle as PO
    d, z, t0, t1 = dataset
    ens = biceps.Ensemble(1.0, z["raw_energies"])
    ens.initialize_restraints(biceps.toolbox.sort_data(d), OPTS)
    s = biceps.PosteriorSampler(ens, seed=1)
    lst = ens.to_list()
    i, idx = 17, [[150], [120, 80]]
    vals = [[lst[i][0].allowed_sigma[150]], [lst[i][1].allowed_sigma[120], lst[i][1].allowed_gamma[80]]]
    total = lst[i][0].energy + s.logZ
    for k in range(2):
        total += lst[i][k].compute_neglogP(vals[k], idx[k], lst[i][k].sse)
    assert abs(total - s.neglogP([i], vals, idx)) < 1e-9 * abs(total)
    # switch the J restraint to a gaussian reference potential and the NOE one to exponential again
    s.build_gaussian_ref(0, use_global_ref_sigma=True)
    s.build_exp_ref(1)
    m, e, w = z["obs0_model"], z["obs0_exp"], z["obs0_weight"]
    mean, sig, ref = PO.gaussian_ref(m, w, use_global_ref_sigma=True)
    assert lst[3][0].ref == "gaussian" and lst[3][0].ref_sigma is not None and np.isfinite(lst[3][0].sum_neglog_gaussian_ref)
    assert np.max(np.abs(s.tables["restraints"][1]["refsum"] - t1["restraints"][1]["refsum"])) < 1e-10
    e_new = s.neglogP([i], vals, idx)
    assert abs((e_new - total) + lst[i][0].sum_neglog_gaussian_ref) < 1e-9 * abs(total)
    assert np.allclose(s.tables["restraints"][0]["refsum"], ref, rtol=1e-11)
    assert np.allclose(lst[0][0].ref_mean, mean, rtol=1e-12) and np.allclose(lst[0][0].ref_sigma, sig, rtol=1e-12)


def test_quickstart_example_runs(tmp_path):
    """examples/quickstart.py: Preparation -> Ensemble -> PosteriorSampler -> Analysis -> Convergence, the reference's
    workflow with only the import line changed; the synthetic experiment must point back at the state it came from."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "quickstart.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BICePs score" in r.stdout and "experiment generated from state 78" in r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("most populated states")][0]
    assert line.split("[")[1].split()[0].strip("],") == "78"
    assert sorted(os.listdir(tmp_path / "results"))[:2] == ["BS.dat", "all_JSD.npy"]
