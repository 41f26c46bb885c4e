"""Generates the golden fixtures under tests/golden/ by importing the UNMODIFIED reference
from /root/reference in the build container (it does not exist on the GPU box).

    python tests/golden/make_golden.py

Fixtures (all small .npz, committed):
  cineromycin_tables.npz   flat tables of the C1 workload (docs/examples/datasets/cineromycin_B,
                           J + NOE, options of CineromycinB.ipynb) for lambda = 0 and 1, plus the
                           raw observables model/exp/weight (inputs of the precompute kernels) and
                           the reference-potential parameters (betas)
  cineromycin_mt_traj.npz  what the reference's own PosteriorSampler.sample produced with
                           np.random.seed(seed) for (lambda, seed) in [(1.0, 0), (0.0, 7)]:
                           snapshots (step, E, accept, state, indices), state_trace, state_counts,
                           sampled_parameters, accepted, sep_accept
  tutorial_mbar.npz        the reference's stored tutorial results (2 lambdas x 1000 snapshots, tables
                           from its pickled samplers, BS.dat and populations.dat written by its own
                           Analysis with the real pymbar) -- the golden for u_kn + MBAR
  pf_grid.npz              HDX protection factors on a reduced 6-D nuisance grid, 6 states x 107 residues:
                           raw <N_c>/<N_h>, dense reference tables (sse, exponential ref, prior) and the
                           MT-seeded reference trajectory (7 parameters moving together)
  apomyoglobin_tables.npz  13 states, cs_H / cs_Ca / cs_N (gaussian / exponential / uniform refs)
                           and their MT-seeded reference trajectory (exercises Restraint_cs and
                           both reference potentials on three restraints)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as H  # noqa: E402

CIN_OPTS = [dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
            dict(ref="exponential", sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02))]


def pack_tables(prefix, tabs, out):
    out[prefix + "energy"] = tabs["energy"]
    out[prefix + "logZ"] = np.float64(tabs["logZ"])
    out[prefix + "lam"] = np.float64(tabs["lam"])
    out[prefix + "n_restraints"] = np.int64(len(tabs["restraints"]))
    for q, rt in enumerate(tabs["restraints"]):
        p = "%sr%d_" % (prefix, q)
        out[p + "kind"] = np.array(rt["kind"])
        out[p + "name"] = np.array(rt["name"])
        out[p + "extension"] = np.array(rt["extension"])
        out[p + "ref"] = np.array(rt["ref"])
        for k in ("ndof", "cnorm"):
            out[p + k] = np.float64(rt[k])
        for k in ("sigma", "nlsig", "two_sig2", "sse"):
            out[p + k] = rt[k]
        out[p + "n_grids"] = np.int64(len(rt["grids"]))
        for g, arr in enumerate(rt["grids"]):
            out[p + "grid%d" % g] = arr
        if rt["refsum"] is not None:
            out[p + "refsum"] = rt["refsum"]
        if "log_normal" in rt:
            out[p + "log_normal"] = np.bool_(rt["log_normal"])


def pack_traj(prefix, s, out):
    tr = s.traj.trajectory
    out[prefix + "step"] = np.array([t[0] for t in tr], np.int64)
    out[prefix + "E"] = np.array([t[1] for t in tr], np.float64)
    out[prefix + "accept"] = np.array([t[2] for t in tr], np.int64)
    out[prefix + "state"] = np.array([int(t[3][0]) for t in tr], np.int64)
    out[prefix + "indices"] = np.array([np.concatenate(t[4]) for t in tr], np.int64)
    out[prefix + "traces"] = np.array(s.traj.traces, np.float64)
    out[prefix + "state_trace"] = np.array(s.traj.state_trace, np.int32)
    out[prefix + "state_counts"] = np.array(s.traj.state_counts, np.float64)
    out[prefix + "sampled_parameters"] = np.concatenate(s.traj.sampled_parameters)
    out[prefix + "accepted"] = np.float64(s.accepted)
    out[prefix + "total"] = np.float64(s.total)
    out[prefix + "sep_accept0"] = np.array(s.traj.sep_accept[0], np.float64)
    out[prefix + "sep_accept1"] = np.float64(s.traj.sep_accept[1])
    out[prefix + "rest_type"] = np.array(s.traj.rest_type)
    out[prefix + "final_E"] = np.float64(s.E)


def make_cineromycin(biceps):
    energies, input_data = H.cineromycin_inputs(biceps)
    tab_out, traj_out = {}, {}
    tab_out["raw_energies"] = energies
    for lam, seed, nsteps, burn in [(1.0, 0, 20000, 100), (0.0, 7, 20000, 0)]:
        s = H.build_reference_sampler(biceps, lam, energies, input_data, CIN_OPTS, seed)
        tabs = H.extract_tables(s)
        tag = "lam%d_" % int(lam)
        pack_tables(tag, tabs, tab_out)
        if lam == 1.0:
            for q, ob in enumerate(H.extract_observables(s)):
                for k, v in ob.items():
                    tab_out["obs%d_%s" % (q, k)] = v
            tab_out["obs1_betas"] = np.array(s.ensemble[0][1].betas, np.float64)
            for q in range(2):
                tab_out["obs%d_restraint_index" % q] = np.array(
                    [r["restraint_index"] for r in s.ensemble[0][q].restraints], np.int64)
        with H.quiet():
            s.sample(nsteps, burn=burn)
        pack_traj(tag, s, traj_out)
        traj_out[tag + "seed"] = np.int64(seed)
        traj_out[tag + "nsteps"] = np.int64(nsteps)
        traj_out[tag + "burn"] = np.int64(burn)
    np.savez_compressed(os.path.join(HERE, "cineromycin_tables.npz"), **tab_out)
    np.savez_compressed(os.path.join(HERE, "cineromycin_mt_traj.npz"), **traj_out)


def make_apomyoglobin(biceps):
    D = os.path.join(H.REFERENCE_ROOT, "docs/examples/datasets/apomyoglobin/")
    energies = np.loadtxt(D + "energy_model_1.txt") if os.path.exists(D + "energy_model_1.txt") else None
    data = biceps.toolbox.sort_data(D + "new_CS_PF")
    # keep the three chemical-shift restraints (cs_H, cs_Ca, cs_N); PF needs the missing prior blob
    data = [[f for f in row if ".cs_" in f] for row in data]
    S = len(data)
    if energies is None or len(energies) != S:
        energies = np.linspace(0.0, 3.0, S)
    energies = energies - energies.min()
    opts = [dict(ref="gaussian", sigma=(0.05, 20.0, 1.02)),
            dict(ref="exponential", sigma=(0.05, 20.0, 1.02)),
            dict(ref="uniform", sigma=(0.05, 20.0, 1.02), use_global_ref_sigma=False)]
    out = {"raw_energies": energies}
    s = H.build_reference_sampler(biceps, 0.5, energies, data, opts, 11)
    pack_tables("t_", H.extract_tables(s), out)
    for q, ob in enumerate(H.extract_observables(s)):
        for k, v in ob.items():
            out["obs%d_%s" % (q, k)] = v
    out["obs0_ref_mean"] = np.array(s.ensemble[0][0].ref_mean, np.float64)
    out["obs0_ref_sigma"] = np.array(s.ensemble[0][0].ref_sigma, np.float64)
    out["obs1_betas"] = np.array(s.ensemble[0][1].betas, np.float64)
    with H.quiet():
        s.sample(10000, burn=0)
    pack_traj("t_", s, out)
    out["t_seed"] = np.int64(11); out["t_nsteps"] = np.int64(10000); out["t_burn"] = np.int64(0)
    np.savez_compressed(os.path.join(HERE, "apomyoglobin_tables.npz"), **out)


def make_tutorial_mbar(biceps):
    """The one self-consistent stored result set of the reference:
    docs/examples/Tutorials/Prep_Rest_Post_Ana/results -- trajectories + pickled samplers for
    lambda = 0 and 1, and the BS.dat / populations.dat its own Analysis (with real pymbar) wrote."""
    import pickle
    D = os.path.join(H.REFERENCE_ROOT, "docs/examples/Tutorials/Prep_Rest_Post_Ana/results/")
    out = {"BS": np.loadtxt(D + "BS.dat"), "populations": np.loadtxt(D + "populations.dat")}
    for k, lam in enumerate(["0.00", "1.00"]):
        traj = np.load(D + "traj_lambda%s.npz" % lam, allow_pickle=True)["arr_0"].item()
        with open(D + "traj_lambda%s.pkl" % lam, "rb") as f:
            sampler = pickle.load(f)
        pack_tables("k%d_" % k, H.extract_tables(sampler), out)
        tr = traj["trajectory"]
        out["k%d_traj_E" % k] = np.array([t[1] for t in tr], np.float64)
        out["k%d_traj_state" % k] = np.array([int(t[3][0]) for t in tr], np.int64)
        out["k%d_traj_indices" % k] = np.array([np.concatenate(t[4]) for t in tr], np.int64)
        out["k%d_state_counts" % k] = np.bincount(np.array(traj["state_trace"], np.int64), minlength=100)
    np.savez_compressed(os.path.join(HERE, "tutorial_mbar.npz"), **out)


PF_OPTS = dict(ref="exponential", sigma=(0.5, 8.0, 1.2), beta_c=(0.05, 0.15, 0.05), beta_h=(0.0, 0.6, 0.2),
               beta_0=(-1.0, 0.0, 0.5), xcs=(5.0, 8.5, 0.5), xhs=(2.0, 2.7, 0.1), bs=(15.0, 16.0, 1.0))


def make_pf(biceps):
    """HDX protection factors on the 6-D nuisance grid (Restraint.py:595-894) with a reduced grid and
    a synthetic prior.  The reference's loader gives every state the LAST listed state's <N_c>/<N_h>
    (Restraint.py:634-644), so per-state arrays are injected after construction and the models / SSE
    recomputed with the reference's own methods."""
    import tempfile
    import pandas as pd
    rng = np.random.default_rng(42)
    S, J = 6, 107
    tmp = tempfile.mkdtemp(prefix="pf_golden_")
    xcs = np.arange(5.0, 8.5, 0.5); xhs = np.arange(2.0, 2.7, 0.1); bs = np.arange(15.0, 16.0, 1.0)
    exp = rng.uniform(0.0, 10.0, J)
    Nc = np.sort(rng.uniform(0.0, 60.0, (S, len(xcs), len(bs), J)), axis=1)      # increasing in x_c
    Nh = np.sort(rng.uniform(0.0, 4.0, (S, len(xhs), len(bs), J)), axis=1)
    for o, xc in enumerate(xcs):
        for q, b in enumerate(bs):
            np.save("%s/Nc_x%0.1f_b%d_state%03d.npy" % (tmp, xc, b, 0), Nc[0, o, q])
    for p_, xh in enumerate(xhs):
        for q, b in enumerate(bs):
            np.save("%s/Nh_x%0.1f_b%d_state%03d.npy" % (tmp, xh, b, 0), Nh[0, p_, q])
    files = []
    for i in range(S):
        fn = "%s/%d.pf" % (tmp, i)
        pd.DataFrame({"atom_index1": np.arange(J), "exp": exp}).to_pickle(fn)
        files.append([fn])
    nshape = (2, 3, 2, len(xcs), len(xhs), len(bs))
    prior = rng.uniform(0.0, 2.0, nshape)
    np.save(tmp + "/prior.npy", prior)
    energies = np.linspace(0.0, 2.5, S)
    opts = dict(PF_OPTS, Ncs_fi=tmp, Nhs_fi=tmp, states=[0], pf_prior=tmp + "/prior.npy")
    ens = biceps.Ensemble(1.0, energies)
    ens.initialize_restraints(files, [opts])
    for i, st in enumerate(ens.to_list()):
        R = st[0]
        assert R.sse.shape == nshape
        R.Ncs, R.Nhs = Nc[i], Nh[i]
        for j in range(J):
            R.restraints[j]["model"] = R.compute_PF_multi(R.Ncs[:, :, j], R.Nhs[:, :, j])
        R.sse = R.compute_sse(R.restraints)
    np.random.seed(5)
    s = biceps.PosteriorSampler(ens)
    out = {"raw_energies": energies, "Nc": Nc, "Nh": Nh, "exp": exp, "weight": np.ones(J)}
    pack_tables("t_", H.extract_tables(s), out)
    out["t_r0_prior"] = prior
    with H.quiet():
        s.sample(4000, burn=0)
    pack_traj("t_", s, out)
    out["t_seed"] = np.int64(5); out["t_nsteps"] = np.int64(4000)
    np.savez_compressed(os.path.join(HERE, "pf_grid.npz"), **out)


if __name__ == "__main__":
    biceps = H.import_reference()
    make_pf(biceps)
    make_cineromycin(biceps)
    make_apomyoglobin(biceps)
    make_tutorial_mbar(biceps)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
