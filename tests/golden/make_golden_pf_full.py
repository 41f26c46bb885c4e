"""Generates tests/golden/pf_full_grid.npz from the LIVE reference (run in the build container, where /root/reference
exists): HDX protection factors on the reference's FULL default 6-D nuisance grid -- beta_c 20 x beta_h 26 x beta_0 50 x
x_c 7 x x_h 8 x b 1 = 1 456 000 cells, the grid BASELINE configs[3] (C4) names -- with S states of 107 residues, the
reference's own dense tables (11.6 MB per state and 1.2 GB of per-residue model arrays per state while it builds them),
its exponential reference potential and its MT19937-seeded sample().  Only the raw <N_c>/<N_h> counts, the grids and the
trajectory go into the fixture; tests/test_mt_stream_gpu.py replays it on the GPU from the raw counts, bit for bit."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_harness as H            # noqa: E402
from make_golden import pack_traj              # noqa: E402

OPTS = dict(ref="exponential", sigma=(0.05, 20.0, 1.02), beta_c=(0.05, 0.25, 0.01), beta_h=(0.0, 5.2, 0.2),
            beta_0=(-10.0, 0.0, 0.2), xcs=(5.0, 8.5, 0.5), xhs=(2.0, 2.7, 0.1), bs=(15.0, 16.0, 1.0))


def main(S=3, J=107, seed=11, nsteps=3000):
    import pandas as pd
    biceps = H.import_reference()
    rng = np.random.default_rng(2024)
    tmp = tempfile.mkdtemp(prefix="pf_full_")
    xcs = np.arange(5.0, 8.5, 0.5); xhs = np.arange(2.0, 2.7, 0.1); bs = np.arange(15.0, 16.0, 1.0)
    exp = rng.uniform(0.0, 10.0, J)
    Nc = np.sort(rng.uniform(0.0, 60.0, (S, len(xcs), len(bs), J)), axis=1)
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
    energies = np.linspace(0.0, 1.5, S)
    ens = biceps.Ensemble(1.0, energies)
    # the prior has the grid's shape (zeros: t += 0.0 leaves the bits alone); the reference needs the file
    probe = biceps.Ensemble(1.0, energies[:1])
    opts0 = dict(OPTS, Ncs_fi=tmp, Nhs_fi=tmp, states=[0])
    grids = [np.arange(*OPTS[k]) for k in ("beta_c", "beta_h", "beta_0", "xcs", "xhs", "bs")]
    nshape = tuple(len(g) for g in grids)
    np.save(tmp + "/prior.npy", np.zeros(nshape))
    opts = dict(opts0, pf_prior=tmp + "/prior.npy")
    ens.initialize_restraints(files, [opts])
    for i, st in enumerate(ens.to_list()):
        R = st[0]
        assert R.sse.shape == nshape, (R.sse.shape, nshape)
        R.Ncs, R.Nhs = Nc[i], Nh[i]
        for j in range(J):
            R.restraints[j]["model"] = R.compute_PF_multi(R.Ncs[:, :, j], R.Nhs[:, :, j])
        R.sse = R.compute_sse(R.restraints)
    np.random.seed(seed)
    s = biceps.PosteriorSampler(ens)
    R0 = ens.to_list()[0][0]
    sig = np.array(R0.allowed_sigma, dtype=np.float64)
    ndof = float(R0.Ndof)
    out = {"raw_energies": energies, "Nc": Nc, "Nh": Nh, "exp": exp, "weight": np.ones(J),
           "energy": np.array([float(st[0].energy) for st in ens.to_list()]), "logZ": np.float64(s.logZ),
           "sigma": sig, "ndof": np.float64(ndof), "nlsig": np.array([ndof * np.log(x) for x in sig]),
           "two_sig2": np.array([2.0 * x ** 2.0 for x in sig]), "cnorm": np.float64(ndof / 2.0 * np.log(2.0 * np.pi)),
           "n_grids": np.int64(6)}
    for g, a in enumerate([np.array(getattr(R0, k), dtype=np.float64) for k in H.PF_GRID_ATTRS]):
        out["grid%d" % g] = a
    with H.quiet():
        s.sample(nsteps, burn=0)
    pack_traj("t_", s, out)
    out["t_seed"] = np.int64(seed); out["t_nsteps"] = np.int64(nsteps)
    np.savez_compressed(os.path.join(HERE, "pf_full_grid.npz"), **out)
    print("pf_full_grid.npz", os.path.getsize(os.path.join(HERE, "pf_full_grid.npz")), "accepted", s.accepted, "of", s.total)


if __name__ == "__main__":
    main()
