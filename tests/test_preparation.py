"""Preparation: the per-state input files the reference's Preparation writes
(/root/reference/biceps/Restraint.py:896-1152) and the one-file-per-restraint-type columnar format."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import ref_harness as RH

S, J = 5, 7


def _inputs(seed):
    rng = np.random.default_rng(seed)
    exp = np.column_stack([np.array([0, 1, 1, 2, 3, 4, 4]), rng.uniform(1, 5, J)])
    return dict(exp=exp, noe=[rng.uniform(2, 6, J) for _ in range(S)], noe_ind=rng.integers(0, 50, (J, 2)),
                J=[list(rng.uniform(2, 6, J)) for _ in range(S)], J_ind=rng.integers(0, 50, (J, 4)).tolist(),
                cs=np.array([rng.uniform(2, 6, J) for _ in range(S)]), cs_ind=rng.integers(0, 50, J))


def _run(P, d, x, **kw):
    p = P(nstates=S, top_file=None, outdir=str(d))
    p.prepare_noe(x["exp"], x["noe"], x["noe_ind"], **kw)
    p.prepare_J(x["exp"], x["J"], x["J_ind"], **kw)
    p.prepare_cs(x["exp"], x["cs"], x["cs_ind"], "H", **kw)
    p.prepare_cs(x["exp"], x["cs"], x["cs_ind"], "Ca", **kw)
    return p


def test_files_columns_and_values(tmp_path):
    from biceps_b200.Restraint import Preparation
    x = _inputs(0)
    p = _run(Preparation, tmp_path, x)
    assert sorted(os.listdir(tmp_path)) == sorted("%d.%s" % (i, e) for i in range(S) for e in ("J", "noe", "cs_H", "cs_Ca"))
    df = pd.read_pickle(tmp_path / "3.J")
    assert list(df.columns) == ["exp", "model", "restraint_index", "atom_index1", "atom_index2", "atom_index3", "atom_index4"]
    assert np.array_equal(df["model"], x["J"][3]) and np.array_equal(df["exp"], x["exp"][:, 1])
    assert df["restraint_index"].tolist() == [0, 1, 1, 2, 3, 4, 4] and df["atom_index4"].tolist() == [r[3] for r in x["J_ind"]]
    assert list(pd.read_pickle(tmp_path / "0.cs_Ca").columns) == ["exp", "model", "restraint_index", "atom_index1"]
    rows = p.to_sorted_list()
    assert len(rows) == S and [os.path.basename(f) for f in rows[2]] == ["2.cs_H", "2.cs_Ca", "2.J", "2.noe"]


def test_errors_like_the_reference(tmp_path):
    from biceps_b200.Restraint import Preparation
    x = _inputs(0)
    p = Preparation(nstates=S + 1, outdir=str(tmp_path))
    with pytest.raises(ValueError, match="number of states"):
        p.prepare_noe(x["exp"], x["noe"], x["noe_ind"])
    p = Preparation(nstates=S, outdir=str(tmp_path))
    with pytest.raises(ValueError, match="does not match the number of restraints"):
        p.prepare_noe(x["exp"][:-1], x["noe"], x["noe_ind"])
    with pytest.raises(TypeError, match="expected a file path"):
        p.prepare_noe(x["exp"], x["noe"], 5)


def test_model_files_by_glob_and_pf(tmp_path):
    from biceps_b200.Restraint import Preparation
    x = _inputs(2)
    d = tmp_path / "in"
    d.mkdir()
    for i in range(S):
        np.savetxt(d / ("model_%d.txt" % i), x["noe"][i])
    np.savetxt(d / "exp.dat", x["exp"])
    np.savetxt(d / "ind.dat", x["noe_ind"], fmt="%d")
    out = tmp_path / "out"
    out.mkdir()
    p = Preparation(nstates=S, outdir=str(out))
    p.prepare_noe(str(d / "exp.dat"), str(d / "model_*.txt"), str(d / "ind.dat"))
    assert np.allclose(pd.read_pickle(out / "4.noe")["model"], x["noe"][4], rtol=1e-15)
    p.prepare_pf(x["exp"], None, x["noe_ind"][:, :1])
    assert list(pd.read_pickle(out / "1.pf").columns) == ["exp", "restraint_index", "atom_index1"]
    p.prepare_pf(x["exp"], x["noe"], x["noe_ind"][:, :1])
    assert list(pd.read_pickle(out / "1.pf").columns) == ["exp", "model", "restraint_index", "atom_index1"]


def test_columnar_file_per_restraint_type(tmp_path):
    from biceps_b200.Restraint import Preparation
    x = _inputs(3)
    _run(Preparation, tmp_path, x, write_as="columnar")
    assert sorted(os.listdir(tmp_path)) == ["ensemble.J", "ensemble.cs_Ca", "ensemble.cs_H", "ensemble.noe"]
    z = np.load(tmp_path / "ensemble.noe")
    assert z["model"].shape == (S, J) and np.array_equal(z["model"][2], x["noe"][2])
    assert np.array_equal(z["atoms"], x["noe_ind"]) and str(z["kind"]) == "noe"


@pytest.mark.skipif(not RH.reference_available(), reason="reference tree not present")
def test_identical_to_the_reference_files(tmp_path):
    ref = RH.import_reference()
    from biceps_b200.Restraint import Preparation
    x = _inputs(1)
    (tmp_path / "ref").mkdir()
    (tmp_path / "mine").mkdir()
    with RH.quiet():
        _run(ref.Preparation, tmp_path / "ref", x)
    _run(Preparation, tmp_path / "mine", x)
    names = sorted(os.listdir(tmp_path / "ref"))
    assert names == sorted(os.listdir(tmp_path / "mine")) and len(names) == 4 * S
    for f in names:
        a, b = pd.read_pickle(tmp_path / "ref" / f), pd.read_pickle(tmp_path / "mine" / f)
        assert list(a.columns) == list(b.columns) and (a.dtypes == b.dtypes).all() and a.equals(b), f


def test_get_files_uses_natsorts_integer_order(tmp_path):
    """natsort's default algorithm splits on unsigned integers (a decimal point is text): the reference's Analysis
    therefore orders lambdas 0.0, 0.5, 0.25, 0.75, 1.0 (toolbox.py:66-80, Analysis.py:112-116); same here, and the
    import shim used to run the live reference does the same."""
    import biceps_b200 as biceps
    from oracle.stubs import natsort
    names = ["traj_lambda%s.npz" % x for x in ("0.75", "1.00", "0.25", "0.00", "0.5", "10.0", "2.0")]
    for nm in names:
        (tmp_path / nm).write_text("")
    got = [os.path.basename(f) for f in biceps.toolbox.get_files(str(tmp_path / "traj_lambda*.npz"))]
    want = ["traj_lambda0.00.npz", "traj_lambda0.5.npz", "traj_lambda0.25.npz", "traj_lambda0.75.npz",
            "traj_lambda1.00.npz", "traj_lambda2.0.npz", "traj_lambda10.0.npz"]
    assert got == want
    assert [os.path.basename(f) for f in natsort.natsorted(str(tmp_path / nm) for nm in names)] == want
