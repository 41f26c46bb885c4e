"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/liboracle.so (oracle/biceps_oracle.c).

Shares the struct mirrors of include/biceps_b200.h with the product binding (the oracle may
import the product's interface definitions; never the other way round)."""
import ctypes as C
import os
import subprocess

from biceps_b200 import _cabi as A
from biceps_b200.engine import alloc_outputs, chain_cfg

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_sample.restype = C.c_int
        L.orc_sample.argtypes = [C.POINTER(A.bcp_tables), C.POINTER(A.bcp_chain_cfg), C.POINTER(A.bcp_sample_cfg),
                                 C.c_uint64, C.POINTER(A.bcp_sample_out), C.c_int]
        L.orc_bcp_exp.restype = C.c_double
        L.orc_bcp_exp.argtypes = [C.c_double]
        L.orc_neglogp.restype = C.c_double
        L.orc_neglogp.argtypes = [C.POINTER(A.bcp_tables), C.c_int, C.c_int, A.c_int32_p]
        _lib = L
    return _lib


def param_layout(packed):
    P = HB = 0
    for q in range(packed.n_restraints):
        r = packed.struct.restraints[q]
        P += 1; HB += r.n_sigma
        if r.kind == A.BCP_KIND_GAMMA:
            P += 1; HB += r.n_gamma
        if r.kind == A.BCP_KIND_PFGRID:
            for k in range(A.BCP_PF_NGRID):
                P += 1; HB += r.pf_n[k]
    return P, HB


def sample(tables, n_chains, seed, n_steps, burn=0, traj_every=0, record_state_trace=False, chain_id0=0,
           chain_lambda=None, chain_group=None, n_groups=0, init_state=None, n_threads=1, step0=0, resume=None):
    """orc_sample with the same argument meaning and output dict as engine.ChainBlock.sample.
    `resume` = output dict of a previous call (continues those chains; pass step0 = steps done)."""
    packed = tables if isinstance(tables, A.PackedTables) else A.PackedTables(tables)
    P, HB = param_layout(packed)
    ng = n_groups if n_groups else packed.n_lambda
    cfg, keep = chain_cfg(n_chains, seed, chain_id0, chain_lambda, chain_group, n_groups, init_state)
    o, s = alloc_outputs(n_chains, ng, packed.n_states, P, HB, n_steps, traj_every, record_state_trace)
    if resume is not None:
        for k in ("state", "indices", "energy", "accepted", "state_counts", "param_counts"):
            o[k][...] = resume[k]
    scfg = A.bcp_sample_cfg(int(n_steps), int(burn), int(traj_every), int(bool(record_state_trace)), 0)
    rc = lib().orc_sample(C.byref(packed.struct), C.byref(cfg), C.byref(scfg), int(step0), C.byref(s), int(n_threads))
    if rc != 0:
        raise RuntimeError("orc_sample failed (%d)" % rc)
    return o
