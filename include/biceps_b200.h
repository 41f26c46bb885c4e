/*
 * biceps_b200 -- C ABI of the B200-native BICePs posterior-sampling hot path.
 *
 * Plain C, opaque handles, caller-owned buffers, no exceptions: every function returns
 * 0 on success or a negative bcp_status; the message is in bcp_last_error() (thread local).
 * The reference (vvoelz/biceps v2.0) is pure Python and has no FFI of its own; each entry
 * point below names the reference function it replaces, as file:line under
 * /root/reference/biceps/.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Memory spaces: arguments called "host" must be host pointers; arguments of the *_dev
 * entry points must be device pointers on the handle's device.  The library never keeps a
 * caller pointer after the call returns.
 *
 * All floating point is IEEE fp64.  All tables are row-major and contiguous.
 */
#ifndef BICEPS_B200_H
#define BICEPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCP_ABI_VERSION 1

typedef enum bcp_status {
    BCP_OK = 0,
    BCP_ERR_INVALID = -1,   /* bad argument (NULL, size, kind) */
    BCP_ERR_CUDA = -2,      /* CUDA runtime error, see bcp_last_error() */
    BCP_ERR_NOMEM = -3,
    BCP_ERR_NOCONV = -4,    /* MBAR did not converge in max_iter */
    BCP_ERR_STATE = -5      /* call order (e.g. fetch before sample) */
} bcp_status;

/* restraint kinds: which compute_neglogP the term follows */
#define BCP_KIND_SCALAR 0   /* Restraint_cs / Restraint_J / Restraint_pf(precomputed): one sigma,
                               sse[S]                      Restraint.py:356-383, 457-477, 881 */
#define BCP_KIND_GAMMA 1    /* Restraint_noe: sigma + gamma, sse[S][G]     Restraint.py:572-592 */
#define BCP_KIND_PFGRID 2   /* Restraint_pf on the 6-D nuisance grid       Restraint.py:862-894 */

#define BCP_MAX_RESTRAINTS 8
#define BCP_PF_NGRID 6

/* One restraint type of the ensemble, flattened over states.
 *   term(i, idx) = ((nlsig[s] + SSE / two_sig2[s]) [+ prior[cell]] + cnorm) [- REF]
 * with s = sigma index.  nlsig = Ndof*ln(sigma), two_sig2 = 2*sigma^2 and
 * cnorm = Ndof/2*ln(2 pi) are evaluated by the caller with the reference's own expressions
 * (Restraint.py:376-378) so that energies are reproducible to the last bit.                */
typedef struct bcp_restraint {
    int32_t kind;
    int32_t n_sigma;
    int32_t n_gamma;            /* GAMMA: G; otherwise 1 */
    int32_t has_ref;            /* 1 if refsum is subtracted (ref == exponential|gaussian) */
    double cnorm;
    const double* nlsig;        /* [n_sigma] */
    const double* two_sig2;     /* [n_sigma] */
    const double* sse;          /* SCALAR: [S]   GAMMA: [S][G]   PFGRID: NULL (see pf_*) */
    const double* refsum;       /* SCALAR/GAMMA: [S] or NULL.  PFGRID: NULL (computed on the fly) */
    /* ---- PFGRID only (ignored otherwise) ------------------------------------------- */
    int32_t pf_n[BCP_PF_NGRID]; /* grid lengths: beta_c, beta_h, beta_0, xcs, xhs, bs */
    int32_t pf_n_obs;           /* J = number of residues */
    int32_t pf_ref_kind;        /* 0 uniform, 1 exponential (Restraint.py:843-849) */
    const double* pf_grid;      /* concatenated grid values, sum(pf_n) doubles */
    const double* pf_Nc;        /* [S][n_xcs][n_bs][J]  <N_c> */
    const double* pf_Nh;        /* [S][n_xhs][n_bs][J]  <N_h> */
    const double* pf_exp;       /* [J] */
    const double* pf_weight;    /* [J] */
    const double* pf_prior;     /* [prod(pf_n)] or NULL */
} bcp_restraint;

typedef struct bcp_tables {
    int32_t n_states;           /* S */
    int32_t n_lambda;           /* K: number of lambda ensembles sharing the restraint tables */
    int32_t n_restraints;       /* R <= BCP_MAX_RESTRAINTS */
    int32_t reserved;
    const double* energy;       /* [K][S]  lambda_k * f_i  (Restraint.py:99) */
    const double* logZ;         /* [K]     PosteriorSampler.py:64-71 */
    const bcp_restraint* restraints; /* [R] in ensemble order */
} bcp_tables;

typedef struct bcp_handle bcp_handle;   /* tables resident in HBM on one device */
typedef struct bcp_chains bcp_chains;   /* a block of independent Metropolis chains */

int bcp_abi_version(void);
const char* bcp_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench gpu_launches) */
uint64_t bcp_kernel_launches(void);

/* Upload the flattened ensemble (host pointers) to `device`.  Replaces the table reads of
 * PosteriorSampler.neglogP (PosteriorSampler.py:233-248). */
int bcp_create(const bcp_tables* tables, int device, bcp_handle** out);
void bcp_destroy(bcp_handle* h);
int bcp_n_params(const bcp_handle* h);      /* P: sigma [+ gamma | + 6 PF grids] per restraint */
int bcp_param_bins(const bcp_handle* h);    /* sum of grid lengths over the P parameters */
/* rest_index[p] and grid_len[p] in the reference's parameter order (PosteriorSampler.py:275-281) */
int bcp_param_layout(const bcp_handle* h, int32_t* rest_index, int32_t* grid_len);

/* ---------------------------------------------------------------------------------------
 * Chains.  Chain c has global id chain_id0 + c; its RNG stream is Philox4x32-10 keyed by
 * `seed` with counter (step, chain id) -- oracle/philox.py is the specification -- so a
 * trajectory does not depend on how chains are split over launches, blocks or GPUs.
 * Initial condition as in the reference: state = RNG draw (PosteriorSampler.py:29) unless
 * init_state is given, indices = int(len(grid)/2) (Restraint.py:240), E = 1e99 (:30).
 * Histograms are accumulated per *group* (default: the chain's lambda index).          */
typedef struct bcp_chain_cfg {
    int64_t n_chains;
    uint64_t seed;
    uint64_t chain_id0;
    int32_t n_groups;               /* histogram groups; 0 -> n_lambda */
    int32_t reserved;
    const int32_t* chain_lambda;    /* host [n_chains] or NULL (all 0) */
    const int32_t* chain_group;     /* host [n_chains] or NULL (= chain_lambda) */
    const int32_t* init_state;      /* host [n_chains] or NULL (Philox draw) */
} bcp_chain_cfg;

int bcp_chains_create(bcp_handle* h, const bcp_chain_cfg* cfg, bcp_chains** out);
void bcp_chains_destroy(bcp_chains* c);

/* One PosteriorSampler.sample(nsteps, burn) call (PosteriorSampler.py:251-374) for every
 * chain of the block.  May be called repeatedly: state, E, indices, accepted, total and the
 * Philox step counter persist, histograms keep accumulating (like calling sample() twice).
 * Snapshots are taken at local steps burn, burn+traj_every, ... (:354).                  */
typedef struct bcp_sample_cfg {
    int64_t n_steps;
    int64_t burn;
    int64_t traj_every;             /* 0: no snapshots */
    int32_t record_state_trace;     /* 1: keep state_trace (4 B/step/chain, :345) */
    int32_t reserved;
} bcp_sample_cfg;

/* Host output buffers; any pointer may be NULL to skip that output.
 * n_snap = traj_every ? ceil(n_steps / traj_every) : 0.                                    */
typedef struct bcp_sample_out {
    /* cumulative over all sample calls of the block */
    int64_t* state_counts;      /* [n_groups][S]   WITHOUT the reference's +1 pseudo-count (:393) */
    int64_t* param_counts;      /* [n_groups][param_bins]  sampled_parameters (:347-348) */
    int64_t* accepted;          /* [n_chains] (:336) */
    int64_t* total;             /* [n_chains] (:337) */
    /* this call only */
    int64_t* sep_accepted;      /* [n_chains][P+1]  per-parameter + state acceptances (:334-335) */
    /* chain state after the call */
    int32_t* state;             /* [n_chains] */
    int32_t* indices;           /* [n_chains][P] */
    double* energy;             /* [n_chains] */
    /* snapshots of this call, snapshot-major */
    double* traj_energy;        /* [n_snap][n_chains] */
    int32_t* traj_state;        /* [n_snap][n_chains] */
    uint8_t* traj_accept;       /* [n_snap][n_chains] */
    int32_t* traj_indices;      /* [n_snap][P][n_chains] */
    int32_t* state_trace;       /* [n_steps][n_chains] if record_state_trace */
} bcp_sample_out;

/* synchronous: launch, wait, copy results to the host buffers */
int bcp_sample(bcp_chains* c, const bcp_sample_cfg* cfg, bcp_sample_out* out);
/* asynchronous on `stream` (a cudaStream_t, NULL = default stream); results stay in HBM */
int bcp_sample_launch(bcp_chains* c, const bcp_sample_cfg* cfg, void* stream);
int bcp_chains_sync(bcp_chains* c);
int bcp_chains_fetch(bcp_chains* c, bcp_sample_out* out);
/* device views of the cumulative histograms, for an NCCL all-reduce by the caller */
int bcp_chains_device_hist(bcp_chains* c, int64_t** d_state_counts, int64_t** d_param_counts);
/* elapsed device time (ms, CUDA events on the launch stream) of the last sample launch */
int bcp_chains_last_kernel_ms(bcp_chains* c, float* ms, int32_t* n_launches);

/* -log P of explicit configurations (PosteriorSampler.neglogP, :233-248): for n
 * configurations, lambda index lam[n], state[n], indices[n][P] -> energy_out[n].         */
int bcp_neglogp(bcp_handle* h, int64_t n, const int32_t* lam, const int32_t* state,
                const int32_t* indices, double* energy_out);

#ifdef __cplusplus
}
#endif
#endif /* BICEPS_B200_H */
