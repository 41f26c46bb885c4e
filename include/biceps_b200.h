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

#include <stddef.h>
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
#define BCP_MAX_PARAMS 24      /* nuisance parameters over all restraints */

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

/* Page-locked host memory for the caller's tables and output buffers (optional; any host pointer
 * works, pinned ones copy at the full PCIe rate).  NULL on failure. */
void* bcp_host_alloc(size_t bytes);
void bcp_host_free(void* p);

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
    int32_t rng_mode;               /* 0: Philox (default); 1: numpy's legacy MT19937 stream, ONE chain, draw for draw
                                       like the reference (PosteriorSampler.py:298-307, 322-326); see
                                       bcp_chains_set_mt_stream.  init_state should then be the reference's own
                                       np.random.randint(nstates) draw (PosteriorSampler.py:29) */
    const int32_t* chain_lambda;    /* host [n_chains] or NULL (all 0) */
    const int32_t* chain_group;     /* host [n_chains] or NULL (= chain_lambda) */
    const int32_t* init_state;      /* host [n_chains] or NULL (Philox draw) */
} bcp_chain_cfg;

int bcp_chains_create(bcp_handle* h, const bcp_chain_cfg* cfg, bcp_chains** out);
void bcp_chains_destroy(bcp_chains* c);

/* Reference-stream mode (rng_mode == 1): the raw 32-bit outputs of numpy's MT19937 from the generator's current
 * position (RandomState._bit_generator.random_raw), at least ~12 per step.  The chain consumes them exactly as the
 * reference's calls to np.random.random / randint would; bcp_chains_mt_consumed tells how many were used (so that
 * the caller can advance the global generator by as many) and fails with BCP_ERR_STATE if the stream ran out.  */
int bcp_chains_set_mt_stream(bcp_chains* c, const uint32_t* raw, int64_t n_raw);
int bcp_chains_mt_consumed(bcp_chains* c, int64_t* n_consumed);

/* One PosteriorSampler.sample(nsteps, burn) call (PosteriorSampler.py:251-374) for every
 * chain of the block.  May be called repeatedly: state, E, indices, accepted, total and the
 * Philox step counter persist, histograms keep accumulating (like calling sample() twice).
 * Snapshots are taken at local steps burn, burn+traj_every, ... (:354).                  */
typedef struct bcp_sample_cfg {
    int64_t n_steps;
    int64_t burn;
    int64_t traj_every;             /* 0: no snapshots */
    int32_t record_state_trace;     /* 1: keep state_trace (4 B/step/chain, :345) */
    int32_t trace_chains;           /* 0: trace every chain; k > 0: only chains 0 .. k-1 (the reference keeps one
                                       trace per sampler; a drop-in sampler with many chains traces chain 0) */
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
    int32_t* state_trace;       /* [n_steps][traced chains] if record_state_trace (traced chains = trace_chains or n_chains) */
} bcp_sample_out;

/* synchronous: launch, wait, copy results to the host buffers */
int bcp_sample(bcp_chains* c, const bcp_sample_cfg* cfg, bcp_sample_out* out);
/* asynchronous on `stream` (a cudaStream_t, NULL = default stream); results stay in HBM */
int bcp_sample_launch(bcp_chains* c, const bcp_sample_cfg* cfg, void* stream);
int bcp_chains_sync(bcp_chains* c);
int bcp_chains_fetch(bcp_chains* c, bcp_sample_out* out);
/* Overwrite the chains' state / parameter indices [n_chains][P] / energy (host arrays, NULL =
 * keep).  Restart from a checkpoint taken with bcp_chains_fetch, or reproduce what the reference
 * does when sample() is called a second time (indices jump back to int(len/2) while state and E
 * persist, PosteriorSampler.py:271-278, 294). */
int bcp_chains_set_state(bcp_chains* c, const int32_t* state, const int32_t* indices, const double* energy);
/* device views of the cumulative histograms, for an NCCL all-reduce by the caller; the two arrays are ONE block
 * (d_param_counts == d_state_counts + n_groups * n_states), so one collective over n_groups * (n_states +
 * bcp_param_bins) int64 sums both (replaces the per-process histograms of PosteriorSampler.py:339-358) */
int bcp_chains_device_hist(bcp_chains* c, int64_t** d_state_counts, int64_t** d_param_counts);
/* elapsed device time (ms, CUDA events on the launch stream) of the last sample launch */
int bcp_chains_last_kernel_ms(bcp_chains* c, float* ms, int32_t* n_launches);
/* which Metropolis kernel the last launch ran ("tab", "ws", "spec", "fast", "coop", "generic", "pf2", "pf", "mt"; "none"
 * before the first launch): the choice is automatic (chain count, layout, acceptance) and invisible in the results */
const char* bcp_chains_last_kernel(bcp_chains* c);

/* -log P of explicit configurations (PosteriorSampler.neglogP, :233-248): for n
 * configurations, lambda index lam[n], state[n], indices[n][P] -> energy_out[n].         */
int bcp_neglogp(bcp_handle* h, int64_t n, const int32_t* lam, const int32_t* state,
                const int32_t* indices, double* energy_out);

/* ---------------------------------------------------------------------------------------
 * Restraint-energy precompute.
 * Raw observables of ONE restraint type over all states.                                 */
typedef struct bcp_obs {
    int32_t kind;                   /* BCP_KIND_SCALAR or BCP_KIND_GAMMA */
    int32_t n_obs;                  /* J */
    int32_t ref_kind;               /* 0 uniform, 1 exponential, 2 gaussian */
    int32_t log_normal;             /* GAMMA: err = ln(model/(gamma*exp))      Restraint.py:561-562 */
    int32_t use_global_ref_sigma;   /* gaussian ref: one sigma for all j       PosteriorSampler.py:139-143 */
    int32_t n_gamma;                /* GAMMA: G */
    const double* model;            /* [S][J] forward-model observables r_j(X_i) */
    const double* exp;              /* [J] experimental values */
    const double* weight;           /* [J] 1/n per equivalency group           Restraint.py:435-442 */
    const double* gamma;            /* [G] */
} bcp_obs;

/* sse[S] (SCALAR: Restraint.py:344-353, 445-454, 736-741) or sse[S][G] (GAMMA: :552-569), and,
 * if ref_kind != 0, refsum[S] = sum_j w_j * (-ln P_ref) with its per-observable parameters:
 *   exponential: ref_par0 = beta_j            PosteriorSampler.py:74-104, Restraint.py:275-286
 *   gaussian   : ref_par0 = mean_j, ref_par1 = sigma_j   PosteriorSampler.py:107-150, Restraint.py:288-298
 * Host pointers in, host pointers out (copies included).  Results agree with the reference to
 * ~1e-13 relative (reduction order and the gamma moment expansion differ), tolerance 1e-6.  */
int bcp_precompute(const bcp_obs* obs, int32_t n_states, int device, double* sse, double* refsum,
                   double* ref_par0, double* ref_par1);
/* same with every pointer (inside obs too) a device pointer; asynchronous on `stream` (scratch comes
 * from the stream-ordered pool, nothing synchronises) */
int bcp_precompute_dev(const bcp_obs* obs, int32_t n_states, int device, void* stream, double* d_sse,
                       double* d_refsum, double* d_ref_par0, double* d_ref_par1);
/* logZ[k] = ln sum_i exp(-energy[k][i])   (PosteriorSampler.py:64-71); host pointers */
int bcp_logz(const double* energy, int32_t n_lambda, int32_t n_states, int device, double* logZ);

/* ---------------------------------------------------------------------------------------
 * Analysis: rescoring and MBAR.
 * u_kn[l][n] = -log P of snapshot n (state, indices) under lambda ensemble l of the handle,
 * i.e. sampler[l].neglogP(...) of Analysis.py:165-183 for every l; layout [K][n].            */
int bcp_ukn(bcp_handle* h, int64_t n, const int32_t* state, const int32_t* indices, double* u_kn);
int bcp_ukn_dev(bcp_handle* h, int64_t n, const int32_t* d_state, const int32_t* d_indices, double* d_u_kn,
                void* stream);

/* MBAR(u_kn, N_k) + compute_free_energy_differences + compute_expectations of state indicators
 * (pymbar calls at Analysis.py:192, 201, 217-223; uncertainty_method='approximate').
 *   u_kn  [K][N]  reduced potentials, samples of run 0 first, then run 1, ... (kln_to_kn order)
 *   N_k   [K]     samples per run (sum = N)
 *   state_n [N]   conformational state of each sample (or NULL: no populations), S states
 * Outputs: f_k [K] (f_0 = 0; BICePs score of lambda_l is f_l - f_0), theta [K][K] = W^T W
 * (dDelta_f_ij = sqrt(theta_ii + theta_jj - 2 theta_ij)), pops / dpops [S][K], iteration count.
 * tol <= 0 -> 1e-12 (max |delta f| relative to max(1, |f|)); max_iter <= 0 -> 10000.          */
int bcp_mbar(const double* u_kn, const int64_t* N_k, int32_t K, int64_t N, const int32_t* state_n, int32_t S,
             int device, double tol, int32_t max_iter, double* f_k, double* theta, double* pops, double* dpops,
             int32_t* iters);
/* same with u_kn / state_n resident in HBM (N_k and the outputs stay host pointers) */
int bcp_mbar_dev(const double* d_u_kn, const int64_t* N_k, int32_t K, int64_t N, const int32_t* d_state_n, int32_t S,
                 int device, void* stream, double tol, int32_t max_iter, double* f_k, double* theta, double* pops,
                 double* dpops, int32_t* iters);
/* The whole Analysis step on the device-resident snapshots of the chains' last sample() call (no file or
 * host round trip): every chain's snapshots are samples of its lambda run
 * (chains of one run concatenated in chain order = Analysis.py:214 kln_to_kn order, several seeds per
 * lambda allowed), u_kn by rescoring under every lambda (Analysis.py:165-183), then MBAR as bcp_mbar.
 * N_k_out [K] (may be NULL) receives the samples per run; pops / dpops [S][K] only if want_pops.  */
int bcp_chains_score(bcp_chains* c, int32_t want_pops, double tol, int32_t max_iter, int64_t* N_k_out,
                     double* f_k, double* theta, double* pops, double* dpops, int32_t* iters);
/* Building blocks of the same step for a block that holds a SHARD of the chains (multi-GPU score pipeline,
 * biceps_b200/distributed.py:score_chains): the number of local samples and the local samples per run, and the
 * local u_kn [K][n_samples] / state_n [n_samples] (may be NULL) written to caller-owned device buffers in the
 * sample order described above.  Replaces the u_kln loop of Analysis.py:145-185 for the local trajectories. */
int bcp_chains_n_samples(bcp_chains* c, int64_t* n_samples, int64_t* N_k_local);
int bcp_chains_ukn_dev(bcp_chains* c, double* d_u_kn, int32_t* d_state_n, void* stream);
/* One streaming reduction over a local shard of samples (multi-GPU MBAR: the caller all-reduces
 * s and M across ranks between passes): s[K] = sum_n W_nk, M[K][K] = sum_n W_nk W_nl.        */
int bcp_mbar_pass_dev(const double* d_u_kn, int64_t n_local, int32_t K, const int64_t* N_k_global,
                      const double* f_k, int32_t want_hess, int device, void* stream, double* s_out, double* M_out);

/* Weighted state histograms of a local shard at free energies f_k (compute_expectations of the state indicators,
 * Analysis.py:217-223, before the sum over shards): P[K][S] = sum_n W_nl [s_n = i], Q[K][S] = sum_n W_nl^2 [s_n = i];
 * host outputs.  After an all-reduce of both: P_i(l) = P, dP_i(l) = |P| sqrt(Q / P^2 + theta_ll - 2 Q / P).          */
int bcp_mbar_pops_dev(const double* d_u_kn, int64_t n_local, int32_t K, const int64_t* N_k_global, const double* f_k,
                      const int32_t* d_state_n, int32_t S, int device, void* stream, double* P_out, double* Q_out);

/* ---------------------------------------------------------------------------------------
 * Convergence diagnostics of the sampled traces.
 * Autocorrelation curves of n_series time series of length T (row-major [n_series][T]):
 *   curves[s][tau] = sum_{t=0}^{T-2-tau} z[t] z[t+tau] / (T - tau),  z = f - mean(f),  tau = 0 .. max_tau,
 * divided by curves[s][0] if normalize (convergence.py:93-111, g(); the last sample is left out of both
 * factors exactly like the reference does), and tau[s] = sum_tau curves[s][tau] (compute_autocorrelation_time,
 * :114-125; may be NULL).  Host pointers; the *_dev variant takes device pointers and is asynchronous on
 * `stream`.  Agrees with the reference's np.dot to ~1e-13 of curves[s][0] (summation order).            */
int bcp_autocorr(const double* series, int32_t n_series, int64_t T, int32_t max_tau, int32_t normalize,
                 int device, double* curves, double* tau);
int bcp_autocorr_dev(const double* d_series, int32_t n_series, int64_t T, int32_t max_tau, int32_t normalize,
                     int device, void* stream, double* d_curves, double* d_tau);

/* The same for every nuisance-parameter trace of every chain of the block, straight from the device-resident
 * snapshots of the last sample() call (no host round trip): series (chain, p) = grid_values[offset_p + index], with
 * grid_values the concatenated parameter grids (bcp_param_bins doubles, bcp_param_layout order).
 * curves [n_chains][P][max_tau+1], tau [n_chains][P] (may be NULL); host pointers.                                 */
int bcp_chains_autocorr(bcp_chains* c, const double* grid_values, int32_t max_tau, int32_t normalize,
                        double* curves, double* tau);

/* One Jensen-Shannon divergence (compute_JSD, convergence.py:145-186) of a parameter-index series
 * idx[offset .. offset + n_total) split into a first part of n_first samples and the rest:
 *   mode 0  first part = the first n_first samples            (Convergence.process :503-507)
 *   mode 1  first part = positions sel[sel_offset .. + n_first) (the caller's np.random.choice draw, :509)
 *   mode 2  first part = the n_first positions with the smallest keys (w0 << 32 | position), w0 the first word
 *           of Philox4x32-10(counter = (position, stream), key = seed): a device-side random half          */
typedef struct bcp_jsd_task {
    int64_t offset;
    int64_t n_total;
    int64_t n_first;
    int64_t sel_offset;
    uint64_t stream;
    int32_t n_bins;             /* length of the parameter's grid (<= 2048) */
    int32_t mode;
} bcp_jsd_task;
int bcp_jsd(const int32_t* idx, int64_t n_idx, const int32_t* sel, int64_t n_sel, const bcp_jsd_task* tasks,
            int64_t n_tasks, uint64_t seed, int device, double* jsd_out);
/* counts[task][max_bins] = histogram of idx[offset .. offset + n_total) (block averaging, :462-480) */
int bcp_param_hist(const int32_t* idx, int64_t n_idx, const bcp_jsd_task* tasks, int64_t n_tasks,
                   int32_t max_bins, int device, int64_t* counts);

#ifdef __cplusplus
}
#endif
#endif /* BICEPS_B200_H */
