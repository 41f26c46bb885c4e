/*
 * TEST INFRASTRUCTURE ONLY -- the product (biceps_b200/) never links or loads this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker or the reported CPU baseline.
 *
 * Plain-C restatement of the reference's Metropolis sampler on the flat tables of
 * include/biceps_b200.h, with the Philox4x32-10 per-step draw budget of oracle/philox.py
 * and the deterministic exp of oracle/sampler_oracle.py (bcp_exp).  It recomputes the full
 * -log P at every step exactly as the reference does (no incremental caching):
 *
 *   PosteriorSampler.sample       /root/reference/biceps/PosteriorSampler.py:251-374
 *   PosteriorSampler.neglogP      /root/reference/biceps/PosteriorSampler.py:233-248
 *   Restraint_*.compute_neglogP   /root/reference/biceps/Restraint.py:356-383,457-477,572-592
 *
 * Pinning: tests/test_oracle_c.py checks it bit for bit against oracle/sampler_oracle.py
 * (PhiloxStepRNG + accept_bcp), which in turn is pinned against the live reference and the
 * golden fixtures under tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (see oracle/Makefile).
 * -ffp-contract=off matters: no FMA may be formed anywhere in the energy or exp arithmetic.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/biceps_b200.h"

#define M0 0xD2511F53u
#define M1 0xCD9E8D57u
#define W0 0x9E3779B9u
#define W1 0xBB67AE85u

static void philox(uint64_t seed, uint64_t chain, uint64_t step, uint32_t w[4]) {
    uint32_t c0 = (uint32_t)step, c1 = (uint32_t)(step >> 32), c2 = (uint32_t)chain, c3 = (uint32_t)(chain >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += W0; k1 += W1;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

/* oracle/sampler_oracle.py: bcp_exp */
static double orc_exp(double x) {
    static const double c[14] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0,
                                 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0,
                                 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0};
    const double LOG2E = 1.4426950408889634, LN2_HI = 6.93147180369123816490e-01,
                 LN2_LO = 1.90821492927058770002e-10, RND = 6755399441055744.0;
    volatile double t = x * LOG2E + RND;     /* volatile: forbid re-association of the rounding trick */
    double k = t - RND;
    double r = (x - k * LN2_HI) - k * LN2_LO;
    double p = c[13];
    for (int n = 12; n >= 0; --n) p = p * r + c[n];
    return ldexp(p, (int)k);                 /* exact: k in [-58, 0], no subnormals */
}

double orc_bcp_exp(double x) { return orc_exp(x); }

static int n_params_of(const bcp_restraint* r) {
    return r->kind == BCP_KIND_GAMMA ? 2 : (r->kind == BCP_KIND_PFGRID ? 1 + BCP_PF_NGRID : 1);
}

/* Restraint_pf on the 6-D nuisance grid, evaluated on the fly from <N_c>, <N_h>
 * (Restraint.py:742-749 sse, :765-801 model, :843-849 exponential reference, :873-894 term).
 * The reference sums the residues left to right; here -- and in the CUDA kernel, which splits the
 * residues over the 32 lanes of a warp -- the order is FIXED as: 32 strided partial sums
 * (j = l, l+32, ...) followed by an xor-butterfly (16, 8, 4, 2, 1).  The result differs from the
 * reference-order sum by rounding only (~1e-15 relative; pinned in tests/test_pf_oracle.py). */
static void pf_eval(const bcp_restraint* r, int n_states, int state, const int* ci, double* sse_out, double* ref_out) {
    (void)n_states;
    const int J = r->pf_n_obs;
    const int nxc = r->pf_n[3], nxh = r->pf_n[4], nb = r->pf_n[5];
    const double* g = r->pf_grid;
    const double bc = g[ci[0]], bh = g[r->pf_n[0] + ci[1]], b0 = g[r->pf_n[0] + r->pf_n[1] + ci[2]];
    const double* Nc = r->pf_Nc + (((size_t)state * nxc + ci[3]) * nb + ci[5]) * J;
    const double* Nh = r->pf_Nh + (((size_t)state * nxh + ci[4]) * nb + ci[5]) * J;
    double ps[32], pr[32], qs[32], qr[32];
    for (int l = 0; l < 32; ++l) { ps[l] = 0.0; pr[l] = 0.0; }
    for (int j = 0; j < J; ++j) {
        const int l = j & 31;
        const double model = (bc * Nc[j] + bh * Nh[j]) + b0;
        const double err = model - r->pf_exp[j];
        ps[l] = ps[l] + r->pf_weight[j] * (err * err);
        const double neg = -1.0 * model;
        pr[l] = pr[l] + r->pf_weight[j] * (neg > 0.0 ? neg : 0.0);
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) { qs[l] = ps[l] + ps[l ^ off]; qr[l] = pr[l] + pr[l ^ off]; }
        memcpy(ps, qs, sizeof(ps)); memcpy(pr, qr, sizeof(pr));
    }
    *sse_out = ps[0];
    *ref_out = pr[0];
}

static size_t pf_cell(const bcp_restraint* r, const int* ci) {
    size_t c = 0;
    for (int k = 0; k < BCP_PF_NGRID; ++k) c = c * (size_t)r->pf_n[k] + (size_t)ci[k];
    return c;
}

/* PosteriorSampler.neglogP (:233-248) + compute_neglogP (Restraint.py:374-383, 583-592) */
static double neglogp(const bcp_tables* t, int lam, int state, const int* idx) {
    double result = 0;
    result += t->energy[(size_t)lam * t->n_states + state] + t->logZ[lam];
    int p = 0;
    for (int q = 0; q < t->n_restraints; ++q) {
        const bcp_restraint* r = &t->restraints[q];
        double term = 0;
        const int sg = idx[p];
        term += r->nlsig[sg];
        if (r->kind == BCP_KIND_PFGRID) {
            double sse, ref;
            pf_eval(r, t->n_states, state, idx + p + 1, &sse, &ref);
            term += sse / r->two_sig2[sg];
            if (r->pf_prior) term += r->pf_prior[pf_cell(r, idx + p + 1)];
            term += r->cnorm;
            if (r->pf_ref_kind == 1) term -= ref;
            result += term;
            p += n_params_of(r);
            continue;
        }
        if (r->kind == BCP_KIND_GAMMA)
            term += r->sse[(size_t)state * r->n_gamma + idx[p + 1]] / r->two_sig2[sg];
        else
            term += r->sse[state] / r->two_sig2[sg];
        term += r->cnorm;
        if (r->has_ref) term -= r->refsum[state];
        result += term;
        p += n_params_of(r);
    }
    return result;
}

double orc_neglogp(const bcp_tables* t, int lam, int state, const int* idx) { return neglogp(t, lam, state, idx); }

typedef struct orc_job {
    const bcp_tables* t;
    int64_t n_chains, c_begin, c_end;
    uint64_t seed, chain_id0, step0;
    const int32_t* chain_lambda;
    const int32_t* chain_group;
    const int32_t* init_state;
    int64_t n_steps, burn, traj_every;
    int P, HB;
    int hist_off[BCP_MAX_PARAMS], glen[BCP_MAX_PARAMS], rest_index[BCP_MAX_PARAMS];
    /* outputs, same layouts as bcp_sample_out; histograms are per worker then merged */
    int64_t* state_counts; int64_t* param_counts;
    int64_t* accepted; int64_t* sep_accepted;
    int32_t* state; int32_t* indices; double* energy;
    double* traj_energy; int32_t* traj_state; uint8_t* traj_accept; int32_t* traj_indices;
    int32_t* state_trace;
    int n_groups;
} orc_job;

static void* worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    const bcp_tables* t = j->t;
    const int R = t->n_restraints, S = t->n_states, P = j->P;
    const uint32_t thresh = (uint32_t)(((uint64_t)R << 32) / (uint64_t)(R + 1));
    const int64_t n = j->n_chains;
    for (int64_t c = j->c_begin; c < j->c_end; ++c) {
        const uint64_t chain = j->chain_id0 + (uint64_t)c;
        const int lam = j->chain_lambda ? j->chain_lambda[c] : 0;
        const int grp = j->chain_group ? j->chain_group[c] : lam;
        int64_t* sc = j->state_counts + (size_t)grp * S;
        int64_t* pc = j->param_counts + (size_t)grp * j->HB;
        uint32_t w[4];
        int state, idx[BCP_MAX_PARAMS], nidx[BCP_MAX_PARAMS];
        double E;
        if (j->step0 == 0) {
            if (j->init_state) state = j->init_state[c];
            else { philox(j->seed, chain, 0xFFFFFFFFFFFFFFFFull, w); state = (int)(((uint64_t)w[1] * (uint32_t)S) >> 32); }
            for (int p = 0; p < P; ++p) idx[p] = j->glen[p] / 2;
            E = 1.0e99;
        } else {                                   /* continue a previous call */
            state = j->state[c]; E = j->energy[c];
            for (int p = 0; p < P; ++p) idx[p] = j->indices[c * P + p];
        }
        int64_t acc_total = 0;
        int64_t* sep = j->sep_accepted ? j->sep_accepted + c * (P + 1) : NULL;
        int64_t snap = 0;
        for (int64_t s = 0; s < j->n_steps + j->burn; ++s) {
            philox(j->seed, chain, j->step0 + (uint64_t)s, w);
            int nstate = state, pick = -1;
            memcpy(nidx, idx, sizeof(int) * P);
            if (w[0] < thresh) {
                uint64_t pr = (uint64_t)w[1] * (uint32_t)R;
                pick = (int)(pr >> 32);
                uint32_t f = (uint32_t)pr;
                for (int p = 0; p < P; ++p)
                    if (j->rest_index[p] == pick) {
                        uint64_t z = (uint64_t)f * 3u;
                        int v = idx[p] + (int)(z >> 32) - 1;
                        f = (uint32_t)z;
                        const int g = j->glen[p];
                        nidx[p] = ((v % g) + g) % g;
                    }
            } else {
                nstate = (int)(((uint64_t)w[1] * (uint32_t)S) >> 32);
            }
            const double En = neglogp(t, lam, nstate, nidx);
            int acc = 0;
            if (En < E) acc = 1;
            else {
                const double d = E - En;
                if (d >= -40.0) {
                    const uint64_t k = ((uint64_t)w[2] << 20) | (uint64_t)(w[3] >> 12);
                    const double u = (double)(int64_t)(2 * k + 1) * 0x1p-53;
                    acc = u < orc_exp(d);
                }
            }
            if (acc) {
                E = En; state = nstate; memcpy(idx, nidx, sizeof(int) * P);
                if (sep) {
                    if (pick < 0) sep[P] += 1;
                    else for (int p = 0; p < P; ++p) if (j->rest_index[p] == pick) sep[p] += 1;
                }
                ++acc_total;
            }
            if (s >= j->burn) {
                const int64_t ls = s - j->burn;
                __atomic_fetch_add(&sc[state], 1, __ATOMIC_RELAXED);
                for (int p = 0; p < P; ++p) __atomic_fetch_add(&pc[j->hist_off[p] + idx[p]], 1, __ATOMIC_RELAXED);
                if (j->state_trace) j->state_trace[(size_t)ls * n + c] = state;
                if (j->traj_every > 0 && ls % j->traj_every == 0) {
                    const size_t o = (size_t)snap * n + c;
                    if (j->traj_energy) j->traj_energy[o] = E;
                    if (j->traj_state) j->traj_state[o] = state;
                    if (j->traj_accept) j->traj_accept[o] = (uint8_t)acc;
                    if (j->traj_indices)
                        for (int p = 0; p < P; ++p) j->traj_indices[((size_t)snap * P + p) * n + c] = idx[p];
                    ++snap;
                }
            }
        }
        if (j->accepted) j->accepted[c] += acc_total;
        if (j->state) j->state[c] = state;
        if (j->energy) j->energy[c] = E;
        if (j->indices) for (int p = 0; p < P; ++p) j->indices[c * P + p] = idx[p];
    }
    return NULL;
}

/* Same contract as bcp_chains_create + bcp_sample, on the host, `n_threads` pthreads over
 * chains.  step0 = number of steps these chains already ran (0: fresh chains; >0: state /
 * indices / energy are read as the starting point).  Histograms accumulate into the given
 * arrays (caller zeroes them).  Returns 0. */
int orc_sample(const bcp_tables* t, const bcp_chain_cfg* cc, const bcp_sample_cfg* sc, uint64_t step0,
               bcp_sample_out* out, int n_threads) {
    if (!t || !cc || !sc || !out) return -1;
    orc_job base;
    memset(&base, 0, sizeof(base));
    base.t = t; base.n_chains = cc->n_chains; base.seed = cc->seed; base.chain_id0 = cc->chain_id0;
    base.step0 = step0; base.chain_lambda = cc->chain_lambda; base.chain_group = cc->chain_group;
    base.init_state = cc->init_state;
    base.n_steps = sc->n_steps; base.burn = sc->burn; base.traj_every = sc->traj_every;
    base.n_groups = cc->n_groups > 0 ? cc->n_groups : t->n_lambda;
    int P = 0, HB = 0;
    for (int q = 0; q < t->n_restraints; ++q) {
        const bcp_restraint* r = &t->restraints[q];
        if (r->kind != BCP_KIND_SCALAR && r->kind != BCP_KIND_GAMMA && r->kind != BCP_KIND_PFGRID) return -1;
        base.rest_index[P] = q; base.glen[P] = r->n_sigma; base.hist_off[P] = HB; HB += r->n_sigma; ++P;
        if (r->kind == BCP_KIND_GAMMA) { base.rest_index[P] = q; base.glen[P] = r->n_gamma; base.hist_off[P] = HB; HB += r->n_gamma; ++P; }
        if (r->kind == BCP_KIND_PFGRID)
            for (int k = 0; k < BCP_PF_NGRID; ++k) {
                if (P >= BCP_MAX_PARAMS) return -1;
                base.rest_index[P] = q; base.glen[P] = r->pf_n[k]; base.hist_off[P] = HB; HB += r->pf_n[k]; ++P;
            }
    }
    base.P = P; base.HB = HB;
    int64_t* sc_tmp = NULL; int64_t* pc_tmp = NULL;
    base.state_counts = out->state_counts; base.param_counts = out->param_counts;
    if (!base.state_counts) base.state_counts = sc_tmp = (int64_t*)calloc((size_t)base.n_groups * t->n_states, 8);
    if (!base.param_counts) base.param_counts = pc_tmp = (int64_t*)calloc((size_t)base.n_groups * HB, 8);
    base.accepted = out->accepted; base.sep_accepted = out->sep_accepted;
    base.state = out->state; base.indices = out->indices; base.energy = out->energy;
    base.traj_energy = out->traj_energy; base.traj_state = out->traj_state; base.traj_accept = out->traj_accept;
    base.traj_indices = out->traj_indices; base.state_trace = sc->record_state_trace ? out->state_trace : NULL;
    if (step0 > 0 && (!base.state || !base.indices || !base.energy)) return -1;
    if (out->sep_accepted) memset(out->sep_accepted, 0, sizeof(int64_t) * (size_t)cc->n_chains * (P + 1));
    if (n_threads < 1) n_threads = 1;
    if (n_threads > cc->n_chains) n_threads = (int)cc->n_chains;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    orc_job* jobs = (orc_job*)malloc(sizeof(orc_job) * n_threads);
    for (int i = 0; i < n_threads; ++i) {
        jobs[i] = base;
        jobs[i].c_begin = cc->n_chains * i / n_threads;
        jobs[i].c_end = cc->n_chains * (i + 1) / n_threads;
        if (n_threads == 1) worker(&jobs[i]);
        else pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    if (n_threads > 1) for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    if (out->total) for (int64_t c = 0; c < cc->n_chains; ++c) out->total[c] = (int64_t)(step0 + sc->n_steps + sc->burn);
    free(th); free(jobs); free(sc_tmp); free(pc_tmp);
    return 0;
}
