// Device-side view of a flattened ensemble (bcp_handle), shared by the sampler and the
// analysis kernels.
#pragma once
#include "common.cuh"

struct bcp_handle;

namespace bcp {

constexpr int MAXR = BCP_MAX_RESTRAINTS;

struct RestraintDev {
    const double* sse;      // [S][G]
    const double* refsum;   // [S] or nullptr
    double cnorm;
    int sig_off;            // doubles into the staged sigma table: nlsig, then two_sig2
    int n_sigma;
    int n_gamma;            // 1 if the restraint has no gamma
    int hist_sigma;         // bin offset of this restraint's sigma histogram
    int hist_gamma;         // bin offset of its gamma histogram (if has_gamma)
    int has_gamma;          // Restraint_noe: gamma moves together with sigma
};

// Restraint_pf on the 6-D nuisance grid (at most one per ensemble): raw <N_c>/<N_h> counts, the term
// is evaluated on the fly (the dense table would be 11.6 MB per state)
struct PfDev {
    const double* grid;     // concatenated grid values beta_c, beta_h, beta_0, xcs, xhs, bs
    const double* Nc;       // [S][n_xcs][n_bs][J]
    const double* Nh;       // [S][n_xhs][n_bs][J]
    const double* exp;      // [J]
    const double* weight;   // [J]
    const double* prior;    // [prod n] or nullptr
    int n[BCP_PF_NGRID];
    int hist[BCP_PF_NGRID]; // histogram bin offsets of the six grid parameters
    int J;
    int ref_kind;           // 0 uniform, 1 exponential
    int q;                  // restraint index, -1 if the ensemble has none
    int pad;
};

struct SamplerTables {
    RestraintDev r[MAXR];
    const double* energy;   // [K][S]
    const double* logZ;     // [K]
    const double* sig_tab;  // staged table in global memory (16 B aligned, padded)
    int sig_bytes;          // multiple of 16
    int R, S, K, P, HB;
    uint32_t nuis_thresh;   // floor(R * 2^32 / (R+1))
    // canonical layout (reference file-extension order: cs, J, noe): every restraint is SCALAR
    // except, optionally, the last one.  rec[i] packs the state-dependent words of state i so
    // that a state move touches one or two sectors: {sse_q, ref_q} for every scalar restraint,
    // then {ref_g, 0} of the gamma restraint (ref = +0.0 where there is no reference potential;
    // t - 0.0 == t bit for bit).
    PfDev pf;
    const double* rec;      // [S][rec_stride]
    // gamma restraint's sse rows with a 2-element periodic halo on each side, so that the +-2
    // window around any gamma index is contiguous and needs no wrap arithmetic:
    // gsse_halo[i][k] = sse[i][(k - 2) mod G], k in [0, G+4)
    const double* gsse_halo;
    // for the speculative kernel (sampler_spec.cu): the same rows with a 4-element periodic halo,
    // gsse_halo4[i][k] = sse[i][(k - 4) mod G], k in [0, gh4_stride), gh4_stride = G + 9 rounded up
    // to even (16-byte aligned rows; any 9-wide window around a gamma index, widened to an aligned
    // 10-wide one, stays inside the row)
    const double* gsse_halo4;
    int gh4_stride;
    // packed sigma entries with a 2-element periodic halo: restraint q owns entries
    // [ent_off, ent_off + n_sigma + 4), entry e <-> sigma index e - 2 (mod n_sigma), each entry
    // the three doubles {Ndof ln sigma, 2 sigma^2, RN(1 / (2 sigma^2))} (24-byte stride: neighbouring
    // entries fall into different shared-memory banks)
    int sig4_bytes;         // multiple of 16
    const double* sig4;
    int ent_off[MAXR];
    int div_safe;           // every sse / 2 sigma^2 is far from under/overflow (Markstein division is exact)
    int pad2;
    int rec_stride;         // 2 * R doubles
    int canonical;          // 1 if the specialised kernel applies
    // table-driven kernel (sampler_tab.cu): lb[l][i] <= -log P(state i, any nuisance parameters) under lambda l, minus a
    // safety far above the rounding of either side (canonical layouts with div_safe only, else nullptr)
    const double* lb;       // [K][S]
};

const SamplerTables& handle_tables(const bcp_handle* h);
int handle_device(const bcp_handle* h);
// energy transposed to [S][K] (built by bcp_create): the K energies of a state are contiguous for the rescoring kernels
const double* handle_energy_t(const bcp_handle* h);

}  // namespace bcp
