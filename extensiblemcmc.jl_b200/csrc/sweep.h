// Host-side interface of the likelihood sweep kernels (sweep_gsn1d.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace extmcmc {

// Which launches carry the programmatic-dependent-launch attribute: bit 0 sweep, bit 1 accept
// (EXTMCMC_PDL environment variable).  Default 2: only the accept kernel starts early (its
// prologue hides behind the sweep).  Letting the sweep start early too was measured 45 % SLOWER at
// cfg 2: its CTAs land on SMs still holding accept blocks, the grid no longer fits in one
// balanced wave of 4 CTAs per SM, and the second wave doubles the tail.
int pdl_mask();

enum { SWEEP_VARIANT_AUTO = 0, SWEEP_VARIANT_CHAINS = 1, SWEEP_VARIANT_OBS = 2 };

struct SweepPlan {
    int variant;       // SWEEP_VARIANT_*
    int R;             // chains per thread ("chains") or chains per CTA pass ("obs")
    int groups;        // chain groups (grid.y for "chains", launches for "obs")
    int S;             // observation segments = rows of partial[S][C] (logistic: partial slots per chain block)
    int launches;      // kernel launches per sweep
    int D;             // observation dimension (general-d Gaussian law)
    int G;             // observation groups (hierarchical law; 1 otherwise)
    int n_cta;         // logistic: persistent CTAs
    const char *name;
};

// Arguments of the 1-D Gaussian sweeps.  Observations of group g occupy
// obs[goff[g] .. goff[g] + glen[g]) (goff even, buffer readable up to the next even index);
// mu[g][C] are the per-chain means; partial[q][G*S][C] receives q = 0: sum (x - mu)^2 and,
// with grad, q = 1: sum (x - mu).
struct Gsn1dArgs {
    const double *obs;
    const int64_t *goff;
    const int64_t *glen;
    int G;
    const double *mu;
    int64_t C;
    double *partial;
    int S;
    // Optional fused tail of the "obs" mapping (G = 1, no gradient): the last CTA to finish adds
    // the per-segment sums in a fixed order and either stores the totals to ssum[C] (tail_mode 1)
    // or pushes them straight into every rank's exchange buffer over NVLink peer mappings and
    // raises this rank's sequence flag there (tail_mode 2) -- compute and collective in one kernel.
    int tail_mode;
    unsigned int *tail_counter;
    double *ssum;
    unsigned long long **peer_rx;   // [world] -> rx[2][world][C][2] of each rank (exchange.cuh)
    int rank, world;
    const void *descs;       // StepDesc array of the block (device), element k names the step
    int k;
    int stream_hint;         // 1: the observations do not fit in L2 -- copy them with L2::evict_first
};

SweepPlan plan_sweep_gsn1d(int64_t C, int64_t n_obs_largest_group, int force_variant, int num_sms, int G);
cudaError_t sweep_gsn1d_init();
void launch_sweep_gsn1d(const SweepPlan &pl, const Gsn1dArgs &a, bool grad, cudaStream_t st);

// General-d GsnTargetLaw: lawc = [mu(d), W = inv(chol(Sigma)) lower-tri row-major, c0],
// obs row-major [n_obs][d], readable up to the next 16-byte boundary.
SweepPlan plan_sweep_gsnmv(int d, int64_t C, int64_t n_obs, int force_variant, int num_sms);
void launch_sweep_gsnmv(const SweepPlan &pl, const double *obs, int64_t n_obs, const double *lawc,
                        int64_t C, double *partial, cudaStream_t st);

// Logistic regression.  X: row-major [n_pad][D] with D = logistic_padded_dim(d) and n_pad a
// multiple of 16 (zero rows / columns beyond n_obs / d); y: [n_pad]; theta: SoA [d][C];
// ll_part [S][C], g_part [S][d][C], S = partial slots per block of 64 chains.
struct LogisticArgs {
    const double *X;
    const double *y;
    int64_t n_obs;
    const double *theta;
    int d;
    int64_t C;
    double *ll_part;
    double *g_part;
    int S;        // partial slots per chain block
    int n_cta;    // persistent CTAs sharing the flattened (chain block, tile) units
    int pair_mode = 0;   // phase-1 variant (set by launch_sweep_logistic; see the kernel)
};
int logistic_padded_dim(int d);
cudaError_t sweep_logistic_init();
SweepPlan plan_sweep_logistic(int d, int64_t C, int64_t n_obs, int num_sms);
// sweep + fixed-order finalize: ll_out[C], grad_out[d][C] (grad_out may be NULL)
void launch_sweep_logistic(const SweepPlan &pl, const LogisticArgs &a, double *ll_out, double *grad_out,
                           cudaStream_t st);

}  // namespace extmcmc
