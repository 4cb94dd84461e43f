/*
 * extmcmc_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * A plain-C restatement of ExtensibleMCMC.jl's single-chain Metropolis-within-
 * Gibbs transition step, applied to C independent chains one chain at a time.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (libextmcmc_cuda.so and
 * the Python host mirror) never does.
 *
 * PARITY STATUS: "parity unpinned" for the numerics of the transition step --
 * the reference's own tests (test/runtests.jl:87-114) run the sampler unseeded
 * with no assertion, Julia is not installed here, and the logpdf arithmetic
 * lives in Distributions.jl / PDMats.jl (unpinned: Project.toml:6-15 has no
 * compat bounds, /Manifest.toml is git-ignored), whose published closed forms
 * are restated here.  Pinned by the reference's tests and re-checked in
 * tests/: the MCMCSchedule golden sequence (test/runtests.jl:13-31) and the
 * AdaptationUnifRW defaults (test/runtests.jl:35-38).  Further pins added by
 * this repository: analytic posteriors, scipy/mpmath cross-checks of every
 * closed form, a hand-computed 3-step trace, Philox4x32-10 known answers.
 *
 * The struct types of the public C ABI (include/extmcmc.h) are reused so that
 * a test can hand the very same configuration to the oracle and to the GPU.
 */
#ifndef EXTMCMC_ORACLE_H_
#define EXTMCMC_ORACLE_H_

#include "../include/extmcmc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_handle *oracle_t;

/* obs: row-major [n_obs][obs_dim] (copied); y may be NULL. */
int32_t oracle_create(const extmcmc_config_t *cfg, const extmcmc_update_t *updates,
                      const double *obs, int64_t n_obs, const double *y,
                      const double *theta_init /* [p][C] */, oracle_t *out);
void    oracle_destroy(oracle_t h);
const char *oracle_last_error(void);

/*
 * Run n_steps schedule elements for every chain (chain loop outermost, as the
 * reference would run C separate samplers).  n_threads > 1 distributes chains
 * over a pthread pool (results do not depend on it).
 *
 * rng_mode EXTMCMC_RNG_PHILOX: draws come from the per-chain Philox stream; if
 *   rec_proposals / rec_exp are non-NULL the local proposals
 *   [n_steps][p_u_max][C] and Exp(1) draws [n_steps][C] are recorded.
 * rng_mode EXTMCMC_RNG_REPLAY: rec_proposals / rec_exp are read instead.
 *
 * History outputs (NULL = skip): theta, theta_prop [n_steps][p][C];
 * ll, ll_prop [n_steps][C]; accepted [n_steps][C]; llr [n_steps][C].
 */
int32_t oracle_run_block(oracle_t h, const extmcmc_step_t *steps, int32_t n_steps,
                         int32_t rng_mode, int32_t p_u_max,
                         double *rec_proposals, double *rec_exp,
                         double *theta_hist, double *theta_prop_hist,
                         double *ll_hist, double *ll_prop_hist,
                         uint8_t *accepted_hist, double *llr_hist,
                         int32_t n_threads);

/* HaarioTypeAdaptation's weight schedule f(lambda, N, iter) of update u (NULL: default). */
int32_t oracle_set_lambda_fn(oracle_t h, int32_t u, extmcmc_lambda_fn f, void *user);
int32_t oracle_get_state(oracle_t h, double *theta, double *ll);
int32_t oracle_get_stats(oracle_t h, double *mean, double *cov, double *rolling_ar,
                         int64_t *n_accept, int64_t *n_prop);
/* Per-chain step-size state of update u, [len][C]: eps (RW_UNIFORM, len = p_u), Sigma (RW_GAUSS,
 * len = p_u^2), [Sigma_A, Sigma_B, lambda] (RW_GAUSS_MIX, len = 2 p_u^2 + 1). */
int32_t oracle_get_eps(oracle_t h, int32_t u, double *eps);
int32_t oracle_get_adapt_state(oracle_t h, int32_t u, double *mean, double *cov);
/* Full-data log-likelihood of arbitrary parameter vectors theta[p][C]
 * (the reference's loglikelihood(P, obs), src/example/gsn_target.jl:23-29). */
int32_t oracle_loglik(oracle_t h, const double *theta, int64_t n_chains_eval,
                      double *ll_out, int32_t n_threads);

/* ll and d ll / d theta (all p parameters) for laws with a gradient: grad_out[p][n]. */
int32_t oracle_loglik_grad(oracle_t h, const double *theta, int64_t n_chains_eval, double *ll_out,
                           double *grad_out);

/* Philox4x32-10 block function, exposed for known-answer tests. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* j-th uniform in (0,1) of the (chain, mcmciter, pidx) substream. */
double oracle_uniform(uint64_t seed, uint64_t chain, int64_t mcmciter, int32_t pidx,
                      uint32_t j);

#ifdef __cplusplus
}
#endif
#endif
