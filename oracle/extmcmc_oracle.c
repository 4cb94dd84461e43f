/*
 * extmcmc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 * See extmcmc_oracle.h for the parity status ("parity unpinned" for the step
 * numerics; schedule + adaptation defaults pinned by the reference's tests).
 *
 * Every function cites the reference file:line (relative to the upstream
 * ExtensibleMCMC.jl tree) whose behaviour it restates.  Compile with
 * -O2 -ffp-contract=off: the reference (Julia) never contracts a*b+c into an
 * FMA, and bit-exact replay against the GPU relies on that.
 */
#include "extmcmc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_MAX_D 16 /* obs_dim limit of the general-d Gaussian law in the oracle */
#define ORC_MAX_P 512 /* parameter-vector limit of the oracle (d = 16: 272 parameters) */

static __thread char g_err[512];
static char g_err_global[512];
static void set_err(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    snprintf(g_err_global, sizeof g_err_global, "%s", msg);
}
const char *oracle_last_error(void) { return g_err_global; }

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11; constants as in Random123 and in     */
/* cuRAND's curand_philox4x32_x.h).  Counter layout of this project:         */
/*   key = (seed lo32, seed hi32)                                            */
/*   ctr = (chain lo32, chain hi32, mcmciter lo32, (pidx << 16) | block)     */
/* The j-th uniform of a (chain, mcmciter, pidx) substream is lane j & 1 of  */
/* block j >> 1: u = (k + 0.5) * 2^-52 with k the top 52 bits of the lane's  */
/* 64-bit word (w[2l+1] << 32 | w[2l]), so u is in (0, 1) exactly.           */
/* ------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double oracle_uniform(uint64_t seed, uint64_t chain, int64_t mcmciter, int32_t pidx,
                      uint32_t j) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)mcmciter,
                       ((uint32_t)pidx << 16) | (j >> 1)};
    uint32_t w[4];
    oracle_philox4x32_10(ctr, key, w);
    uint32_t l = j & 1u;
    uint64_t word = ((uint64_t)w[2 * l + 1] << 32) | w[2 * l];
    uint64_t k = word >> 12;
    return ((double)k + 0.5) * 0x1.0p-52;
}

/* ------------------------------------------------------------------------- */
/* State                                                                     */
/* ------------------------------------------------------------------------- */
typedef struct {
    int32_t kernel, n_coords, prior, n_prior_params;
    int32_t *coords;
    uint8_t *pos;
    double *prior_params;  /* [n_prior_params] */
    extmcmc_lambda_fn lambda_fn;  /* HaarioTypeAdaptation's f(lambda, N, iter), adaptation.jl:385; NULL = default */
    void *lambda_user;
    extmcmc_adapt_t adapt;
    int32_t step_len; /* doubles of step-size state per chain */
    double *step0;    /* initial step-size state [step_len] */
    double *hmean;    /* HaarioTypeAdaptation.mean [C][n]   (adaptation.jl:373) */
    double *hcov;     /* HaarioTypeAdaptation.cov  [C][n*n] column-major (:374) */
    int64_t *hM;      /* HaarioTypeAdaptation.M    [C]      (:378) */
} orc_update_t;

struct oracle_handle {
    extmcmc_config_t cfg;
    orc_update_t *upd;
    double *obs; /* [n_obs][d] */
    double *y;
    int64_t n_obs;
    int64_t C;
    int32_t p, NU, W;
    /* per chain, AoS by chain */
    double *theta;      /* [C][p] */
    double *ll;         /* [C]    */
    double *mean;       /* [C][p] */
    double *cov;        /* [C][p*p] column-major */
    int64_t *statN;     /* [C]    */
    double **step;      /* [NU] -> [C][step_len] */
    int64_t *proposed;  /* [C][NU] adaptation counters (adaptation.jl:52-53) */
    int64_t *accepted;  /* [C][NU] */
    int64_t *tot_prop;  /* [C][NU] */
    int64_t *tot_acc;   /* [C][NU] */
    double *ra_val;     /* [C][NU] latest rolling_ar value */
    int64_t *ra_iter;   /* [C][NU] mcmciter at which ra_val was written (0 = never) */
    uint8_t *acc_ring;  /* [C][NU][W] */
    int64_t *acc_tag;   /* [C][NU][W] mcmciter stored in the slot (0 = never) */
    int32_t domain_err;
};

static void *xcalloc(size_t n, size_t s) {
    void *p = calloc(n ? n : 1, s);
    return p;
}

int32_t oracle_create(const extmcmc_config_t *cfg, const extmcmc_update_t *updates,
                      const double *obs, int64_t n_obs, const double *y,
                      const double *theta_init, oracle_t *out) {
    if (!cfg || !updates || !out || !theta_init) { set_err("null argument"); return EXTMCMC_EINVAL; }
    if (cfg->law != EXTMCMC_LAW_GSN_IID_1D && cfg->law != EXTMCMC_LAW_GSN_MV &&
        cfg->law != EXTMCMC_LAW_HIER_NORMAL && cfg->law != EXTMCMC_LAW_LOGISTIC) {
        set_err("oracle: law not implemented");
        return EXTMCMC_EUNSUPPORTED;
    }
    int32_t d = cfg->obs_dim;
    if (cfg->law == EXTMCMC_LAW_GSN_IID_1D && (d != 1 || cfg->n_params != 2)) {
        set_err("GSN_IID_1D needs obs_dim = 1 and n_params = 2");
        return EXTMCMC_EINVAL;
    }
    if (cfg->law == EXTMCMC_LAW_GSN_MV && (d < 1 || d > ORC_MAX_D || cfg->n_params != d * (d + 1))) {
        set_err("GSN_MV needs n_params = d(d+1), 1 <= d <= 16");
        return EXTMCMC_EINVAL;
    }
    if (cfg->n_params > ORC_MAX_P || cfg->n_params < 1 || cfg->n_chains < 1 || cfg->n_updates < 1) {
        set_err("oracle: need 1 <= n_params <= 512, n_chains >= 1, n_updates >= 1");
        return EXTMCMC_EINVAL;
    }
    for (int u = 0; u < cfg->n_updates; ++u)
        if (updates[u].n_coords < 1 || updates[u].n_coords > 256) { set_err("oracle: 1 <= n_coords <= 256"); return EXTMCMC_EINVAL; }
    if (cfg->law == EXTMCMC_LAW_LOGISTIC && (d != cfg->n_params || d < 1 || d > 256 || !y)) {
        set_err("LOGISTIC needs obs_dim = n_params = d <= 256 and responses in y");
        return EXTMCMC_EINVAL;
    }
    if (cfg->law == EXTMCMC_LAW_HIER_NORMAL && (d != 1 || cfg->n_params < 3 || !y)) {
        set_err("HIER_NORMAL needs obs_dim = 1, n_params = G + 2 and group indices in y");
        return EXTMCMC_EINVAL;
    }
    struct oracle_handle *h = xcalloc(1, sizeof *h);
    h->cfg = *cfg;
    h->C = cfg->n_chains; h->p = cfg->n_params; h->NU = cfg->n_updates;
    h->W = cfg->roll_window > 0 ? cfg->roll_window : 100;
    h->n_obs = n_obs;
    h->obs = xcalloc((size_t)n_obs * d, sizeof(double));
    if (n_obs) memcpy(h->obs, obs, (size_t)n_obs * d * sizeof(double));
    if (y) { h->y = xcalloc((size_t)n_obs, sizeof(double)); memcpy(h->y, y, (size_t)n_obs * sizeof(double)); }
    int64_t C = h->C; int32_t p = h->p, NU = h->NU;
    h->upd = xcalloc(NU, sizeof(orc_update_t));
    h->step = xcalloc(NU, sizeof(double *));
    for (int u = 0; u < NU; ++u) {
        const extmcmc_update_t *s = &updates[u];
        orc_update_t *t = &h->upd[u];
        if (s->kernel != EXTMCMC_KERNEL_RW_UNIFORM && s->kernel != EXTMCMC_KERNEL_RW_GAUSS &&
            s->kernel != EXTMCMC_KERNEL_RW_GAUSS_MIX && s->kernel != EXTMCMC_KERNEL_MALA) {
            set_err("oracle: transition kernel not implemented");
            oracle_destroy(h);
            return EXTMCMC_EUNSUPPORTED;
        }
        if ((s->adapt.kind == EXTMCMC_ADAPT_UNIF_RW && s->kernel != EXTMCMC_KERNEL_RW_UNIFORM) ||
            (s->adapt.kind == EXTMCMC_ADAPT_HAARIO && s->kernel != EXTMCMC_KERNEL_RW_GAUSS_MIX) ||
            (s->adapt.kind == EXTMCMC_ADAPT_MALA && s->kernel != EXTMCMC_KERNEL_MALA) ||
            s->adapt.kind > EXTMCMC_ADAPT_MALA) {
            /* readjust!(rw, adpt, iter) exists only for (UniformRandomWalk, AdaptationUnifRW) and
             * (GaussianRandomWalkMix, HaarioTypeAdaptation): adaptation.jl:273,422 */
            set_err("oracle: adaptation does not match the transition kernel");
            oracle_destroy(h);
            return EXTMCMC_EUNSUPPORTED;
        }
        if (s->prior < EXTMCMC_PRIOR_IMPROPER || s->prior > EXTMCMC_PRIOR_MVNORMAL) {
            /* reference: error("logpdf not implemented for prior ...") src/priors.jl:11-13 */
            set_err("oracle: prior not implemented");
            oracle_destroy(h);
            return EXTMCMC_EUNSUPPORTED;
        }
        t->kernel = s->kernel; t->n_coords = s->n_coords; t->prior = s->prior;
        t->n_prior_params = s->n_prior_params; t->adapt = s->adapt;
        t->coords = xcalloc(s->n_coords, sizeof(int32_t));
        t->pos = xcalloc(s->n_coords, 1);
        for (int i = 0; i < s->n_coords; ++i) {
            t->coords[i] = s->coords[i];
            t->pos[i] = s->pos ? s->pos[i] : 0;
            if (s->coords[i] < 0 || s->coords[i] >= p) { set_err("coord out of range"); oracle_destroy(h); return EXTMCMC_EINVAL; }
            /* UniformRandomWalk asserts all(eps .> 0), random_walk.jl:50 */
            if (s->kernel == EXTMCMC_KERNEL_RW_UNIFORM && !(s->step[i] > 0.0)) { set_err("eps must be > 0"); oracle_destroy(h); return EXTMCMC_EINVAL; }
        }
        t->prior_params = xcalloc(s->n_prior_params > 0 ? s->n_prior_params : 1, sizeof(double));
        for (int i = 0; i < s->n_prior_params; ++i) t->prior_params[i] = s->prior_params[i];
        if (s->prior == EXTMCMC_PRIOR_MVNORMAL && s->n_prior_params != s->n_coords * (s->n_coords + 1)) { set_err("oracle: MvNormal prior needs {mu[n], L[n*n]}"); oracle_destroy(h); return EXTMCMC_EINVAL; }
        {
            const int nn = s->n_coords * s->n_coords;
            t->step_len = s->kernel == EXTMCMC_KERNEL_RW_UNIFORM ? s->n_coords
                        : s->kernel == EXTMCMC_KERNEL_MALA ? 1
                        : s->kernel == EXTMCMC_KERNEL_RW_GAUSS ? nn : 2 * nn + 1;
            if ((s->kernel == EXTMCMC_KERNEL_RW_GAUSS || s->kernel == EXTMCMC_KERNEL_RW_GAUSS_MIX) && s->n_coords > 32) { set_err("oracle: Gaussian walks need n_coords <= 32"); oracle_destroy(h); return EXTMCMC_EINVAL; }
            if (s->adapt.kind == EXTMCMC_ADAPT_HAARIO) {
                t->hmean = xcalloc((size_t)C * s->n_coords, sizeof(double));   /* zero(state), adaptation.jl:388 */
                t->hcov = xcalloc((size_t)C * nn, sizeof(double));
                t->hM = xcalloc((size_t)C, sizeof(int64_t));
            }
        }
        t->step0 = xcalloc(t->step_len, sizeof(double));
        memcpy(t->step0, s->step, t->step_len * sizeof(double));
        h->step[u] = xcalloc((size_t)C * t->step_len, sizeof(double));
        for (int64_t c = 0; c < C; ++c) memcpy(h->step[u] + c * t->step_len, t->step0, t->step_len * sizeof(double));
    }
    h->theta = xcalloc((size_t)C * p, sizeof(double));
    h->ll = xcalloc((size_t)C, sizeof(double));
    h->mean = xcalloc((size_t)C * p, sizeof(double));
    h->cov = xcalloc((size_t)C * p * p, sizeof(double));
    h->statN = xcalloc((size_t)C, sizeof(int64_t));
    h->proposed = xcalloc((size_t)C * NU, sizeof(int64_t));
    h->accepted = xcalloc((size_t)C * NU, sizeof(int64_t));
    h->tot_prop = xcalloc((size_t)C * NU, sizeof(int64_t));
    h->tot_acc = xcalloc((size_t)C * NU, sizeof(int64_t));
    h->ra_val = xcalloc((size_t)C * NU, sizeof(double));
    h->ra_iter = xcalloc((size_t)C * NU, sizeof(int64_t));
    h->acc_ring = xcalloc((size_t)C * NU * h->W, 1);
    h->acc_tag = xcalloc((size_t)C * NU * h->W, sizeof(int64_t));
    for (int64_t c = 0; c < C; ++c) {
        for (int k = 0; k < p; ++k) h->theta[c * p + k] = theta_init[(int64_t)k * C + c];
        h->ll[c] = -INFINITY; /* StandardLocalSubworkspace: ll = -Inf, src/workspaces.jl:425 */
        h->statN[c] = 1;      /* GenericChainStats: N = 1, mean = 0, cov = 0, chain_statistics.jl:29-35 */
    }
    *out = h;
    return EXTMCMC_OK;
}

void oracle_destroy(oracle_t h) {
    if (!h) return;
    if (h->upd) for (int u = 0; u < h->NU; ++u) { free(h->upd[u].coords); free(h->upd[u].pos); free(h->upd[u].step0); free(h->upd[u].prior_params);
                                                  free(h->upd[u].hmean); free(h->upd[u].hcov); free(h->upd[u].hM); }
    if (h->step) for (int u = 0; u < h->NU; ++u) free(h->step[u]);
    free(h->upd); free(h->step); free(h->obs); free(h->y); free(h->theta); free(h->ll);
    free(h->mean); free(h->cov); free(h->statN); free(h->proposed); free(h->accepted);
    free(h->tot_prop); free(h->tot_acc); free(h->ra_val); free(h->ra_iter);
    free(h->acc_ring); free(h->acc_tag);
    free(h);
}

/* ------------------------------------------------------------------------- */
/* Target law: loglikelihood(P::GsnTargetLaw, observs)                       */
/*   src/example/gsn_target.jl:23-29 -- ll = 0.0; for obs: ll += logpdf(P.P, obs) */
/* with P.P = MvNormal(mu, Symmetric(triu(Sigma))) rebuilt by set_parameters! */
/*   src/example/gsn_target.jl:15-21.                                        */
/* logpdf(::MvNormal, x) is Distributions.jl (unpinned, absent): published   */
/* form  c0 - sqmahal/2,  c0 = -(d*log(2pi) + logdet(Sigma))/2,              */
/* logdet = 2*sum(log(diag(chol))), sqmahal = sum(abs2, L \ (x - mu)).        */
/* Returns NaN (and *bad = 1) when Sigma is not positive definite; the       */
/* reference throws PosDefException there.                                   */
/* ------------------------------------------------------------------------- */
static const double LOG2PI = 1.8378770664093454835606594728112; /* log(2*pi) */

static double loglik_gsn_1d(const double *x, int64_t n, double mu, double var, int *bad) {
    if (!(var > 0.0) || !isfinite(var)) { *bad = 1; return NAN; }
    double s = sqrt(var);                       /* cholesky of the 1x1 Sigma */
    double c0 = -(1.0 * LOG2PI + 2.0 * log(s)) / 2.0;
    double ll = 0.0;
    for (int64_t i = 0; i < n; ++i) {           /* sequential left-to-right sum */
        double z = (x[i] - mu) / s;             /* L \ (x - mu) */
        ll += c0 - (z * z) / 2.0;
    }
    return ll;
}

static double loglik_gsn_mv(const double *x, int64_t n, int d, const double *theta, int *bad) {
    const double *mu = theta;
    const double *Sg = theta + d; /* column-major d x d; only triu is used (Symmetric(triu(S))) */
    double L[ORC_MAX_D * ORC_MAX_D];
    memset(L, 0, sizeof L);
    /* lower Cholesky factor of the symmetric matrix whose upper triangle is Sg's */
    for (int j = 0; j < d; ++j) {
        double s = Sg[j + j * d]; /* A[j][j] */
        for (int k = 0; k < j; ++k) s -= L[j + k * d] * L[j + k * d];
        if (!(s > 0.0) || !isfinite(s)) { *bad = 1; return NAN; }
        double ljj = sqrt(s);
        L[j + j * d] = ljj;
        for (int i = j + 1; i < d; ++i) {
            double a = Sg[j + i * d]; /* A[i][j] = A[j][i] = upper entry (row j, col i) */
            for (int k = 0; k < j; ++k) a -= L[i + k * d] * L[j + k * d];
            L[i + j * d] = a / ljj;
        }
    }
    double logdet = 0.0;
    for (int j = 0; j < d; ++j) logdet += log(L[j + j * d]);
    logdet = 2.0 * logdet;
    double c0 = -((double)d * LOG2PI + logdet) / 2.0;
    double ll = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double *xi = x + i * d;
        double z[ORC_MAX_D];
        double sq = 0.0;
        for (int r = 0; r < d; ++r) { /* forward substitution L z = x - mu */
            double a = xi[r] - mu[r];
            for (int k = 0; k < r; ++k) a -= L[r + k * d] * z[k];
            z[r] = a / L[r + r * d];
            sq += z[r] * z[r];
        }
        ll += c0 - sq / 2.0;
    }
    return ll;
}

/* Hierarchical normal law (BASELINE cfg 4; no reference law, build-defined):
 * theta = [th_1..th_G, mu, tau];  y_gj ~ N(th_g, 1),  th_g ~ N(mu, tau^2).  The hierarchical
 * term is part of the law because the reference's priors only see the update's own coordinates
 * (src/run.jl:374-385).  One sequential sum over the observations, like gsn_target.jl:23-29. */
static double loglik_hier(const struct oracle_handle *h, const double *th, double *grad, int *bad) {
    const int G = h->p - 2;
    const double mu = th[G], tau = th[G + 1];
    if (grad) for (int k = 0; k < h->p; ++k) grad[k] = 0.0;
    if (!(tau > 0.0) || !isfinite(tau)) { *bad = 1; return NAN; }
    double ll = 0.0;
    for (int64_t i = 0; i < h->n_obs; ++i) {
        const int g = (int)h->y[i];
        const double r = h->obs[i] - th[g];
        ll += -0.5 * LOG2PI - (r * r) / 2.0;
        if (grad) grad[g] += r;
    }
    for (int g = 0; g < G; ++g) {
        const double dv = th[g] - mu;
        ll += -0.5 * LOG2PI - log(tau) - (dv * dv) / (2.0 * tau * tau);
        if (grad) {
            grad[g] += -dv / (tau * tau);
            grad[G] += dv / (tau * tau);
            grad[G + 1] += -1.0 / tau + (dv * dv) / (tau * tau * tau);
        }
    }
    return ll;
}

/* Bayesian logistic regression (BASELINE cfg 3; no reference law, build-defined):
 * theta = beta[d];  y_i ~ Bernoulli(sigmoid(x_i . beta)).
 * ll = sum_i [ y_i z_i - softplus(z_i) ],  grad = sum_i (y_i - sigmoid(z_i)) x_i, sequential sums. */
static double loglik_logistic(const struct oracle_handle *h, const double *th, double *grad) {
    const int d = h->p;
    if (grad) for (int k = 0; k < d; ++k) grad[k] = 0.0;
    double ll = 0.0;
    for (int64_t i = 0; i < h->n_obs; ++i) {
        const double *x = h->obs + i * d;
        double z = 0.0;
        for (int k = 0; k < d; ++k) z += x[k] * th[k];
        const double e = exp(-fabs(z));
        const double sp = (z > 0.0 ? z : 0.0) + log1p(e);
        const double sg = (z >= 0.0 ? 1.0 : e) / (1.0 + e);
        ll += h->y[i] * z - sp;
        if (grad) { const double r = h->y[i] - sg; for (int k = 0; k < d; ++k) grad[k] += r * x[k]; }
    }
    return ll;
}

/* d/dmu and d/dvar of the 1-D Gaussian log-likelihood, sequential per-observation sums */
static void grad_gsn_1d(const double *x, int64_t n, double mu, double var, double *grad) {
    double gm = 0.0, gv = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double r = x[i] - mu;
        gm += r / var;
        gv += -1.0 / (2.0 * var) + (r * r) / (2.0 * var * var);
    }
    grad[0] = gm; grad[1] = gv;
}

static double law_loglik(const struct oracle_handle *h, const double *theta, int *bad) {
    if (h->cfg.law == EXTMCMC_LAW_GSN_IID_1D)
        return loglik_gsn_1d(h->obs, h->n_obs, theta[0], theta[1], bad);
    if (h->cfg.law == EXTMCMC_LAW_HIER_NORMAL)
        return loglik_hier(h, theta, NULL, bad);
    if (h->cfg.law == EXTMCMC_LAW_LOGISTIC)
        return loglik_logistic(h, theta, NULL);
    return loglik_gsn_mv(h->obs, h->n_obs, h->cfg.obs_dim, theta, bad);
}

/* log-likelihood and its gradient w.r.t. all p parameters (laws with a gradient only) */
static double law_loglik_grad(const struct oracle_handle *h, const double *theta, double *grad, int *bad) {
    if (h->cfg.law == EXTMCMC_LAW_HIER_NORMAL) return loglik_hier(h, theta, grad, bad);
    if (h->cfg.law == EXTMCMC_LAW_LOGISTIC) return loglik_logistic(h, theta, grad);
    double ll = loglik_gsn_1d(h->obs, h->n_obs, theta[0], theta[1], bad);
    if (*bad) { grad[0] = grad[1] = NAN; return ll; }
    grad_gsn_1d(h->obs, h->n_obs, theta[0], theta[1], grad);
    return ll;
}

/* ------------------------------------------------------------------------- */
/* Priors: logpdf(prior, theta_loc) on the update's own coordinates only     */
/*   src/updates.jl:104, src/run.jl:374-385, src/priors.jl:18-39             */
/* ------------------------------------------------------------------------- */
static double log_prior_family(int kind, const double *pp, const double *th, int n) {
    switch (kind) {
    case EXTMCMC_PRIOR_IMPROPER: /* logpdf(::ImproperPrior, th) = 0.0, priors.jl:19 */
        return 0.0;
    case EXTMCMC_PRIOR_IMPROPER_POS: { /* -sum(log.(th)), priors.jl:26 */
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += log(th[i]);
        return -s;
    }
    case EXTMCMC_PRIOR_NORMAL: { /* StandardPrior(dist), priors.jl:35-39; Normal: -(z^2 + log2pi)/2 - log(s) */
        double m = pp[0], sd = pp[1], s = 0.0;
        for (int i = 0; i < n; ++i) { double z = (th[i] - m) / sd; s += -(z * z + LOG2PI) / 2.0 - log(sd); }
        return s;
    }
    case EXTMCMC_PRIOR_GAMMA: { /* Gamma(k, scale): -lgamma(k) - k log(scale) + (k-1) log x - x/scale */
        double k = pp[0], sc = pp[1], s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0)) return -INFINITY;
            s += -lgamma(k) - k * log(sc) + (k - 1.0) * log(th[i]) - th[i] / sc;
        }
        return s;
    }
    case EXTMCMC_PRIOR_UNIFORM: { /* Uniform(a, b): -log(b - a) inside, -Inf outside */
        double a = pp[0], b = pp[1], s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] >= a && th[i] <= b)) return -INFINITY;
            s += -log(b - a);
        }
        return s;
    }
    case EXTMCMC_PRIOR_EXPONENTIAL: { /* Exponential(scale): log(rate) - rate x, rate = 1/scale; -Inf for x < 0 */
        const double rate = 1.0 / pp[0];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] >= 0.0)) return -INFINITY;
            s += log(rate) - rate * th[i];
        }
        return s;
    }
    case EXTMCMC_PRIOR_INV_GAMMA: { /* InverseGamma(a, sc): a log sc - lgamma(a) - (a + 1) log x - sc/x */
        const double a = pp[0], sc = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0)) return -INFINITY;
            s += a * log(sc) - lgamma(a) - (a + 1.0) * log(th[i]) - sc / th[i];
        }
        return s;
    }
    case EXTMCMC_PRIOR_BETA: { /* Beta(a, b): (a-1) log x + (b-1) log1p(-x) - logbeta(a, b) on (0, 1) */
        const double a = pp[0], b = pp[1];
        const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0 && th[i] < 1.0)) return -INFINITY;
            s += (a - 1.0) * log(th[i]) + (b - 1.0) * log1p(-th[i]) - lbeta;
        }
        return s;
    }
    case EXTMCMC_PRIOR_LOGNORMAL: { /* LogNormal(m, sd): logpdf(Normal(m, sd), log x) - log x */
        const double m = pp[0], sd = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0)) return -INFINITY;
            const double lx = log(th[i]), z = (lx - m) / sd;
            s += (-(z * z + LOG2PI) / 2.0 - log(sd)) - lx;
        }
        return s;
    }
    case EXTMCMC_PRIOR_CAUCHY: { /* Cauchy(m, sc): -(log1p(z^2) + log(pi) + log(sc)) */
        const double m = pp[0], sc = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double z = (th[i] - m) / sc;
            s += -(log1p(z * z) + 1.1447298858494001741434273513531 + log(sc));
        }
        return s;
    }
    }
    return NAN;
}

static double mvn_logpdf_chol(const double *L, int n, const double *mu, const double *x);

static double log_prior(const orc_update_t *u, const double *th) {
    /* StandardPrior(MvNormal(mu, Sigma)): logpdf by whitening with the Cholesky factor (PDMats), priors.jl:35-39 */
    if (u->prior == EXTMCMC_PRIOR_MVNORMAL)
        return mvn_logpdf_chol(u->prior_params + u->n_coords, u->n_coords, u->prior_params, th);
    if (u->prior != EXTMCMC_PRIOR_PRODUCT) return log_prior_family(u->prior, u->prior_params, th, u->n_coords);
    /* logpdf(prior::ProductPrior, th): lp = 0.0; lp += logpdf(dist_k, th[idx_k]), priors.jl:82-88 */
    const int K = (int)u->prior_params[0];
    double lp = 0.0;
    int off = 0;
    for (int k = 0; k < K; ++k) {
        const double *f = u->prior_params + 1 + 4 * k;
        const int dim = (int)f[1];
        lp += log_prior_family((int)f[0], f + 2, th + off, dim);
        off += dim;
    }
    return lp;
}

/* logpdf(rw::UniformRandomWalk, a, b): sum over i of pos[i] ? -log(2 eps_i) - log(b_i) : 0.0
 *   src/transition_kernels/random_walk.jl:88-94 (mapreduce, left fold) */
static double log_q_unif(const orc_update_t *u, const double *eps, const double *to) {
    double s = 0.0;
    for (int i = 0; i < u->n_coords; ++i) {
        double t = u->pos[i] ? (-log(2.0 * eps[i]) - log(to[i])) : 0.0;
        s = (i == 0) ? t : s + t;
    }
    return s;
}

/* ------------------------------------------------------------------------- */
/* Gaussian random walks (src/transition_kernels/random_walk.jl:123-232)     */
/* ------------------------------------------------------------------------- */
/* Lower Cholesky factor of Symmetric(S) (upper triangle of the column-major n x n S;
 * GaussianRandomWalk stores Symmetric(Sigma), random_walk.jl:132).  Returns 0 if S is not
 * positive definite (MvNormal(theta, Sigma) throws PosDefException in the reference). */
static int chol_lower_sym_upper(const double *S, int n, double *L) {
    for (int j = 0; j < n; ++j) {
        double s = S[j + j * n];
        for (int k = 0; k < j; ++k) s -= L[j + k * n] * L[j + k * n];
        if (!(s > 0.0) || !isfinite(s)) return 0;
        double ljj = sqrt(s);
        L[j + j * n] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double a = S[j + i * n];
            for (int k = 0; k < j; ++k) a -= L[i + k * n] * L[j + k * n];
            L[i + j * n] = a / ljj;
        }
    }
    return 1;
}

/* logpdf(MvNormal(mu, L L'), x) = -(n log 2pi + 2 sum log L_ii)/2 - |L \ (x - mu)|^2 / 2 */
static double mvn_logpdf_chol(const double *L, int n, const double *mu, const double *x) {
    double z[32], sq = 0.0, logdet = 0.0;
    for (int r = 0; r < n; ++r) {
        double a = x[r] - mu[r];
        for (int k = 0; k < r; ++k) a -= L[r + k * n] * z[k];
        z[r] = a / L[r + r * n];
        sq += z[r] * z[r];
        logdet += log(L[r + r * n]);
    }
    return -((double)n * LOG2PI + 2.0 * logdet) / 2.0 - sq / 2.0;
}

/* logpdf(rw::GaussianRandomWalk, from, to) random_walk.jl:163-171 on transformed COPIES
 * (the reference log/exp-transforms in place and perturbs the state by ulps). */
static double log_q_gauss(const uint8_t *pos, int n, const double *Sigma, const double *from,
                          const double *to, int *bad) {
    double L[1024], tf[32], tt[32], logJ = 0.0;
    if (!chol_lower_sym_upper(Sigma, n, L)) { *bad = 1; return NAN; }
    double s = 0.0;
    for (int i = 0; i < n; ++i) if (pos[i]) s += log(to[i]);     /* _logjacobian: -sum(log.(to[pos])) */
    logJ = -s;
    for (int i = 0; i < n; ++i) { tf[i] = pos[i] ? log(from[i]) : from[i]; tt[i] = pos[i] ? log(to[i]) : to[i]; }
    return mvn_logpdf_chol(L, n, tf, tt) + logJ;
}

/* log transition density of an update's kernel, from -> to */
static double log_q(const orc_update_t *u, const double *step, const double *from, const double *to, int *bad) {
    const int n = u->n_coords;
    if (u->kernel == EXTMCMC_KERNEL_RW_UNIFORM) return log_q_unif(u, step, to);
    if (u->kernel == EXTMCMC_KERNEL_RW_GAUSS) return log_q_gauss(u->pos, n, step, from, to, bad);
    /* GaussianRandomWalkMix, random_walk.jl:229-232 (no log-sum-exp guard, as in the reference) */
    const double lam = step[2 * n * n];
    const double lpA = log_q_gauss(u->pos, n, step, from, to, bad);
    const double lpB = log_q_gauss(u->pos, n, step + n * n, from, to, bad);
    return log((1.0 - lam) * exp(lpA) + lam * exp(lpB));
}

/* rand(rw::GaussianRandomWalk, theta) random_walk.jl:145-151: theta° = exp-back(log-transform(theta) + L z) */
static void gauss_propose(const uint8_t *pos, int n, const double *Sigma, const double *th,
                          const double *z, double *out, int *bad) {
    double L[1024];
    if (!chol_lower_sym_upper(Sigma, n, L)) { *bad = 1; for (int i = 0; i < n; ++i) out[i] = NAN; return; }
    for (int i = 0; i < n; ++i) {
        double t = pos[i] ? log(th[i]) : th[i];
        double a = 0.0;
        for (int k = 0; k <= i; ++k) a += L[i + k * n] * z[k];   /* unwhiten: L * z */
        t = a + t;                                               /* + mu */
        out[i] = pos[i] ? exp(t) : t;
    }
}

/* ------------------------------------------------------------------------- */
/* One schedule element for one chain: the body of __run! src/run.jl:70-82   */
/* ------------------------------------------------------------------------- */
static double mala_prior_grad(const orc_update_t *u, double th) {
    if (u->prior == EXTMCMC_PRIOR_NORMAL) return -(th - u->prior_params[0]) / (u->prior_params[1] * u->prior_params[1]);
    return 0.0;
}
static double mala_prior_logpdf1(const orc_update_t *u, double th) {
    if (u->prior == EXTMCMC_PRIOR_NORMAL) {
        const double z = (th - u->prior_params[0]) / u->prior_params[1];
        return -(z * z + LOG2PI) / 2.0 - log(u->prior_params[1]);
    }
    return 0.0;
}

typedef struct {
    double *rec_prop, *rec_exp;
    double *theta_hist, *prop_hist, *ll_hist, *llp_hist, *llr_hist;
    uint8_t *acc_hist;
    int32_t rng_mode, p_u_max;
} orc_io_t;

static void chain_step(struct oracle_handle *h, int64_t c, const extmcmc_step_t *st,
                       int64_t s_idx, const orc_io_t *io) {
    const int32_t p = h->p, NU = h->NU, W = h->W;
    const int64_t C = h->C;
    const int32_t u_idx = st->pidx;
    const orc_update_t *u = &h->upd[u_idx];
    const int n = u->n_coords;
    double *theta = h->theta + c * p;
    double *eps = h->step[u_idx] + c * u->step_len;
    const uint64_t gchain = (uint64_t)(h->cfg.chain_offset + c);

    /* update_workspaces! src/run.jl:101-112: local state <- global state[coords];
     * local ll <- ll_history of the previously executed update (here: the carried
     * ll); on the very first element prev is `nothing` and ll stays -Inf. */
    double th_loc[ORC_MAX_P], th_prop[ORC_MAX_P];
    double g_cur[ORC_MAX_P], g_prop[ORC_MAX_P];
    for (int i = 0; i < n; ++i) th_loc[i] = theta[u->coords[i]];
    double ll_cur = (st->prev_pidx < 0) ? -INFINITY : h->ll[c];
    if (st->prev_pidx < 0) h->ll[c] = -INFINITY;

    /* compute_gradients_and_momenta!(updt, ws, Previous) -- the reference's hook for
     * gradient-based updates, called on every step (src/run.jl:110, src/updates.jl:129-133) */
    if (u->kernel == EXTMCMC_KERNEL_MALA) {
        int badg = 0;
        law_loglik_grad(h, theta, g_cur, &badg);
        if (badg) h->domain_err = 1;
    }

    /* proposal! src/updates.jl:191-196 + rand(::UniformRandomWalk) random_walk.jl:65-73 */
    uint32_t j = 0; /* index into this chain-step's uniform substream */
    if (io->rng_mode == EXTMCMC_RNG_REPLAY) {
        for (int i = 0; i < n; ++i)
            th_prop[i] = io->rec_prop[((int64_t)s_idx * io->p_u_max + i) * C + c];
    } else {
        for (;;) {
            if (u->kernel == EXTMCMC_KERNEL_MALA) {
                /* theta° = theta + (tau^2/2) grad(ll + log prior)(theta) + tau z */
                const double tau = eps[0], h2 = tau * tau / 2.0;
                for (int q = 0; q < n; q += 2) {
                    double u1 = oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++);
                    double u2 = oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++);
                    double rad = sqrt(-2.0 * log(u1));
                    double zz[2] = {rad * cos(6.283185307179586476925286766559 * u2),
                                    rad * sin(6.283185307179586476925286766559 * u2)};
                    for (int w = 0; w < 2 && q + w < n; ++w) {
                        const int i = q + w;
                        const double g = g_cur[u->coords[i]] + mala_prior_grad(u, th_loc[i]);
                        th_prop[i] = th_loc[i] + h2 * g + tau * zz[w];
                    }
                }
                break;   /* no prior-support redraw: MALA priors have full support */
            } else if (u->kernel == EXTMCMC_KERNEL_RW_UNIFORM) {
                for (int i = 0; i < n; ++i) {
                    double r = oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++);
                    /* rand(Uniform(a, b)) = a + (b - a) * rand() with a = -eps, b = eps */
                    double a = -eps[i], b = eps[i];
                    double U = a + (b - a) * r;
                    /* theta .* (exp.(U).*pos .+ 1.0.*.!pos) .+ U.*.!pos, random_walk.jl:72 */
                    th_prop[i] = u->pos[i] ? th_loc[i] * exp(U) : th_loc[i] + U;
                }
            } else {
                /* Mix: pick_kernel, rand(Bernoulli(lambda)) = rand() <= lambda, random_walk.jl:225-227 */
                const double *Sig = eps;
                if (u->kernel == EXTMCMC_KERNEL_RW_GAUSS_MIX) {
                    double r = oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++);
                    if (r <= eps[2 * n * n]) Sig = eps + n * n;
                }
                double z[32];
                for (int q = 0; q < n; q += 2) {          /* randn via Box-Muller on the uniform stream */
                    double u1 = oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++);
                    double u2 = oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++);
                    double rad = sqrt(-2.0 * log(u1));
                    z[q] = rad * cos(6.283185307179586476925286766559 * u2);
                    if (q + 1 < n) z[q + 1] = rad * sin(6.283185307179586476925286766559 * u2);
                }
                int badp = 0;
                gauss_propose(u->pos, n, Sig, th_loc, z, th_prop, &badp);
                if (badp) { h->domain_err = 1; break; }
            }
            /* redraw the whole vector while the prior is exactly -Inf, updates.jl:193-195 */
            if (!(log_prior(u, th_prop) == -INFINITY)) break;
            if (j > 60000u) break; /* substream exhausted: keep the -Inf proposal (will be rejected) */
        }
    }
    if (io->rec_prop && io->rng_mode != EXTMCMC_RNG_REPLAY)
        for (int i = 0; i < n; ++i)
            io->rec_prop[((int64_t)s_idx * io->p_u_max + i) * C + c] = th_prop[i];

    /* set_proposal! src/run.jl:221-240: full proposal = global state with the
     * update's coordinates replaced; pushed into P° (gsn_target.jl:15-21) */
    double full_prop[ORC_MAX_P];
    for (int k = 0; k < p; ++k) full_prop[k] = theta[k];
    for (int i = 0; i < n; ++i) full_prop[u->coords[i]] = th_prop[i];

    /* compute_ll! src/run.jl:251-260 -> loglikelihood(P°, obs) */
    int bad = 0;
    /* compute_ll! + compute_gradients_and_momenta!(updt, ws, Proposal), src/run.jl:257-259 */
    double ll_prop = (u->kernel == EXTMCMC_KERNEL_MALA) ? law_loglik_grad(h, full_prop, g_prop, &bad)
                                                         : law_loglik(h, full_prop, &bad);
    if (bad) h->domain_err = 1;

    /* accept_reject! src/run.jl:268-281; strict left-to-right association */
    double llr = ll_prop - ll_cur;
    if (u->kernel == EXTMCMC_KERNEL_MALA) {
        /* log q(a -> b) = -|b - a - (tau^2/2) g(a)|^2 / (2 tau^2); constant omitted (same both ways) */
        const double tau = eps[0], h2 = tau * tau / 2.0;
        double qf = 0.0, qb = 0.0, lpp = 0.0, lpc = 0.0;
        for (int i = 0; i < n; ++i) {
            const double a = th_loc[i], b = th_prop[i];
            const double ga = g_cur[u->coords[i]] + mala_prior_grad(u, a);
            const double gb = g_prop[u->coords[i]] + mala_prior_grad(u, b);
            const double rf = b - a - h2 * ga, rb = a - b - h2 * gb;
            qf += rf * rf; qb += rb * rb;
            lpp += mala_prior_logpdf1(u, b); lpc += mala_prior_logpdf1(u, a);
        }
        const double inv = 1.0 / (2.0 * tau * tau);
        qf = -qf * inv; qb = -qb * inv;
        llr = llr + qb;
        llr = llr - qf;
        llr = llr + lpp;
        llr = llr - lpc;
    } else {
        int badq = 0;
        llr = llr + log_q(u, eps, th_prop, th_loc, &badq);   /* ltd(Proposal): theta° -> theta, run.jl:360-367 */
        llr = llr - log_q(u, eps, th_loc, th_prop, &badq);   /* ltd(Previous): theta -> theta°, run.jl:344-351 */
        if (badq) h->domain_err = 1;
        llr = llr + log_prior(u, th_prop);
        llr = llr - log_prior(u, th_loc);
    }
    double E;
    if (io->rng_mode == EXTMCMC_RNG_REPLAY) {
        E = io->rec_exp[(int64_t)s_idx * C + c];
    } else {
        /* rand(Exponential(1.0)), run.jl:278, by inversion of the next uniform */
        E = -log(oracle_uniform(h->cfg.seed, gchain, st->mcmciter, u_idx, j++));
        if (io->rec_exp) io->rec_exp[(int64_t)s_idx * C + c] = E;
    }
    int accepted = E > -llr; /* NaN llr compares false => reject */

    /* register_accept_reject_results! run.jl:329-335 and set_chain_param! run.jl:312-320 */
    double ll_new = accepted ? ll_prop : ll_cur;
    if (accepted) for (int i = 0; i < n; ++i) theta[u->coords[i]] = th_prop[i];
    h->ll[c] = ll_new;
    if (io->theta_hist) for (int k = 0; k < p; ++k) io->theta_hist[((int64_t)s_idx * p + k) * C + c] = theta[k];
    if (io->prop_hist) for (int k = 0; k < p; ++k) io->prop_hist[((int64_t)s_idx * p + k) * C + c] = full_prop[k];
    if (io->ll_hist) io->ll_hist[(int64_t)s_idx * C + c] = ll_new;
    if (io->llp_hist) io->llp_hist[(int64_t)s_idx * C + c] = ll_prop;
    if (io->llr_hist) io->llr_hist[(int64_t)s_idx * C + c] = llr;
    if (io->acc_hist) io->acc_hist[(int64_t)s_idx * C + c] = (uint8_t)accepted;

    /* update_stats! src/chain_statistics.jl:41-66 (verbatim arithmetic) */
    {
        double *mean = h->mean + c * p, *cov = h->cov + c * p * p;
        int64_t N = h->statN[c];
        double f_old = (double)(N - 1) / (double)N;
        double f_mean = (double)N / (double)(N + 1);
        double f_new = (double)(N + 1) / (double)N;
        double old_mean[ORC_MAX_P];
        for (int k = 0; k < p; ++k) old_mean[k] = mean[k];
        for (int k = 0; k < p; ++k) mean[k] = old_mean[k] * f_mean + theta[k] / (double)(N + 1);
        for (int b = 0; b < p; ++b)
            for (int a = 0; a < p; ++a) {
                double old_sum_sq = f_old * cov[a + b * p] + old_mean[a] * old_mean[b];
                double new_sum_sq = old_sum_sq + (theta[a] * theta[b]) / (double)N;
                cov[a + b * p] = new_sum_sq - f_new * (mean[a] * mean[b]);
            }
        int64_t it = st->mcmciter;
        int64_t *ra_iter = h->ra_iter + c * NU;
        double *ra_val = h->ra_val + c * NU;
        /* ra_prev = rolling_ar[max(1, it-1)][pidx]; zeros unless written */
        int64_t prev_it = it - 1 > 1 ? it - 1 : 1;
        double ra_prev = (ra_iter[u_idx] == prev_it && prev_it != it) ? ra_val[u_idx] : 0.0;
        uint8_t *ring = h->acc_ring + ((size_t)c * NU + u_idx) * W;
        int64_t *tag = h->acc_tag + ((size_t)c * NU + u_idx) * W;
        int slot = (int)(it % W);
        int acc_out = 0;
        /* acceptance_history[it - W] (undef memory in the reference if that update
         * was excluded at it - W; the restatement reads false there) */
        if (it > W && tag[slot] == it - W) acc_out = ring[slot];
        int64_t mn = (int64_t)W < N ? (int64_t)W : N;
        double ra = (ra_prev * (double)W + (double)(accepted - acc_out)) / (double)mn;
        ra_val[u_idx] = ra; ra_iter[u_idx] = it;
        ring[slot] = (uint8_t)accepted; tag[slot] = it;
        h->statN[c] = N + 1;
    }

    /* update_adaptation! src/run.jl:136-173 (only the update whose turn it is
     * registers, run.jl:176-177) -> register!/time_to_update/readjust!
     * src/transition_kernels/adaptation.jl:273-329 */
    h->tot_prop[c * NU + u_idx] += 1;
    h->tot_acc[c * NU + u_idx] += accepted;
    if (u->adapt.kind == EXTMCMC_ADAPT_UNIF_RW || u->adapt.kind == EXTMCMC_ADAPT_MALA) {
        const int n_eps = u->kernel == EXTMCMC_KERNEL_MALA ? 1 : n;
        int64_t *prop = &h->proposed[c * NU + u_idx], *acc = &h->accepted[c * NU + u_idx];
        *acc += accepted; *prop += 1;                       /* register! :292-295 */
        if (*prop >= u->adapt.adapt_every_k_steps) {        /* time_to_update :302-304 */
            /* compute_delta :312-319 (Int / Int -> Float64 division) */
            double r = (double)st->mcmciter / (double)u->adapt.adapt_every_k_steps - u->adapt.offset;
            double delta = u->adapt.scale / sqrt(r > 1.0 ? r : 1.0);
            double a_r = (*prop == 0) ? 0.0 : (double)*acc / (double)*prop; /* :242-244 */
            *prop = 0; *acc = 0;                            /* reset! :263-266 */
            double sgn = (a_r > u->adapt.target_accpt_rate) ? 1.0 : -1.0;
            for (int i = 0; i < n_eps; ++i) {               /* compute_eps :326-329 */
                double e = eps[i] + sgn * delta;
                e = e < u->adapt.max ? e : u->adapt.max;    /* min(e, max) */
                e = e > u->adapt.min ? e : u->adapt.min;    /* max(., min) */
                eps[i] = e;
            }
        }
    }
    /* HaarioTypeAdaptation registers on EVERY update step of ANY update
     * (register_only_on_my_turn(::Val{false}, ::Haario) = false, adaptation.jl:399-404), on the
     * update's view of the global state, log-transformed (here: a transformed copy). */
    for (int v = 0; v < NU; ++v) {
        const orc_update_t *w = &h->upd[v];
        if (w->adapt.kind != EXTMCMC_ADAPT_HAARIO) continue;
        const int m = w->n_coords;
        double *hm = w->hmean + c * m, *hc = w->hcov + c * m * m;
        if (v == u_idx) w->hM[c] += 1;                           /* my turn: M += 1, :401-404 */
        double t[32], old_m[32];
        for (int i = 0; i < m; ++i) { double x = theta[w->coords[i]]; t[i] = w->pos[i] ? log(x) : x; }
        const int64_t N = h->statN[c] - 1;                       /* adpt.N: starts at 1, +1 per register */
        const double f_old = (double)(N - 1) / (double)N, f_mean = (double)N / (double)(N + 1);
        const double f_new = (double)(N + 1) / (double)N;
        for (int i = 0; i < m; ++i) old_m[i] = hm[i];
        for (int i = 0; i < m; ++i) hm[i] = old_m[i] * f_mean + t[i] / (double)(N + 1);     /* :409 */
        for (int b = 0; b < m; ++b)
            for (int a = 0; a < m; ++a) {
                double old_sum_sq = f_old * hc[a + b * m] + old_m[a] * old_m[b];          /* :408 */
                double new_sum_sq = old_sum_sq + (t[a] * t[b]) / (double)N;                /* :410 */
                hc[a + b * m] = new_sum_sq - f_new * (hm[a] * hm[b]);                      /* :411 */
            }
        if (v == u_idx && w->hM[c] >= w->adapt.adapt_every_k_steps) {   /* time_to_update :416-420 */
            w->hM[c] = 0;
            double *SigB = h->step[v] + c * w->step_len + m * m;
            for (int k = 0; k < m * m; ++k) SigB[k] = (2.38 * 2.38) / (double)m * hc[k];   /* 2.38^2/length(rw)*cov :423 */
            /* rw.lambda = adpt.f_lambda(rw.lambda, adpt.N, mcmc_iter) (:425); adpt.N was incremented by
             * this step's register! (:413); the default closure returns lambda unchanged (:385) */
            if (w->lambda_fn) {
                double *lam = h->step[v] + c * w->step_len + 2 * m * m;
                *lam = w->lambda_fn(*lam, h->statN[c], st->mcmciter, w->lambda_user);
            }
        }
    }
}

/* Chains are independent, so the optional thread pool hands out chain ids from
 * a shared counter; results do not depend on the thread count. */
typedef struct {
    struct oracle_handle *h;
    const extmcmc_step_t *steps;
    int32_t n_steps;
    const orc_io_t *io;
    const double *theta_eval;
    double *ll_out;
    int64_t n_eval;
    int any_bad;
} run_job_t;

typedef struct {
    void (*fn)(void *, int64_t);
    void *arg;
    int64_t n;
    int64_t next;
    pthread_mutex_t mu;
} pool_t;

static void *pool_main(void *a) {
    pool_t *pl = a;
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        int64_t c = pl->next < pl->n ? pl->next++ : -1;
        pthread_mutex_unlock(&pl->mu);
        if (c < 0) return NULL;
        pl->fn(pl->arg, c);
    }
}

static void parallel_chains(void (*fn)(void *, int64_t), void *arg, int64_t n, int n_threads) {
    if (n_threads <= 1 || n <= 1) { for (int64_t c = 0; c < n; ++c) fn(arg, c); return; }
    if (n_threads > 256) n_threads = 256;
    pool_t pl = {fn, arg, n, 0, PTHREAD_MUTEX_INITIALIZER};
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads; ++t) if (pthread_create(&th[started], NULL, pool_main, &pl) == 0) ++started;
    if (started == 0) pool_main(&pl);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

static void run_worker(void *a, int64_t c) {
    run_job_t *j = a;
    for (int s = 0; s < j->n_steps; ++s) chain_step(j->h, c, &j->steps[s], s, j->io);
}

static void loglik_worker(void *a, int64_t c) {
    run_job_t *j = a;
    double th[ORC_MAX_P];
    for (int k = 0; k < j->h->p; ++k) th[k] = j->theta_eval[(int64_t)k * j->n_eval + c];
    int bad = 0;
    j->ll_out[c] = law_loglik(j->h, th, &bad);
    if (bad) j->any_bad = 1;
}

int32_t oracle_run_block(oracle_t h, const extmcmc_step_t *steps, int32_t n_steps,
                         int32_t rng_mode, int32_t p_u_max,
                         double *rec_proposals, double *rec_exp,
                         double *theta_hist, double *theta_prop_hist,
                         double *ll_hist, double *ll_prop_hist,
                         uint8_t *accepted_hist, double *llr_hist, int32_t n_threads) {
    if (!h || !steps) { set_err("null argument"); return EXTMCMC_EINVAL; }
    if (rng_mode == EXTMCMC_RNG_REPLAY && (!rec_proposals || !rec_exp)) { set_err("replay needs buffers"); return EXTMCMC_EINVAL; }
    for (int s = 0; s < n_steps; ++s) {
        if (steps[s].pidx < 0 || steps[s].pidx >= h->NU) { set_err("pidx out of range"); return EXTMCMC_EINVAL; }
        if (h->upd[steps[s].pidx].n_coords > p_u_max && (rec_proposals != NULL)) { set_err("p_u_max too small"); return EXTMCMC_EINVAL; }
    }
    orc_io_t io = {rec_proposals, rec_exp, theta_hist, theta_prop_hist, ll_hist,
                   ll_prop_hist, llr_hist, accepted_hist, rng_mode, p_u_max};
    run_job_t job = {h, steps, n_steps, &io, NULL, NULL, 0, 0};
    parallel_chains(run_worker, &job, h->C, n_threads);
    return h->domain_err ? EXTMCMC_EDOMAIN : EXTMCMC_OK;
}

int32_t oracle_set_lambda_fn(oracle_t h, int32_t u, extmcmc_lambda_fn f, void *user) {
    if (!h || u < 0 || u >= h->NU) return EXTMCMC_EINVAL;
    h->upd[u].lambda_fn = f;
    h->upd[u].lambda_user = user;
    return EXTMCMC_OK;
}

int32_t oracle_get_state(oracle_t h, double *theta, double *ll) {
    for (int64_t c = 0; c < h->C; ++c) {
        if (theta) for (int k = 0; k < h->p; ++k) theta[(int64_t)k * h->C + c] = h->theta[c * h->p + k];
        if (ll) ll[c] = h->ll[c];
    }
    return EXTMCMC_OK;
}

int32_t oracle_get_stats(oracle_t h, double *mean, double *cov, double *rolling_ar,
                         int64_t *n_accept, int64_t *n_prop) {
    int64_t C = h->C; int p = h->p, NU = h->NU;
    for (int64_t c = 0; c < C; ++c) {
        if (mean) for (int k = 0; k < p; ++k) mean[(int64_t)k * C + c] = h->mean[c * p + k];
        if (cov) for (int k = 0; k < p * p; ++k) cov[(int64_t)k * C + c] = h->cov[c * p * p + k];
        for (int u = 0; u < NU; ++u) {
            if (rolling_ar) rolling_ar[(int64_t)u * C + c] = h->ra_val[c * NU + u];
            if (n_accept) n_accept[(int64_t)u * C + c] = h->tot_acc[c * NU + u];
            if (n_prop) n_prop[(int64_t)u * C + c] = h->tot_prop[c * NU + u];
        }
    }
    return EXTMCMC_OK;
}

int32_t oracle_get_eps(oracle_t h, int32_t u, double *eps) {
    if (u < 0 || u >= h->NU) return EXTMCMC_EINVAL;
    int L = h->upd[u].step_len;
    for (int64_t c = 0; c < h->C; ++c)
        for (int i = 0; i < L; ++i) eps[(int64_t)i * h->C + c] = h->step[u][c * L + i];
    return EXTMCMC_OK;
}

int32_t oracle_loglik_grad(oracle_t h, const double *theta, int64_t n_eval, double *ll_out, double *grad_out) {
    if (h->cfg.law == EXTMCMC_LAW_GSN_MV) return EXTMCMC_EUNSUPPORTED;
    if (h->p < 1 || h->p > ORC_MAX_P) return EXTMCMC_EINVAL;
    for (int64_t c = 0; c < n_eval; ++c) {
        double th[ORC_MAX_P] = {0.0}, g[ORC_MAX_P] = {0.0};
        int bad = 0;
        for (int k = 0; k < h->p; ++k) th[k] = theta[(int64_t)k * n_eval + c];
        ll_out[c] = law_loglik_grad(h, th, g, &bad);
        for (int k = 0; k < h->p; ++k) grad_out[(int64_t)k * n_eval + c] = g[k];
    }
    return EXTMCMC_OK;
}

int32_t oracle_get_adapt_state(oracle_t h, int32_t u, double *mean, double *cov) {
    if (u < 0 || u >= h->NU || h->upd[u].adapt.kind != EXTMCMC_ADAPT_HAARIO) return EXTMCMC_EINVAL;
    const int m = h->upd[u].n_coords;
    for (int64_t c = 0; c < h->C; ++c) {
        if (mean) for (int i = 0; i < m; ++i) mean[(int64_t)i * h->C + c] = h->upd[u].hmean[c * m + i];
        if (cov) for (int k = 0; k < m * m; ++k) cov[(int64_t)k * h->C + c] = h->upd[u].hcov[c * m * m + k];
    }
    return EXTMCMC_OK;
}

int32_t oracle_loglik(oracle_t h, const double *theta, int64_t n_eval, double *ll_out,
                      int32_t n_threads) {
    run_job_t job = {h, NULL, 0, NULL, theta, ll_out, n_eval, 0};
    parallel_chains(loglik_worker, &job, n_eval, n_threads);
    return job.any_bad ? EXTMCMC_EDOMAIN : EXTMCMC_OK;
}
