"""A second, independent restatement of the reference sampler -- pure Python / numpy, ONE chain.

TEST INFRASTRUCTURE.  Written from the reference's Julia sources (cited per function), not from
oracle/extmcmc_oracle.c: it keeps the reference's own object structure (global workspace with
`state_history[M][NU]`, one local workspace per update with `ll_history` / `acceptance_history`,
a `GenericChainStats`, adaptation objects holding their own counters) and its broadcast
expressions, so that the C oracle and this file are two lineages of the same algorithm.  The
tests replay one chain of the oracle through it: decisions and trajectories must agree exactly,
log-likelihoods to 1e-13 relative, step sizes and running moments exactly.

Only what the reference defines (or DESIGN.md defines for MALA) is here: uniform / Gaussian /
mixture random walks, priors, GsnTargetLaw, accept/reject, chain statistics, AdaptationUnifRW,
HaarioTypeAdaptation, the schedule with exclusions.  Randomness is injected (a `draws` object),
because the reference's own source is Julia's global MersenneTwister.
"""
import math

import numpy as np

LOG2PI = math.log(2.0 * math.pi)


# ------------------------------------------------------------------ schedule (src/schedule.jl:24-89)
class Schedule:
    def __init__(self, num_mcmc_steps, num_updates, exclude_updates=()):
        self.M, self.NU = num_mcmc_steps, num_updates
        self.excl = {}                                     # DefaultDict(0:0): nothing matches
        for idxs, rng in exclude_updates:
            for i in np.atleast_1d(idxs):
                self.excl[int(i)] = rng

    def _transition(self, st):                             # schedule.jl:77-89
        prev_it, prev_p, it, p = st
        reset = p == self.NU
        it, p = it + (1 if reset else 0), (1 if reset else p + 1)
        if it in self.excl.get(p, range(0, 1)):
            return self._transition((prev_it, prev_p, it, p))
        return (prev_it, prev_p, it, p)

    def __iter__(self):                                    # schedule.jl:56-66
        st = (None, None, 1, 1)
        while st[2] <= self.M:
            yield st
            st = self._transition((st[2], st[3], st[2], st[3]))


# ------------------------------------------------------------------ priors (src/priors.jl)
def _lgamma(x):
    return math.lgamma(x)


def logpdf_dist(dist, x):
    """logpdf of a univariate Distributions.jl family (published closed forms)."""
    kind, a, b = dist
    if kind == "Normal":
        z = (x - a) / b
        return -(z * z + LOG2PI) / 2.0 - math.log(b)
    if kind == "Gamma":                                    # shape a, scale b
        if not x > 0.0:
            return -math.inf
        return -_lgamma(a) - a * math.log(b) + (a - 1.0) * math.log(x) - x / b
    if kind == "Uniform":
        return -math.log(b - a) if a <= x <= b else -math.inf
    if kind == "Exponential":                              # scale a
        rate = 1.0 / a
        return math.log(rate) - rate * x if x >= 0.0 else -math.inf
    if kind == "InverseGamma":
        if not x > 0.0:
            return -math.inf
        return a * math.log(b) - _lgamma(a) - (a + 1.0) * math.log(x) - b / x
    if kind == "Beta":
        if not 0.0 < x < 1.0:
            return -math.inf
        return (a - 1.0) * math.log(x) + (b - 1.0) * math.log1p(-x) - (_lgamma(a) + _lgamma(b) - _lgamma(a + b))
    if kind == "LogNormal":
        if not x > 0.0:
            return -math.inf
        lx = math.log(x)
        z = (lx - a) / b
        return (-(z * z + LOG2PI) / 2.0 - math.log(b)) - lx
    if kind == "Cauchy":
        z = (x - a) / b
        return -(math.log1p(z * z) + math.log(math.pi) + math.log(b))
    raise NotImplementedError(kind)


def chol_lower(S):
    """Lower Cholesky factor of Symmetric(S) (the UPPER triangle of S is the data, as in
    Symmetric(Sigma), random_walk.jl:132, gsn_target.jl:19); None if not positive definite."""
    n = S.shape[0]
    L = np.zeros((n, n))
    for j in range(n):
        s = S[j, j]
        for k in range(j):
            s -= L[j, k] * L[j, k]
        if not (s > 0.0) or math.isinf(s):
            return None
        L[j, j] = math.sqrt(s)
        for i in range(j + 1, n):
            a = S[j, i]
            for k in range(j):
                a -= L[i, k] * L[j, k]
            L[i, j] = a / L[j, j]
    return L


def mvn_logpdf(L, mu, x):
    """logpdf(MvNormal(mu, L L'), x) = -(n log 2pi + logdet)/2 - |L \\ (x - mu)|^2 / 2 (PDMats)."""
    n = len(mu)
    z = np.zeros(n)
    sq = logdet = 0.0
    for r in range(n):
        a = x[r] - mu[r]
        for k in range(r):
            a -= L[r, k] * z[k]
        z[r] = a / L[r, r]
        sq += z[r] * z[r]
        logdet += math.log(L[r, r])
    return -(n * LOG2PI + 2.0 * logdet) / 2.0 - sq / 2.0


def log_prior(prior, th):
    """prior: ("improper",) | ("improper_pos",) | ("std", dist) | ("mvn", mu, L) |
    ("product", [(prior_k, dim_k), ...]) -- priors.jl:19,26,39,82-88."""
    kind = prior[0]
    if kind == "improper":
        return 0.0
    if kind == "improper_pos":
        s = 0.0
        for v in th:
            s += math.log(v)                               # log of a negative throws, as in the reference
        return -s
    if kind == "std":                                      # logpdf(dist, th): iid product over the coordinates
        s = 0.0
        for v in th:
            t = logpdf_dist(prior[1], v)
            if t == -math.inf:
                return -math.inf
            s += t
        return s
    if kind == "mvn":
        return mvn_logpdf(prior[2], prior[1], th)
    if kind == "product":
        lp, off = 0.0, 0
        for pr, dim in prior[1]:
            lp += log_prior(pr, th[off:off + dim])
            off += dim
        return lp
    raise NotImplementedError(kind)


# ------------------------------------------------------------------ transition kernels
class UniformRW:                                           # random_walk.jl:45-94
    def __init__(self, eps, pos=None):
        self.eps = np.array(eps, dtype=float)
        self.pos = np.zeros(len(self.eps), bool) if pos is None else np.array(pos, bool)

    def __len__(self):
        return len(self.eps)

    def rand(self, th, draws):
        U = np.array([-e + (e - -e) * draws.uniform() for e in self.eps])     # rand(Uniform(-eps, eps))
        out = np.empty(len(th))
        for i in range(len(th)):
            out[i] = th[i] * math.exp(U[i]) if self.pos[i] else th[i] + U[i]  # random_walk.jl:72
        return out

    def logpdf(self, th, th_prop):                         # mapreduce(+), random_walk.jl:88-94
        s = None
        for i in range(len(self.eps)):
            t = (-math.log(2.0 * self.eps[i]) - math.log(th_prop[i])) if self.pos[i] else 0.0
            s = t if s is None else s + t
        return s


class GaussianRW:                                          # random_walk.jl:123-171
    def __init__(self, Sigma, pos=None):
        self.Sigma = np.array(Sigma, dtype=float)
        self.pos = np.zeros(self.Sigma.shape[0], bool) if pos is None else np.array(pos, bool)

    def __len__(self):
        return self.Sigma.shape[0]

    def _t(self, th):                                      # remove_constraints! on a copy
        return np.array([math.log(v) if p else v for v, p in zip(th, self.pos)])

    def rand(self, th, draws):
        n = len(th)
        z = draws.normals(n)
        L = chol_lower(self.Sigma)
        if L is None:
            raise FloatingPointError("PosDefException")
        t = self._t(th)
        out = np.empty(n)
        for i in range(n):
            a = 0.0
            for k in range(i + 1):
                a += L[i, k] * z[k]                        # unwhiten, then + mu
            v = a + t[i]
            out[i] = math.exp(v) if self.pos[i] else v
        return out

    def logpdf(self, th, th_prop):
        L = chol_lower(self.Sigma)
        if L is None:
            raise FloatingPointError("PosDefException")
        s = 0.0
        for v, p in zip(th_prop, self.pos):
            if p:
                s += math.log(v)
        return mvn_logpdf(L, self._t(th), self._t(th_prop)) + (-s)


class GaussianRWMix:                                       # random_walk.jl:193-232
    def __init__(self, Sigma_A, Sigma_B, lam=0.5, pos=None):
        self.A, self.B, self.lam = GaussianRW(Sigma_A, pos), GaussianRW(Sigma_B, pos), float(lam)
        self.pos = self.A.pos

    def __len__(self):
        return len(self.A)

    def rand(self, th, draws):
        rw = self.B if draws.uniform() <= self.lam else self.A       # rand(Bernoulli(lambda))
        return rw.rand(th, draws)

    def logpdf(self, th, th_prop):
        return math.log((1 - self.lam) * math.exp(self.A.logpdf(th, th_prop))
                        + self.lam * math.exp(self.B.logpdf(th, th_prop)))


# ------------------------------------------------------------------ adaptation (adaptation.jl)
class NoAdapt:
    pass


class AdaptUnifRW:                                         # adaptation.jl:51-71, 242-329
    def __init__(self, k=100, target=0.234, scale=1.0, vmin=1e-12, vmax=1e7, offset=1e2):
        self.proposed = self.accepted = 0
        self.k, self.target, self.scale, self.min, self.max, self.offset = k, target, scale, vmin, vmax, offset

    def register(self, accepted):
        self.accepted += int(accepted)
        self.proposed += 1

    def time_to_update(self):
        return self.proposed >= self.k

    def readjust(self, eps, mcmc_iter):
        delta = self.scale / math.sqrt(max(1.0, mcmc_iter / self.k - self.offset))       # compute_delta
        a_r = 0.0 if self.proposed == 0 else self.accepted / self.proposed
        self.proposed = self.accepted = 0
        new = eps + (2 * (a_r > self.target) - 1) * delta                               # compute_eps
        return np.maximum(np.minimum(new, self.max), self.min)


class Haario:                                              # adaptation.jl:372-426
    def __init__(self, n, k=100, f=None):
        self.mean, self.cov = np.zeros(n), np.zeros((n, n))
        self.k, self.N, self.M = k, 1, 0
        self.f = f if f is not None else (lambda lam, N, it: lam)

    def register(self, th_transformed):
        N, t = self.N, th_transformed
        old_sum_sq = (N - 1) / N * self.cov + np.outer(self.mean, self.mean)
        self.mean = self.mean * (N / (N + 1)) + t / (N + 1)
        new_sum_sq = old_sum_sq + np.outer(t, t) / N
        self.cov = new_sum_sq - (N + 1) / N * np.outer(self.mean, self.mean)
        self.N += 1


# ------------------------------------------------------------------ target laws
def ll_gsn(theta, obs, d):
    """loglikelihood(P::GsnTargetLaw, obs), gsn_target.jl:15-29: theta = [mu; vec(Sigma)], sequential
    sum of logpdf(MvNormal(mu, Symmetric(triu(Sigma))), x)."""
    mu, S = theta[:d], theta[d:].reshape(d, d).T           # column-major vec
    L = chol_lower(S)
    if L is None:
        return math.nan
    ll = 0.0
    for x in obs.reshape(-1, d):
        ll += mvn_logpdf(L, mu, x)
    return ll


def ll_grad_gsn1d(theta, obs):
    mu, var = theta
    ll = ll_gsn(theta, obs, 1)
    r = obs.ravel() - mu
    s1, s2 = float(np.cumsum(r)[-1]), float(np.cumsum(r * r)[-1])
    return ll, np.array([s1 / var, -len(r) / (2.0 * var) + s2 / (2.0 * var * var)])


def ll_hier(theta, obs, grp, G, grad=False):
    """DESIGN.md (build-defined law of BASELINE cfg 4): y_gj ~ N(th_g, 1), th_g ~ N(mu, tau^2)."""
    th, mu, tau = theta[:G], theta[G], theta[G + 1]
    if not tau > 0.0:
        return (math.nan, None) if grad else math.nan
    ll = 0.0
    T = np.zeros(G)
    for y, g in zip(obs.ravel(), grp):
        r = y - th[int(g)]
        ll += -0.5 * LOG2PI - r * r / 2.0
        T[int(g)] += r
    for g in range(G):
        ll += -0.5 * LOG2PI - math.log(tau) - (th[g] - mu) ** 2 / (2.0 * tau * tau)
    if not grad:
        return ll
    gr = np.zeros(G + 2)
    gr[:G] = T - (th - mu) / (tau * tau)
    gr[G] = np.sum((th - mu) / (tau * tau))
    gr[G + 1] = -G / tau + np.sum((th - mu) ** 2) / tau ** 3
    return ll, gr


# ------------------------------------------------------------------ the sampler
class Update:
    def __init__(self, kind, rw_or_tau, coords, prior=("improper",), adpt=None):
        self.kind = kind                                   # "rw" | "mala"
        self.rw = rw_or_tau if kind == "rw" else None
        self.tau = float(rw_or_tau) if kind == "mala" else None
        self.coords = [c - 1 for c in coords]              # 1-based in, 0-based inside
        self.prior, self.adpt = prior, adpt if adpt is not None else NoAdapt()


class Replay:
    """Randomness source replaying recorded proposals / Exp(1) draws of one chain."""
    def __init__(self, proposals, exp_draws):
        self.proposals, self.exp_draws, self.k = proposals, exp_draws, 0

    def proposal(self, n):
        return np.array(self.proposals[self.k][:n], dtype=float)

    def exponential(self):
        e = self.exp_draws[self.k]
        self.k += 1
        return e


def run_chain(law, obs, updates, theta0, M, draws, exclude=(), roll_window=100, grp=None):
    """run!/__run! for ONE chain (run.jl:34-83).  law = ("gsn", d) | ("hier", G).  Returns the
    reference's histories and final adaptation / statistics state."""
    NU, p = len(updates), len(theta0)
    theta = np.array(theta0, dtype=float)                  # global_ws.sub_ws.state
    state_hist = np.full((M, NU, p), np.nan)
    prop_hist = np.full((M, NU, p), np.nan)
    ll_hist = np.zeros((NU, M))                            # per local workspace (workspaces.jl:426)
    llp_hist = np.zeros((NU, M))
    acc_hist = np.zeros((NU, M), bool)
    executed = np.zeros((NU, M), bool)
    local_ll = np.full(NU, -math.inf)                      # StandardLocalSubworkspace.ll = -Inf (:425)
    # GenericChainStats (chain_statistics.jl:16-38)
    cs_mean, cs_cov, cs_N = np.zeros(p), np.zeros((p, p)), 1
    rolling_ar = np.zeros((M, NU))
    grad_cur = [None] * NU

    def loglik(th, grad=False):
        if law[0] == "gsn":
            if grad:
                return ll_grad_gsn1d(th, obs)
            return ll_gsn(th, obs, law[1])
        return ll_hier(th, obs, grp, law[1], grad)

    trace = []
    for prev_it, prev_p, it, pj in Schedule(M, NU, exclude):
        u = updates[pj - 1]
        j = pj - 1
        # update_workspaces! (run.jl:101-112)
        th_loc = theta[u.coords].copy()
        if prev_p is not None:
            local_ll[j] = ll_hist[prev_p - 1][prev_it - 1]
        ll_cur = local_ll[j]
        if u.kind == "mala":
            _, g_cur = loglik(theta, grad=True)            # compute_gradients_and_momenta!(.., Previous) run.jl:110
        # proposal! (updates.jl:191-196)
        if isinstance(draws, Replay):
            th_prop = draws.proposal(len(u.coords))
        elif u.kind == "mala":
            h2 = u.tau * u.tau / 2.0
            z = draws.normals(len(u.coords))
            th_prop = np.array([th_loc[i] + h2 * (g_cur[u.coords[i]] + _prior_grad(u.prior, th_loc[i])) + u.tau * z[i]
                                for i in range(len(u.coords))])
        else:
            th_prop = u.rw.rand(th_loc, draws)
            while log_prior(u.prior, th_prop) == -math.inf:
                th_prop = u.rw.rand(th_loc, draws)
        # set_proposal! (run.jl:221-240)
        full_prop = theta.copy()
        full_prop[u.coords] = th_prop
        prop_hist[it - 1, j] = full_prop
        # compute_ll! (run.jl:251-260)
        if u.kind == "mala":
            ll_prop, g_prop = loglik(full_prop, grad=True)
        else:
            ll_prop = loglik(full_prop)
        # accept_reject! (run.jl:268-281)
        if u.kind == "mala":
            h2 = u.tau * u.tau / 2.0
            qf = qb = lpp = lpc = 0.0
            for i, c in enumerate(u.coords):
                a, b = th_loc[i], th_prop[i]
                ga = g_cur[c] + _prior_grad(u.prior, a)
                gb = g_prop[c] + _prior_grad(u.prior, b)
                rf, rb = b - a - h2 * ga, a - b - h2 * gb
                qf += rf * rf
                qb += rb * rb
                lpp += log_prior(u.prior, [b]) if u.prior[0] == "std" else 0.0
                lpc += log_prior(u.prior, [a]) if u.prior[0] == "std" else 0.0
            inv = 1.0 / (2.0 * u.tau * u.tau)
            llr = ll_prop - ll_cur + (-qb * inv) - (-qf * inv) + lpp - lpc
        else:
            llr = (ll_prop - ll_cur
                   + u.rw.logpdf(th_prop, th_loc)          # ltd(Proposal): theta° -> theta (run.jl:360-367)
                   - u.rw.logpdf(th_loc, th_prop)          # ltd(Previous): theta -> theta°
                   + log_prior(u.prior, th_prop)
                   - log_prior(u.prior, th_loc))
        E = draws.exponential()
        accepted = E > -llr
        # register_accept_reject_results! + set_chain_param! (run.jl:299-335)
        llp_hist[j][it - 1] = ll_prop
        ll_hist[j][it - 1] = ll_prop if accepted else ll_cur
        acc_hist[j][it - 1] = accepted
        executed[j][it - 1] = True
        if accepted:
            theta[u.coords] = th_prop
        state_hist[it - 1, j] = theta
        # update_stats! (chain_statistics.jl:41-66)
        N = cs_N
        old_sum_sq = (N - 1) / N * cs_cov + np.outer(cs_mean, cs_mean)
        cs_mean = cs_mean * (N / (N + 1)) + theta / (N + 1)
        new_sum_sq = old_sum_sq + np.outer(theta, theta) / N
        cs_cov = new_sum_sq - (N + 1) / N * np.outer(cs_mean, cs_mean)
        ra_prev = rolling_ar[max(1, it - 1) - 1][j]
        W = roll_window
        out_win = bool(acc_hist[j][it - W - 1]) if (it > W and executed[j][it - W - 1]) else False
        rolling_ar[it - 1][j] = (ra_prev * W + (int(accepted) - int(out_win))) / min(W, N)
        cs_N += 1
        # update_adaptation! (run.jl:136-173): every update is visited; only Haario registers off-turn
        for i, v in enumerate(updates):
            a = v.adpt
            my_turn = i == j
            if isinstance(a, NoAdapt):
                continue
            if isinstance(a, AdaptUnifRW):
                if not my_turn:
                    continue
                a.register(accepted)
                if a.time_to_update():
                    if v.kind == "mala":
                        v.tau = float(a.readjust(np.array([v.tau]), it)[0])
                    else:
                        v.rw.eps = a.readjust(v.rw.eps, it)
            elif isinstance(a, Haario):
                if my_turn:
                    a.M += 1
                t = np.array([math.log(theta[c]) if pz else theta[c] for c, pz in zip(v.coords, v.rw.pos)])
                a.register(t)
                if my_turn and a.M >= a.k:
                    a.M = 0
                    v.rw.B.Sigma = 2.38 ** 2 / len(v.rw) * a.cov
                    v.rw.lam = a.f(v.rw.lam, a.N, it)
        trace.append((it, pj, bool(accepted), float(ll_prop), float(llr)))
    return dict(theta=state_hist, theta_prop=prop_hist, ll=ll_hist, ll_prop=llp_hist, accepted=acc_hist,
                executed=executed, mean=cs_mean, cov=cs_cov, rolling_ar=rolling_ar, trace=trace, final=theta)


def _prior_grad(prior, x):
    if prior[0] == "std" and prior[1][0] == "Normal":
        return -(x - prior[1][1]) / (prior[1][2] * prior[1][2])
    return 0.0


# ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11)
def philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


class PhiloxDraws:
    """The project's counter stream (csrc/philox.cuh): key = seed, ctr = (chain lo, chain hi,
    mcmciter, pidx0 << 16 | block); uniform j = lane (j & 1) of block (j >> 1), top 52 bits + 0.5."""
    def __init__(self, seed, chain):
        self.seed, self.chain, self.it, self.p0, self.j = seed, chain, 1, 0, 0

    def start_step(self, mcmciter, pidx0):
        self.it, self.p0, self.j = mcmciter, pidx0, 0

    def uniform(self):
        j = self.j
        self.j += 1
        w = philox4x32_10([self.chain & 0xFFFFFFFF, self.chain >> 32, self.it & 0xFFFFFFFF, (self.p0 << 16) | (j >> 1)],
                          [self.seed & 0xFFFFFFFF, self.seed >> 32])
        lane = j & 1
        word = (w[2 * lane + 1] << 32) | w[2 * lane]
        return ((word >> 12) + 0.5) * 2.0 ** -52

    def normals(self, n):
        z = np.empty(n + (n & 1))
        for q in range(0, n, 2):                           # Box-Muller on the uniform stream
            u1, u2 = self.uniform(), self.uniform()
            rad = math.sqrt(-2.0 * math.log(u1))
            z[q], z[q + 1] = rad * math.cos(2.0 * math.pi * u2), rad * math.sin(2.0 * math.pi * u2)
        return z[:n]

    def exponential(self):
        return -math.log(self.uniform())
