/* A plain-C client of libextmcmc_cuda.so: what a Julia `ccall` (or any FFI) does, without Python.
 * Runs the Gaussian mean/variance sampler (BASELINE cfg 1 shape, 64 chains) for 600 iterations in
 * blocks of 50 schedule elements and prints the posterior mean of (mu, sigma^2) over the second
 * half of the run, the acceptance rate and the adapted eps of chain 0.
 *   gcc -I include tests/c_abi_demo.c -o demo -ldl     (the library is dlopen'ed at run time) */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "extmcmc.h"

#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }
#define CK(call) do { int32_t rc_ = (call); if (rc_ != EXTMCMC_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_extmcmc_last_error(h)); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s path/to/libextmcmc_cuda.so\n", argv[0]); return 2; }
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    SYM(extmcmc_create) SYM(extmcmc_destroy) SYM(extmcmc_last_error) SYM(extmcmc_upload_obs)
    SYM(extmcmc_set_update) SYM(extmcmc_set_state) SYM(extmcmc_run_block) SYM(extmcmc_sync)
    SYM(extmcmc_get_history) SYM(extmcmc_get_stats) SYM(extmcmc_get_eps)

    enum { C = 64, N = 1000, M = 600, NU = 2, BLOCK = 50 };
    static double x[N], theta0[2 * C], th[BLOCK * 2 * C], eps[C];
    unsigned long long s = 88172645463325252ull;   /* xorshift + Box-Muller: x ~ N(1, 2^2) */
    double sum = 0, ss = 0;
    for (int i = 0; i < N; i += 2) {
        double u[2];
        for (int k = 0; k < 2; ++k) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; u[k] = ((s >> 11) + 0.5) / 9007199254740992.0; }
        double r = sqrt(-2 * log(u[0]));
        x[i] = 1 + 2 * r * cos(6.283185307179586 * u[1]);
        x[i + 1] = 1 + 2 * r * sin(6.283185307179586 * u[1]);
    }
    for (int i = 0; i < N; ++i) sum += x[i];
    for (int i = 0; i < N; ++i) ss += (x[i] - sum / N) * (x[i] - sum / N);
    for (int c = 0; c < C; ++c) { theta0[c] = 0.0; theta0[C + c] = 1.0; }

    extmcmc_config_t cfg = {0};
    cfg.abi_version = EXTMCMC_ABI_VERSION; cfg.n_chains = C; cfg.n_params = 2; cfg.n_updates = NU;
    cfg.law = EXTMCMC_LAW_GSN_IID_1D; cfg.obs_dim = 1; cfg.seed = 2024; cfg.world_size = 1;
    cfg.history_window = 2 * BLOCK; cfg.roll_window = 100; cfg.use_graphs = 1;
    extmcmc_t h = NULL;
    if (p_extmcmc_create(&cfg, &h) != EXTMCMC_OK) { fprintf(stderr, "create: %s\n", p_extmcmc_last_error(NULL)); return 1; }
    int32_t c0[1] = {0}, c1[1] = {1};
    double e0[1] = {0.5};
    uint8_t pos1[1] = {1};
    extmcmc_adapt_t ad = {EXTMCMC_ADAPT_UNIF_RW, 50, 0.234, 0.1, 1e-12, 1e7, 1e2};
    extmcmc_update_t u0 = {EXTMCMC_KERNEL_RW_UNIFORM, 1, c0, e0, NULL, EXTMCMC_PRIOR_IMPROPER, 0, NULL, ad};
    extmcmc_update_t u1 = {EXTMCMC_KERNEL_RW_UNIFORM, 1, c1, e0, pos1, EXTMCMC_PRIOR_IMPROPER_POS, 0, NULL, ad};
    CK(p_extmcmc_set_update(h, 0, &u0));
    CK(p_extmcmc_set_update(h, 1, &u1));
    CK(p_extmcmc_upload_obs(h, x, N, 1, NULL));
    CK(p_extmcmc_set_state(h, theta0));

    double mean[2] = {0, 0};
    long n_kept = 0, seq = 0;
    extmcmc_step_t steps[BLOCK];
    int filled = 0;
    for (int it = 1; it <= M; ++it)
        for (int pj = 0; pj < NU; ++pj) {
            int first = it == 1 && pj == 0;
            steps[filled].mcmciter = it; steps[filled].pidx = pj;
            steps[filled].prev_pidx = first ? -1 : (pj ? pj - 1 : NU - 1);
            steps[filled].prev_mcmciter = first ? 0 : (pj ? it : it - 1);
            if (++filled == BLOCK || (it == M && pj == NU - 1)) {
                CK(p_extmcmc_run_block(h, steps, filled));
                CK(p_extmcmc_sync(h));
                CK(p_extmcmc_get_history(h, seq, seq + filled, th, NULL, NULL, NULL, NULL));
                for (int k = 0; k < filled; ++k)
                    if (steps[k].mcmciter > M / 2 && steps[k].pidx == NU - 1) {
                        for (int c = 0; c < C; ++c) { mean[0] += th[(k * 2 + 0) * C + c]; mean[1] += th[(k * 2 + 1) * C + c]; }
                        n_kept += C;
                    }
                seq += filled; filled = 0;
            }
        }
    static int64_t na[NU * C], np_[NU * C];
    CK(p_extmcmc_get_stats(h, NULL, NULL, NULL, na, np_));
    CK(p_extmcmc_get_eps(h, 0, eps));
    long a = 0, p = 0;
    for (int i = 0; i < NU * C; ++i) { a += na[i]; p += np_[i]; }
    printf("%.6f %.6f %.6f %.6f %.4f %.6f\n", mean[0] / n_kept, mean[1] / n_kept, sum / N, ss / (N - 3), (double)a / p, eps[0]);
    p_extmcmc_destroy(h);
    return 0;
}
