// Host-side interface of the persistent block kernels (block_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "dev_state.cuh"

namespace extmcmc {

constexpr int kBlkMaxUpd = 8;       // most updates of a model served by the block kernels
constexpr uint32_t kNotStaged = 0xFFFFFFFFu;

// Shared-memory layout of the per-chain state a CTA keeps for the whole block (byte offsets from the
// start of the view area; computed on the host for `n_cap` chains, used by the device as is).
struct ViewLayout {
    uint32_t theta, ll, prop_loc, prop_full, lawc, n_used, ll_prop, grad_cur, grad_prop, mean, cov;
    uint32_t eps[kBlkMaxUpd], adapt_prop[kBlkMaxUpd], adapt_acc[kBlkMaxUpd], tot_prop[kBlkMaxUpd],
        tot_acc[kBlkMaxUpd], ra_val[kBlkMaxUpd], acc_ring[kBlkMaxUpd];
    uint32_t upd_table, dv;
    uint32_t bytes;
    int32_t pl_rows, cov_rows, eps_rows[kBlkMaxUpd];
    int32_t n_cap;
};
// Which arrays a model needs in the view (any_mala: gradient buffers); false if it cannot be served.
bool make_view_layout(const DevState &d, const DevUpdate *upd_host, int n_cap, ViewLayout *vl);

// ---- team-resident block kernel (1-D Gaussian family laws, uniform random walks and MALA) ------
struct TeamPlan {
    int n_cta, ts;        // CTAs, CTAs per team (n_cta % (2 ts) == 0 when there are two phases)
    int phases;           // 2: the second half of the CTAs starts half a sweep late
    int n_team;           // teams
    int64_t base, rem;    // team t owns base + (t < rem) chains, contiguous
    int R, cg;            // chains per thread, chain groups per CTA (cg * ns = 256 threads)
    int stage_doubles;    // per staging array of the cooperative covariance update (0: none)
    int stages;           // TMA ring depth (tiles of 1024 observations)
    ViewLayout vl;
    size_t smem_bytes;
    bool cooperative;
};
struct TeamArgs {
    DevState d;
    const StepDesc *descs;
    int n_steps;
    int n_sweeps;         // likelihood sweeps the block runs (1 per random-walk element, 1-2 per MALA element)
    const double *obs;
    const int64_t *goff, *glen;
    int G;
    double *ll_scratch;   // [C]
    unsigned int *sync;   // [n_team] team barrier counters, then [n_cta / 2] half-sweep flags (zeroed per launch)
    long long *prof;      // diagnostics (EXTMCMC_TEAM_PROF): [n_cta][8] cycle counters, or nullptr
    // from the plan
    int ts, phases, n_team, cg, stage_doubles, stages;
    int64_t base, rem;
    ViewLayout vl;
};
// false: the shape is not served (too few chains unless `force`, too many groups / parameters / updates)
bool plan_team(const DevState &d, const DevUpdate *upd_host, int num_sms, bool force, TeamPlan *pl);
size_t team_sync_words(const TeamPlan &pl);
cudaError_t launch_team_block(const TeamPlan &pl, TeamArgs a, cudaStream_t st);

// ---- observation-mapped block kernel (GSN_IID_1D, uniform random walks, C <= 32) ---------------
struct ObsBlockPlan {
    int cb, grid;
    ViewLayout vl;
    size_t smem_bytes;
};
struct ObsBlockArgs {
    DevState d;
    const StepDesc *descs;
    int n_steps;
    const double *obs;    // this rank's observations (16-byte aligned, readable up to the next pair)
    int64_t n_obs;
    unsigned long long *go;   // exchange steps completed so far (= xseq of the next step to run)
    unsigned int *counter;    // CTAs that have delivered their sums of the step in flight
    ViewLayout vl;
};
cudaError_t plan_obs_block(const DevState &d, const DevUpdate *upd_host, int num_sms, int64_t n_obs, ObsBlockPlan *pl);
cudaError_t launch_obs_block(const ObsBlockPlan &pl, ObsBlockArgs a, cudaStream_t st);

}  // namespace extmcmc
