// Host-side interface of the persistent block kernels (block_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "dev_state.cuh"

namespace extmcmc {

// ---- chain-resident block kernel (1-D Gaussian family laws, uniform random walks and MALA) -----
struct ResidentPlan {
    int n_cta;            // CTAs (a multiple of the SM count when there are enough chains)
    int64_t base, rem;    // CTA b owns base + (b < rem) chains, contiguous
    int R, cg;            // chains per thread, chain groups per 256-thread group
    int stage_doubles;    // per staging array of the cooperative covariance update (0: none)
    size_t smem_per_group, smem_bytes;
};
struct ResidentArgs {
    DevState d;           // d.S must be 1: the kernel leaves its sums in partial[2][G][C]
    const StepDesc *descs;
    int n_steps;
    int n_sweeps;         // likelihood sweeps the block runs (1 per random-walk element, 1-2 per MALA element)
    const double *obs;
    const int64_t *goff, *glen;
    int G;
    double *ll_scratch;   // [C]
    // filled by launch_resident_block from the plan
    int cg, stage_doubles;
    int64_t base, rem;
    unsigned int smem_per_group;
};
// false: the shape is not served (too few chains unless `force`, too many observation groups)
bool plan_resident(const DevState &d, int num_sms, bool force, ResidentPlan *pl);
cudaError_t launch_resident_block(const ResidentPlan &pl, ResidentArgs a, cudaStream_t st);

// ---- observation-mapped block kernel (GSN_IID_1D, uniform random walks, C <= 32) ---------------
struct ObsBlockArgs {
    DevState d;
    const StepDesc *descs;
    int n_steps;
    const double *obs;    // this rank's observations (16-byte aligned, readable up to the next pair)
    int64_t n_obs;
    unsigned long long *go;   // exchange steps completed so far (= xseq of the next step to run)
    unsigned int *counter;    // CTAs that have delivered their sums of the step in flight
};
cudaError_t plan_obs_block(int cb, int num_sms, int64_t n_obs, int *grid);
// proposal of the block's first element + arming of the go flag (one small CTA)
void launch_obs_block_first(const DevState &d, const StepDesc *descs, unsigned long long *go,
                            unsigned int *counter, cudaStream_t st);
cudaError_t launch_obs_block(int cb, int grid, ObsBlockArgs a, cudaStream_t st);

}  // namespace extmcmc
