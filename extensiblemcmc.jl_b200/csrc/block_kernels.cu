// Persistent block kernels: a whole BLOCK of schedule elements (src/run.jl:70-82, the body of
// __run!) in ONE launch -- proposal, likelihood sweep, accept/reject, chain statistics and
// adaptation of every element, back to back on the SMs, no kernel boundary and no grid-wide
// reduction pass in between.  Two shapes:
//
//  resident_block_kernel<R>  ("chain-resident", many chains: BASELINE cfg 2 and cfg 4)
//      A CTA owns a contiguous range of chains for the whole block and streams ALL observations
//      through its own TMA ring (they are L2-resident: 8 MB / 256 KB), so a chain's sums never
//      leave the SM: thread <-> (R chains in registers) x (observation slice), a transposing
//      warp butterfly + one shared-memory pass reduce over the slices, and the chain's own thread
//      takes the decision.  Chains never interact, so CTAs never synchronise with each other.
//      A CTA is two independent 256-thread groups (named barriers), each with its own chains and
//      ring, started half a sweep apart: while one group is in its latency-bound scalar phase
//      (decision, statistics, next proposal) the other owns the FP64 pipe.  Equal work keeps the
//      offset locked, so the pipe idles only while BOTH groups are scalar -- never, in steady state.
//
//  obs_block_kernel<CB>  ("observation-mapped", a handful of chains, huge N: BASELINE cfg 5)
//      Every CTA streams its own observation segment for all chains (HBM-bound) and keeps its TMA
//      ring running ACROSS steps (observations are constant), so HBM stays busy while the step is
//      decided.  The last CTA to deliver its partial sums is the step's leader: it adds the
//      segments in a fixed order, exchanges the totals with the other ranks over NVLink peer
//      mappings when observations are sharded (stores into every peer + flag, ordered sum on
//      arrival: compute and collective in one kernel), takes the decision for every chain, issues
//      the next proposal and releases a go-flag the other CTAs spin on.  Cooperative launch
//      guarantees co-residency.
//
// Both reuse the per-chain functions of step_device.cuh unchanged, so results are bit-identical to
// the per-step kernels up to the association of the observation sums.
#include <cstdint>
#include <cuda_runtime.h>
#include "dev_state.cuh"
#include "philox.cuh"
#include "step_device.cuh"
#include "block_kernels.h"
#include "tma.cuh"

namespace extmcmc {

namespace {

constexpr int kGroupThreads = 256;          // threads of one chain group of the resident kernel
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kResTile = 2048;              // observations per TMA tile (16 KB)
constexpr int kResStages = 3;

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
struct GroupSync {
    int id;
    __device__ __forceinline__ void operator()() const { named_bar_sync(id, kGroupThreads); }
};

template <int V> struct Pow2Ceil { static constexpr int value = V <= 1 ? 1 : V <= 2 ? 2 : V <= 4 ? 4 : V <= 8 ? 8 : 16; };

// Sum over the 32 lanes of NV values per lane in ~NV + log2(32/NV) adds instead of 5 NV: at each
// butterfly step a lane keeps one half of its values and hands the other half to its partner.
// On return v[0] of lane L is the total of value (L >> log2(32 / NV)); fixed order.
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(double (&v)[NV], int lane) {
    int o = 16;
#pragma unroll
    for (int len = NV; len > 1; len >>= 1, o >>= 1) {
        const int half = len >> 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double keep = upper ? v[k + half] : v[k];
            const double send = upper ? v[k] : v[k + half];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

// Per-group view of the shared memory and of the group's place in the CTA.
struct Group {
    double *tile;          // [kResStages][kResTile]
    uint64_t *full;        // [kResStages] "tile landed" mbarriers
    unsigned int *done;    // [kResStages] warps that have finished the stage (the last one refills it)
    double *red;           // [G][2][cap][warps per chain group]
    StepCtx *ctx;
    double *sh_t, *sh_m, *sh_n;   // staging of the cooperative covariance update
    const int *tile_start; // [G + 1] first tile of every observation group within a sweep
    int tid, bar_id;
    int64_t c0;            // first chain of the group
    int n;                 // chains of the group
    int cg, ns;            // chain groups per thread group, observation slices (cg * ns = 256)
    int cgi, sl;           // this thread's chain group and slice
    uint32_t tiles_done;   // tiles consumed so far (kernel lifetime)
    uint32_t tiles_total;  // tiles the group consumes in this launch
};

// issue the TMA copy of lifetime-tile tt into its stage
__device__ __forceinline__ void res_issue(const ResidentArgs &a, const Group &g, uint32_t tt) {
    const int T = g.tile_start[a.G];
    const int ti = (int)(tt % (uint32_t)T);
    int og = 0;
    while (g.tile_start[og + 1] <= ti) ++og;
    const int64_t off = (int64_t)(ti - g.tile_start[og]) * kResTile;
    const int64_t len = a.glen[og];
    const int cnt = (int)((len - off) < (int64_t)kResTile ? (len - off) : (int64_t)kResTile);
    const uint32_t bytes = (uint32_t)((cnt + 1) >> 1) * 16u;   // padded device buffer
    const int st = (int)(tt % kResStages);
    mbar_expect_tx(&g.full[st], bytes);
    bulk_g2s(g.tile + st * kResTile, a.obs + a.goff[og] + off, bytes, &g.full[st]);
}

// One likelihood sweep of the group's chains over all observations.  mu_src[g][C]: per-chain mean
// of observation group g.  Leaves sum (x - mu)^2 in partial[g][c] and, with GRAD, sum (x - mu)
// in partial[G + g][c] (the S = 1 layout the per-chain finalize functions read).
template <int R, bool GRAD>
__device__ __forceinline__ void res_sweep(const ResidentArgs &a, Group &g, const double *__restrict__ mu_src,
                                          bool signal_half) {
    constexpr int NQ = GRAD ? 2 : 1;
    constexpr int NV = Pow2Ceil<R * NQ>::value;
    const DevState &d = a.d;
    const int64_t C = d.C;
    const int lane = g.tid & 31, warp = g.tid >> 5;
    const int wpc = kGroupWarps / g.cg;          // warps per chain group
    const int cap = g.cg * R;
    const int T = g.tile_start[a.G];
    for (int og = 0; og < a.G; ++og) {
        double m[R], acc[R], accT[GRAD ? R : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int local = g.cgi * R + r;
            m[r] = local < g.n ? mu_src[(int64_t)og * C + g.c0 + local] : 0.0;
            acc[r] = 0.0;
            if (GRAD) accT[r] = 0.0;
        }
        auto eat = [&](const double2 x) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d0 = x.x - m[r];
                acc[r] = fma(d0, d0, acc[r]);
                if (GRAD) accT[r] += d0;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d1 = x.y - m[r];
                acc[r] = fma(d1, d1, acc[r]);
                if (GRAD) accT[r] += d1;
            }
        };
        const int64_t len = a.glen[og];
        const int nt = g.tile_start[og + 1] - g.tile_start[og];
        for (int t = 0; t < nt; ++t) {
            const uint32_t tt = g.tiles_done;
            const int st = (int)(tt % kResStages);
            mbar_wait(&g.full[st], (tt / kResStages) & 1u);
            const int64_t off = (int64_t)t * kResTile;
            const int cnt = (int)((len - off) < (int64_t)kResTile ? (len - off) : (int64_t)kResTile);
            const double2 *xs = reinterpret_cast<const double2 *>(g.tile + st * kResTile);
            if (cnt == kResTile) {
                const int iters = (kResTile / 2) / g.ns;
#pragma unroll 4
                for (int k = 0; k < iters; ++k) eat(xs[k * g.ns + g.sl]);   // consecutive threads, consecutive 16 B
            } else {
                const int np = cnt >> 1;
                for (int i = g.sl; i < np; i += g.ns) eat(xs[i]);
                if ((cnt & 1) && g.sl == 0) {
                    const double x = g.tile[st * kResTile + cnt - 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double d0 = x - m[r];
                        acc[r] = fma(d0, d0, acc[r]);
                        if (GRAD) accT[r] += d0;
                    }
                }
            }
            // this warp is done with the stage; the last warp of the group to say so refills it
            __syncwarp();
            if (lane == 0) {
                const unsigned int prev = atomicAdd(&g.done[st], 1u);
                if (prev == kGroupWarps - 1) {
                    g.done[st] = 0u;
                    if (tt + kResStages < g.tiles_total) res_issue(a, g, tt + kResStages);
                }
            }
            g.tiles_done = tt + 1;
            // let the sibling group start: half a sweep of offset keeps its scalar phases inside
            // our sweeps and ours inside its sweeps
            if (signal_half && (int)(tt % (uint32_t)T) == (T >> 1)) { named_bar_arrive(3, 2 * kGroupThreads); signal_half = false; }
        }
        // reduce over the slices: lanes of a warp (butterfly), then the warps of the chain group
        double v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = i < R ? acc[i] : (GRAD && i < 2 * R) ? accT[i - R] : 0.0;
        warp_transpose_reduce<NV>(v, lane);
        constexpr int LPV = 32 / NV;    // lanes holding the same value
        const int idx = lane / LPV;
        if ((lane % LPV) == 0 && idx < R * NQ) {
            const int q = idx / R, r = idx % R;
            g.red[(((og * 2 + q) * cap) + g.cgi * R + r) * wpc + (warp % wpc)] = v[0];
        }
    }
    if (signal_half) named_bar_arrive(3, 2 * kGroupThreads);   // (single-tile sweeps)
    named_bar_sync(g.bar_id, kGroupThreads);
    const int nout = NQ * a.G * g.n;
    for (int j = g.tid; j < nout; j += kGroupThreads) {
        const int local = j % g.n, row = j / g.n;        // row = q * G + og
        const int q = row / a.G, og = row % a.G;
        const double *p = g.red + (((og * 2 + q) * cap) + local) * wpc;
        double s = 0.0;
        for (int w = 0; w < wpc; ++w) s += p[w];
        d.partial[(int64_t)row * C + g.c0 + local] = s;
    }
    named_bar_sync(g.bar_id, kGroupThreads);
}

// sum of the per-group sums of one chain, group order (what reduce_segments does with S = 1)
__device__ __forceinline__ double res_total(const DevState &d, int G, int64_t c) {
    double s = 0.0;
    for (int og = 0; og < G; ++og) s += d.partial[(int64_t)og * d.C + c];
    return s;
}

}  // namespace

template <int R>
__global__ void __launch_bounds__(2 * kGroupThreads, 1)
resident_block_kernel(ResidentArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DevState &d = a.d;
    const int gi = threadIdx.x / kGroupThreads;
    Group g;
    g.tid = threadIdx.x % kGroupThreads;
    g.bar_id = 1 + gi;
    g.cg = a.cg;
    g.ns = kGroupThreads / a.cg;
    g.cgi = g.tid / g.ns;
    g.sl = g.tid % g.ns;
    // chains of this CTA and of this group
    const int64_t b = blockIdx.x;
    const int64_t c_lo = b * a.base + (b < a.rem ? b : a.rem);
    const int n_cta = (int)(a.base + (b < a.rem ? 1 : 0));
    const int n0 = (n_cta + 1) / 2;
    g.c0 = c_lo + (gi ? n0 : 0);
    g.n = gi ? n_cta - n0 : n0;
    // shared memory carve-up (host: resident_smem_bytes)
    const int cap = a.cg * R;
    unsigned char *sp = smem_raw;
    int *tile_start = reinterpret_cast<int *>(sp); sp += ((a.G + 1) * 4 + 127) / 128 * 128;
    const size_t per_group = a.smem_per_group;
    unsigned char *gp = sp + (size_t)gi * per_group;
    g.tile = reinterpret_cast<double *>(gp); gp += (size_t)kResStages * kResTile * 8;
    g.full = reinterpret_cast<uint64_t *>(gp); gp += 64;
    g.done = reinterpret_cast<unsigned int *>(gp); gp += 64;
    g.ctx = reinterpret_cast<StepCtx *>(gp); gp += (sizeof(StepCtx) + 127) / 128 * 128;
    g.red = reinterpret_cast<double *>(gp); gp += (size_t)a.G * 2 * cap * kGroupWarps * 8;
    g.sh_t = reinterpret_cast<double *>(gp); gp += (size_t)a.stage_doubles * 8;
    g.sh_m = reinterpret_cast<double *>(gp); gp += (size_t)a.stage_doubles * 8;
    g.sh_n = reinterpret_cast<double *>(gp);
    g.tile_start = tile_start;
    g.tiles_done = 0;

    if (threadIdx.x == 0) {
        int acc = 0;
        for (int og = 0; og < a.G; ++og) { tile_start[og] = acc; acc += (int)((a.glen[og] + kResTile - 1) / kResTile); }
        tile_start[a.G] = acc;
    }
    if (g.tid == 0) {
        for (int s = 0; s < kResStages; ++s) { mbar_init(&g.full[s], 1); g.done[s] = 0u; }
        mbar_fence_init();
    }
    __syncthreads();
    g.tiles_total = (uint32_t)a.n_sweeps * (uint32_t)tile_start[a.G];
    if (g.n == 0) return;   // (an odd chain left the second group empty)
    if (g.tid == 0)
        for (uint32_t t = 0; t < (uint32_t)kResStages && t < g.tiles_total; ++t) res_issue(a, g, t);
    // the second group starts half a sweep after the first one
    bool signal_half = false;
    if (n_cta - n0 > 0) {
        if (gi == 1) named_bar_sync(3, 2 * kGroupThreads);
        else signal_half = true;
    }

    const GroupSync gsync{g.bar_id};
    const bool owner = g.tid < g.n;
    const int64_t c = g.c0 + g.tid;
    const bool iid = d.law == EXTMCMC_LAW_GSN_IID_1D;
    const bool stage = d.p > 4 && d.p <= kCoopP && a.stage_doubles > 0;   // group-uniform
    const bool coop = stage && d.stats_mode == 0;
    const CoopStage cs{g.sh_t, coop ? g.sh_m : nullptr, g.sh_n, g.n, g.tid};
    for (int k = 0; k < a.n_steps; ++k) {
        load_step_ctx(g.ctx, d, a.descs, k, g.tid, kGroupThreads, gsync);
        const StepDesc &sd = g.ctx->sd;
        const DevUpdate &u = g.ctx->u;
        if (u.kernel == EXTMCMC_KERNEL_MALA) {
            if (sd.need_cur_grad) {
                // gradient at the current state (another update moved it since it was last computed)
                if (iid) { if (owner) law_prepare(d, c, d.theta + c, d.C); gsync(); }
                res_sweep<R, true>(a, g, iid ? d.lawc : d.theta, signal_half);
                signal_half = false;
                if (owner) grad_finalize_chain(d, c, d.theta, a.ll_scratch, d.grad_cur);
            }
            if (owner) mala_propose_chain(d, sd, u, c);
            gsync();
            res_sweep<R, true>(a, g, iid ? d.lawc : d.prop_full, signal_half);
            signal_half = false;
            if (owner) {
                grad_finalize_chain(d, c, d.prop_full, d.ll_prop, d.grad_prop);
                mala_decide(d, sd, u, c, stage ? &cs : nullptr);
            }
        } else {
            if (owner) propose_chain(d, sd, u, c);
            gsync();
            res_sweep<R, false>(a, g, iid ? d.lawc : d.prop_full, signal_half);
            signal_half = false;
            if (owner) {
                const RwPre pre = rw_accept_prologue(d, sd, u, c);
                rw_accept_finish(d, sd, u, c, pre, res_total(d, a.G, c), stage ? &cs : nullptr);
            }
        }
        if (coop) {
            gsync();
            update_cov_coop(d, sd.stat_n, g.c0, g.n, g.sh_t, g.sh_m, g.sh_n, g.tid, kGroupThreads);
        }
        gsync();   // the step context and the staging area are rewritten by the next element
    }
}

// =====================================================================================
// observation-mapped block kernel
// =====================================================================================
namespace {
constexpr int kObNT = 256, kObTile = 2048, kObStages = 4;
constexpr unsigned long long kGoAbort = ~0ull;

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
}  // namespace

template <int CB>
__global__ void __launch_bounds__(kObNT, CB <= 8 ? 3 : CB <= 16 ? 2 : 1)
obs_block_kernel(ObsBlockArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);                       // [kObStages][kObTile]
    uint64_t *full = reinterpret_cast<uint64_t *>(tile + kObStages * kObTile); // [kObStages]
    unsigned int *done = reinterpret_cast<unsigned int *>(full + kObStages);   // [kObStages]
    double *red = reinterpret_cast<double *>(done + kObStages + (kObStages & 1)); // [kObNT/32][CB], then [kObNT]
    __shared__ StepCtx ctx, ctx_next;
    __shared__ int sh_go, sh_last, sh_alive;
    const DevState &d = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kObNT / 32;
    const int seg = blockIdx.x, S = gridDim.x;
    const int64_t C = d.C;

    // this CTA's observation segment (even boundaries: 16-byte units of the bulk copies)
    const int64_t n_pairs = (a.n_obs + 1) >> 1;
    const int64_t lo = 2 * ((int64_t)seg * n_pairs / S);
    int64_t hi = 2 * ((int64_t)(seg + 1) * n_pairs / S);
    if (hi > a.n_obs) hi = a.n_obs;
    const int64_t len = hi - lo;
    const int n_tiles = (int)((len + kObTile - 1) / kObTile);
    const uint32_t tiles_total = (uint32_t)n_tiles * (uint32_t)a.n_steps;

    if (tid == 0) {
        for (int s = 0; s < kObStages; ++s) { mbar_init(&full[s], 1); done[s] = 0u; }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](uint32_t tt) {
        const int t = (int)(tt % (uint32_t)n_tiles), st = (int)(tt % kObStages);
        const int64_t off = (int64_t)t * kObTile;
        const int cnt = (int)((len - off) < (int64_t)kObTile ? (len - off) : (int64_t)kObTile);
        const uint32_t bytes = (uint32_t)((cnt + 1) >> 1) * 16u;
        mbar_expect_tx(&full[st], bytes);
        bulk_g2s(tile + st * kObTile, a.obs + lo + off, bytes, &full[st]);
    };
    if (tid == 0)
        for (uint32_t t = 0; t < (uint32_t)kObStages && t < tiles_total; ++t) issue(t);

    uint32_t tt = 0;
    for (int k = 0; k < a.n_steps; ++k) {
        // wait until the proposal of step k is out: go counts the exchange steps completed so far.
        // The ring keeps filling meanwhile (the observations do not depend on the decision).
        const long long xseq = a.descs[k].xseq;
        if (tid == 0) {
            const unsigned long long t0 = global_timer_ns();
            unsigned long long v;
            int ok = 1;
            for (;;) {
                v = ld_acquire_gpu(a.go);
                if (v == kGoAbort) { ok = 0; break; }
                if (v >= (unsigned long long)xseq) break;
                if (global_timer_ns() - t0 > 2 * d.p2p_timeout_ns) { atomicExch(d.err_flag, 2); ok = 0; break; }
            }
            sh_go = ok;
        }
        __syncthreads();
        if (!sh_go) return;
        double m[CB], acc[CB];
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) { m[cc] = cc < C ? __ldcg(d.lawc + cc) : 0.0; acc[cc] = 0.0; }
        auto eat = [&](const double2 x) {
#pragma unroll
            for (int cc = 0; cc < CB; ++cc) {
                const double d0 = x.x - m[cc];
                acc[cc] = fma(d0, d0, acc[cc]);
                const double d1 = x.y - m[cc];
                acc[cc] = fma(d1, d1, acc[cc]);
            }
        };
        for (int t = 0; t < n_tiles; ++t, ++tt) {
            const int st = (int)(tt % kObStages);
            mbar_wait(&full[st], (tt / kObStages) & 1u);
            const int64_t off = (int64_t)t * kObTile;
            const int cnt = (int)((len - off) < (int64_t)kObTile ? (len - off) : (int64_t)kObTile);
            const double2 *xs = reinterpret_cast<const double2 *>(tile + st * kObTile);
            if (cnt == kObTile) {
#pragma unroll
                for (int q = 0; q < kObTile / 2 / kObNT; ++q) eat(xs[q * kObNT + tid]);
            } else {
                const int np = cnt >> 1;
                for (int i = tid; i < np; i += kObNT) eat(xs[i]);
                if ((cnt & 1) && tid == 0) {
                    const double x = tile[st * kObTile + cnt - 1];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) { const double d0 = x - m[cc]; acc[cc] = fma(d0, d0, acc[cc]); }
                }
            }
            __syncwarp();
            if (lane == 0) {
                const unsigned int prev = atomicAdd(&done[st], 1u);
                if (prev == NW - 1) {
                    done[st] = 0u;
                    if (tt + kObStages < tiles_total) issue(tt + kObStages);
                }
            }
        }
        // fixed-order block reduction: xor-shuffle tree inside each warp, then warp 0..NW-1
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
            double v = acc[cc];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp * CB + cc] = v;
        }
        __syncthreads();
        if (tid < CB && tid < C) {
            double v = 0.0;
            for (int w = 0; w < NW; ++w) v += red[w * CB + tid];
            __stcg(d.partial + (int64_t)seg * C + tid, v);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) sh_last = atomicAdd(a.counter, 1u) == (unsigned int)S - 1 ? 1 : 0;
        __syncthreads();
        if (!sh_last) continue;

        // ---- leader of step k: every segment's sums are in ------------------------------------
        __threadfence();   // (acquire side of the counter; also drops stale L1 lines of the chain state)
        constexpr int NS = kObNT / CB;
        const int ch = tid % CB, sli = tid / CB;
        double part = 0.0;
        if (ch < C) {
            // independent loads in batches, adds in increasing row order
            int i = sli;
            for (; i + 3 * NS < S; i += 4 * NS) {
                double q[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = __ldcg(d.partial + (int64_t)(i + j * NS) * C + ch);
#pragma unroll
                for (int j = 0; j < 4; ++j) part += q[j];
            }
            for (; i < S; i += NS) part += __ldcg(d.partial + (int64_t)i * C + ch);
        }
        double *sh = red + NW * CB;    // [NS][CB]
        sh[sli * CB + ch] = part;
        __syncthreads();
        double tot = 0.0;
        if (tid < CB) for (int j = 0; j < NS; ++j) tot += sh[j * CB + tid];
        bool alive = true;
        if (d.p2p) {
            // cross-rank exchange: store this rank's totals into every rank's slot, raise the flag
            // there, wait for everybody's flag here, add the slots in rank order
            const int parity = (int)(xseq & 1);
            if (tid < CB && tid < C) {
                const int64_t slot = ((int64_t)parity * d.world + d.rank) * C + tid;
                for (int q = 0; q < d.world; ++q) d.peer_rx[q][slot] = tot;
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                const unsigned long long tag = (unsigned long long)(xseq + 1);
                for (int q = 0; q < d.world; ++q) {
                    unsigned long long *f = d.peer_flag[q] + (parity * d.world + d.rank);
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(tag) : "memory");
                }
                sh_alive = wait_peer_flags(d, parity, tag) ? 1 : 0;
            }
            __syncthreads();
            alive = sh_alive != 0;
            if (alive && tid < CB && tid < C) {
                tot = 0.0;
                for (int r = 0; r < d.world; ++r) tot += __ldcg(d.my_rx + ((int64_t)parity * d.world + r) * C + tid);
            }
        }
        if (alive) {
            load_step_ctx(&ctx, d, a.descs, k, tid, kObNT, CtaSync{});
            const bool more = k + 1 < a.n_steps;
            if (more) load_step_ctx(&ctx_next, d, a.descs, k + 1, tid, kObNT, CtaSync{});
            if (tid < CB && tid < C) {
                const RwPre pre = rw_accept_prologue(d, ctx.sd, ctx.u, tid);
                rw_accept_finish(d, ctx.sd, ctx.u, tid, pre, tot);
                if (more) propose_chain(d, ctx_next.sd, ctx_next.u, tid);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            *a.counter = 0u;
            __threadfence();
            st_release_gpu(a.go, alive ? (unsigned long long)(xseq + 1) : kGoAbort);
        }
        if (!alive) return;
    }
}

// proposal of the block's first element (the elements after it are proposed by the leaders) and
// arming of the flags: go = exchange steps completed before this block
__global__ void __launch_bounds__(32)
obs_block_first_kernel(DevState d, const StepDesc *__restrict__ descs, unsigned long long *go, unsigned int *counter) {
    __shared__ StepCtx ctx;
    load_step_ctx(&ctx, d, descs, 0, threadIdx.x, blockDim.x, CtaSync{});
    if (*reinterpret_cast<volatile int32_t *>(d.err_flag) == 2) {   // an exchange failed earlier: stay down
        if (threadIdx.x == 0) *go = kGoAbort;
        return;
    }
    if (threadIdx.x < d.C) propose_chain(d, ctx.sd, ctx.u, threadIdx.x);
    if (threadIdx.x == 0) { *go = (unsigned long long)ctx.sd.xseq; *counter = 0u; }
}

// =====================================================================================
// host side
// =====================================================================================
void launch_obs_block_first(const DevState &d, const StepDesc *descs, unsigned long long *go,
                            unsigned int *counter, cudaStream_t st) {
    obs_block_first_kernel<<<1, 32, 0, st>>>(d, descs, go, counter);
}

static size_t resident_smem_per_group(int G, int cap, int stage_doubles) {
    size_t b = (size_t)kResStages * kResTile * 8 + 64 + 64;
    b += (sizeof(StepCtx) + 127) / 128 * 128;
    b += (size_t)G * 2 * cap * kGroupWarps * 8;
    b += (size_t)3 * stage_doubles * 8;
    return (b + 127) / 128 * 128;
}

bool plan_resident(const DevState &d, int num_sms, bool force, ResidentPlan *pl) {
    const int64_t C = d.C;
    if (d.G > 32) return false;
    if (!force && C < (int64_t)num_sms * 8) return false;
    // waves of CTAs (one CTA per SM at a time), chains split evenly over the CTAs
    const int64_t max_per_cta = 2 * 64;
    const int64_t waves = (C + (int64_t)num_sms * max_per_cta - 1) / ((int64_t)num_sms * max_per_cta);
    int64_t n_cta = (int64_t)num_sms * waves;
    const int64_t most = (C + 7) / 8;              // at least 8 chains per CTA (4 per group)
    if (n_cta > most) n_cta = most > 0 ? most : 1;
    pl->n_cta = (int)n_cta;
    pl->base = C / n_cta;
    pl->rem = C % n_cta;
    const int need = (int)((pl->base + (pl->rem ? 1 : 0) + 1) / 2);   // chains of the larger group
    int best_cap = 1 << 30, best_r = 0, best_cg = 0;
    for (int cg = 1; cg <= 8; cg <<= 1)
        for (int r = 4; r <= 8; ++r) {
            const int cap = cg * r;
            if (cap < need) continue;
            if (cap < best_cap || (cap == best_cap && r > best_r)) { best_cap = cap; best_r = r; best_cg = cg; }
        }
    if (!best_r) return false;
    pl->R = best_r;
    pl->cg = best_cg;
    const bool stage = d.p > 4 && d.p <= kCoopP;
    pl->stage_doubles = stage ? d.p * need : 0;
    pl->smem_per_group = resident_smem_per_group(d.G, best_cap, pl->stage_doubles);
    pl->smem_bytes = ((size_t)(d.G + 1) * 4 + 127) / 128 * 128 + 2 * pl->smem_per_group;
    return pl->smem_bytes <= 220 * 1024;
}

template <int R>
static cudaError_t launch_resident_r(const ResidentPlan &pl, const ResidentArgs &a, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(resident_block_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    resident_block_kernel<R><<<pl.n_cta, 2 * kGroupThreads, pl.smem_bytes, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_resident_block(const ResidentPlan &pl, ResidentArgs a, cudaStream_t st) {
    a.cg = pl.cg;
    a.base = pl.base;
    a.rem = pl.rem;
    a.stage_doubles = pl.stage_doubles;
    a.smem_per_group = (unsigned int)pl.smem_per_group;
    switch (pl.R) {
    case 4: return launch_resident_r<4>(pl, a, st);
    case 5: return launch_resident_r<5>(pl, a, st);
    case 6: return launch_resident_r<6>(pl, a, st);
    case 7: return launch_resident_r<7>(pl, a, st);
    default: return launch_resident_r<8>(pl, a, st);
    }
}

static size_t obs_block_smem(int cb) {
    return (size_t)kObStages * kObTile * 8 + kObStages * 8 + (kObStages + (kObStages & 1)) * 4 +
           (size_t)(kObNT / 32) * cb * 8 + (size_t)kObNT * 8;
}

template <int CB>
static cudaError_t obs_block_grid(int num_sms, int64_t n_obs, int *grid) {
    cudaError_t e = cudaFuncSetAttribute(obs_block_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)obs_block_smem(CB));
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, obs_block_kernel<CB>, kObNT, obs_block_smem(CB));
    if (e != cudaSuccess) return e;
    if (per_sm > 3) per_sm = 3;     // 3 x 64 KB of staging per SM saturate HBM
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    int64_t S = (int64_t)num_sms * per_sm;
    const int64_t max_S = ((n_obs + 1) / 2 + 1023) / 1024;   // at least one tile per segment
    if (S > max_S) S = max_S;
    if (S < 1) S = 1;
    *grid = (int)S;
    return cudaSuccess;
}

cudaError_t plan_obs_block(int cb, int num_sms, int64_t n_obs, int *grid) {
    switch (cb) {
    case 1: return obs_block_grid<1>(num_sms, n_obs, grid);
    case 2: return obs_block_grid<2>(num_sms, n_obs, grid);
    case 4: return obs_block_grid<4>(num_sms, n_obs, grid);
    case 8: return obs_block_grid<8>(num_sms, n_obs, grid);
    case 16: return obs_block_grid<16>(num_sms, n_obs, grid);
    default: return obs_block_grid<32>(num_sms, n_obs, grid);
    }
}

template <int CB>
static cudaError_t launch_obs_block_cb(int grid, ObsBlockArgs &a, cudaStream_t st) {
    void *args[] = {&a};
    // cooperative: all CTAs must be co-resident (they wait for each other's sums every step)
    return cudaLaunchCooperativeKernel((const void *)obs_block_kernel<CB>, dim3(grid), dim3(kObNT), args,
                                       obs_block_smem(CB), st);
}

cudaError_t launch_obs_block(int cb, int grid, ObsBlockArgs a, cudaStream_t st) {
    switch (cb) {
    case 1: return launch_obs_block_cb<1>(grid, a, st);
    case 2: return launch_obs_block_cb<2>(grid, a, st);
    case 4: return launch_obs_block_cb<4>(grid, a, st);
    case 8: return launch_obs_block_cb<8>(grid, a, st);
    case 16: return launch_obs_block_cb<16>(grid, a, st);
    default: return launch_obs_block_cb<32>(grid, a, st);
    }
}

}  // namespace extmcmc
