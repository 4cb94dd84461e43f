// Persistent block kernels: a whole BLOCK of schedule elements (src/run.jl:70-82, the body of
// __run!) in ONE launch -- proposal, likelihood sweep, accept/reject, chain statistics and
// adaptation of every element, back to back on the SMs, no kernel boundary in between.  A CTA
// keeps the state of the chains it decides for in SHARED MEMORY for the whole block (a "view" of
// DevState whose pointers lead into shared memory), so the latency-bound scalar phases run at
// shared-memory latency instead of L2 / HBM latency; the per-chain functions of step_device.cuh
// are reused unchanged on that view.  Two shapes:
//
//  team_block_kernel<R>  (many chains: BASELINE cfg 2 and cfg 4)
//      Two CTAs per SM.  CTAs form teams of `ts` (4): a team owns a contiguous range of chains;
//      every member streams ITS QUARTER of the observations (TMA ring) for ALL chains of the team
//      -- thread <-> (R chains in registers) x (observation slice) -- so each observation byte read
//      from L2 feeds ~110-220 FP64 instructions instead of ~28 (one CTA streaming everything for its
//      own chains is L2-bandwidth-bound: measured).  The members leave their sums in L2, meet at a
//      team barrier (one atomic counter), and each member then decides for ITS share of the chains:
//      fixed-order sum over the members, accept/reject, statistics, adaptation, next proposal, all
//      on the shared-memory view.  Teams never talk to each other.  The second half of the CTAs
//      starts half a sweep late: while one CTA of an SM is in its scalar phase or at a barrier, the
//      other one owns the FP64 pipe; equal work keeps that offset locked.
//
//  obs_block_kernel<CB>  (a handful of chains, huge N: BASELINE cfg 5)
//      Every CTA streams its own observation segment for all chains (HBM-bound) and keeps its TMA
//      ring running ACROSS steps (observations are constant), so HBM stays busy while the step is
//      decided.  CTA 0 holds the chains' state in shared memory and is the decider: when every
//      segment's sums are in, it adds them in a fixed order, exchanges the totals with the other
//      ranks over NVLink peer mappings when observations are sharded (stores into every peer +
//      flag, ordered sum on arrival: compute and collective in one kernel), decides, issues the next
//      proposal and releases a go-flag the other CTAs spin on.  Cooperative launch guarantees
//      co-residency.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "dev_state.cuh"
#include "philox.cuh"
#include "step_device.cuh"
#include "block_kernels.h"
#include "tma.cuh"

namespace extmcmc {

namespace {

constexpr int kBT = 256;                 // threads per CTA (both kernels)
constexpr int kBW = kBT / 32;
constexpr int kTile = 1024;              // observations per TMA tile (8 KB)
constexpr int kStages = 4;               // stages of the observation-mapped kernel's ring
constexpr int kMaxStages = 12;           // the team kernel sizes its ring to the shared memory left (a.stages)
constexpr int kMaxG = 32;                // observation groups served by the team kernel
constexpr unsigned long long kGoAbort = ~0ull;

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// segment s of S over n observations, boundaries on even indices (16-byte units of the bulk copies)
__device__ __forceinline__ void seg_bounds(int64_t n_obs, int s, int S, int64_t &lo, int64_t &hi) {
    const int64_t n_pairs = (n_obs + 1) >> 1;
    lo = 2 * ((int64_t)s * n_pairs / S);
    hi = 2 * ((int64_t)(s + 1) * n_pairs / S);
    if (hi > n_obs) hi = n_obs;
}

// ---------------------------------------------------------------------------------------------
// shared-memory view of the per-chain state of chains [c0, c0 + n)
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void copy_in(T *s, const T *g, int rows, int64_t C, int64_t c0, int n, int tid, int nt) {
    for (int i = tid; i < rows * n; i += nt) s[i] = g[(int64_t)(i / n) * C + c0 + (i % n)];
}
template <class T>
__device__ __forceinline__ void copy_out(const T *s, T *g, int rows, int64_t C, int64_t c0, int n, int tid, int nt) {
    for (int i = tid; i < rows * n; i += nt) g[(int64_t)(i / n) * C + c0 + (i % n)] = s[i];
}
template <class T>
__device__ __forceinline__ T *at(unsigned char *base, uint32_t off) { return reinterpret_cast<T *>(base + off); }

// Copies the state in and writes the view structs (the caller synchronises afterwards).
__device__ __noinline__ void view_build(const DevState &d, const ViewLayout &vl, unsigned char *base, int64_t c0,
                                           int n, int tid, int nt) {
    const int64_t C = d.C;
    copy_in(at<double>(base, vl.theta), d.theta, d.p, C, c0, n, tid, nt);
    copy_in(at<double>(base, vl.ll), d.ll, 1, C, c0, n, tid, nt);
    if (vl.grad_cur != kNotStaged) copy_in(at<double>(base, vl.grad_cur), d.grad_cur, d.p, C, c0, n, tid, nt);
    if (vl.mean != kNotStaged) copy_in(at<double>(base, vl.mean), d.mean, d.p, C, c0, n, tid, nt);
    if (vl.cov != kNotStaged) copy_in(at<double>(base, vl.cov), d.cov, vl.cov_rows, C, c0, n, tid, nt);
    for (int u = 0; u < d.NU; ++u) {
        const DevUpdate &g = d.upd[u];
        copy_in(at<double>(base, vl.eps[u]), g.eps, vl.eps_rows[u], C, c0, n, tid, nt);
        copy_in(at<int32_t>(base, vl.adapt_prop[u]), g.adapt_prop, 1, C, c0, n, tid, nt);
        copy_in(at<int32_t>(base, vl.adapt_acc[u]), g.adapt_acc, 1, C, c0, n, tid, nt);
        copy_in(at<int64_t>(base, vl.tot_prop[u]), g.tot_prop, 1, C, c0, n, tid, nt);
        copy_in(at<int64_t>(base, vl.tot_acc[u]), g.tot_acc, 1, C, c0, n, tid, nt);
        copy_in(at<double>(base, vl.ra_val[u]), g.ra_val, 1, C, c0, n, tid, nt);
        copy_in(at<uint8_t>(base, vl.acc_ring[u]), g.acc_ring, d.W, C, c0, n, tid, nt);
    }
    // the view structs: word copies of the global ones, then the pointers are redirected
    {
        DevUpdate *uv = at<DevUpdate>(base, vl.upd_table);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(d.upd);
        uint32_t *dst = reinterpret_cast<uint32_t *>(uv);
        for (int i = tid; i < (int)(sizeof(DevUpdate) / 4) * d.NU; i += nt) dst[i] = src[i];
    }
    __syncthreads();
    if (tid == 0) {
        DevState *dv = at<DevState>(base, vl.dv);
        *dv = d;
        dv->C = n;
        dv->chain_offset = d.chain_offset + c0;
        dv->gC = C;
        dv->g0 = c0;
        dv->theta = at<double>(base, vl.theta);
        dv->ll = at<double>(base, vl.ll);
        dv->prop_loc = at<double>(base, vl.prop_loc);
        dv->prop_full = at<double>(base, vl.prop_full);
        dv->lawc = at<double>(base, vl.lawc);
        dv->n_used = at<uint32_t>(base, vl.n_used);
        dv->ll_prop = at<double>(base, vl.ll_prop);
        if (vl.grad_cur != kNotStaged) { dv->grad_cur = at<double>(base, vl.grad_cur); dv->grad_prop = at<double>(base, vl.grad_prop); }
        if (vl.mean != kNotStaged) dv->mean = at<double>(base, vl.mean);
        if (vl.cov != kNotStaged) dv->cov = at<double>(base, vl.cov);
        DevUpdate *uv = at<DevUpdate>(base, vl.upd_table);
        dv->upd = uv;
        for (int u = 0; u < d.NU; ++u) {
            uv[u].eps = at<double>(base, vl.eps[u]);
            uv[u].adapt_prop = at<int32_t>(base, vl.adapt_prop[u]);
            uv[u].adapt_acc = at<int32_t>(base, vl.adapt_acc[u]);
            uv[u].tot_prop = at<int64_t>(base, vl.tot_prop[u]);
            uv[u].tot_acc = at<int64_t>(base, vl.tot_acc[u]);
            uv[u].ra_val = at<double>(base, vl.ra_val[u]);
            uv[u].acc_ring = at<uint8_t>(base, vl.acc_ring[u]);
        }
    }
}

// Writes everything a block may have changed back to the global state.
__device__ __noinline__ void view_flush(const DevState &d, const ViewLayout &vl, unsigned char *base, int64_t c0,
                                           int n, int tid, int nt) {
    const int64_t C = d.C;
    copy_out(at<double>(base, vl.theta), d.theta, d.p, C, c0, n, tid, nt);
    copy_out(at<double>(base, vl.ll), d.ll, 1, C, c0, n, tid, nt);
    if (vl.grad_cur != kNotStaged) copy_out(at<double>(base, vl.grad_cur), d.grad_cur, d.p, C, c0, n, tid, nt);
    if (vl.mean != kNotStaged) copy_out(at<double>(base, vl.mean), d.mean, d.p, C, c0, n, tid, nt);
    if (vl.cov != kNotStaged) copy_out(at<double>(base, vl.cov), d.cov, vl.cov_rows, C, c0, n, tid, nt);
    for (int u = 0; u < d.NU; ++u) {
        const DevUpdate &g = d.upd[u];
        copy_out(at<double>(base, vl.eps[u]), g.eps, vl.eps_rows[u], C, c0, n, tid, nt);
        copy_out(at<int32_t>(base, vl.adapt_prop[u]), g.adapt_prop, 1, C, c0, n, tid, nt);
        copy_out(at<int32_t>(base, vl.adapt_acc[u]), g.adapt_acc, 1, C, c0, n, tid, nt);
        copy_out(at<int64_t>(base, vl.tot_prop[u]), g.tot_prop, 1, C, c0, n, tid, nt);
        copy_out(at<int64_t>(base, vl.tot_acc[u]), g.tot_acc, 1, C, c0, n, tid, nt);
        copy_out(at<double>(base, vl.ra_val[u]), g.ra_val, 1, C, c0, n, tid, nt);
        copy_out(at<uint8_t>(base, vl.acc_ring[u]), g.acc_ring, d.W, C, c0, n, tid, nt);
    }
}

// The scalar phases are latency-insensitive and register-hungry; the sweeps are the opposite.  Kept
// out of line, the scalar code does not weigh on the register allocation (and therefore on the
// instruction scheduling) of the sweep loops.
__device__ __noinline__ void blk_rw_propose(const DevState *dv, const StepCtx *ctx, int lc) {
    propose_chain(*dv, ctx->sd, ctx->u, lc);
}
__device__ __noinline__ void blk_rw_accept(const DevState *dv, const StepCtx *ctx, int lc, double S, const CoopStage *cs) {
    const RwPre pre = rw_accept_prologue(*dv, ctx->sd, ctx->u, lc);
    rw_accept_finish(*dv, ctx->sd, ctx->u, lc, pre, S, cs);
}
__device__ __noinline__ void blk_prepare_current(const DevState *dv, int lc) { law_prepare(*dv, lc, dv->theta + lc, dv->C); }
__device__ __noinline__ void blk_grad_current(const DevState *dv, int lc, double *ll_scratch) {
    grad_finalize_chain(*dv, lc, dv->theta, ll_scratch, dv->grad_cur);
}
__device__ __noinline__ void blk_mala_propose(const DevState *dv, const StepCtx *ctx, int lc) {
    mala_propose_chain(*dv, ctx->sd, ctx->u, lc);
}
__device__ __noinline__ void blk_mala_accept(const DevState *dv, const StepCtx *ctx, int lc, const CoopStage *cs) {
    grad_finalize_chain(*dv, lc, dv->prop_full, dv->ll_prop, dv->grad_prop);
    mala_decide(*dv, ctx->sd, ctx->u, lc, cs);
}
__device__ __noinline__ void blk_cov_coop(const DevState *d, int64_t N, int64_t c0, int n, const double *sh_t,
                                          const double *sh_m, const double *sh_n, int tid) {
    update_cov_coop(*d, N, c0, n, sh_t, sh_m, sh_n, tid, kBT);
}

template <int V> struct Pow2Ceil { static constexpr int value = V <= 1 ? 1 : V <= 2 ? 2 : V <= 4 ? 4 : V <= 8 ? 8 : 16; };

// Sum of NV values per lane over each aligned group of 8 lanes: at every butterfly step a lane
// keeps one half of its values and hands the other half to its partner, so the whole reduction
// costs ~NV adds instead of 3 NV.  On return a lane holds NV / 8 totals (at least one), value
// index = octet_value_base() + k; fixed order.
template <int NV>
__device__ __forceinline__ void octet_transpose_reduce(double (&v)[NV], int lane) {
    int len = NV;
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
        if (len > 1) {
            const int half = len >> 1;
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int k = 0; k < NV / 2; ++k)
                if (k < half) {
                    const double keep = upper ? v[k + half] : v[k];
                    const double send = upper ? v[k] : v[k + half];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            len = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        }
    }
}
template <int NV>
__device__ __forceinline__ int octet_value_base(int lane) {
    int base = 0, len = NV;
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1)
        if (len > 1) { len >>= 1; if (lane & o) base += len; }
    return base;
}

// CTA-uniform state of a team-kernel CTA.  It lives in SHARED memory on purpose: values kept in
// registers across the (out-of-line, register-hungry) scalar phases would add to those phases'
// register demand, push the kernel over its 128-register budget and make ptxas fall back to a
// register-minimising schedule for the sweep loops too (DADD -> DFMA chained through one temporary,
// FP64 pipe 58 % busy: measured).  Read from shared memory where needed, nothing but two counters is
// live across a call.
struct TeamU {
    double *tile;          // [kStages][kTile]
    uint64_t *full;        // [kStages] "tile landed" mbarriers
    unsigned int *done;    // [kStages] warps that have finished the stage (the last one refills it)
    int64_t *seg_lo;       // [G] first observation of this member's slice of group g (offset in the obs array)
    int *seg_len;          // [G]
    int *tile_start;       // [G + 1] first tile of every observation group within a sweep
    double *red;           // [G][16][cg][octets per chain group]
    StepCtx *ctx;
    double *sh_t, *sh_m, *sh_n;
    unsigned char *view;
    const DevState *dv;    // the shared-memory view of this member's own chains
    int flag;
    int team, member, phase;
    int64_t tc0;           // first chain of the team
    int nt;                // chains of the team
    int64_t oc0;           // first chain this member decides for
    int n;                 // ... and how many
    int cg, ns, noct;
    uint32_t tiles_total;
    unsigned int *ctr;     // the team's barrier counter
    unsigned int *half_set;   // flag this CTA raises halfway through its first sweep (phase 0), or nullptr
};

__device__ __forceinline__ void team_issue(const TeamArgs &a, const TeamU &u, uint32_t tt) {
    const int T = u.tile_start[a.G];
    const int ti = (int)(tt % (uint32_t)T);
    int og = 0;
    while (u.tile_start[og + 1] <= ti) ++og;
    const int off = (ti - u.tile_start[og]) * kTile;
    const int rest = u.seg_len[og] - off;
    const int cnt = rest < kTile ? rest : kTile;
    const uint32_t bytes = (uint32_t)((cnt + 1) >> 1) * 16u;   // padded device buffer
    const int st = (int)(tt % (uint32_t)a.stages);
    mbar_expect_tx(&u.full[st], bytes);
    bulk_g2s(u.tile + st * kTile, a.obs + u.seg_lo[og] + off, bytes, &u.full[st]);
}

// All members of the team have arrived (and their global writes are visible).  false on timeout.
// (scalars by value: nothing of the caller's state is forced into local memory by this call)
__device__ __noinline__ bool team_barrier_wait(unsigned int *ctr, unsigned int target, unsigned long long timeout_ns,
                                               int32_t *err_flag) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const unsigned long long t0 = global_timer_ns();
    bool ok = true;
    while (ld_acquire_gpu(ctr) < target)
        if (global_timer_ns() - t0 > timeout_ns) { atomicExch(err_flag, 3); ok = false; break; }
    __threadfence();
    return ok;
}
__device__ __forceinline__ bool team_barrier(const TeamArgs &a, TeamU &u, unsigned int &bar_count) {
    if (a.ts == 1) { __syncthreads(); return true; }
    bar_count += 1;
    __syncthreads();
    if (threadIdx.x == 0)
        u.flag = team_barrier_wait(u.ctr, bar_count * (unsigned int)a.ts, a.d.p2p_timeout_ns, a.d.err_flag) ? 1 : 0;
    __syncthreads();
    return u.flag != 0;
}

// One likelihood sweep: this member's slice of every observation group, for ALL chains of the team.
// mu[g][C] (global): per-chain mean of observation group g, published by the chains' deciders.
// Leaves sum (x - mu)^2 in partial[g * ts + member][c] and, with GRAD, sum (x - mu) in
// partial[(G + g) * ts + member][c] -- the layout the per-chain finalize functions read with S = ts.
template <int R, bool GRAD>
__device__ __forceinline__ void team_sweep(const TeamArgs &a, const TeamU &u, uint32_t &tiles_done,
                                           const double *__restrict__ mu) {
    constexpr int NQ = GRAD ? 2 : 1;
    constexpr int NV = Pow2Ceil<R * NQ>::value;
    constexpr int NKEEP = NV >= 8 ? NV / 8 : 1;
    const int64_t C = a.d.C;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ns = u.ns, cg = u.cg, noct = u.noct, nt_ch = u.nt;
    const int cgi = tid / ns, sl = tid % ns;
    const int64_t tc0 = u.tc0;
    const int T = u.tile_start[a.G];
    const uint32_t tiles_total = u.tiles_total;
    const uint32_t stages = (uint32_t)a.stages;
    unsigned int *half_set = u.half_set;
    double *const tile = u.tile;
    for (int og = 0; og < a.G; ++og) {
        double m[R], acc[R], accT[GRAD ? R : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int local = cgi * R + r;
            m[r] = local < nt_ch ? __ldcg(mu + (int64_t)og * C + tc0 + local) : 0.0;
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
        const int len = u.seg_len[og];
        const int ntile = u.tile_start[og + 1] - u.tile_start[og];
        for (int ti = 0; ti < ntile; ++ti) {
            const uint32_t tt = tiles_done;
            const int st = (int)(tt % stages);
            mbar_wait(&u.full[st], (tt / stages) & 1u);
            const int rest = len - ti * kTile;
            const int cnt = rest < kTile ? rest : kTile;
            const double2 *xs = reinterpret_cast<const double2 *>(tile + st * kTile);
            if (cnt == kTile) {
                const int iters = (kTile / 2) / ns;
#pragma unroll 4
                for (int k = 0; k < iters; ++k) eat(xs[k * ns + sl]);
            } else {
                const int np = cnt >> 1;
                for (int i = sl; i < np; i += ns) eat(xs[i]);
                if ((cnt & 1) && sl == 0) {
                    const double x = tile[st * kTile + cnt - 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double d0 = x - m[r];
                        acc[r] = fma(d0, d0, acc[r]);
                        if (GRAD) accT[r] += d0;
                    }
                }
            }
            // this warp is done with the stage; the last warp of the CTA to say so refills it
            __syncwarp();
            if (lane == 0) {
                const unsigned int prev = atomicAdd(&u.done[st], 1u);
                if (prev == kBW - 1) {
                    u.done[st] = 0u;
                    if (tt + stages < tiles_total) team_issue(a, u, tt + stages);
                }
            }
            tiles_done = tt + 1;
            // halfway through the first sweep: let the CTA that shares this SM start
            if (half_set && tid == 0 && (int)tt == (T >> 1)) st_release_gpu(half_set, 1u);
        }
        // reduce over the observation slices: lanes of an octet (butterfly), octets later
        double v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = i < R ? acc[i] : (GRAD && i < 2 * R) ? accT[i - R] : 0.0;
        octet_transpose_reduce<NV>(v, lane);
        const int vbase = octet_value_base<NV>(lane);
        const int oct = sl >> 3;
        if (NV >= 8 || !(lane & 4)) {
#pragma unroll
            for (int k = 0; k < NKEEP; ++k) {
                const int idx = vbase + k;
                if (idx < R * NQ) u.red[((og * 16 + idx) * cg + cgi) * noct + oct] = v[k];
            }
        }
    }
    __syncthreads();
    // octets of a chain group in order -> this member's partial sums, for every chain of the team
    const int nout = NQ * a.G * nt_ch;
    for (int j = tid; j < nout; j += kBT) {
        const int local = j % nt_ch, row = j / nt_ch;          // row = q * G + og
        const int q = row / a.G, og = row % a.G;
        const int lcg = local / R, r = local % R;
        const double *p = u.red + ((og * 16 + q * R + r) * cg + lcg) * noct;
        double sum = 0.0;
        for (int w = 0; w < noct; ++w) sum += p[w];
        __stcg(a.d.partial + ((int64_t)row * a.ts + u.member) * C + tc0 + local, sum);
    }
}

// sum over the observation groups and members of one chain, row order (the sums of a random-walk step)
__device__ __forceinline__ double team_total(const DevState &dv, int G, int ts, int lc) {
    const double *p = dv.partial + dv.g0 + lc;
    double s = 0.0;
    for (int row = 0; row < G * ts; ++row) s += __ldcg(p + (int64_t)row * dv.gC);
    return s;
}

// publish the means the team sweeps with: mu (iid law) or theta_1..G (hierarchical law) of `src`
__device__ __forceinline__ void team_publish(const TeamArgs &a, const TeamU &u, const double *src_view, int lc) {
    const DevState &d = a.d;
    if (d.law == EXTMCMC_LAW_GSN_IID_1D) __stcg(d.lawc + u.oc0 + lc, u.dv->lawc[lc]);
    else
        for (int og = 0; og < a.G; ++og) __stcg(d.lawc + (int64_t)og * d.C + u.oc0 + lc, src_view[og * u.n + lc]);
}

// set-up of the CTA-uniform state (one thread)
__device__ __noinline__ void team_setup(const TeamArgs &a, TeamU &u, unsigned char *smem_raw) {
    unsigned char *sp = smem_raw;
    u.tile = reinterpret_cast<double *>(sp); sp += (size_t)a.stages * kTile * 8;
    u.full = reinterpret_cast<uint64_t *>(sp); sp += kMaxStages * 8;
    u.done = reinterpret_cast<unsigned int *>(sp); sp += kMaxStages * 4 + 16;
    u.seg_lo = reinterpret_cast<int64_t *>(sp); sp += kMaxG * 8;
    u.seg_len = reinterpret_cast<int *>(sp); sp += kMaxG * 4;
    u.tile_start = reinterpret_cast<int *>(sp); sp += (kMaxG + 1) * 4 + 124;
    sp = smem_raw + (((sp - smem_raw) + 127) & ~(size_t)127);
    u.ctx = reinterpret_cast<StepCtx *>(sp); sp += (sizeof(StepCtx) + 127) / 128 * 128;
    u.red = reinterpret_cast<double *>(sp); sp += (size_t)a.G * 16 * (kBT / 8) * 8;
    u.sh_t = reinterpret_cast<double *>(sp); sp += (size_t)a.stage_doubles * 8;
    u.sh_m = reinterpret_cast<double *>(sp); sp += (size_t)a.stage_doubles * 8;
    u.sh_n = reinterpret_cast<double *>(sp); sp += (size_t)a.stage_doubles * 8;
    sp = smem_raw + (((sp - smem_raw) + 15) & ~(size_t)15);
    u.view = sp;
    u.dv = at<DevState>(u.view, a.vl.dv);
    u.flag = 1;
    u.team = blockIdx.x / a.ts;
    u.member = blockIdx.x % a.ts;
    u.phase = (a.phases == 2 && blockIdx.x >= gridDim.x / 2) ? 1 : 0;
    u.tc0 = (int64_t)u.team * a.base + (u.team < a.rem ? u.team : a.rem);
    u.nt = (int)(a.base + (u.team < a.rem ? 1 : 0));
    const int ob = u.nt / a.ts, orem = u.nt % a.ts;
    u.n = ob + (u.member < orem ? 1 : 0);
    u.oc0 = u.tc0 + (int64_t)u.member * ob + (u.member < orem ? u.member : orem);
    u.cg = a.cg; u.ns = kBT / a.cg; u.noct = u.ns >> 3;
    u.ctr = a.sync + u.team;
    u.half_set = (a.phases == 2 && u.phase == 0) ? a.sync + a.n_team + blockIdx.x : nullptr;
    int acc = 0;
    for (int og = 0; og < a.G; ++og) {
        int64_t lo, hi;
        seg_bounds(a.glen[og], u.member, a.ts, lo, hi);
        u.seg_lo[og] = a.goff[og] + lo;
        u.seg_len[og] = (int)(hi - lo);
        u.tile_start[og] = acc;
        acc += (int)((hi - lo + kTile - 1) / kTile);
    }
    u.tile_start[a.G] = acc;
    u.tiles_total = (uint32_t)a.n_sweeps * (uint32_t)acc;
    for (int q = 0; q < a.stages; ++q) { mbar_init(&u.full[q], 1); u.done[q] = 0u; }
    mbar_fence_init();
}

}  // namespace

// (no minimum-blocks clause in the launch bounds: with one, ptxas schedules the whole kernel for
// minimum register use once the out-of-line scalar code reaches the cap, and chains the sweep's
// DADD -> DFMA pairs through one temporary; without it the kernel still fits two CTAs per SM)
template <int R>
__global__ void __launch_bounds__(kBT)
team_block_kernel(const __grid_constant__ TeamArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ TeamU u;
    const DevState &d = a.d;
    const int tid = threadIdx.x;
    if (tid == 0) team_setup(a, u, smem_raw);
    __syncthreads();
    view_build(d, a.vl, u.view, u.oc0, u.n, tid, kBT);
    __syncthreads();
    // (a member whose slices are all empty -- fewer observations than members -- has no tiles at all)
    if (tid == 0)
        for (uint32_t q = 0; q < (uint32_t)a.stages && q < u.tiles_total; ++q) team_issue(a, u, q);
    // the second half of the CTAs starts when its SM-mate is halfway through its first sweep
    if (u.phase == 1) {
        if (tid == 0) {
            const unsigned int *f = a.sync + a.n_team + (blockIdx.x - gridDim.x / 2);
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_gpu(f) == 0u)
                if (global_timer_ns() - t0 > a.d.p2p_timeout_ns) break;   // (only the overlap is lost)
        }
        __syncthreads();
    }
    if (u.half_set && u.tile_start[a.G] == 0 && tid == 0) st_release_gpu(u.half_set, 1u);

    const CtaSync csync{};
    const bool iid = d.law == EXTMCMC_LAW_GSN_IID_1D;
    const bool stage = d.p > 4 && d.p <= kCoopP && a.stage_doubles > 0;   // CTA-uniform
    const bool coop = stage && d.stats_mode == 0;
    uint32_t tiles_done = 0;
    unsigned int bar_count = 0;
    bool alive = true;
    // optional per-CTA cycle breakdown (diagnostics, a.prof != nullptr): sweep / barrier 1 / barrier 2 / scalar
    long long pc[7] = {0, 0, 0, 0, 0, 0, 0}, pt = a.prof ? clock64() : 0;
    auto lap = [&](int i) { if (a.prof) { const long long now = clock64(); pc[i] += now - pt; pt = now; } };
    for (int k = 0; k < a.n_steps && alive; ++k) {
        load_step_ctx(u.ctx, *u.dv, a.descs, k, tid, kBT, csync);
        lap(6);
        const int kernel = u.ctx->u.kernel;
        const bool owner = tid < u.n;
        const int lc = tid;                                 // local chain index in the view
        const CoopStage cs{u.sh_t, coop ? u.sh_m : nullptr, u.sh_n, u.n, tid};
        if (kernel == EXTMCMC_KERNEL_MALA) {
            if (u.ctx->sd.need_cur_grad) {
                // gradient at the current state (another update moved it since it was last computed)
                if (owner) {
                    if (iid) blk_prepare_current(u.dv, lc);
                    team_publish(a, u, u.dv->theta, lc);
                }
                lap(3);
                if (!(alive = team_barrier(a, u, bar_count))) break;
                lap(1);
                team_sweep<R, true>(a, u, tiles_done, d.lawc);
                lap(0);
                if (!(alive = team_barrier(a, u, bar_count))) break;
                lap(2);
                if (owner) blk_grad_current(u.dv, lc, a.ll_scratch + u.oc0);
            }
            if (owner) {
                blk_mala_propose(u.dv, u.ctx, lc);
                team_publish(a, u, u.dv->prop_full, lc);
            }
            lap(3);
            if (!(alive = team_barrier(a, u, bar_count))) break;
            lap(1);
            team_sweep<R, true>(a, u, tiles_done, d.lawc);
            lap(0);
            if (!(alive = team_barrier(a, u, bar_count))) break;
            lap(2);
            if (owner) blk_mala_accept(u.dv, u.ctx, lc, stage ? &cs : nullptr);
            lap(4);
        } else {
            if (owner) {
                blk_rw_propose(u.dv, u.ctx, lc);
                team_publish(a, u, u.dv->prop_full, lc);
            }
            lap(3);
            if (!(alive = team_barrier(a, u, bar_count))) break;
            lap(1);
            team_sweep<R, false>(a, u, tiles_done, d.lawc);
            lap(0);
            if (!(alive = team_barrier(a, u, bar_count))) break;
            lap(2);
            if (owner) blk_rw_accept(u.dv, u.ctx, lc, team_total(*u.dv, a.G, a.ts, lc), stage ? &cs : nullptr);
            lap(4);
        }
        if (coop) {
            __syncthreads();
            blk_cov_coop(&d, u.ctx->sd.stat_n, u.oc0, u.n, u.sh_t, u.sh_m, u.sh_n, tid);   // (covariance stays global)
        }
        __syncthreads();   // the step context and the staging area are rewritten by the next element
        lap(5);
    }
    if (!alive) return;
    view_flush(d, a.vl, u.view, u.oc0, u.n, tid, kBT);
    if (a.prof && tid == 0) {
        long long *o = a.prof + (int64_t)blockIdx.x * 8;
        for (int i = 0; i < 7; ++i) o[i] += pc[i];
        o[7] += 1;
    }
}

// =====================================================================================
// observation-mapped block kernel
// =====================================================================================
namespace {
constexpr int kObTile = 2048, kObStages = 4;
}

template <int CB>
__global__ void __launch_bounds__(kBT)
obs_block_kernel(ObsBlockArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);                       // [kObStages][kObTile]
    uint64_t *full = reinterpret_cast<uint64_t *>(tile + kObStages * kObTile); // [kObStages]
    unsigned int *done = reinterpret_cast<unsigned int *>(full + kObStages);   // [kObStages]
    double *red = reinterpret_cast<double *>(done + kObStages + (kObStages & 1)); // [kBW][CB], then [kBT]
    unsigned char *view = reinterpret_cast<unsigned char *>(red + kBW * CB + kBT);   // CTA 0 only
    __shared__ StepCtx ctx, ctx_next;
    __shared__ int sh_go, sh_alive;
    const DevState &d = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = blockIdx.x, S = gridDim.x;
    const int64_t C = d.C;
    const bool decider = blockIdx.x == 0;

    // this CTA's observation segment
    int64_t lo, hi;
    seg_bounds(a.n_obs, seg, S, lo, hi);
    const int64_t len = hi - lo;
    const int n_tiles = (int)((len + kObTile - 1) / kObTile);
    const uint32_t tiles_total = (uint32_t)n_tiles * (uint32_t)a.n_steps;

    if (tid == 0) {
        for (int q = 0; q < kObStages; ++q) { mbar_init(&full[q], 1); done[q] = 0u; }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](uint32_t tt) {
        const int ti = (int)(tt % (uint32_t)n_tiles), st = (int)(tt % kObStages);
        const int64_t off = (int64_t)ti * kObTile;
        const int cnt = (int)((len - off) < (int64_t)kObTile ? (len - off) : (int64_t)kObTile);
        const uint32_t bytes = (uint32_t)((cnt + 1) >> 1) * 16u;
        mbar_expect_tx(&full[st], bytes);
        bulk_g2s(tile + st * kObTile, a.obs + lo + off, bytes, &full[st]);
    };
    if (tid == 0)
        for (uint32_t q = 0; q < (uint32_t)kObStages && q < tiles_total; ++q) issue(q);

    // The decider keeps the chains in shared memory for the whole block and issues the first
    // proposal; go = number of proposals published so far = xseq of the step they belong to, + 1.
    const DevState *dvp = nullptr;
    if (decider) {
        view_build(d, a.vl, view, 0, (int)C, tid, kBT);
        __syncthreads();
        dvp = at<DevState>(view, a.vl.dv);
        load_step_ctx(&ctx, *dvp, a.descs, 0, tid, kBT, CtaSync{});
        const bool dead = *reinterpret_cast<volatile int32_t *>(d.err_flag) >= 2;   // an exchange failed earlier: stay down
        if (!dead && tid < C) {
            blk_rw_propose(dvp, &ctx, tid);
            __stcg(d.lawc + tid, dvp->lawc[tid]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(a.go, dead ? kGoAbort : (unsigned long long)(a.descs[0].xseq + 1));
        if (dead) return;
    }

    uint32_t tt = 0;
    for (int k = 0; k < a.n_steps; ++k) {
        // wait for the proposal of step k; the ring keeps filling meanwhile
        const long long xseq = a.descs[k].xseq;
        if (tid == 0) {
            const unsigned long long t0 = global_timer_ns();
            int ok = 1;
            for (;;) {
                const unsigned long long v = ld_acquire_gpu(a.go);
                if (v == kGoAbort) { ok = 0; break; }
                if (v >= (unsigned long long)(xseq + 1)) break;
                if (global_timer_ns() - t0 > 2 * d.p2p_timeout_ns) { atomicExch(d.err_flag, 3); ok = 0; break; }
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
        for (int ti = 0; ti < n_tiles; ++ti, ++tt) {
            const int st = (int)(tt % kObStages);
            mbar_wait(&full[st], (tt / kObStages) & 1u);
            const int64_t off = (int64_t)ti * kObTile;
            const int cnt = (int)((len - off) < (int64_t)kObTile ? (len - off) : (int64_t)kObTile);
            const double2 *xs = reinterpret_cast<const double2 *>(tile + st * kObTile);
            if (cnt == kObTile) {
#pragma unroll
                for (int q = 0; q < kObTile / 2 / kBT; ++q) eat(xs[q * kBT + tid]);
            } else {
                const int np = cnt >> 1;
                for (int i = tid; i < np; i += kBT) eat(xs[i]);
                if ((cnt & 1) && tid == 0) {
                    const double x = tile[st * kObTile + cnt - 1];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) { const double d0 = x - m[cc]; acc[cc] = fma(d0, d0, acc[cc]); }
                }
            }
            __syncwarp();
            if (lane == 0) {
                const unsigned int prev = atomicAdd(&done[st], 1u);
                if (prev == kBW - 1) {
                    done[st] = 0u;
                    if (tt + kObStages < tiles_total) issue(tt + kObStages);
                }
            }
        }
        // fixed-order block reduction: xor-shuffle tree inside each warp, then warp 0..7
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
            for (int w = 0; w < kBW; ++w) v += red[w * CB + tid];
            __stcg(d.partial + (int64_t)seg * C + tid, v);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(a.counter, 1u);
        if (!decider) continue;

        // ---- the decider: wait for every segment's sums, add them in segment order ----------------
        if (tid == 0) {
            const unsigned long long t0 = global_timer_ns();
            int ok = 1;
            while (ld_acquire_gpu(a.counter) < (unsigned int)S)
                if (global_timer_ns() - t0 > 2 * d.p2p_timeout_ns) { atomicExch(d.err_flag, 3); ok = 0; break; }
            *a.counter = 0u;        // (nobody adds again before go is released)
            __threadfence();
            sh_alive = ok;
        }
        __syncthreads();
        bool alive = sh_alive != 0;
        constexpr int NS = kBT / CB;
        const int ch = tid % CB, sli = tid / CB;
        double part = 0.0;
        if (alive && ch < C) {
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
        double *sh = red + kBW * CB;    // [NS][CB]
        sh[sli * CB + ch] = part;
        __syncthreads();
        double tot = 0.0;
        if (tid < CB) for (int j = 0; j < NS; ++j) tot += sh[j * CB + tid];
        if (alive && d.p2p) {
            // cross-rank exchange: store this rank's totals into every rank's slot, raise the flag
            // there, wait for everybody's flag here, add the slots in rank order
            const int parity = (int)(xseq & 1);
            const uint32_t tag = (uint32_t)(xseq + 1);
            int mine = 1;
            if (tid < CB && tid < C) {
                const int64_t cell = ll_cell(parity, d.world, d.rank, C, tid);
                for (int q = 0; q < d.world; ++q) ll_store(d.peer_rx[q] + cell, tot, tag);
                const unsigned long long t0 = global_timer_ns();
                tot = 0.0;
                for (int r = 0; r < d.world && mine; ++r) {
                    double v;
                    if (ll_load(d.my_rx + ll_cell(parity, d.world, r, C, tid), tag, v, t0, d.p2p_timeout_ns)) tot += v;
                    else { atomicExch(d.err_flag, 2); mine = 0; }
                }
            }
            alive = __syncthreads_and(mine) != 0;
        }
        const bool more = k + 1 < a.n_steps;
        if (alive) {
            load_step_ctx(&ctx, *dvp, a.descs, k, tid, kBT, CtaSync{});
            if (more) load_step_ctx(&ctx_next, *dvp, a.descs, k + 1, tid, kBT, CtaSync{});
            if (tid < CB && tid < C) {
                blk_rw_accept(dvp, &ctx, tid, tot, nullptr);
                if (more) {
                    blk_rw_propose(dvp, &ctx_next, tid);
                    __stcg(d.lawc + tid, dvp->lawc[tid]);
                }
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0 && (more || !alive)) st_release_gpu(a.go, alive ? (unsigned long long)(xseq + 2) : kGoAbort);
        if (!alive) return;
    }
    if (decider) view_flush(d, a.vl, view, 0, (int)C, tid, kBT);
}

// =====================================================================================
// host side
// =====================================================================================
bool make_view_layout(const DevState &d, const DevUpdate *upd_host, int n_cap, ViewLayout *vl) {
    if (d.NU > kBlkMaxUpd || d.n_haario > 0) return false;
    std::memset(vl, 0, sizeof *vl);
    uint32_t off = 0;
    auto take = [&](size_t bytes) { const uint32_t o = off; off += (uint32_t)((bytes + 15) & ~(size_t)15); return o; };
    bool any_mala = false;
    int pl_rows = 1;
    for (int u = 0; u < d.NU; ++u) {
        if (upd_host[u].kernel == EXTMCMC_KERNEL_MALA) any_mala = true;
        else if (upd_host[u].kernel == EXTMCMC_KERNEL_RW_UNIFORM) pl_rows = std::max(pl_rows, upd_host[u].n_coords);
        else return false;
    }
    const size_t n = (size_t)n_cap;
    vl->n_cap = n_cap;
    vl->pl_rows = pl_rows;
    vl->theta = take(8 * d.p * n);
    vl->ll = take(8 * n);
    vl->prop_loc = take(8 * pl_rows * n);
    vl->prop_full = take(8 * d.p * n);
    vl->lawc = take(8 * d.lawc_k * n);
    vl->n_used = take(4 * n);
    vl->ll_prop = take(8 * n);
    vl->grad_cur = any_mala ? take(8 * d.p * n) : kNotStaged;
    vl->grad_prop = any_mala ? take(8 * d.p * n) : kNotStaged;
    vl->mean = d.stats_mode != 2 ? take(8 * d.p * n) : kNotStaged;
    // covariance: variances, or a full matrix of a handful of parameters, live in the view; a larger
    // full matrix stays global and is updated cooperatively (update_cov_coop)
    if (d.stats_mode == 1) { vl->cov_rows = d.p; vl->cov = take(8 * d.p * n); }
    else if (d.stats_mode == 0 && d.p <= 4) { vl->cov_rows = d.p * d.p; vl->cov = take(8 * d.p * d.p * n); }
    else if (d.stats_mode == 0 && d.p > kCoopP) return false;
    else { vl->cov_rows = 0; vl->cov = kNotStaged; }
    for (int u = 0; u < d.NU; ++u) {
        vl->eps_rows[u] = upd_host[u].kernel == EXTMCMC_KERNEL_MALA ? 1 : upd_host[u].n_coords;
        vl->eps[u] = take(8 * vl->eps_rows[u] * n);
        vl->adapt_prop[u] = take(4 * n);
        vl->adapt_acc[u] = take(4 * n);
        vl->tot_prop[u] = take(8 * n);
        vl->tot_acc[u] = take(8 * n);
        vl->ra_val[u] = take(8 * n);
        vl->acc_ring[u] = take((size_t)d.W * n);
    }
    vl->upd_table = take(sizeof(DevUpdate) * d.NU);
    vl->dv = take(sizeof(DevState));
    vl->bytes = off;
    return true;
}

template <int R>
static bool team_fits_r(size_t smem, int n_cta, int num_sms) {
    if (cudaFuncSetAttribute(team_block_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, team_block_kernel<R>, kBT, smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return (int64_t)per_sm * num_sms >= n_cta;     // all CTAs co-resident (they wait for each other)
}
static bool team_fits(int R, size_t smem, int n_cta, int num_sms) {
    switch (R) {
    case 4: return team_fits_r<4>(smem, n_cta, num_sms);
    case 5: return team_fits_r<5>(smem, n_cta, num_sms);
    case 6: return team_fits_r<6>(smem, n_cta, num_sms);
    case 7: return team_fits_r<7>(smem, n_cta, num_sms);
    default: return team_fits_r<8>(smem, n_cta, num_sms);
    }
}

bool plan_team(const DevState &d, const DevUpdate *upd_host, int num_sms, bool force, TeamPlan *pl) {
    const int64_t C = d.C;
    if (d.G > kMaxG) return false;
    const int max_cta = 2 * num_sms;
    if (!force && C < (int64_t)max_cta * 4) return false;
    // CTAs: two per SM when there are enough chains, teams of 4
    int n_cta, ts;
    if (C >= (int64_t)max_cta * 4) { n_cta = max_cta - max_cta % 8; ts = 4; }
    else if (C >= 32) { n_cta = (int)std::min<int64_t>(max_cta, C / 4); n_cta -= n_cta % 8; ts = 4; }
    else { n_cta = C >= 2 ? 2 : 1; ts = 1; }
    if (const char *e = getenv("EXTMCMC_TEAM_TS")) {      // diagnostics: 1, 2 or 4 CTAs per team
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4) && n_cta % (2 * v) == 0) ts = v;
    }
    pl->n_cta = n_cta;
    pl->ts = ts;
    pl->phases = n_cta >= 2 ? 2 : 1;
    if (const char *e = getenv("EXTMCMC_TEAM_PHASES")) if (atoi(e) == 1) pl->phases = 1;
    pl->n_team = n_cta / ts;
    pl->base = C / pl->n_team;
    pl->rem = C % pl->n_team;
    const int64_t team_max = pl->base + (pl->rem ? 1 : 0);
    int best_cap = 1 << 30, best_r = 0, best_cg = 0;
    for (int cg = 1; cg <= 32; cg <<= 1)
        for (int r = 4; r <= 8; ++r) {
            const int cap = cg * r;
            if (cap < team_max) continue;
            if (cap < best_cap || (cap == best_cap && r > best_r)) { best_cap = cap; best_r = r; best_cg = cg; }
        }
    if (!best_r) return false;
    pl->R = best_r;
    pl->cg = best_cg;
    const int own_max = (int)((team_max + ts - 1) / ts);
    if (!make_view_layout(d, upd_host, own_max, &pl->vl)) return false;
    const bool stage = d.p > 4 && d.p <= kCoopP;
    pl->stage_doubles = stage ? d.p * own_max : 0;
    auto bytes_for = [&](int stages) {
        size_t b = (size_t)stages * kTile * 8 + kMaxStages * 8 + kMaxStages * 4 + 16 + kMaxG * 8 + kMaxG * 4 + (kMaxG + 1) * 4 + 124;
        b = (b + 127) & ~(size_t)127;
        b += (sizeof(StepCtx) + 127) / 128 * 128;
        b += (size_t)d.G * 16 * (kBT / 8) * 8;
        b += (size_t)3 * pl->stage_doubles * 8;
        b = (b + 15) & ~(size_t)15;
        return b + pl->vl.bytes;
    };
    // the ring takes what two CTAs per SM leave: a deep ring absorbs the skew between the warps of a
    // CTA (the slowest warp refills a stage) and the L2 round trip of the refill
    int stages = kMaxStages;
    if (const char *e = getenv("EXTMCMC_TEAM_STAGES")) stages = std::max(2, std::min(kMaxStages, atoi(e)));
    while (stages > 3 && bytes_for(stages) > 110 * 1024) --stages;
    pl->stages = stages;
    const size_t b = bytes_for(stages);
    pl->smem_bytes = b;
    pl->cooperative = ts > 1 || pl->phases > 1;
    if (b > 110 * 1024) return false;     // two CTAs per SM
    return team_fits(pl->R, b, pl->n_cta, num_sms);
}

size_t team_sync_words(const TeamPlan &pl) { return (size_t)pl.n_team + (size_t)pl.n_cta; }

template <int R>
static cudaError_t launch_team_r(const TeamPlan &pl, TeamArgs &a, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(team_block_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    if (pl.cooperative) {
        // the members of a team wait for each other: all CTAs must be co-resident
        void *args[] = {&a};
        return cudaLaunchCooperativeKernel((const void *)team_block_kernel<R>, dim3(pl.n_cta), dim3(kBT), args, pl.smem_bytes, st);
    }
    team_block_kernel<R><<<pl.n_cta, kBT, pl.smem_bytes, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_team_block(const TeamPlan &pl, TeamArgs a, cudaStream_t st) {
    a.ts = pl.ts; a.phases = pl.phases; a.n_team = pl.n_team; a.cg = pl.cg;
    a.stage_doubles = pl.stage_doubles; a.base = pl.base; a.rem = pl.rem; a.vl = pl.vl; a.stages = pl.stages;
    switch (pl.R) {
    case 4: return launch_team_r<4>(pl, a, st);
    case 5: return launch_team_r<5>(pl, a, st);
    case 6: return launch_team_r<6>(pl, a, st);
    case 7: return launch_team_r<7>(pl, a, st);
    default: return launch_team_r<8>(pl, a, st);
    }
}

static size_t obs_block_smem(int cb, size_t view_bytes) {
    size_t b = (size_t)kObStages * kObTile * 8 + kObStages * 8 + (kObStages + (kObStages & 1)) * 4 +
               (size_t)kBW * cb * 8 + (size_t)kBT * 8;
    return b + view_bytes + 16;
}

template <int CB>
static cudaError_t obs_block_grid(int num_sms, int64_t n_obs, size_t smem, int *grid) {
    cudaError_t e = cudaFuncSetAttribute(obs_block_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, obs_block_kernel<CB>, kBT, smem);
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

cudaError_t plan_obs_block(const DevState &d, const DevUpdate *upd_host, int num_sms, int64_t n_obs, ObsBlockPlan *pl) {
    int cb = 1;
    while (cb < d.C) cb <<= 1;
    if (cb > 32) return cudaErrorInvalidValue;
    pl->cb = cb;
    if (!make_view_layout(d, upd_host, (int)d.C, &pl->vl)) return cudaErrorInvalidValue;
    pl->smem_bytes = obs_block_smem(cb, pl->vl.bytes);
    switch (cb) {
    case 1: return obs_block_grid<1>(num_sms, n_obs, pl->smem_bytes, &pl->grid);
    case 2: return obs_block_grid<2>(num_sms, n_obs, pl->smem_bytes, &pl->grid);
    case 4: return obs_block_grid<4>(num_sms, n_obs, pl->smem_bytes, &pl->grid);
    case 8: return obs_block_grid<8>(num_sms, n_obs, pl->smem_bytes, &pl->grid);
    case 16: return obs_block_grid<16>(num_sms, n_obs, pl->smem_bytes, &pl->grid);
    default: return obs_block_grid<32>(num_sms, n_obs, pl->smem_bytes, &pl->grid);
    }
}

template <int CB>
static cudaError_t launch_obs_block_cb(const ObsBlockPlan &pl, ObsBlockArgs &a, cudaStream_t st) {
    void *args[] = {&a};
    // cooperative: all CTAs must be co-resident (they wait for the decider every step)
    return cudaLaunchCooperativeKernel((const void *)obs_block_kernel<CB>, dim3(pl.grid), dim3(kBT), args, pl.smem_bytes, st);
}

cudaError_t launch_obs_block(const ObsBlockPlan &pl, ObsBlockArgs a, cudaStream_t st) {
    a.vl = pl.vl;
    switch (pl.cb) {
    case 1: return launch_obs_block_cb<1>(pl, a, st);
    case 2: return launch_obs_block_cb<2>(pl, a, st);
    case 4: return launch_obs_block_cb<4>(pl, a, st);
    case 8: return launch_obs_block_cb<8>(pl, a, st);
    case 16: return launch_obs_block_cb<16>(pl, a, st);
    default: return launch_obs_block_cb<32>(pl, a, st);
    }
}

}  // namespace extmcmc
