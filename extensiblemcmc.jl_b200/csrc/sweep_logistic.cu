// K4 -- Bayesian logistic regression (BASELINE cfg 3): log-likelihood AND its gradient for all
// chains in one pass over the design matrix, on the FP64 tensor-core path.
//
//   z_ic = x_i . theta_c,   ll_c = sum_i [ y_i z_ic - log(1 + e^{z_ic}) ],
//   grad_c = sum_i (y_i - sigmoid(z_ic)) x_i.
//
// No reference equivalent exists (the reference ships only GsnTargetLaw and an empty MALAUpdate
// stub, src/updates.jl:216-218); the gradient feeds the hook compute_gradients_and_momenta!
// (src/updates.jl:129-133).
//
// Both contractions are GEMM-shaped and run as FP64 DMMA (mma.sync.m8n8k4.f64 -- tcgen05 has no
// FP64 kind, so this legacy path IS Blackwell's FP64 tensor pipe):
//   phase 1  Z[16 obs x 64 chains]  = X_tile[16 x D] . Theta_blk^T[D x 64]
//   epilogue r = y - sigmoid(z), ll += y z - softplus(z)   (registers -> shared R tile)
//   phase 2  G[64 chains x D]      += R^T[64 x 16] . X_tile[16 x D]      (accumulators in registers)
// A CTA owns 64 chains and a contiguous range of 16-observation tiles; the X tile is TMA
// bulk-copied (UBLKCP, one 1-D copy per row into a padded, bank-conflict-free layout) into a
// 2-stage ring and is used by BOTH phases, so X is read from L2/HBM once per (chain block,
// sweep).  Theta_blk stays resident in shared memory for the whole kernel.  Partial ll / G per
// (segment, chain) are reduced in fixed order by logistic_finalize_kernel: deterministic.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>
#include "sweep.h"
#include "tma.cuh"

namespace extmcmc {

namespace {
constexpr int kBM = 16;    // observations per tile (2 x 16 x (D+4) + 64 x (D+4) doubles fit 227 KB at D = 256)
constexpr int kBN = 64;    // chains per CTA
constexpr int kNT = 512;   // 16 warps = 8 chain blocks x 2 halves (see the kernel comment)
constexpr int kPad = 4;    // row padding (doubles) of the X tiles: 128-bit loads indexed (row tq, col 2 gq)
                           // and (row rho(gq), col 2 tq) are both bank-conflict free with stride D + 4
constexpr int kPadT = 8;   // row padding of the Theta tile: 128-bit loads indexed (row gq, col 2 tq)
constexpr int kRPad = 68;  // row stride of the R tile

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// softplus(z) = log(1 + e^z) and sigmoid(z) in one go, ~31 FP64 instructions + 3 table loads (the
// libm route exp + log1p + divide costs ~8x more and made the epilogue, not the tensor pipe, the
// critical path of a warp; the FP64 pipe shares its datapath with DMMA, so every instruction saved
// here is MMA time).
//   e = exp(-|z|) = 2^m * 2^(j/32) * exp(r):  k = rint(-|z| 32/ln2) by the magic-number add,
//       j = k mod 32 (table E2), m = k div 32 (into the exponent field), |r| <= ln2/64, degree 6.
//   1 + e in [1, 2]: interval idx of width 1/128, c = 10-bit reciprocal of its midpoint (table RC,
//       RC[0] = 1), rr = (1 + e) c - 1 = fma(e, c, c - 1) exactly formed, |rr| <= 2^-7;
//       log(1 + e) = -log c (table LG) + log1p(rr), degree 7;
//       1/(1 + e) = c / (1 + rr) = c (1 - rr)(1 + rr^2)(1 + rr^4)   (error rr^8 <= 2^-56).
// Max relative error vs long-double libm over z in [-700, 700] (4e7 samples incl. the tails, host
// replica of this code): softplus 5.0e-16, sigmoid 6.5e-16.
constexpr int kTabE2 = 0, kTabRC = 32, kTabLG = 160, kTabN = 288;
__device__ double g_sp_tab[kTabN];

__device__ __forceinline__ void softplus_sigmoid(double z, double &sp, double &sg, const double *tab) {
    const double x = fmax(-fabs(z), -700.0);
    const double kMagic = 6755399441055744.0;   // 1.5 * 2^52: the low word of t is rint(x * 32/ln2)
    const double t = fma(x, 46.166241308446828384, kMagic);
    const int ki = __double2loint(t);
    const double kf = t - kMagic;
    double r = fma(-kf, 6.93147180369123816490e-01 / 32.0, x);
    r = fma(-kf, 1.90821492927058770002e-10 / 32.0, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double ep = tab[kTabE2 + (ki & 31)] * p;   // in [1, 2)
    const double e = __hiloint2double(__double2hiint(ep) + ((ki >> 5) << 20), __double2loint(ep));
    const int hi = __double2hiint(1.0 + e);
    const int idx = hi >= 0x40000000 ? 127 : (hi >> 13) & 127;
    const double c = tab[kTabRC + idx];
    const double rr = fma(e, c, c - 1.0), r2 = rr * rr;
    double q = 1.0 / 7.0;
    q = fma(q, rr, -1.0 / 6.0);
    q = fma(q, rr, 1.0 / 5.0);
    q = fma(q, rr, -1.0 / 4.0);
    q = fma(q, rr, 1.0 / 3.0);
    q = fma(q, rr, -0.5);
    sp = fmax(z, 0.0) + (tab[kTabLG + idx] + fma(q, r2, rr));
    const double a1 = 1.0 - rr, b1 = fma(a1, r2, a1);
    const double inv = c * fma(b1, r2 * r2, b1);   // 1 / (1 + e)
    sg = z >= 0.0 ? inv : e * inv;
}

template <int D>
constexpr size_t logistic_smem() {
    return (size_t)kBN * (D + kPadT) * 8 + (size_t)2 * kBM * (D + kPad) * 8 + (size_t)2 * kBM * kRPad * 8 +
           (size_t)(kNT / 32) * 64 * 8 + 2 * kBM * 8 + kTabN * 8 + 4 * 8;
}

// Work split.  The (chain block, observation tile) pairs are flattened block-major into
// U = blocks * T units and CTA c takes units [c U / n_cta, (c + 1) U / n_cta): every SM gets the same
// number of tiles whatever the number of chain blocks (cfg 3: 16 blocks on 148 SMs; a (segments x
// blocks) grid would leave 4 SMs idle), at the price of at most one extra Theta load for a CTA
// whose range straddles two blocks.  A CTA's part of one block is a "piece"; piece j of a block
// writes partial slot j, and the finalize kernel adds the slots of a block in order.
__host__ __device__ __forceinline__ int64_t unit_lo(int64_t c, int64_t U, int n_cta) { return c * U / n_cta; }
__host__ __device__ __forceinline__ int cta_of_unit(int64_t x, int64_t U, int n_cta) {
    int64_t c = x * n_cta / U;
    while (c + 1 < n_cta && unit_lo(c + 1, U, n_cta) <= x) ++c;
    while (c > 0 && unit_lo(c, U, n_cta) > x) --c;
    return (int)c;
}

__device__ __forceinline__ void pair_sync(int id) {  // named barrier for the two warps of a chain block
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
}  // namespace

// 16 warps: warp (w8 = w % 8, h = w / 8).  The two warps of a pair share chain block w8 (8 chains):
//   phase 1  PM = 2 (default): each contracts its own 8 observations (rows 8h .. 8h+7) with the 8 chains
//            over the whole feature range -- no exchange, no barrier;
//            PM = 0: each takes half of the feature range for the whole 16 x 8 Z block and the halves are
//            exchanged through shared memory (pair barrier);
//   epilogue each handles its own 8 observations, writes its R rows, pair barrier;
//   phase 2  each accumulates G for its half of the feature columns over all 16 observations.
// Four warps per SM sub-partition instead of two: while one pair sits in its epilogue or at a
// barrier, the others keep the MMA pipe busy.  G costs 64 registers per thread (D = 256), which is
// what makes 512 threads x 128 registers possible.
template <int D, int PM>
__global__ void __launch_bounds__(kNT, 1) sweep_logistic_kernel(LogisticArgs a) {
    constexpr int LD = D + kPad, LDT = D + kPadT;
    constexpr int NW = kNT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *Th = reinterpret_cast<double *>(smem_raw);   // [kBN][LDT]  Theta block, row = chain
    double *Xs = Th + kBN * LDT;                          // [2][kBM][LD] observation tiles
    double *Rs = Xs + 2 * kBM * LD;                       // [2][kBM][kRPad] residuals y - sigmoid(z), by tile parity (PM >= 2)
    double *Zp = Rs + 2 * kBM * kRPad;                    // [NW][64]     partial Z handed to the partner
    double *ys = Zp + NW * 64;                            // [2][kBM]
    double *tab = ys + 2 * kBM;                           // [kTabN]      softplus/sigmoid tables
    uint64_t *bar = reinterpret_cast<uint64_t *>(tab + kTabN);   // [2] full (TMA landed)
    unsigned int *done = reinterpret_cast<unsigned int *>(bar + 2);  // [2] warps done with a stage

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int w8 = w & 7, h = w >> 3;   // partners w and w + 8 sit on the same SM sub-partition (w % 4): measured
                                        // 9 % faster than partners on neighbouring sub-partitions (2 w8, 2 w8 + 1)
    const int gq = lane >> 2, tq = lane & 3;  // mma fragment coordinates: group id, thread in group
    const int64_t C = a.C;
    const int64_t T = (a.n_obs + kBM - 1) / kBM;            // tiles per chain block
    const int64_t U = (int64_t)((C + kBN - 1) / kBN) * T;   // flattened (block, tile) units
    const int64_t u_end = unit_lo((int64_t)blockIdx.x + 1, U, a.n_cta);

    if (tid < kTabN) tab[tid] = g_sp_tab[tid];
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        done[0] = done[1] = 0u;
        mbar_fence_init();
    }
    constexpr int NPH = D / 32;       // column pairs (16 features each) of this warp's half
    // fragment row gq stands for observation rho(gq) of the 8-block (bits 0 and 1 swapped): the two
    // rows a quarter-warp touches are then 2 apart, which the D + 4 stride separates (128-bit loads)
    const int rq = (gq & 4) | ((gq & 1) << 1) | ((gq >> 1) & 1);
    int tt = 0;   // tiles consumed so far by this CTA: the ring's stage and phase run on across pieces

  for (int64_t u = unit_lo(blockIdx.x, U, a.n_cta); u < u_end;) {
    // ---- one piece: chain block blk, tiles [t0, t0 + n_tiles) ----------------------------------
    const int64_t blk = u / T, t0 = u - blk * T;
    const int n_tiles = (int)((T - t0 < u_end - u) ? (T - t0) : (u_end - u));
    const int seg = (int)blockIdx.x - cta_of_unit(blk * T, U, a.n_cta);   // partial slot of this piece
    const int64_t cbase = blk * kBN;
    u += n_tiles;

    __syncthreads();   // the previous piece is completely done with Theta, the stages and Zp
    // Theta block -> shared (coalesced over chains), zero beyond d or beyond C
    for (int idx = tid; idx < kBN * D; idx += kNT) {
        const int k = idx / kBN, c = idx % kBN;
        const int64_t gc = cbase + c;
        Th[c * LDT + k] = (k < a.d && gc < C) ? a.theta[(int64_t)k * C + gc] : 0.0;
    }
    __syncthreads();

    // One bulk copy per observation row (+ one for the y slice) into a 2-stage ring.  There is no
    // CTA-wide barrier in the main loop: each warp counts itself out of a stage when it is done
    // with it, and the warp that completes the count (the last one) refills the stage.
    auto issue = [&](int t) {
        const int st = (tt + t) & 1;
        const int64_t row0 = (t0 + t) * kBM;
        if (lane == 0) mbar_expect_tx(&bar[st], (uint32_t)(kBM * D * 8 + kBM * 8));
        __syncwarp();
        if (lane < kBM)
            bulk_g2s(Xs + (st * kBM + lane) * LD, a.X + (row0 + lane) * D, (uint32_t)(D * 8), &bar[st]);
        if (lane == 0) bulk_g2s(ys + st * kBM, a.y + row0, (uint32_t)(kBM * 8), &bar[st]);
    };
    if (w == 0)
        for (int t = 0; t < 2 && t < n_tiles; ++t) issue(t);

    double G[2 * NPH][2];
#pragma unroll
    for (int nb = 0; nb < 2 * NPH; ++nb) G[nb][0] = G[nb][1] = 0.0;
    double ll0 = 0.0, ll1 = 0.0;  // chains cbase + 8 w8 + 2 tq + {0, 1}, observation rows 8h + rho(gq)

    for (int t = 0; t < n_tiles; ++t) {
        const int st = (tt + t) & 1;
        mbar_wait(&bar[st], (uint32_t)((tt + t) >> 1) & 1u);
        const double *X = Xs + st * kBM * LD;
        const int64_t row0 = (t0 + t) * kBM;
        const int tile_no = tt + t;   // tiles this CTA has consumed before this one
        // PM >= 2 has no barrier between phase 2 of a tile and the epilogue of the next one: the R tile is
        // double-buffered by tile parity (a buffer is rewritten two tiles later, when the partner has passed
        // the hand-over of the tile in between, i.e. finished reading it)
        double *R = Rs + (PM >= 2 ? (tile_no & 1) * kBM * kRPad : 0);

        // ---- phase 1: half of the k range for Z[16 x 8(w8)] ---------------------------------------
        // The k index of the MMA fragments is a relabelling of the real feature index: fragment slot
        // tq of MMA e (e = 0, 1) of pair s stands for feature 8s + 2 tq + e, so that one LDS.128 per
        // operand feeds two MMAs, and the two MMAs go to independent accumulators.
        constexpr int MB = kBM / 8, KH = D / 2;
        double z[2];
      if (PM >= 2) {
        // ---- phase 1 (default): my own 8 observations over the FULL feature range -- no exchange with the
        // partner, one pair barrier less per tile, at 16 more operand loads per warp and tile; four
        // independent accumulator pairs, added in a fixed order ----------
        double zc[4][2];
#pragma unroll
        for (int m = 0; m < 4; ++m) zc[m][0] = zc[m][1] = 0.0;
        const double *Bq = Th + (8 * w8 + gq) * LDT + 2 * tq;
        const double *Aq = X + (8 * h + rq) * LD + 2 * tq;
        double2 b_cur = *reinterpret_cast<const double2 *>(Bq);
        double2 a_cur = *reinterpret_cast<const double2 *>(Aq);
#pragma unroll 8
        for (int k0 = 0; k0 < D; k0 += 16) {
            const double2 b_mid = *reinterpret_cast<const double2 *>(Bq + k0 + 8);
            const double2 a_mid = *reinterpret_cast<const double2 *>(Aq + k0 + 8);
            const int kn = k0 + 16 < D ? k0 + 16 : k0;
            const double2 b_nxt = *reinterpret_cast<const double2 *>(Bq + kn);
            const double2 a_nxt = *reinterpret_cast<const double2 *>(Aq + kn);
            dmma(zc[0][0], zc[0][1], a_cur.x, b_cur.x);
            dmma(zc[1][0], zc[1][1], a_cur.y, b_cur.y);
            dmma(zc[2][0], zc[2][1], a_mid.x, b_mid.x);
            dmma(zc[3][0], zc[3][1], a_mid.y, b_mid.y);
            b_cur = b_nxt;
            a_cur = a_nxt;
        }
        z[0] = (zc[0][0] + zc[1][0]) + (zc[2][0] + zc[3][0]);
        z[1] = (zc[0][1] + zc[1][1]) + (zc[2][1] + zc[3][1]);
      } else {
        double za[MB][2], zb[MB][2];
#pragma unroll
        for (int m = 0; m < MB; ++m) za[m][0] = za[m][1] = zb[m][0] = zb[m][1] = 0.0;
        const double *Bp = Th + (8 * w8 + gq) * LDT + h * KH + 2 * tq;
        const double *Ap = X + rq * LD + h * KH + 2 * tq;
        // fragments are fetched one k-pair ahead into their own registers (software pipeline)
        double2 b_cur = *reinterpret_cast<const double2 *>(Bp);
        double2 a_cur[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) a_cur[m] = *reinterpret_cast<const double2 *>(Ap + (8 * m) * LD);
#pragma unroll 8
        for (int k0 = 0; k0 < KH; k0 += 8) {
            const int kn = k0 + 8 < KH ? k0 + 8 : k0;
            const double2 b_nxt = *reinterpret_cast<const double2 *>(Bp + kn);
            double2 a_nxt[MB];
#pragma unroll
            for (int m = 0; m < MB; ++m) a_nxt[m] = *reinterpret_cast<const double2 *>(Ap + (8 * m) * LD + kn);
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                dmma(za[m][0], za[m][1], a_cur[m].x, b_cur.x);
                dmma(zb[m][0], zb[m][1], a_cur[m].y, b_cur.y);
            }
            b_cur = b_nxt;
#pragma unroll
            for (int m = 0; m < MB; ++m) a_cur[m] = a_nxt[m];
        }
        // hand the partner its observations' half-sums, take mine (selects, no dynamic indexing)
        const double s00 = za[0][0] + zb[0][0], s01 = za[0][1] + zb[0][1];
        const double s10 = za[1][0] + zb[1][0], s11 = za[1][1] + zb[1][1];
        {
            double2 v;   // the partner's observation block is 1 - h
            v.x = h ? s00 : s10;
            v.y = h ? s01 : s11;
            *reinterpret_cast<double2 *>(Zp + w * 64 + gq * 8 + 2 * tq) = v;
        }
        pair_sync(1 + w8);
        {
            const double2 v = *reinterpret_cast<const double2 *>(Zp + (w ^ 8) * 64 + gq * 8 + 2 * tq);
            z[0] = (h ? s10 : s00) + v.x;
            z[1] = (h ? s11 : s01) + v.y;
        }
      }
        // ---- epilogue: residuals and log-likelihood of my 8 observations ---------------------------
        {
            const int r = 8 * h + rq;
            const bool live = row0 + r < a.n_obs;   // zero-padded rows must not contribute
            const double yv = ys[st * kBM + r];
            double2 res;
            double sp, sg;
            softplus_sigmoid(z[0], sp, sg, tab);
            res.x = live ? yv - sg : 0.0;
            ll0 += live ? yv * z[0] - sp : 0.0;
            softplus_sigmoid(z[1], sp, sg, tab);
            res.y = live ? yv - sg : 0.0;
            ll1 += live ? yv * z[1] - sp : 0.0;
            *reinterpret_cast<double2 *>(R + r * kRPad + 8 * w8 + 2 * tq) = res;
        }
        pair_sync(1 + w8);   // both halves of R[:, 8 w8 ..] are in place

        // ---- phase 2: G[8(w8) x my half of D] += R[:, 8 w8..]^T . X[16 x my half] -------------------
        // Column relabelling: column gq of MMA e of pair P stands for feature 16 P + 2 gq + e
        // (one LDS.128 of X feeds two MMAs); undone when the partials are written.
#pragma unroll
        for (int i0 = 0; i0 < kBM; i0 += 4) {
            const double af = R[(i0 + tq) * kRPad + 8 * w8 + gq];
            const double *Xr = X + (i0 + tq) * LD + 16 * NPH * h + 2 * gq;
            constexpr int AH = NPH < 4 ? NPH : 4;   // X fragments fetched AH pairs ahead of their MMAs
            double2 xq[AH];
#pragma unroll
            for (int P = 0; P < AH; ++P) xq[P] = *reinterpret_cast<const double2 *>(Xr + 16 * P);
#pragma unroll
            for (int P = 0; P < NPH; ++P) {
                const double2 xv = xq[P % AH];
                if (P + AH < NPH) xq[P % AH] = *reinterpret_cast<const double2 *>(Xr + 16 * (P + AH));
                dmma(G[2 * P][0], G[2 * P][1], af, xv.x);
                dmma(G[2 * P + 1][0], G[2 * P + 1][1], af, xv.y);
            }
        }
        __syncwarp();
        unsigned int prev = 0u;
        if (lane == 0) prev = atomicAdd(&done[st], 1u);
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == NW - 1) {          // last warp out of this stage: refill it
            if (lane == 0) done[st] = 0u;
            __syncwarp();
            if (t + 2 < n_tiles) issue(t + 2);
        }
    }

    // ---- write the partials of this (segment, chain block) -----------------------------------
    // ll: my 8 observation rows are spread over gq (lane bits 2..4): xor-shuffle tree, then the
    // two halves of the pair are added in a fixed order (h = 0 first)
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        ll0 += __shfl_xor_sync(0xffffffffu, ll0, o);
        ll1 += __shfl_xor_sync(0xffffffffu, ll1, o);
    }
    __syncthreads();   // Zp is free: reuse it for the ll hand-over
    if (h == 1 && gq == 0) { Zp[w8 * 8 + 2 * tq] = ll0; Zp[w8 * 8 + 2 * tq + 1] = ll1; }
    __syncthreads();
    if (h == 0 && gq == 0) {
        const int64_t c = cbase + 8 * w8 + 2 * tq;
        if (c < C) a.ll_part[(int64_t)seg * C + c] = ll0 + Zp[w8 * 8 + 2 * tq];
        if (c + 1 < C) a.ll_part[(int64_t)seg * C + c + 1] = ll1 + Zp[w8 * 8 + 2 * tq + 1];
    }
    const int64_t c = cbase + 8 * w8 + gq;
    if (c < C) {
#pragma unroll
        for (int nb = 0; nb < 2 * NPH; ++nb) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                // accumulator (P = nb / 2, e = nb & 1), fragment column 2 tq + j -> feature index
                const int k = 16 * (NPH * h + (nb >> 1)) + 2 * (2 * tq + j) + (nb & 1);
                if (k < a.d) a.g_part[((int64_t)seg * a.d + k) * C + c] = G[nb][j];
            }
        }
    }
    tt += n_tiles;
  }   // pieces
}

// ll[c] = sum_s ll_part[s][c];  grad[k][c] = sum_s g_part[s][k][c]  (fixed order)
__global__ void __launch_bounds__(256)
logistic_finalize_kernel(LogisticArgs a, double *__restrict__ ll_out, double *__restrict__ grad_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t C = a.C;
    if (idx >= (int64_t)(a.d + 1) * C) return;
    const int64_t row = idx / C, c = idx % C;
    // pieces of this chain's block (see the work split above)
    const int64_t T = (a.n_obs + kBM - 1) / kBM, U = (int64_t)((C + kBN - 1) / kBN) * T, blk = c / kBN;
    const int np = cta_of_unit((blk + 1) * T - 1, U, a.n_cta) - cta_of_unit(blk * T, U, a.n_cta) + 1;
    double s = 0.0;
    if (row == a.d) {
        for (int i = 0; i < np; ++i) s += a.ll_part[(int64_t)i * C + c];
        ll_out[c] = s;
    } else {
        for (int i = 0; i < np; ++i) s += a.g_part[((int64_t)i * a.d + row) * C + c];
        if (grad_out) grad_out[row * C + c] = s;
    }
}

namespace {
template <int D>
cudaError_t prep() {
    cudaError_t e = cudaFuncSetAttribute(sweep_logistic_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)logistic_smem<D>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(sweep_logistic_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)logistic_smem<D>());
}
template <int D>
void launch(const SweepPlan &pl, const LogisticArgs &a0, cudaStream_t st) {
    // EXTMCMC_LOGI_PHASE1: 2 (default) = every warp contracts its own 8 observations over all features;
    // 0 = the two warps of a pair split the feature range and exchange half-sums (one more pair barrier
    // per tile): 34.19 -> 33.92 ms per launch at cfg 3 with 2
    static const int pair_mode = [] { const char *e = getenv("EXTMCMC_LOGI_PHASE1"); return e ? atoi(e) : 2; }();
    LogisticArgs a = a0;
    a.pair_mode = pair_mode;
    if (pair_mode == 2) sweep_logistic_kernel<D, 2><<<a.n_cta, kNT, logistic_smem<D>(), st>>>(a);
    else sweep_logistic_kernel<D, 0><<<a.n_cta, kNT, logistic_smem<D>(), st>>>(a);
}
}  // namespace

int logistic_padded_dim(int d) {
    for (int D : {32, 64, 128, 256})
        if (d <= D) return D;
    return 0;
}

cudaError_t sweep_logistic_init() {
    cudaError_t e;
    {
        // tables of softplus_sigmoid, rounded from extended precision
        double tabh[kTabN];
        for (int j = 0; j < 32; ++j) tabh[kTabE2 + j] = (double)exp2l((long double)j / 32.0L);
        for (int j = 0; j < 128; ++j) {
            double c = (double)(1.0L / (1.0L + ((long double)j + 0.5L) / 128.0L));
            uint64_t b;
            memcpy(&b, &c, 8);
            b = (b + (1ull << 41)) & ~((1ull << 42) - 1);   // 10 significant bits
            memcpy(&c, &b, 8);
            if (j == 0) c = 1.0;                              // rr = e exactly on the first interval
            tabh[kTabRC + j] = c;
            tabh[kTabLG + j] = (double)(-logl((long double)c));
        }
        if ((e = cudaMemcpyToSymbol(g_sp_tab, tabh, sizeof tabh)) != cudaSuccess) return e;
    }
    if ((e = prep<32>()) != cudaSuccess) return e;
    if ((e = prep<64>()) != cudaSuccess) return e;
    if ((e = prep<128>()) != cudaSuccess) return e;
    if ((e = prep<256>()) != cudaSuccess) return e;
    return cudaSuccess;
}

SweepPlan plan_sweep_logistic(int d, int64_t C, int64_t n_obs, int num_sms) {
    SweepPlan pl{};
    pl.D = logistic_padded_dim(d);
    pl.variant = SWEEP_VARIANT_CHAINS;
    pl.R = kBN;
    pl.groups = (int)((C + kBN - 1) / kBN);
    const int64_t n_tiles = (n_obs + kBM - 1) / kBM, U = (int64_t)pl.groups * n_tiles;
    // one persistent CTA per SM (the tiles take ~215 KB of shared memory), each with an equal share
    // of the flattened (chain block, tile) units; S = partial slots per chain block
    pl.n_cta = (int)(U < num_sms ? U : num_sms);
    pl.S = (pl.n_cta + pl.groups - 1) / pl.groups + 1;
    pl.launches = 2;  // sweep + finalize
    pl.name = "logistic_dmma";
    return pl;
}

void launch_sweep_logistic(const SweepPlan &pl, const LogisticArgs &a, double *ll_out, double *grad_out,
                           cudaStream_t st) {
    switch (pl.D) {
    case 32: launch<32>(pl, a, st); break;
    case 64: launch<64>(pl, a, st); break;
    case 128: launch<128>(pl, a, st); break;
    default: launch<256>(pl, a, st); break;
    }
    const int64_t n = (int64_t)(a.d + 1) * a.C;
    logistic_finalize_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(a, ll_out, grad_out);
}

}  // namespace extmcmc
