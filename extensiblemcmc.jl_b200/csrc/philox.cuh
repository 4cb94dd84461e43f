// Per-chain counter RNG: Philox4x32-10 (Salmon et al., SC'11).
//
// Replaces the reference's use of Random.GLOBAL_RNG (rand(rng, Uniform) in
// src/transition_kernels/random_walk.jl:70-71 and rand(Exponential(1.0)) in
// src/run.jl:278).  A counter RNG makes every draw a pure function of
// (seed, global chain id, mcmciter, pidx, draw index): no RNG state lives in
// HBM, chains can be sharded over GPUs without changing their streams, and a
// run is checkpointable by construction.
//
//   key = (seed lo32, seed hi32)
//   ctr = (chain lo32, chain hi32, mcmciter lo32, (pidx << 16) | block)
//   uniform j of a chain-step = lane (j & 1) of block (j >> 1):
//   u = (k + 0.5) * 2^-52, k = top 52 bits of the lane's 64-bit word -> u in (0, 1).
#pragma once
#include <cstdint>

namespace extmcmc {

struct Philox4 { uint32_t w[4]; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

__host__ __device__ __forceinline__ double u52_to_unit(uint32_t lo, uint32_t hi) {
    const uint64_t word = ((uint64_t)hi << 32) | lo;
    return ((double)(word >> 12) + 0.5) * 0x1.0p-52;
}

// Sequential reader of one (chain, mcmciter, pidx) substream.
struct ChainStepStream {
    uint32_t c0, c1, c2, c3hi, k0, k1;
    uint32_t j;       // index of the next uniform
    Philox4 blk;      // cached block (valid when (j & 1) == 1)
    __device__ __forceinline__ ChainStepStream(uint64_t seed, uint64_t chain, int64_t mcmciter,
                                               int32_t pidx, uint32_t j0 = 0)
        : c0((uint32_t)chain), c1((uint32_t)(chain >> 32)), c2((uint32_t)mcmciter),
          c3hi((uint32_t)pidx << 16), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), j(j0) {
        if (j & 1u) blk = philox4x32_10(c0, c1, c2, c3hi | (j >> 1), k0, k1);
    }
    __device__ __forceinline__ double next() {
        const uint32_t lane = j & 1u;
        if (lane == 0u) blk = philox4x32_10(c0, c1, c2, c3hi | (j >> 1), k0, k1);
        ++j;
        return lane ? u52_to_unit(blk.w[2], blk.w[3]) : u52_to_unit(blk.w[0], blk.w[1]);
    }
};

}  // namespace extmcmc
