// Cross-rank exchange of per-chain sums under observation sharding: self-validating 16-byte cells.
//
// A cell holds one double split over two 8-byte words, each word = 32 bits of the value + the
// 32-bit sequence tag of the exchange (xseq + 1).  The sender stores the cell straight into every
// rank's buffer over the NVLink peer mappings -- no fence, no separate flag: 8-byte stores are
// delivered whole, so a reader that sees the expected tag in BOTH words has the value, whatever
// order the words land in.  One one-way NVLink trip per exchange instead of data + fence + flag per
// peer (a release at system scope waits for an NVLink round trip; issued peer after peer by one
// thread it cost ~20 us per update step on 8 GPUs).  The cells of a parity are rewritten two
// exchanges later, after every rank has consumed them (a rank reaches exchange k + 2 only after all
// ranks have delivered k + 1, which each does after consuming k); xseq is never reset.
#pragma once
#include <cstdint>

namespace extmcmc {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void ll_store(unsigned long long *cell, double v, uint32_t tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = (b & 0xffffffff00000000ull) | (unsigned long long)tag;
    const unsigned long long w1 = (b << 32) | (unsigned long long)tag;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(w0), "l"(w1) : "memory");
}

// Spin until the cell carries `tag`; false when `timeout_ns` has passed since t0.
__device__ __forceinline__ bool ll_load(const unsigned long long *cell, uint32_t tag, double &v,
                                        unsigned long long t0, unsigned long long timeout_ns) {
    for (;;) {
        unsigned long long w0, w1;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(cell) : "memory");
        if ((uint32_t)w0 == tag && (uint32_t)w1 == tag) {
            v = __longlong_as_double((long long)((w0 & 0xffffffff00000000ull) | (w1 >> 32)));
            return true;
        }
        if (global_timer_ns() - t0 > timeout_ns) return false;
    }
}

// cell of (parity, source rank, chain) in a rank's buffer rx[2][world][C][2 words]
__device__ __forceinline__ int64_t ll_cell(int parity, int world, int rank, int64_t C, int64_t c) {
    return 2 * ((((int64_t)parity * world + rank) * C) + c);
}

}  // namespace extmcmc
