// Philox4x32-10 counter-based generator (Salmon et al., SC'11), host + device.
//
// Draw-stream definition shared by every sampler in this library (and restated independently by the test oracle):
//   key     = {seed_lo, seed_hi}
//   counter = {block_lo, block_hi, stream_lo, stream_hi}
//   word n of stream s = lane (n & 3) of block (n >> 2)
// A sampler consumes words strictly in order, one per reference random_gen()/rng call:
//   probability draw u = k * 2^-32   -> the compare `u < prob` is the exact integer test k < ceil(prob * 2^32)
//   index draw over n = umulhi32(k, n)
// Stream ids: [0, 2^62) sampler streams (one per warp), 2^62 + t table-init streams, 2^63 the host shuffle stream.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SMORE_HD __host__ __device__ __forceinline__
#else
#define SMORE_HD inline
#endif

namespace smore {

constexpr uint64_t kInitStreamBase = 1ull << 62;
constexpr uint64_t kShuffleStream = 1ull << 63;

struct U4 {
    uint32_t x, y, z, w;
};

SMORE_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

SMORE_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

SMORE_HD U4 philox_block(uint64_t seed, uint64_t stream, uint64_t block) {
    return philox4x32_10((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)stream, (uint32_t)(stream >> 32),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Host-side sequential reader (shuffle stream, table init on the host side of tests).
struct HostStream {
    uint64_t seed, stream, pos = 0;
    uint64_t cached = ~0ull;
    U4 buf{};
    HostStream(uint64_t s, uint64_t st) : seed(s), stream(st) {}
    uint32_t next() {
        uint64_t blk = pos >> 2;
        if (blk != cached) {
            buf = philox_block(seed, stream, blk);
            cached = blk;
        }
        uint32_t lane = (uint32_t)(pos++ & 3);
        return lane == 0 ? buf.x : lane == 1 ? buf.y : lane == 2 ? buf.z : buf.w;
    }
};

}  // namespace smore
