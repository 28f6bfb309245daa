// TEST INFRASTRUCTURE ONLY -- the oracle's own copy of the Philox4x32-10 draw stream (Salmon et al.,
// "Parallel random numbers: as easy as 1, 2, 3", SC'11; constants from Random123). The product has an
// independent copy in smore_b200/csrc/philox.cuh; tests check both against the Random123 known answers.
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <stdint.h>

#define ORACLE_SHUFFLE_STREAM 0x8000000000000000ull

static inline void oracle_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// word `pos` of stream `stream` under `seed`.
static inline uint32_t oracle_philox_word(uint64_t seed, uint64_t stream, uint64_t pos) {
    uint64_t blk = pos >> 2;
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    oracle_philox4x32_10(ctr, key, out);
    return out[pos & 3];
}

// Sequential reader with a one-block cache.
typedef struct {
    uint64_t seed, stream, pos, cached_blk;
    uint32_t buf[4];
    int valid;
} oracle_stream;

static inline void oracle_stream_init(oracle_stream* s, uint64_t seed, uint64_t stream, uint64_t pos) {
    s->seed = seed; s->stream = stream; s->pos = pos; s->cached_blk = 0; s->valid = 0;
}

static inline uint32_t oracle_stream_next(oracle_stream* s) {
    uint64_t blk = s->pos >> 2;
    if (!s->valid || blk != s->cached_blk) {
        uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)s->stream, (uint32_t)(s->stream >> 32)};
        uint32_t key[2] = {(uint32_t)s->seed, (uint32_t)(s->seed >> 32)};
        oracle_philox4x32_10(ctr, key, s->buf);
        s->cached_blk = blk; s->valid = 1;
    }
    return s->buf[(s->pos++) & 3];
}

#endif
