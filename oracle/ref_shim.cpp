// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Link-time replacement for the reference's src/random.cpp (reference: src/random.cpp:5-13,
// the sole entropy source of the samplers; `nm proNet.o` shows `U random_gen`) plus an
// interposed libc rand() (reference uses unseeded rand() for init, src/model/LINE.cpp:83, and for
// the DeepWalk/Walklets Fisher-Yates shuffle, src/model/DeepWalk.cpp:124-131).
//
// Every call pops ONE 32-bit word k from a Philox4x32-10 stream (key = seed, counter =
// {block_lo, block_hi, stream_lo, stream_hi}, word n of a stream = lane n&3 of block n>>2):
//   * probability draw (max-min == 1):  min + k * 2^-32
//   * index draw       (max-min != 1):  min + floor(k*(max-min) / 2^32) as an exact integer double
// so the reference's `(long)random_gen(0,n)` reproduces umulhi32(k,n) exactly and
// `random_gen(0,1) < prob` is the exact integer test k < ceil(prob * 2^32)  (SURVEY.md §7 hard part 2).
// rand() pops from the separate SHUFFLE stream and returns k >> 1 (31 bits, RAND_MAX = 2^31-1).
#include <cstdint>
#include <cstdlib>
#include <omp.h>
#include "oracle_philox.h"

namespace {
// One sampler stream per OpenMP thread (stream id = base + omp thread number): -threads 1 == stream base.
uint64_t g_seed = 0;
uint64_t g_stream_base = 0;
thread_local oracle_stream t_stream;
thread_local uint64_t t_epoch = ~0ull;
uint64_t g_epoch = 0;
uint64_t g_shuffle_pos = 0;
}  // namespace

extern "C" void ref_shim_seed(uint64_t seed, uint64_t stream_base) {
    g_seed = seed;
    g_stream_base = stream_base;
    g_shuffle_pos = 0;
    ++g_epoch;  // lazily resets every thread's position
}

extern "C" uint64_t ref_shim_pos(void) {
    if (t_epoch != g_epoch) return 0;
    return t_stream.pos;
}

static inline uint32_t next_word() {
    if (t_epoch != g_epoch) {
        t_epoch = g_epoch;
        oracle_stream_init(&t_stream, g_seed, g_stream_base + (uint64_t)omp_get_thread_num(), 0);
    }
    return oracle_stream_next(&t_stream);
}

double random_gen(const int& min, const int& max) {
    uint32_t k = next_word();
    long span = (long)max - (long)min;
    if (span == 1) return (double)min + (double)k * (1.0 / 4294967296.0);
    uint64_t idx = ((uint64_t)k * (uint64_t)span) >> 32;
    return (double)min + (double)idx;
}

// Unused by the hot path but declared in src/random.h; kept so every reference object links.
double ran_uniform() { return rand() / ((double)RAND_MAX + 1); }
double ran_gaussian() { return 0.0; }
double ran_gaussian(double mean, double) { return mean; }

// Interposed libc rand(): the .so is linked with -Wl,-Bsymbolic-functions so the reference objects bind here.
extern "C" int rand(void) {
    uint32_t k = oracle_philox_word(g_seed, ORACLE_SHUFFLE_STREAM, g_shuffle_pos++);
    return (int)(k >> 1);
}
