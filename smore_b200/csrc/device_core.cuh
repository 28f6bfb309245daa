// Device-side building blocks of the sampled-SGD hot path (sm_100a).
//
//   DrawRing   per-warp Philox sub-stream staged in shared memory (256-word ring, refilled 128 words at a time
//              with one Philox block per lane), consumed strictly sequentially -> a warp is one reference "worker".
//   samplers   alias draws with packed {u32 threshold, u32 alias} entries: integer-only, bit-exact with the
//              reference's `random_gen(0,1) < prob` under the replayed stream (include/smore_b200.h).
//   RowT       one embedding row spread over the 32 lanes as 128-bit (or narrower) vectors: coalesced
//              ld.global.cg / st.global.cg, dot products by xor-shuffle butterflies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace smore {

constexpr int kMonitor = 10000;  // reference MONITOR (src/proNet.h:32)
constexpr int kSigmoidTable = 1000;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------------
// Graph view
// ---------------------------------------------------------------------------------------------------------------
struct GraphDev {
    int64_t V, E;
    const int64_t* row_off;    // V+1
    const int32_t* col;        // E
    const uint2* vertex_at;    // V   {thr, alias}
    const uint2* negative_at;  // V
    const uint2* ctx_at;       // E   C++ semantics: per-vertex sub-tables, alias already a vertex id
    const double* prefix;      // E   Go semantics: running sum of weights inside each vertex' slice
    const int32_t* field;      // V   (HOP-Rec)
    const int32_t* col_sorted; // E   node2vec: every adjacency slice sorted ascending (membership tests by binary search)
    const double* w;           // E   node2vec: raw edge weights in adjacency order
    int sem;                   // SMORE_SEM_*
    // Row sharding (multi-GPU, one process per GPU): vertex v is owned by rank v & (world-1), its local row is
    // v >> shard_shift. The rank that owns the POSITIVE CONTEXT of a sample computes it: edge_at is an alias table over
    // the CSR entries whose target this rank owns (probability = P(source) * P(target | source) of the unsharded
    // samplers, so the union over ranks is the global edge distribution), edge_src / edge_dst give the endpoints, and
    // negative_at covers the owned vertices only (n_neg entries, local indices): negatives come from the same shard as
    // the positive. world == 1: edge_at == nullptr, n_neg = V, shard_shift = shard_rank = 0.
    const uint2* edge_at;
    const int32_t* edge_src;
    const int32_t* edge_dst;
    uint32_t n_edge_local;
    uint32_t n_neg;
    int shard_shift, shard_rank;
    // negative_at index -> vertex id: (l << neg_shift) + neg_rank. Shard-local negatives: neg_shift/neg_rank = the shard's;
    // a table over all vertices (unsharded, or the sharded "global negatives" variant): 0 / 0.
    int neg_shift, neg_rank;
    // split samples (kernels.cuh update_pair_split): the table the second vertex of a sample is drawn from -- the global
    // vertex table (peer-access mode, ids = vertex ids) or the table of the resident sub-part (rotating shards, ids = rows)
    const uint2* vsrc_at;
    uint32_t n_vsrc;
    __device__ __forceinline__ uint32_t global_id(uint32_t local) const { return (local << neg_shift) + (uint32_t)neg_rank; }
};

// A (possibly row-sharded) embedding table: base[r] = rank r's shard (the local cudaMalloc, or a CUDA-IPC peer mapping
// reached over NVLink). Bases live in shared memory: a dynamically indexed kernel-parameter array would be demoted
// to local memory.
template <typename T>
struct TableView {
    T* const* base;
    int shift, mask, dim;
    __device__ __forceinline__ T* row(int id) const { return base[id & mask] + (size_t)(id >> shift) * dim; }
};
// The unsharded table: one base pointer, no indirection (keeps the single-GPU kernels free of the shard arithmetic).
template <typename T>
struct DirectView {
    T* base0;
    int dim;
    static constexpr int mask = 0;
    __device__ __forceinline__ T* row(int id) const { return base0 + (size_t)id * dim; }
};

// Bulk-exchange mode of a row-sharded table, vertex side: id >= 0 is a vertex owned by this rank (row id >> shift of the
// local shard), id < 0 names row (-2 - id) of the staging table the remote rows of the super-batch were gathered into.
// A non-negative id of another rank's vertex (a HOT row, see ExchDev::hot) goes through that rank's peer mapping.
template <typename T>
struct ExchView {
    T* const* base;  // shard bases per rank (shared memory); only this rank's is dereferenced unless hot rows exist
    T* wrk;
    int shift, mask, dim;
    __device__ __forceinline__ T* row(int id) const {
        return id >= 0 ? base[id & mask] + (size_t)(id >> shift) * dim : wrk + (size_t)(-2 - id) * dim;
    }
};
// ... context side: context rows of the samples a rank computes are always its own.
template <typename T>
struct OwnedView {
    T* local;
    int shift, dim;
    static constexpr int mask = 0;
    __device__ __forceinline__ T* row(int id) const { return local + (size_t)(id >> shift) * dim; }
};

// ---------------------------------------------------------------------------------------------------------------
// Per-warp draw ring
// ---------------------------------------------------------------------------------------------------------------
struct DrawRing {
    uint32_t* buf;  // 256 words of shared memory owned by this warp
    uint64_t pos;   // next unread word (warp-uniform)
    uint64_t base;  // words [base, base+256) are resident, base % 128 == 0
    uint64_t seed, stream;
    int lane;

    __device__ __forceinline__ void fill_half(uint64_t first) {
        U4 r = philox_block(seed, stream, (first >> 2) + (uint64_t)lane);
        uint32_t slot = ((uint32_t)first & 255u) + 4u * (uint32_t)lane;
        *reinterpret_cast<uint4*>(buf + slot) = make_uint4(r.x, r.y, r.z, r.w);
    }
    __device__ __forceinline__ void init(uint32_t* smem, uint64_t seed_, uint64_t stream_, uint64_t pos0, int lane_) {
        buf = smem;
        seed = seed_;
        stream = stream_;
        lane = lane_;
        pos = pos0;
        base = pos0 & ~127ull;
        fill_half(base);
        fill_half(base + 128);
        __syncwarp();
    }
    // After ensure() at least 128 words starting at pos are resident.
    __device__ __forceinline__ void ensure() {
        while (pos >= base + 128) {
            __syncwarp();
            fill_half(base + 256);
            base += 128;
            __syncwarp();
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t i) const { return buf[((uint32_t)pos + i) & 255u]; }
    __device__ __forceinline__ void advance(uint32_t n) { pos += n; }
};

__device__ __forceinline__ uint32_t index_draw(uint32_t k, uint32_t n) { return __umulhi(k, n); }

__device__ __forceinline__ uint32_t alias_pick(const uint2* __restrict__ at, uint32_t idx, uint32_t kp) {
    uint2 e = __ldg(at + idx);
    return kp < e.x ? idx : e.y;
}

// SourceSample: C++ draws p then index (src/proNet.cpp:649-650); Go index then p (alias.go:99-100). 2 words.
__device__ __forceinline__ uint32_t source_sample(const GraphDev& g, uint32_t w0, uint32_t w1) {
    uint32_t kp = g.sem == 0 ? w0 : w1;
    uint32_t ki = g.sem == 0 ? w1 : w0;
    return alias_pick(g.vertex_at, index_draw(ki, (uint32_t)g.V), kp);
}

// NegativeSample: index then p in both trees (src/proNet.cpp:625-626; alias.go:99-100). 2 words.
__device__ __forceinline__ uint32_t negative_sample(const GraphDev& g, uint32_t w0, uint32_t w1) {
    return g.global_id(alias_pick(g.negative_at, index_draw(w0, g.n_neg), w1));
}

// TargetSample(v). C++ (src/proNet.cpp:671-683): p, then index into the vertex' alias slice: 2 words, O(1).
// Go (pronet.go:257-284): one Float64, CDF scan with `r <= cum`: 1 word; here a binary search over the running sums
// (same fp64 additions in the same order, so the same boundary decisions). Returns -1 on a sink (no words used).
__device__ __forceinline__ int64_t target_sample(const GraphDev& g, int64_t v, uint32_t w0, uint32_t w1, int& used) {
    int64_t off = __ldg(g.row_off + v);
    int64_t br = __ldg(g.row_off + v + 1) - off;
    if (br == 0) {
        used = 0;
        return -1;
    }
    if (g.sem == 0) {
        used = 2;
        int64_t j = off + (int64_t)index_draw(w1, (uint32_t)br);
        uint2 e = __ldg(g.ctx_at + j);
        return w0 < e.x ? (int64_t)__ldg(g.col + j) : (int64_t)e.y;
    }
    used = 1;
    double total = __ldg(g.prefix + off + br - 1);
    double r = __dmul_rn((double)w0 * (1.0 / 4294967296.0), total);
    int64_t lo = 0, hi = br - 1;  // first e with r <= prefix[e]; falls back to the last neighbour
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (r <= __ldg(g.prefix + off + mid)) hi = mid;
        else lo = mid + 1;
    }
    return (int64_t)__ldg(g.col + off + lo);
}

// ---------------------------------------------------------------------------------------------------------------
// Sigmoid LUT (src/proNet.cpp:52-71): 1001 entries over [-8, 8], truncating index, hard 0/1 outside.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T fast_sigmoid(const T* __restrict__ lut, T x);

template <>
__device__ __forceinline__ double fast_sigmoid<double>(const double* __restrict__ lut, double x) {
    if (x < -8.0) return 0.0;
    if (x > 8.0) return 1.0;
    // (x + 8) * 1000 / 8 / 2 : the two divisions are exact (powers of two), so this is the reference's value
    return lut[(int)(__dmul_rn(__dadd_rn(x, 8.0), 1000.0) * 0.0625)];
}
template <>
__device__ __forceinline__ float fast_sigmoid<float>(const float* __restrict__ lut, float x) {
    if (x < -8.0f) return 0.0f;
    if (x > 8.0f) return 1.0f;
    return lut[(int)(__fmul_rn(__fadd_rn(x, 8.0f), 1000.0f) * 0.0625f)];
}

// ---------------------------------------------------------------------------------------------------------------
// Rows
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int VEC>
struct VecIO;
template <>
struct VecIO<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, float* x) {
        float4 v = __ldcg(reinterpret_cast<const float4*>(p));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    static __device__ __forceinline__ void ld_ca(const float* p, float* x) {
        float4 v = __ldca(reinterpret_cast<const float4*>(p));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* x) {
        __stcg(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
    }
};
template <>
struct VecIO<float, 2> {
    static __device__ __forceinline__ void ld(const float* p, float* x) {
        float2 v = __ldcg(reinterpret_cast<const float2*>(p));
        x[0] = v.x; x[1] = v.y;
    }
    static __device__ __forceinline__ void ld_ca(const float* p, float* x) {
        float2 v = __ldca(reinterpret_cast<const float2*>(p));
        x[0] = v.x; x[1] = v.y;
    }
    static __device__ __forceinline__ void st(float* p, const float* x) {
        __stcg(reinterpret_cast<float2*>(p), make_float2(x[0], x[1]));
    }
};
template <>
struct VecIO<float, 1> {
    static __device__ __forceinline__ void ld(const float* p, float* x) { x[0] = __ldcg(p); }
    static __device__ __forceinline__ void ld_ca(const float* p, float* x) { x[0] = __ldca(p); }
    static __device__ __forceinline__ void st(float* p, const float* x) { __stcg(p, x[0]); }
};
template <>
struct VecIO<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double* x) {
        double2 v = __ldcg(reinterpret_cast<const double2*>(p));
        x[0] = v.x; x[1] = v.y;
    }
    static __device__ __forceinline__ void ld_ca(const double* p, double* x) {
        double2 v = __ldca(reinterpret_cast<const double2*>(p));
        x[0] = v.x; x[1] = v.y;
    }
    static __device__ __forceinline__ void st(double* p, const double* x) {
        __stcg(reinterpret_cast<double2*>(p), make_double2(x[0], x[1]));
    }
};
template <>
struct VecIO<double, 1> {
    static __device__ __forceinline__ void ld(const double* p, double* x) { x[0] = __ldcg(p); }
    static __device__ __forceinline__ void ld_ca(const double* p, double* x) { x[0] = __ldca(p); }
    static __device__ __forceinline__ void st(double* p, const double* x) { __stcg(p, x[0]); }
};

// Row layout over a warp: element index of lane l, chunk c, slot j = (c*32 + l)*VEC + j. A given element of ANY row is
// always owned by the same lane, so in-place read-modify-write sequences on aliasing rows keep the reference's
// per-element program order without fences.
template <typename T_, int VEC_, int NCH_, bool MASKED_>
struct RowCfg {
    using T = T_;
    static constexpr int VEC = VEC_;
    static constexpr int NCH = NCH_;
    static constexpr bool MASKED = MASKED_;  // dim not a multiple of 32*VEC (VEC == 1 then)
    static constexpr int EPL = VEC_ * NCH_;
};

// Run-time switch of the L1-allocating gathers (Row::load_ca), one copy per translation unit, set by the host before a
// launch (host_common.h: set_l1_gather). See the hazard note at load_ca.
static __constant__ int c_l1_gather = 1;

template <class C>
struct Row {
    typename C::T x[C::EPL];

    __device__ __forceinline__ void load(const typename C::T* base, int lane, int dim) {
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {
            int idx = (c * 32 + lane) * C::VEC;
            if (!C::MASKED || idx < dim) VecIO<typename C::T, C::VEC>::ld(base + idx, x + c * C::VEC);
            else {
#pragma unroll
                for (int j = 0; j < C::VEC; ++j) x[c * C::VEC + j] = 0;
            }
        }
    }
    // L1-allocating gather (ld.global.ca): a row that many warps of the SM keep re-reading (popular items of a
    // Zipf-distributed catalogue) is served from L1 instead of queueing at its L2 slice behind everybody's writes.
    // Stale by at most the L1 residence time; a lane's own stores still invalidate the line (same-SM coherence).
    // HAZARD: a row that stays L1-resident on an SM is a private replica of that SM for as long as it stays -- the SM
    // keeps training its own copy and its full-row stores overwrite what the other SMs wrote (L1s are only invalidated
    // at kernel boundaries). Harmless for the few genuinely hot rows of a large table (Hogwild loses most concurrent
    // updates of such rows anyway), but a table small enough to live in the L1s turns into 148 diverging replicas:
    // measured on a 12 k-vertex graph in 4 shards (384 KB of context rows per shard), held-out AUC 0.52 instead of
    // 0.92. The host therefore enables the switch only for tables far larger than the aggregate L1 (set_l1_gather).
    __device__ __forceinline__ void load_ca(const typename C::T* base, int lane, int dim) {
        if (!c_l1_gather) {
            load(base, lane, dim);
            return;
        }
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {
            int idx = (c * 32 + lane) * C::VEC;
            if (!C::MASKED || idx < dim) VecIO<typename C::T, C::VEC>::ld_ca(base + idx, x + c * C::VEC);
            else {
#pragma unroll
                for (int j = 0; j < C::VEC; ++j) x[c * C::VEC + j] = 0;
            }
        }
    }
    __device__ __forceinline__ void store(typename C::T* base, int lane, int dim) const {
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {
            int idx = (c * 32 + lane) * C::VEC;
            if (!C::MASKED || idx < dim) VecIO<typename C::T, C::VEC>::st(base + idx, x + c * C::VEC);
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) x[e] = 0;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Arithmetic policy.
//   float  : throughput path. Products may contract into FMAs, dot products are lane-partial sums + xor butterfly.
//   double : PARITY path. Every operation is the reference's operation: separately rounded multiply and add (the
//            reference is built for baseline x86-64, which has no FMA) and dot products summed in index order
//            d = 0..D-1 (src/proNet.cpp:1317-1318), so the deterministic fp64 mode reproduces the compiled reference
//            bit for bit. It is slower (a serial chain of D additions per dot) and is not what bench.py measures.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct Ar;
template <>
struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ float madd(float c, float a, float b) { return c + a * b; }  // c + a*b
    static __device__ __forceinline__ float msub(float c, float a, float b) { return c - a * b; }  // c - a*b
};
template <>
struct Ar<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double madd(double c, double a, double b) { return __dadd_rn(c, __dmul_rn(a, b)); }
    static __device__ __forceinline__ double msub(double c, double a, double b) { return __dsub_rn(c, __dmul_rn(a, b)); }
};

// Asynchronous global -> shared staging of one row (each lane copies exactly the vectors it owns, so a lane only ever
// waits for its own copies: cp.async.wait_group, no warp barrier). Works on peer-mapped (NVLink) addresses, which the
// local L2 cannot prefetch.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class C>
__device__ __forceinline__ void row_stage_async(typename C::T* slot, const typename C::T* src, int lane, int dim) {
#pragma unroll
    for (int c = 0; c < C::NCH; ++c) {
        int idx = (c * 32 + lane) * C::VEC;
        if (!C::MASKED || idx < dim) cp_async<C::VEC * (int)sizeof(typename C::T)>(slot + idx, src + idx);
    }
}
template <class C>
__device__ __forceinline__ void row_from_smem(Row<C>& r, const typename C::T* slot, int lane, int dim) {
#pragma unroll
    for (int c = 0; c < C::NCH; ++c) {
        int idx = (c * 32 + lane) * C::VEC;
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) r.x[c * C::VEC + j] = (!C::MASKED || idx < dim) ? slot[idx + j] : (typename C::T)0;
    }
}

// Compiler barrier on a row's registers: forces the gathers that fill them to be ISSUED before this point. Without it
// ptxas, squeezed by the launch-bounds register cap, sinks each row load next to its first use, which turns one
// overlapped gather of n rows into n serialized DRAM round trips (measured on k_bpr_cpp: 4x slower).
template <class C>
__device__ __forceinline__ void pin(Row<C>& r) {
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) {
        if constexpr (sizeof(typename C::T) == 4) asm volatile("" : "+f"(r.x[e]));
        else asm volatile("" : "+d"(r.x[e]));
    }
}

// Fire-and-forget accumulation of a row delta into (possibly peer-mapped) memory: red.global.add, performed by the L2 of
// the GPU that owns the row (atomics travel over NVLink), so concurrent pushes from several GPUs never lose an update.
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <class C>
__device__ __forceinline__ void row_red_add(typename C::T* base, const Row<C>& delta, int lane, int dim) {
#pragma unroll
    for (int c = 0; c < C::NCH; ++c) {
        int idx = (c * 32 + lane) * C::VEC;
        if (!C::MASKED || idx < dim) {
            if constexpr (sizeof(typename C::T) == 4 && C::VEC == 4) {
                red_add4(base + idx, delta.x[c * 4], delta.x[c * 4 + 1], delta.x[c * 4 + 2], delta.x[c * 4 + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < C::VEC; ++j) red_add(base + idx + j, delta.x[c * C::VEC + j]);
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// N dot products v . c[r] at once (r < n valid rows; the others return 0). Result is warp-uniform.
template <class C, int N>
__device__ __forceinline__ void dots(const Row<C>& v, const Row<C>* c, int n, typename C::T* f) {
    using T = typename C::T;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
            T s = 0;
            if (r < n) {
#pragma unroll
                for (int e = 0; e < C::EPL; ++e) s += v.x[e] * c[r].x[e];
            }
            f[r] = s;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < N; ++r) f[r] += __shfl_xor_sync(kFull, f[r], o);
        }
    } else {
        // index order: chunk, lane, slot (element = (chunk*32 + lane)*VEC + slot); every lane replays the same chain
        T prod[N][C::EPL];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            f[r] = 0;
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) prod[r][e] = r < n ? Ar<T>::mul(v.x[e], c[r].x[e]) : (T)0;
        }
#pragma unroll
        for (int ch = 0; ch < C::NCH; ++ch) {
            for (int l = 0; l < 32; ++l) {
#pragma unroll
                for (int j = 0; j < C::VEC; ++j) {
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        if (r < n) f[r] = Ar<T>::add(f[r], __shfl_sync(kFull, prod[r][ch * C::VEC + j], l));
                    }
                }
            }
        }
    }
}

template <class C>
__device__ __forceinline__ typename C::T dot(const Row<C>& a, const Row<C>& b) {
    typename C::T f[1];
    dots<C, 1>(a, &b, 1, f);
    return f[0];
}

// ---------------------------------------------------------------------------------------------------------------
// Learning-rate schedule. The reference refreshes alpha every MONITOR samples of a worker from a shared progress
// counter (src/model/LINE.cpp:177-187; line.go:133-142). With W concurrent warps each warp estimates global progress
// as count*W; tick t happens when count*W crosses t*MONITOR and sets
//     alpha_t = max(alpha*1e-4, alpha * (1 - (t - lag)*MONITOR / total))
// lag = 1 for the C++ LINE/BPR/WARP/HBPR loops (they read current_sample before bumping it), 0 for the C++ walk models
// and every Go loop. W = 1 reproduces the reference's single-thread schedule exactly.
// ---------------------------------------------------------------------------------------------------------------
struct Sched {
    double alpha0;
    double total;  // as the reference converts it: (double)total
    uint64_t W;
    int lag;
    double offset;  // samples of the (possibly longer) schedule already done before this call; 0 for a whole run
    // host-mapped {schedule units done, alpha bits}: warp 0 of the grid stores its estimate at every tick, smore_progress
    // reads it from any host thread while the train call blocks (the reference's progress line, LINE.cpp:179-187)
    unsigned long long* live;
};

struct WarpState {
    uint64_t pos;        // stream position
    uint64_t count;      // reference loop counter
    uint64_t next_tick;  // count at which the next LR refresh happens
    double alpha;        // current learning rate
    uint64_t pairs;      // pair updates done (stats)
    uint64_t tries;      // WARP: negatives scanned (stats)
};

__device__ __forceinline__ void sched_tick(WarpState& st, const Sched& s) {
    if (st.count >= st.next_tick) {
        uint64_t t = (st.count * s.W) / kMonitor;
        double a = s.alpha0 * (1.0 - (s.offset + (double)((t - (uint64_t)s.lag) * kMonitor)) / s.total);
        double amin = s.alpha0 * 0.0001;
        st.alpha = a < amin ? amin : a;
        st.next_tick = ((t + 1) * kMonitor + s.W - 1) / s.W;
        if (s.live && blockIdx.x == 0 && threadIdx.x == 0) {
            volatile unsigned long long* live = s.live;
            live[0] = (unsigned long long)s.offset + t * (unsigned long long)kMonitor;
            live[1] = (unsigned long long)__double_as_longlong(st.alpha);
        }
    }
}

}  // namespace smore
