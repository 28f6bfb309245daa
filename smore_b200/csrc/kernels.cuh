// Hot-path kernels (sm_100a): warp-per-worker Hogwild SGD over embedding rows in HBM.
//
// One warp == one reference worker thread: it owns a Philox sub-stream, draws (source, target, negatives) from the
// alias tables, gathers the 2+K rows with 128-bit loads, does dot / sigmoid-LUT / axpy in registers and scatters
// the rows back. DETERMINISTIC mode is the same kernel with a single warp.
//
// Row aliasing: the reference updates rows in place and in order, so a sample whose rows repeat (negative == positive,
// LINE-1 self pair, BPR pos == neg ...) sees its own earlier writes. Each update therefore has a FAST path (all rows
// of the sample distinct: every gather is issued up front, maximum memory-level parallelism) and an ORDERED path
// (some rows repeat: rows are re-read from memory in the reference's order; an element is always owned by the same
// lane, so plain program order reproduces the reference's in-place semantics).
#pragma once
#include "device_core.cuh"

namespace smore {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlockThreads = kWarpsPerBlock * 32;
constexpr int kCtxChunk = 6;      // context rows gathered per batch (1 positive + 5 negatives = the default K)
constexpr int kMaxWalkLen = 256;  // walk_steps + 1 <= kMaxWalkLen
constexpr int kMaxWorld = 8;      // ranks of one NVSwitch box
#ifndef SMORE_LINE_PREFETCH
#define SMORE_LINE_PREFETCH 2
#endif
constexpr int kLinePrefetch = SMORE_LINE_PREFETCH;  // L2 prefetch distance (samples) inside a batch

// Bulk-exchange mode of the row-sharded store (batch_kernels.cuh, "exchange" section): the remote vertex rows a
// super-batch needs are requested from their owners, moved with an all-to-all (NCCL over NVLink) into a contiguous local
// staging table `wrk`, updated there, and returned to the owner, who adds the difference to its row. All peer traffic is
// bulk and sequential; the random accesses stay in local HBM.
struct ExchDev {
    int32_t* hkey;       // open-addressing hash of the remote vertex ids of this super-batch: global id, -1 = empty
    int32_t* hval;       // (owner << 28) | index into that owner's request list
    uint32_t hmask;      // table size - 1 (power of two >= 2 x samples of a super-batch)
    int32_t* req;        // [world][req_stride]: local row ids requested from owner o
    int64_t req_stride;
    int32_t* cnt;        // [kMaxWorld] requests per owner
    const int32_t* off;  // [kMaxWorld] first wrk row of owner o's region (exclusive prefix sum of cnt)
    void* wrk;           // staged vertex rows [sum cnt][dim]
    int* errors;         // samples whose source was missing from the hash (must stay 0; checked by the host)
    const uint32_t* hot; // bitmap over vertex ids (null = none): HOT vertices are sampled so often that per-super-batch
                         // copies on several ranks would each run a long update sequence whose sum overshoots; their rows
                         // stay single-copy and are reached through the peer mappings as in the peer-access mode
    __device__ __forceinline__ bool is_hot(int v) const { return hot && ((__ldg(hot + (v >> 5)) >> (v & 31)) & 1u); }
};

template <typename T>
struct TrainArgs {
    GraphDev g;
    T* Wv;  // local shard (== the whole table when world == 1)
    T* Wc;
    T* peer_v[kMaxWorld];  // rank r's shard of the vertex / context table (peer_x[rank] == Wx)
    T* peer_c[kMaxWorld];
    int world_shift, world_mask;
    T* replica_v;  // row-sharded + replica mode: full-size local read copy of the vertex table (global ids), else null
    int dim;
    int same_table;  // Wv == Wc
    const T* lut;    // 1001-entry sigmoid table in global memory (copied to shared)
    uint64_t seed, stream_base;
    Sched sched;
    WarpState* state;
    int n_warps;
    uint64_t jobs;  // loop trips per warp in this launch
    int jobs_rem;   // ... plus one more for warps w < jobs_rem (block trainer: no sample of a small block is lost to rounding)
    int K;
    int order;
    T lambda;
    // walk models
    const int32_t* keys;  // start vertices of this launch (already shuffled / identity)
    int64_t n_walks;      // walks in this launch; warp w takes walks w, w+W, ...
    int steps, w0, w1, walklets;
    ExchDev x;  // bulk-exchange mode only
    T xi, omega;  // Skew-OPT
    int eta;
    int n2v;   // node2vec: biased second-order walks (1/p, 1/q below)
    double n2v_pinv, n2v_qinv;
    int vred;  // row-sharded peer-access mode: vertex rows take their delta with red.global.add instead of a full-row store
    // CPR / TPR (Go tree): adjacency of the SECOND graph over its own vids and the third table (CPR: source-domain item rows,
    // read only; TPR: word rows, trained); a vertex of the first graph is looked up there by its vid (cpr.go:148-169, tpr.go:108)
    const int64_t* aux_off;
    const int32_t* aux_col;
    int64_t aux_V;
    T* aux_tab;
    T item_reg, margin, text_w;
};

// ---------------------------------------------------------------------------------------------------------------
// skip-gram pair update, C++ semantics: proNet::UpdatePair + Opt_SigmoidSGD (src/proNet.cpp:1784-1809, :1312-1330).
// my_id: lane 0 holds the positive context, lane 1+n negative n; nrows = K+1 <= 32.
// ---------------------------------------------------------------------------------------------------------------
// CTX_CA: gather the context rows with ld.global.ca (hub rows served from L1; used when sharding concentrates them)
// Atomic rows (kAtomicRows<C>: the fp32 throughput tables): every row takes its DELTA with red.global.add.v4.f32 instead of
// a store of the updated copy. Concurrent updates of a row are then never lost, only computed from slightly stale values --
// with plain stores a table of 24 k rows trained by 3 552 resident warps loses so many updates that held-out AUC drops from
// 0.92 to 0.81 (profiles/r2b_ab_sharded_quality.txt) -- and the kernel gets faster (803 vs 775 M updates/s on configs[1]:
// the updated copy no longer has to stay in registers until the store, and the write needs no ownership of the line).
// fp64 tables keep the store path: it reproduces the reference's in-place arithmetic bit for bit (DETERMINISTIC parity).
template <class C>
constexpr bool kAtomicRows = sizeof(typename C::T) == 4;

template <class C, class TV, class TC, bool CTX_CA = false, bool RED = kAtomicRows<C>>
__device__ __forceinline__ void update_pair_cpp(const TV& tv,
                                                const TC& tc, int dim, bool same_table,
                                                const typename C::T* lut, int v1, int my_id, int nrows,
                                                typename C::T alpha, int lane, const Row<C>* vpre = nullptr,
                                                typename C::T* vpush = nullptr, bool vred = false) {
    using T = typename C::T;
    using A = Ar<T>;
    bool active = lane < nrows;
    unsigned peers = __match_any_sync(kFull, active ? my_id : (-1 - lane));
    bool dup = __any_sync(kFull, (active && __popc(peers) > 1) || (active && same_table && my_id == v1));
    T* pv = tv.row(v1);
    if (!dup) {
        Row<C> v, back;
        if (vpre) v = *vpre;  // staged ahead of time (remote shard): a few samples stale, Hogwild-tolerated
        else v.load(pv, lane, dim);
        back.zero();
        for (int base = 0; base < nrows; base += kCtxChunk) {
            Row<C> c[kCtxChunk];
            int ids[kCtxChunk];
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                ids[r] = __shfl_sync(kFull, my_id, (base + r) & 31);
                if (base + r < nrows) {
                    if constexpr (CTX_CA) c[r].load_ca(tc.row(ids[r]), lane, dim);
                    else c[r].load(tc.row(ids[r]), lane, dim);
                }
            }
            T f[kCtxChunk];
            dots<C, kCtxChunk>(v, c, nrows - base, f);
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                if (base + r < nrows) {
                    T label = (base + r == 0) ? (T)1 : (T)0;
                    T g = A::mul(A::sub(label, fast_sigmoid<T>(lut, f[r])), alpha);  // (label - f) * alpha
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        back.x[e] = A::madd(back.x[e], g, c[r].x[e]);  // loss_vertex += g * w_context
                        if constexpr (RED) c[r].x[e] = A::mul(g, v.x[e]);
                        else c[r].x[e] = A::madd(c[r].x[e], g, v.x[e]);  // w_context   += g * w_vertex (in place)
                    }
                    if constexpr (RED) row_red_add<C>(tc.row(ids[r]), c[r], lane, dim);
                    else c[r].store(tc.row(ids[r]), lane, dim);
                }
            }
        }
        // vred (row-sharded table, rows shared with other GPUs): the row takes its DELTA with red.global.add, performed by
        // the owner's L2 -- a full-row store of the (staged, possibly stale) copy would overwrite what the other ranks, or an
        // earlier sample of this very warp, added to the row in the meantime
        if (RED || vred) row_red_add<C>(pv, back, lane, dim);
        else {
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], back.x[e]);
            v.store(pv, lane, dim);
        }
        if (vpush) row_red_add<C>(vpush, back, lane, dim);  // replica mode: the delta also goes to the owner's row
    } else {
        Row<C> back;
        back.zero();
        for (int r = 0; r < nrows; ++r) {
            int cid = __shfl_sync(kFull, my_id, r);
            T* pc = tc.row(cid);
            Row<C> v, c;
            v.load(pv, lane, dim);
            c.load(pc, lane, dim);
            T f = dot(v, c);
            T label = (r == 0) ? (T)1 : (T)0;
            T g = A::mul(A::sub(label, fast_sigmoid<T>(lut, f)), alpha);
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                back.x[e] = A::madd(back.x[e], g, c.x[e]);
                if constexpr (RED) c.x[e] = A::mul(g, v.x[e]);
                else c.x[e] = A::madd(c.x[e], g, v.x[e]);
            }
            // (a lane's accesses to one address stay in program order, so the next round's load sees this update)
            if constexpr (RED) row_red_add<C>(pc, c, lane, dim);
            else c.store(pc, lane, dim);
        }
        if (RED || vred) row_red_add<C>(pv, back, lane, dim);
        else {
            Row<C> v;
            v.load(pv, lane, dim);
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], back.x[e]);
            v.store(pv, lane, dim);
        }
        if (vpush) row_red_add<C>(vpush, back, lane, dim);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// skip-gram pair update, Go semantics: ProNet.UpdatePair + sgdUpdate (pkg/pronet/optimizer.go:21-84) and, with
// skip_source, LINE.updateFirstOrder (internal/models/line/line.go:153-200). Negatives equal to the context (or the
// source) are skipped; the positive context row is written last.
// ---------------------------------------------------------------------------------------------------------------
template <class C, class TV, class TC, bool CTX_CA = false, bool RED = kAtomicRows<C>>
__device__ __forceinline__ void update_pair_go(const TV& tv,
                                                const TC& tc, int dim, bool same_table,
                                               bool skip_source, const typename C::T* lut, int v1, int my_id,
                                               int nrows, typename C::T alpha, int lane, const Row<C>* vpre = nullptr,
                                               typename C::T* vpush = nullptr, bool vred = false) {
    using T = typename C::T;
    using A = Ar<T>;
    int ctx = __shfl_sync(kFull, my_id, 0);
    bool active = lane < nrows;
    bool skipped = active && lane > 0 && (my_id == ctx || (skip_source && my_id == v1));
    unsigned skipmask = __ballot_sync(kFull, skipped);
    bool live = active && !skipped;
    unsigned peers = __match_any_sync(kFull, live ? my_id : (-1 - lane));
    bool dup = __any_sync(kFull, (live && __popc(peers) > 1) || (live && same_table && my_id == v1));
    T* pv = tv.row(v1);
    T* pp = tc.row(ctx);
    if (!dup) {
        Row<C> v, vgrad, pos, cgrad;
        if (vpre) v = *vpre;
        else v.load(pv, lane, dim);
        if constexpr (CTX_CA) pos.load_ca(pp, lane, dim);
        else pos.load(pp, lane, dim);
        {
            T g = A::mul(alpha, A::sub((T)1, fast_sigmoid<T>(lut, dot(v, pos))));  // alpha * (label - pred)
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                vgrad.x[e] = A::mul(g, pos.x[e]);
                cgrad.x[e] = A::mul(g, v.x[e]);
            }
        }
        for (int base = 1; base < nrows; base += kCtxChunk) {
            Row<C> c[kCtxChunk];
            int ids[kCtxChunk];
            bool ok[kCtxChunk];
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                ids[r] = __shfl_sync(kFull, my_id, (base + r) & 31);
                ok[r] = (base + r < nrows) && !((skipmask >> ((base + r) & 31)) & 1u);
                if (ok[r]) {
                    if constexpr (CTX_CA) c[r].load_ca(tc.row(ids[r]), lane, dim);
                    else c[r].load(tc.row(ids[r]), lane, dim);
                }
                else c[r].zero();
            }
            T f[kCtxChunk];
            dots<C, kCtxChunk>(v, c, nrows - base, f);
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                if (ok[r]) {
                    T g = A::mul(alpha, A::sub((T)0, fast_sigmoid<T>(lut, f[r])));
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        vgrad.x[e] = A::madd(vgrad.x[e], g, c[r].x[e]);
                        if constexpr (RED) c[r].x[e] = A::mul(g, v.x[e]);
                        else c[r].x[e] = A::add(c[r].x[e], A::mul(g, v.x[e]));  // wContext[neg] += negGrad
                    }
                    if constexpr (RED) row_red_add<C>(tc.row(ids[r]), c[r], lane, dim);
                    else c[r].store(tc.row(ids[r]), lane, dim);
                }
            }
        }
        if constexpr (RED) {
            row_red_add<C>(pv, vgrad, lane, dim);
            row_red_add<C>(pp, cgrad, lane, dim);
        } else {
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                v.x[e] = A::add(v.x[e], vgrad.x[e]);
                pos.x[e] = A::add(pos.x[e], cgrad.x[e]);
            }
            if (vred) row_red_add<C>(pv, vgrad, lane, dim);  // see update_pair_cpp
            else v.store(pv, lane, dim);
            pos.store(pp, lane, dim);
        }
        if (vpush) row_red_add<C>(vpush, vgrad, lane, dim);
    } else {
        Row<C> vgrad, cgrad;
        {
            Row<C> v, pos;
            v.load(pv, lane, dim);
            pos.load(pp, lane, dim);
            T g = A::mul(alpha, A::sub((T)1, fast_sigmoid<T>(lut, dot(v, pos))));
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                vgrad.x[e] = A::mul(g, pos.x[e]);
                cgrad.x[e] = A::mul(g, v.x[e]);
            }
        }
        for (int r = 1; r < nrows; ++r) {
            if ((skipmask >> r) & 1u) continue;
            int cid = __shfl_sync(kFull, my_id, r);
            T* pc = tc.row(cid);
            Row<C> v, c;
            v.load(pv, lane, dim);
            c.load(pc, lane, dim);
            T g = A::mul(alpha, A::sub((T)0, fast_sigmoid<T>(lut, dot(v, c))));
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                vgrad.x[e] = A::madd(vgrad.x[e], g, c.x[e]);
                if constexpr (RED) c.x[e] = A::mul(g, v.x[e]);
                else c.x[e] = A::add(c.x[e], A::mul(g, v.x[e]));
            }
            if constexpr (RED) row_red_add<C>(pc, c, lane, dim);
            else c.store(pc, lane, dim);
        }
        // per element: vertex row first, then context row (optimizer.go:54-57), through memory so that
        // a self pair on a shared table accumulates both.
        if (RED || vred) row_red_add<C>(pv, vgrad, lane, dim);
        else {
            Row<C> v;
            v.load(pv, lane, dim);
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], vgrad.x[e]);
            v.store(pv, lane, dim);
        }
        if (vpush) row_red_add<C>(vpush, vgrad, lane, dim);
        if constexpr (RED) row_red_add<C>(pp, cgrad, lane, dim);
        else {
            Row<C> pos;
            pos.load(pp, lane, dim);
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) pos.x[e] = A::add(pos.x[e], cgrad.x[e]);
            pos.store(pp, lane, dim);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Split sample (row-sharded modes). With shard-local negatives the K negatives of a sample come from the shard of its
// positive context; applied to the SAME vertex v1 they make the noise distribution depend on v1 (a vertex is only ever
// pushed away from contexts in the shards of its own neighbours -- with 8 shards and ~20 neighbours that already costs
// 3 points of held-out AUC, profiles/r2a_ab_sharded_quality.txt). The split sample restores the unsharded distribution
// exactly: the positive pair (v1, ctx) is one skip-gram step with label 1, and the K negatives are one step with label 0
// for a SECOND vertex v2 drawn from the source distribution independently of the edge. Per pair the arithmetic is
// Opt_SigmoidSGD (src/proNet.cpp:1312-1330 / optimizer.go:61-84); both trees reduce to the same thing here because the
// skip rules of the Go tree (negative == context) concern a pairing that no longer exists.
// my_id: lane 0 = positive context, lane 1+n = negative n; nrows = K + 1. One more row per sample: 2 (K+3) D s bytes.
// ---------------------------------------------------------------------------------------------------------------
template <class C, class TV, class TC>
__device__ __forceinline__ void update_pair_split(const TV& tv, const TC& tc, int dim, const typename C::T* lut, int v1,
                                                  int v2, int my_id, int nrows, typename C::T alpha, int lane,
                                                  const Row<C>* vpre, bool vred) {
    using T = typename C::T;
    using A = Ar<T>;
    constexpr bool RED = kAtomicRows<C>;
    const bool active = lane < nrows;
    const unsigned peers = __match_any_sync(kFull, active ? my_id : (-1 - lane));
    const bool dup = __any_sync(kFull, active && __popc(peers) > 1) || v1 == v2;
    T* pv1 = tv.row(v1);
    T* pv2 = tv.row(v2);
    const int ctx = __shfl_sync(kFull, my_id, 0);
    T* pp = tc.row(ctx);
    // one pair: row `c` of the context table against vertex row `v`; the vertex delta accumulates in `back`
    auto pair = [&](const Row<C>& v, Row<C>& c, T* pc, T label, T f, Row<C>& back) {
        const T g = A::mul(A::sub(label, fast_sigmoid<T>(lut, f)), alpha);
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            back.x[e] = A::madd(back.x[e], g, c.x[e]);
            if constexpr (RED) c.x[e] = A::mul(g, v.x[e]);
            else c.x[e] = A::madd(c.x[e], g, v.x[e]);
        }
        if constexpr (RED) row_red_add<C>(pc, c, lane, dim);
        else c.store(pc, lane, dim);
    };
    auto vertex = [&](Row<C>& v, T* pv, const Row<C>& back) {
        if (RED || vred) row_red_add<C>(pv, back, lane, dim);
        else {
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], back.x[e]);
            v.store(pv, lane, dim);
        }
    };
    if (!dup) {
        Row<C> a, b, pos, back;
        if (vpre) a = *vpre;
        else a.load(pv1, lane, dim);
        b.load(pv2, lane, dim);
        pos.load(pp, lane, dim);
        back.zero();
        pair(a, pos, pp, (T)1, dot(a, pos), back);  // positive pair: v1 <-> ctx
        vertex(a, pv1, back);
        back.zero();
        for (int base = 1; base < nrows; base += kCtxChunk) {
            Row<C> c[kCtxChunk];
            int ids[kCtxChunk];
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                ids[r] = __shfl_sync(kFull, my_id, (base + r) & 31);
                if (base + r < nrows) c[r].load(tc.row(ids[r]), lane, dim);
            }
            T f[kCtxChunk];
            dots<C, kCtxChunk>(b, c, nrows - base, f);
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r)
                if (base + r < nrows) pair(b, c[r], tc.row(ids[r]), (T)0, f[r], back);
        }
        if (nrows > 1) vertex(b, pv2, back);
    } else {
        // some rows coincide: one pair at a time through memory
        Row<C> back;
        {
            Row<C> a, pos;
            a.load(pv1, lane, dim);
            pos.load(pp, lane, dim);
            back.zero();
            pair(a, pos, pp, (T)1, dot(a, pos), back);
            vertex(a, pv1, back);
        }
        back.zero();
        for (int r = 1; r < nrows; ++r) {
            const int cid = __shfl_sync(kFull, my_id, r);
            T* pc = tc.row(cid);
            Row<C> b, c;
            b.load(pv2, lane, dim);
            c.load(pc, lane, dim);
            pair(b, c, pc, (T)0, dot(b, c), back);
        }
        if (nrows > 1) {
            Row<C> b;
            if (!(RED || vred)) b.load(pv2, lane, dim);
            vertex(b, pv2, back);
        }
    }
}

// Draw the K negatives of one pair from ring words [off, off+2K) into lanes 1..K; lane 0 keeps `ctx`.
__device__ __forceinline__ int draw_pair_ids(const GraphDev& g, const DrawRing& ring, uint32_t off, int ctx,
                                             int K, int lane) {
    int my = ctx;
    if (lane >= 1 && lane <= K) {
        uint32_t w0 = ring.peek(off + 2u * (uint32_t)(lane - 1));
        uint32_t w1 = ring.peek(off + 2u * (uint32_t)(lane - 1) + 1u);
        my = (int)negative_sample(g, w0, w1);
    }
    return my;
}

template <typename T>
__device__ __forceinline__ const T* stage_lut(const T* lut_global, T* lut_shared) {
    for (int i = threadIdx.x; i <= kSigmoidTable; i += blockDim.x) lut_shared[i] = lut_global[i];
    __syncthreads();
    return lut_shared;
}

// Peer shard bases -> shared memory (static indices only: a dynamically indexed parameter array goes to local memory),
// and the two table views built on them. Must be called by the whole CTA.
template <typename T>
__device__ __forceinline__ void stage_views(const TrainArgs<T>& a, TableView<T>& tv, TableView<T>& tc) {
    __shared__ T* s_base[2][kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
        if (threadIdx.x == r) {
            s_base[0][r] = a.peer_v[r];
            s_base[1][r] = a.peer_c[r];
        }
    __syncthreads();
    tv = TableView<T>{s_base[0], a.world_shift, a.world_mask, a.dim};
    tc = TableView<T>{s_base[1], a.world_shift, a.world_mask, a.dim};
}

// Shared memory layout of every training kernel: [kWarpsPerBlock][256] u32 draw rings, then the LUT, then
// (walk kernels) per-warp walk buffers.
extern __shared__ __align__(16) unsigned char smem_raw[];

// ---------------------------------------------------------------------------------------------------------------
// On-device walks: RandomWalk (src/proNet.cpp:704-724 / pronet.go:292-307). Lane 0 walks; ids go to shared memory.
// Returns the walk length (warp-uniform).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int random_walk(const GraphDev& g, DrawRing& ring, int64_t start, int steps, int32_t* walk,
                                           int lane) {
    int len = 1;
    if (lane == 0) walk[0] = (int32_t)start;
    int64_t next = start;
    const bool go = g.sem != 0;
    for (int s0 = 0; s0 < steps; s0 += 32) {  // <= 64 words per batch of 32 steps: one ensure() covers it
        ring.ensure();
        int nb = min(32, steps - s0);
        uint32_t used = 0;
        int stop = 0;
        if (lane == 0) {
            for (int s = 0; s < nb; ++s) {
                if (!go) {
                    int64_t br = __ldg(g.row_off + next + 1) - __ldg(g.row_off + next);
                    if (br == 0) {
                        if (next == start) { stop = 1; break; }
                        next = start;
                    }
                }
                int u;
                int64_t n = target_sample(g, next, ring.peek(used), ring.peek(used + 1), u);
                if (n < 0) { stop = 1; break; }  // Go: dead end stops the walk
                used += (uint32_t)u;
                walk[len++] = (int32_t)n;
                next = n;
            }
        }
        used = __shfl_sync(kFull, used, 0);
        stop = __shfl_sync(kFull, stop, 0);
        next = __shfl_sync(kFull, next, 0);
        len = __shfl_sync(kFull, len, 0);
        ring.advance(used);
        if (stop) break;
    }
    __syncwarp();
    return len;
}

// ---------------------------------------------------------------------------------------------------------------
// node2vec (Go tree only): biasedRandomWalk / biasedTargetSample / areNeighbors (internal/models/node2vec/node2vec.go:82-173).
// The first step is a plain TargetSample; every further step weighs the neighbours of the current vertex by
// weight * bias (bias = 1/p for the previous vertex, 1 for a neighbour of the previous vertex, 1/q otherwise), sums them in
// adjacency order and scans `r <= cum`: one word per step. The warp computes the biased weights 32 neighbours at a time
// (membership in the previous vertex' adjacency by binary search in a sorted copy: the reference scans linearly, the answer
// is the same) and lane 0 accumulates them in adjacency order -- the same fp64 additions in the same order as the
// reference, so the boundary decisions are the reference's. `buf`: 32 doubles of shared memory owned by this warp.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_neighbor_sorted(const GraphDev& g, int64_t a, int32_t b) {
    int64_t lo = __ldg(g.row_off + a);
    const int64_t end = __ldg(g.row_off + a + 1);
    int64_t hi = end;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(g.col_sorted + mid) < b) lo = mid + 1;
        else hi = mid;
    }
    return lo < end && __ldg(g.col_sorted + lo) == b;
}

__device__ __forceinline__ double biased_weight(const GraphDev& g, int64_t e, int64_t prev, double pinv, double qinv) {
    const int32_t nb = __ldg(g.col + e);
    const double bias = nb == (int32_t)prev ? pinv : (is_neighbor_sorted(g, prev, nb) ? 1.0 : qinv);
    return __dmul_rn(__ldg(g.w + e), bias);
}

__device__ __forceinline__ int biased_walk(const GraphDev& g, DrawRing& ring, int64_t start, int steps, int32_t* walk,
                                           double* buf, int lane, double pinv, double qinv) {
    if (lane == 0) walk[0] = (int32_t)start;
    if (steps == 0) {
        __syncwarp();
        return 1;
    }
    ring.ensure();
    int64_t first = -1;
    uint32_t used = 0;
    if (lane == 0) {
        int u;
        first = target_sample(g, start, ring.peek(0), ring.peek(1), u);
        used = (uint32_t)u;
    }
    first = __shfl_sync(kFull, first, 0);
    used = __shfl_sync(kFull, used, 0);
    if (first < 0) {
        __syncwarp();
        return 1;
    }
    ring.advance(used);
    if (lane == 0) walk[1] = (int32_t)first;
    int len = 2;
    int64_t prev = start, cur = first;
    for (int i = 1; i < steps; ++i) {
        const int64_t off = __ldg(g.row_off + cur);
        const int64_t deg = __ldg(g.row_off + cur + 1) - off;
        if (deg == 0) break;
        ring.ensure();
        double total = 0.0;
        for (int64_t c0 = 0; c0 < deg; c0 += 32) {
            buf[lane] = c0 + lane < deg ? biased_weight(g, off + c0 + lane, prev, pinv, qinv) : 0.0;
            __syncwarp();
            if (lane == 0) {
                const int n = (int)min((int64_t)32, deg - c0);
                for (int k = 0; k < n; ++k) total = __dadd_rn(total, buf[k]);
            }
            __syncwarp();
        }
        total = __shfl_sync(kFull, total, 0);
        const uint32_t word = ring.peek(0);
        ring.advance(1u);
        int64_t pick = deg - 1;  // falls back to the last neighbour (node2vec.go:161)
        if (total == 0.0) pick = (int64_t)index_draw(word, (uint32_t)deg);
        else {
            const double r = __dmul_rn((double)word * (1.0 / 4294967296.0), total);
            double cum = 0.0;
            for (int64_t c0 = 0; c0 < deg; c0 += 32) {
                buf[lane] = c0 + lane < deg ? biased_weight(g, off + c0 + lane, prev, pinv, qinv) : 0.0;
                __syncwarp();
                int hit = -1;
                if (lane == 0) {
                    const int n = (int)min((int64_t)32, deg - c0);
                    for (int k = 0; k < n; ++k) {
                        cum = __dadd_rn(cum, buf[k]);
                        if (r <= cum) {
                            hit = k;
                            break;
                        }
                    }
                }
                hit = __shfl_sync(kFull, hit, 0);
                __syncwarp();
                if (hit >= 0) {
                    pick = c0 + hit;
                    break;
                }
            }
        }
        const int64_t nxt = (int64_t)__ldg(g.col + off + pick);
        if (lane == 0) walk[len] = (int32_t)nxt;
        ++len;
        prev = cur;
        cur = nxt;
    }
    __syncwarp();
    return len;
}

// SkipGrams' per-centre window draw (src/proNet.cpp:783): reduce[i] = index(window)+1, one word per position.
__device__ __forceinline__ void draw_windows(DrawRing& ring, int len, int window, uint8_t* reduce, int lane) {
    for (int i0 = 0; i0 < len; i0 += 128) {
        ring.ensure();
        int nb = min(128, len - i0);
        for (int i = lane; i < nb; i += 32) reduce[i0 + i] = (uint8_t)(index_draw(ring.peek((uint32_t)i), (uint32_t)window) + 1u);
        ring.advance((uint32_t)nb);
    }
    __syncwarp();
}

// DeepWalk / Walklets: DeepWalk::Train (src/model/DeepWalk.cpp:133-150), Walklets::Train (Walklets.cpp:42-60),
// DeepWalk.Train (deepwalk.go:110-134). One warp per walk; pairs are enumerated in-warp and never touch HBM.
#ifndef SMORE_WALK_MINBLOCKS
#define SMORE_WALK_MINBLOCKS 3
#endif
// resident CTAs per SM the register allocator must allow for the walk-type kernels (rows <= 16 B per lane: fp32 dim <= 128)
template <class C>
constexpr int walk_min_blocks() {
    return C::EPL * (int)sizeof(typename C::T) <= 16 ? SMORE_WALK_MINBLOCKS : C::EPL * (int)sizeof(typename C::T) <= 32 ? 2 : 1;
}
template <class C>
__global__ void __launch_bounds__(kBlockThreads, walk_min_blocks<C>()) k_walk(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    int32_t* walks = reinterpret_cast<int32_t*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t) + 1008 * sizeof(T));
    uint8_t* reduces = reinterpret_cast<uint8_t*>(walks + kWarpsPerBlock * kMaxWalkLen);
    const T* lut = stage_lut<T>(a.lut, lut_s);
    const DirectView<T> tv{a.Wv, a.dim}, tc{a.Wc, a.dim};
    int lane = threadIdx.x & 31;
    int wib = threadIdx.x >> 5;
    int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    int32_t* walk = walks + wib * kMaxWalkLen;
    uint8_t* reduce = reduces + wib * kMaxWalkLen;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const bool go = g.sem != 0;
    const int nrows = a.K + 1;
    for (int64_t wi = w; wi < a.n_walks; wi += a.n_warps) {
        int64_t start = (int64_t)__ldg(a.keys + wi);
        // (the Go tree draws no windows, so node2vec's per-warp scratch of 32 doubles shares the bytes of `reduce`)
        int len = a.n2v ? biased_walk(g, ring, start, a.steps, walk, reinterpret_cast<double*>(reduce), lane, a.n2v_pinv, a.n2v_qinv)
                        : random_walk(g, ring, start, a.steps, walk, lane);
        if (!go && !a.walklets) draw_windows(ring, len, a.w1, reduce, lane);
        T alpha = (T)st.alpha;
        for (int i = 0; i < len; ++i) {
            int vi = walk[i];
            // the (up to two) index ranges of contexts for centre i
            int lo[2], hi[2];
            if (a.walklets) {  // ScaleSkipGrams (src/proNet.cpp:939-978)
                lo[0] = max(i - a.w1, 0); hi[0] = max(i - a.w0, 0);
                lo[1] = min(i + a.w0, len - 1); hi[1] = min(i + a.w1, len - 1);
            } else if (!go) {  // SkipGrams (src/proNet.cpp:781-801)
                int r = reduce[i];
                lo[0] = max(i - r, 0); hi[0] = min(i + r, len - 1);
                lo[1] = 1; hi[1] = 0;
            } else {  // pronet.go:314-329
                lo[0] = max(i - a.w1, 0); hi[0] = min(i + a.w1 + 1, len) - 1;
                lo[1] = 1; hi[1] = 0;
            }
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                for (int j = lo[part]; j <= hi[part]; ++j) {
                    if (j == i) continue;
                    ring.ensure();
                    int my = draw_pair_ids(g, ring, 0u, walk[j], a.K, lane);
                    ring.advance(2u * (uint32_t)a.K);
                    if (!go) update_pair_cpp<C>(tv, tc, a.dim, a.same_table != 0, lut, vi, my, nrows, alpha, lane);
                    else update_pair_go<C>(tv, tc, a.dim, a.same_table != 0, false, lut, vi, my, nrows, alpha, lane);
                    st.pairs++;
                }
            }
        }
        st.count++;
        sched_tick(st, a.sched);
        __syncwarp();
    }
    st.pos = ring.pos;
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// HPE. One step of proNet::UpdateCommunity (src/proNet.cpp:3028-3051) with Opt_SigmoidRegSGD
// (:1332-1351): the regularised skip-gram step -- g = label - sigmoid(v.c) carries no alpha, the whole bracket does:
//     loss_vertex += alpha * (g * c - reg * v)        c += alpha * (g * v - reg * c)
// and the vertex row takes loss_vertex at the end of the step. my_id: lane 0 = the step's context, lane 1+n negative n.
// The two tables are distinct objects in the reference (w_vertex / w_context), so only context rows can repeat.
// ---------------------------------------------------------------------------------------------------------------
template <class C, class TV, class TC>
__device__ __forceinline__ void update_community_step(const TV& tv, const TC& tc, int dim, const typename C::T* lut, int v1,
                                                      int my_id, int nrows, typename C::T alpha, typename C::T reg,
                                                      int lane) {
    using T = typename C::T;
    using A = Ar<T>;
    const bool active = lane < nrows;
    const unsigned peers = __match_any_sync(kFull, active ? my_id : (-1 - lane));
    const bool dup = __any_sync(kFull, active && __popc(peers) > 1);
    T* pv = tv.row(v1);
    Row<C> v, back;
    v.load(pv, lane, dim);
    back.zero();
    if (!dup) {
        for (int base = 0; base < nrows; base += kCtxChunk) {
            Row<C> c[kCtxChunk];
            int ids[kCtxChunk];
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                ids[r] = __shfl_sync(kFull, my_id, (base + r) & 31);
                if (base + r < nrows) c[r].load(tc.row(ids[r]), lane, dim);
            }
            T f[kCtxChunk];
            dots<C, kCtxChunk>(v, c, nrows - base, f);
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                if (base + r < nrows) {
                    const T label = (base + r == 0) ? (T)1 : (T)0;
                    const T g = A::sub(label, fast_sigmoid<T>(lut, f[r]));
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        const T ce = c[r].x[e];
                        back.x[e] = A::add(back.x[e], A::mul(alpha, A::sub(A::mul(g, ce), A::mul(reg, v.x[e]))));
                        const T dc = A::mul(alpha, A::sub(A::mul(g, v.x[e]), A::mul(reg, ce)));
                        c[r].x[e] = kAtomicRows<C> ? dc : A::add(ce, dc);
                    }
                    if constexpr (kAtomicRows<C>) row_red_add<C>(tc.row(ids[r]), c[r], lane, dim);
                    else c[r].store(tc.row(ids[r]), lane, dim);
                }
            }
        }
    } else {
        for (int r = 0; r < nrows; ++r) {  // a context repeats inside the step: rows re-read in the reference's order
            const int cid = __shfl_sync(kFull, my_id, r);
            T* pc = tc.row(cid);
            Row<C> c;
            c.load(pc, lane, dim);
            const T label = (r == 0) ? (T)1 : (T)0;
            const T g = A::sub(label, fast_sigmoid<T>(lut, dot(v, c)));
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                const T ce = c.x[e];
                back.x[e] = A::add(back.x[e], A::mul(alpha, A::sub(A::mul(g, ce), A::mul(reg, v.x[e]))));
                const T dc = A::mul(alpha, A::sub(A::mul(g, v.x[e]), A::mul(reg, ce)));
                c.x[e] = kAtomicRows<C> ? dc : A::add(ce, dc);
            }
            if constexpr (kAtomicRows<C>) row_red_add<C>(pc, c, lane, dim);
            else c.store(pc, lane, dim);
        }
    }
    if constexpr (kAtomicRows<C>) row_red_add<C>(pv, back, lane, dim);
    else {
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], back.x[e]);
        v.store(pv, lane, dim);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// MF: proNet::UpdateFactorizedPair (src/proNet.cpp:2591-2614) with Opt_SGD (:991-1012): linear
// prediction, labels +1 / -1, L2 term inside the bracket:
//     g = label - v.c      loss_vertex += alpha * (g * c - reg * v)      c += alpha * (g * v - reg * c)
// MF::Train passes ONE table in both roles (src/model/MF.cpp:78), so a context row may BE the vertex row: the ORDERED
// path re-reads both rows from memory for every context, as the reference's references into the live table do.
// ---------------------------------------------------------------------------------------------------------------
template <class C, class TV, class TC>
__device__ __forceinline__ void update_factorized_pair(const TV& tv, const TC& tc, int dim, bool same_table, int v1,
                                                       int my_id, int nrows, typename C::T alpha, typename C::T reg,
                                                       int lane) {
    using T = typename C::T;
    using A = Ar<T>;
    const bool active = lane < nrows;
    const unsigned peers = __match_any_sync(kFull, active ? my_id : (-1 - lane));
    const bool dup = __any_sync(kFull, (active && __popc(peers) > 1) || (active && same_table && my_id == v1));
    T* pv = tv.row(v1);
    Row<C> back;
    back.zero();
    if (!dup) {
        Row<C> v;
        v.load(pv, lane, dim);
        for (int base = 0; base < nrows; base += kCtxChunk) {
            Row<C> c[kCtxChunk];
            int ids[kCtxChunk];
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                ids[r] = __shfl_sync(kFull, my_id, (base + r) & 31);
                if (base + r < nrows) c[r].load(tc.row(ids[r]), lane, dim);
            }
            T f[kCtxChunk];
            dots<C, kCtxChunk>(v, c, nrows - base, f);
#pragma unroll
            for (int r = 0; r < kCtxChunk; ++r) {
                if (base + r < nrows) {
                    const T label = (base + r == 0) ? (T)1 : (T)-1;
                    const T g = A::sub(label, f[r]);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        const T ce = c[r].x[e];
                        back.x[e] = A::add(back.x[e], A::mul(alpha, A::sub(A::mul(g, ce), A::mul(reg, v.x[e]))));
                        const T dc = A::mul(alpha, A::sub(A::mul(g, v.x[e]), A::mul(reg, ce)));
                        c[r].x[e] = kAtomicRows<C> ? dc : A::add(ce, dc);
                    }
                    if constexpr (kAtomicRows<C>) row_red_add<C>(tc.row(ids[r]), c[r], lane, dim);
                    else c[r].store(tc.row(ids[r]), lane, dim);
                }
            }
        }
        if constexpr (kAtomicRows<C>) row_red_add<C>(pv, back, lane, dim);
        else {
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], back.x[e]);
            v.store(pv, lane, dim);
        }
    } else {
        for (int r = 0; r < nrows; ++r) {
            const int cid = __shfl_sync(kFull, my_id, r);
            T* pc = tc.row(cid);
            Row<C> v, c;
            v.load(pv, lane, dim);
            c.load(pc, lane, dim);
            const T label = (r == 0) ? (T)1 : (T)-1;
            const T g = A::sub(label, dot(v, c));
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                const T ce = c.x[e];
                back.x[e] = A::add(back.x[e], A::mul(alpha, A::sub(A::mul(g, ce), A::mul(reg, v.x[e]))));
                const T dc = A::mul(alpha, A::sub(A::mul(g, v.x[e]), A::mul(reg, ce)));
                c.x[e] = kAtomicRows<C> ? dc : A::add(ce, dc);
            }
            if constexpr (kAtomicRows<C>) row_red_add<C>(pc, c, lane, dim);
            else c.store(pc, lane, dim);
        }
        if constexpr (kAtomicRows<C>) row_red_add<C>(pv, back, lane, dim);
        else {
            Row<C> v;
            v.load(pv, lane, dim);
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) v.x[e] = A::add(v.x[e], back.x[e]);
            v.store(pv, lane, dim);
        }
    }
}

// HPE::Train (src/model/HPE.cpp:121-143): SourceSample, TargetSample, UpdateCommunity(v1, v2: the context walks on for
// walk_steps steps, one positive + K negatives per step), then UpdatePair with the roles swapped (vertex v2, context
// v1). The number of words a sample consumes depends on the walk (a sink ends it), so draws come from the sequential
// ring, one warp = one reference worker.
template <class C>
__global__ void __launch_bounds__(kBlockThreads, walk_min_blocks<C>()) k_hpe(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    const DirectView<T> tv{a.Wv, a.dim}, tc{a.Wc, a.dim};
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int nrows = a.K + 1;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        ring.ensure();
        int64_t v1 = 0, v2 = 0;
        uint32_t used = 0;
        if (lane == 0) {
            v1 = (int64_t)source_sample(g, ring.peek(0), ring.peek(1));
            int u;
            v2 = target_sample(g, v1, ring.peek(2), ring.peek(3), u);
            used = 2u + (uint32_t)u;
        }
        v1 = __shfl_sync(kFull, v1, 0);
        v2 = __shfl_sync(kFull, v2, 0);
        ring.advance(__shfl_sync(kFull, used, 0));
        if (v2 >= 0) {  // (a source always has a neighbour: zero out-weight gets alias probability 0)
            const T alpha = (T)st.alpha;
            int64_t ctx = v2;
            for (int s = 0; s < a.steps; ++s) {
                if (s != 0) {
                    ring.ensure();
                    int64_t nx = -1;
                    uint32_t un = 0;
                    if (lane == 0) {
                        int u;
                        nx = target_sample(g, ctx, ring.peek(0), ring.peek(1), u);
                        un = (uint32_t)u;
                    }
                    ctx = __shfl_sync(kFull, nx, 0);
                    ring.advance(__shfl_sync(kFull, un, 0));
                    if (ctx < 0) break;  // proNet.cpp:3033: a sink ends the walk
                }
                ring.ensure();
                const int my = draw_pair_ids(g, ring, 0u, (int)ctx, a.K, lane);
                ring.advance(2u * (uint32_t)a.K);
                update_community_step<C>(tv, tc, a.dim, lut, (int)v1, my, nrows, alpha, a.lambda, lane);
                st.pairs++;
            }
            ring.ensure();
            const int my = draw_pair_ids(g, ring, 0u, (int)v1, a.K, lane);
            ring.advance(2u * (uint32_t)a.K);
            update_pair_cpp<C>(tv, tc, a.dim, false, lut, (int)v2, my, nrows, alpha, lane);
            st.pairs++;
        }
        st.count++;
        sched_tick(st, a.sched);
        __syncwarp();
    }
    st.pos = ring.pos;
    if (lane == 0) a.state[w] = st;
}

// HPE, fp32 Hogwild throughput path. k_hpe above consumes its stream word for word like the reference worker, which chains
// ~24 dependent memory round trips per sample (source, target, then per step: the next context, K alias entries, the rows).
// None of HPE's draws depends on an embedding, so under Hogwild -- compared with the reference statistically, not word for
// word -- every sample gets a FIXED slice of its warp's stream and the chain is reorganised:
//   * 8 lanes draw 8 consecutive samples at once, each walking its own context chain (source, target, steps - 1 more hops);
//   * the (steps + 1) * K negatives of a sample are looked up at once, one lane each (needs (steps + 1) * K <= 32);
//   * the update steps then run back to back with every id already in registers.
// Slice layout (words): [0,4) source + target, [4, 4 + 2 (steps - 1)) the further hops, then from the next multiple of 4 on
// two words per negative: step 0's K, step 1's K, ..., the final UpdatePair's K.
constexpr int kHpeGroup = 8;
constexpr int kHpeMaxSteps = 6;
__host__ __device__ constexpr int hpe_neg_offset(int steps) { return (4 + 2 * (steps - 1) + 3) & ~3; }
__host__ __device__ constexpr int hpe_slice_words(int steps, int K) { return (hpe_neg_offset(steps) + 2 * (steps + 1) * K + 3) & ~3; }

template <class C>
__global__ void __launch_bounds__(kBlockThreads, walk_min_blocks<C>()) k_hpe_fast(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    const T* lut = stage_lut<T>(a.lut, reinterpret_cast<T*>(smem_raw));
    const DirectView<T> tv{a.Wv, a.dim}, tc{a.Wc, a.dim};
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    const GraphDev& g = a.g;
    const int K = a.K, nrows = a.K + 1, steps = a.steps;
    const uint64_t stream = a.stream_base + (uint64_t)w;
    const uint64_t slice = (uint64_t)hpe_slice_words(steps, K);
    const uint64_t s_base = (st.pos + slice - 1) / slice;  // samples this stream has already served
    const uint32_t neg_blk0 = (uint32_t)hpe_neg_offset(steps) >> 2;
    for (uint64_t it0 = 0; it0 < a.jobs; it0 += kHpeGroup) {
        const int ng = (int)min((uint64_t)kHpeGroup, a.jobs - it0);
        int m1 = -1, mc[kHpeMaxSteps];
#pragma unroll
        for (int s = 0; s < kHpeMaxSteps; ++s) mc[s] = -1;
        if (lane < ng) {
            const uint64_t blk = (slice >> 2) * (s_base + it0 + (uint64_t)lane);
            U4 r = philox_block(a.seed, stream, blk);
            m1 = (int)source_sample(g, r.x, r.y);
            int used;
            int64_t ctx = target_sample(g, (int64_t)m1, r.z, r.w, used);
            mc[0] = (int)ctx;
#pragma unroll
            for (int s = 1; s < kHpeMaxSteps; ++s) {
                if (s < steps && ctx >= 0) {  // (a sink ends the walk: proNet.cpp:3033)
                    if (s & 1) r = philox_block(a.seed, stream, blk + 1u + (uint32_t)((s - 1) >> 1));
                    ctx = target_sample(g, ctx, (s & 1) ? r.x : r.z, (s & 1) ? r.y : r.w, used);
                    mc[s] = (int)ctx;
                }
            }
        }
        for (int k = 0; k < ng; ++k) {
            const int v1 = __shfl_sync(kFull, m1, k);
            int cdist = -1;  // lane s: the context of step s (runtime-indexed by shuffle: one copy of the update code)
#pragma unroll
            for (int s = 0; s < kHpeMaxSteps; ++s) {
                const int c = __shfl_sync(kFull, mc[s], k);
                if (lane == s) cdist = c;
            }
            const int v2 = __shfl_sync(kFull, cdist, 0);
            if (v2 >= 0) {
                // negative `lane` of this sample (lanes < (steps + 1) * K)
                int neg = 0;
                if (lane < (steps + 1) * K) {
                    const U4 r = philox_block(a.seed, stream, (slice >> 2) * (s_base + it0 + (uint64_t)k) + neg_blk0 + (uint32_t)(lane >> 1));
                    neg = (int)negative_sample(g, (lane & 1) ? r.z : r.x, (lane & 1) ? r.w : r.y);
                }
                const T alpha = (T)st.alpha;
                for (int s = 0; s < steps; ++s) {
                    const int c = __shfl_sync(kFull, cdist, s);
                    if (c < 0) break;
                    const int nid = __shfl_sync(kFull, neg, (s * K + lane - 1) & 31);
                    update_community_step<C>(tv, tc, a.dim, lut, v1, lane == 0 ? c : nid, nrows, alpha, a.lambda, lane);
                    st.pairs++;
                }
                const int nid = __shfl_sync(kFull, neg, (steps * K + lane - 1) & 31);
                update_pair_cpp<C>(tv, tc, a.dim, false, lut, v2, lane == 0 ? v1 : nid, nrows, alpha, lane);
                st.pairs++;
            }
            st.count++;
            sched_tick(st, a.sched);
        }
    }
    st.pos = (s_base + a.jobs) * slice;
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// Parity hooks
// ---------------------------------------------------------------------------------------------------------------
// One warp replays n sampler calls on stream (seed, stream). which: 0 source, 1 negative, 2 target(arg), 3 source+target.
static __global__ void k_sample_debug(GraphDev g, int which, uint64_t seed, uint64_t stream, int64_t n, const int64_t* arg,
                               int64_t* out, uint64_t* words_used) {
    __shared__ __align__(16) uint32_t ringbuf[256];
    int lane = threadIdx.x;
    DrawRing ring;
    ring.init(ringbuf, seed, stream, 0, lane);
    for (int64_t i = 0; i < n; ++i) {
        ring.ensure();
        int64_t r0 = -1, r1 = -1;
        uint32_t used = 0;
        if (lane == 0) {
            if (which == 0) { r0 = source_sample(g, ring.peek(0), ring.peek(1)); used = 2; }
            else if (which == 1) { r0 = negative_sample(g, ring.peek(0), ring.peek(1)); used = 2; }
            else if (which == 2) { int u; r0 = target_sample(g, arg[i], ring.peek(0), ring.peek(1), u); used = u; }
            else if (g.edge_at) {  // row-sharded graph: one draw over this rank's edge table
                uint32_t le = alias_pick(g.edge_at, index_draw(ring.peek(0), g.n_edge_local), ring.peek(1));
                r0 = __ldg(g.edge_src + le);
                r1 = __ldg(g.edge_dst + le);
                used = 2;
            } else {
                r0 = source_sample(g, ring.peek(0), ring.peek(1));
                int u;
                r1 = target_sample(g, r0, ring.peek(2), ring.peek(3), u);
                used = 2 + u;
            }
            if (which == 3) { out[2 * i] = r0; out[2 * i + 1] = r1; }
            else out[i] = r0;
        }
        used = __shfl_sync(kFull, used, 0);
        ring.advance(used);
    }
    if (lane == 0 && words_used) *words_used = ring.pos;
}

// One warp: RandomWalk + pair enumeration exactly as k_walk does it, pairs written out instead of applied.
static __global__ void k_walk_debug(GraphDev g, uint64_t seed, uint64_t stream, int64_t start, int steps, int mode, int w0,
                             int w1, int64_t* walk_out, int64_t* walk_len, int64_t* pv, int64_t* pc, int64_t cap,
                             int64_t* n_pairs) {
    __shared__ __align__(16) uint32_t ringbuf[256];
    __shared__ int32_t walk[kMaxWalkLen];
    __shared__ uint8_t reduce[kMaxWalkLen];
    int lane = threadIdx.x;
    DrawRing ring;
    ring.init(ringbuf, seed, stream, 0, lane);
    int len = random_walk(g, ring, start, steps, walk, lane);
    const bool go = g.sem != 0;
    if (!go && mode == 0) draw_windows(ring, len, w0, reduce, lane);
    if (lane == 0) {
        *walk_len = len;
        for (int i = 0; i < len; ++i) walk_out[i] = walk[i];
        int64_t np = 0;
        for (int i = 0; i < len; ++i) {
            int lo[2], hi[2];
            if (mode == 1) {
                lo[0] = max(i - w1, 0); hi[0] = max(i - w0, 0);
                lo[1] = min(i + w0, len - 1); hi[1] = min(i + w1, len - 1);
            } else if (!go) {
                int r = reduce[i];
                lo[0] = max(i - r, 0); hi[0] = min(i + r, len - 1);
                lo[1] = 1; hi[1] = 0;
            } else {
                lo[0] = max(i - w0, 0); hi[0] = min(i + w0 + 1, len) - 1;
                lo[1] = 1; hi[1] = 0;
            }
            for (int part = 0; part < 2; ++part)
                for (int j = lo[part]; j <= hi[part]; ++j) {
                    if (j == i) continue;
                    if (np < cap) { pv[np] = walk[i]; pc[np] = walk[j]; }
                    ++np;
                }
        }
        *n_pairs = np;
    }
}

// Table init: (U - 0.5) / dim, U = k * 2^-32, k = word e of stream (seed, kInitStreamBase + table) for GLOBAL element e
// (row-sharded tables: local row l is global row (l << shift) + rank, so the values do not depend on the sharding).
template <typename T>
__global__ void k_init_table(T* W, int64_t n_elems, int dim, uint64_t seed, uint64_t stream, int random, int shift, int rank,
                             int64_t row0 /* local row of W[0] */) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_elems; i += stride) {
        if (!random) {
            W[i] = (T)0;
            continue;
        }
        const int64_t l = i / dim, d = i - l * dim;
        const uint64_t ge = (uint64_t)((((l + row0) << shift) + rank) * dim + d);
        U4 r = philox_block(seed, stream, ge >> 2);
        const uint32_t lane = (uint32_t)(ge & 3);
        const uint32_t k = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
        W[i] = (T)(((double)k * (1.0 / 4294967296.0) - 0.5) / (double)dim);
    }
}

// Replica refresh: every row of the full-size local replica is pulled from its owner's shard (peer loads over NVLink for
// remote shards), 128 bits per thread. Row bytes must be a multiple of 16.
template <typename T>
struct PeerBases {
    T* b[kMaxWorld];
};
template <typename T>
__global__ void k_refresh_replica(T* replica, PeerBases<T> bases, int shift, int mask, int dim, int64_t V) {
    const int64_t n_vec = (int64_t)dim * (int64_t)sizeof(T) / 16;
    const int64_t total = V * n_vec;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t v = i / n_vec, k = i - v * n_vec;
        const int r = (int)(v & mask);
        T* src_base = bases.b[0];
#pragma unroll
        for (int q = 1; q < kMaxWorld; ++q)
            if (q == r) src_base = bases.b[q];
        const uint4* src = reinterpret_cast<const uint4*>(src_base + (size_t)(v >> shift) * dim) + k;
        reinterpret_cast<uint4*>(replica + (size_t)v * dim)[k] = __ldcg(src);
    }
}

template <typename TD, typename TS>
__global__ void k_convert(TD* dst, const TS* src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (TD)src[i];
}

}  // namespace smore
