// Pairwise-ranking kernels with data-dependent draw counts: WARP and HOP-Rec (warp-per-worker over a DrawRing), plus
// the helpers the ranking updates share. (BPR has a fixed draw count per sample and lives in batch_kernels.cuh.)
//
// Every variant first resolves the draws of a sample (they depend on the alias tables / field array only, never on the
// embeddings), then gathers every row it will touch in one batch (FAST path, rows distinct) or replays the reference's
// in-place order through memory (ORDERED path, some rows coincide).
#pragma once
#include "kernels.cuh"

namespace smore {

enum RankKind { RANK_BPR = 0, RANK_WARP = 1, RANK_HOPREC = 2, RANK_SKEWOPT = 3 };

template <typename T>
__device__ __forceinline__ T ldv(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <typename T>
__device__ __forceinline__ void stv(T* p, T v) { *reinterpret_cast<volatile T*>(p) = v; }

// Calls f(e, idx) for every element this lane owns: e = register slot, idx = element index inside the row.
template <class C, class F>
__device__ __forceinline__ void for_owned(int lane, int dim, F&& f) {
#pragma unroll
    for (int c = 0; c < C::NCH; ++c) {
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) {
            int idx = (c * 32 + lane) * C::VEC + j;
            if (!C::MASKED || idx < dim) f(c * C::VEC + j, idx);
        }
    }
}

// Lane-parallel rejection sampling: attempt t uses ring words [t*wpa, (t+1)*wpa). `draw(w0, w1)` proposes a vertex,
// `ok(v)` accepts. Consumes exactly the words the sequential reference loop would. wpa in {1, 2}.
template <class Draw, class Ok>
__device__ __forceinline__ int64_t reject_sample(DrawRing& ring, int lane, uint32_t wpa, Draw&& draw, Ok&& ok) {
    for (;;) {
        ring.ensure();
        uint32_t w0 = ring.peek((uint32_t)lane * wpa);
        uint32_t w1 = ring.peek((uint32_t)lane * wpa + (wpa - 1u));
        int64_t cand = draw(w0, w1);
        unsigned acc = __ballot_sync(kFull, ok(cand));
        if (acc) {
            int first = __ffs(acc) - 1;
            ring.advance((uint32_t)(first + 1) * wpa);
            return __shfl_sync(kFull, cand, first);
        }
        ring.advance(32u * wpa);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One BPR-style round of the C++ ranking updates, ORDERED flavour (rows re-read from memory, per-element writes in the
// reference's order: src/proNet.cpp:1426-1440 / :1485-1504). Returns false if the round was margin-gated away
// (Opt_FBPRSGD, :1023). verr accumulates in registers (it is a local vector in the reference too).
// ---------------------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ bool ordered_round(typename C::T* pv, typename C::T* pi, typename C::T* pj, int dim, int lane,
                                              const typename C::T* lut, typename C::T alpha, bool gated,
                                              typename C::T margin, Row<C>& verr) {
    using T = typename C::T;
    using A = Ar<T>;
    Row<C> v, ri, rj, cvec;
    v.load_ca(pv, lane, dim);
    ri.load_ca(pi, lane, dim);
    rj.load_ca(pj, lane, dim);
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj.x[e]);
    const T f = dot(v, cvec);
    if (gated && f > margin) return false;
    const T g = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
    const T c = A::mul(alpha, (T)0.0025);
    for_owned<C>(lane, dim, [&](int e, int idx) {
        verr.x[e] = A::madd(verr.x[e], g, cvec.x[e]);
        const T cerr = A::mul(g, v.x[e]);
        stv(pi + idx, A::msub(ldv(pi + idx), c, ldv(pi + idx)));
        stv(pj + idx, A::msub(ldv(pj + idx), c, ldv(pj + idx)));
        stv(pi + idx, A::add(ldv(pi + idx), cerr));
        stv(pj + idx, A::sub(ldv(pj + idx), cerr));
    });
    return true;
}

// One round of proNet::UpdateSBPRPair (src/proNet.cpp:1531-1556) with Opt_SBPRSGD (:1070-1098), through memory in the
// reference's order (the rows of one shared table may coincide): g = (f - xi) / omega is gated at g > 2 and clamped at
// -2, the gradient is sigmoid(-g^eta) * g^(eta-1) / omega * alpha. Returns false when the round was gated away.
template <class C>
__device__ __forceinline__ bool sbpr_round(typename C::T* pv, typename C::T* pi, typename C::T* pj, int dim, int lane,
                                           const typename C::T* lut, typename C::T alpha, typename C::T xi,
                                           typename C::T omega, int eta, Row<C>& verr) {
    using T = typename C::T;
    using A = Ar<T>;
    Row<C> v, ri, rj, cvec;
    v.load(pv, lane, dim);
    ri.load(pi, lane, dim);
    rj.load(pj, lane, dim);
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj.x[e]);
    const T f = dot(v, cvec);
    T g = A::div(A::sub(f, xi), omega);
    if (g > (T)2) return false;
    if (g < (T)-2) g = (T)-2;
    T g_in = (T)1;
    for (int i = 0; i < eta; ++i) g_in = A::mul(g_in, g);
    const T chain = A::div(g_in, g);
    g = A::div(A::mul(fast_sigmoid<T>(lut, A::mul((T)-1, g_in)), chain), omega);
    g = A::mul(g, alpha);
    const T c = A::mul(alpha, (T)0.01);  // alpha*0.01*w evaluates left to right
    for_owned<C>(lane, dim, [&](int e, int idx) {
        verr.x[e] = A::madd(verr.x[e], g, cvec.x[e]);
        const T cerr = A::mul(g, v.x[e]);
        stv(pi + idx, A::msub(ldv(pi + idx), c, ldv(pi + idx)));
        stv(pj + idx, A::msub(ldv(pj + idx), c, ldv(pj + idx)));
        stv(pi + idx, A::add(ldv(pi + idx), cerr));
        stv(pj + idx, A::sub(ldv(pj + idx), cerr));
    });
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// WARP: WARP::Train (src/model/WARP.cpp:85-103) + UpdateWARPPair (src/proNet.cpp:1353-1403): scan up to 32
// negatives, the first one with margin f < 1 triggers a single BPR step on three rows. Words: source (p, idx),
// target (p, idx), negative (idx, p), then (idx, p) per further negative actually scanned.
// Candidates are evaluated in batches of kWarpBatch rows gathered together; rows do not change before the (single)
// update, so speculative gathers beyond the first violator are exact, only wasted.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpBatch = 4;

template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_warp(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    int lane = threadIdx.x & 31;
    int wib = threadIdx.x >> 5;
    int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int dim = a.dim;
    T* W = a.Wv;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        ring.ensure();
        int64_t v1 = -1, v2 = -1;
        if (lane == 0) {
            v1 = (int64_t)source_sample(g, ring.peek(0), ring.peek(1));
            int u;
            v2 = target_sample(g, v1, ring.peek(2), ring.peek(3), u);
        }
        v1 = __shfl_sync(kFull, v1, 0);
        v2 = __shfl_sync(kFull, v2, 0);
        if (v2 < 0) {
            ring.advance(6u);
            continue;
        }
        const T alpha = (T)st.alpha;
        T* pv = W + v1 * dim;
        T* pi = W + v2 * dim;
        Row<C> v, ri;
        v.load_ca(pv, lane, dim);
        ri.load_ca(pi, lane, dim);
        int scanned = 0;  // negatives evaluated so far
        bool hit = false;
        // candidate n lives at ring words 4+2n, 5+2n
        for (int base = 0; base < 32 && !hit; base += (base == 0 ? 1 : kWarpBatch)) {
            const int nb = base == 0 ? 1 : min(kWarpBatch, 32 - base);
            int64_t cand = -1;
            if (lane < nb) cand = (int64_t)negative_sample(g, ring.peek(4u + 2u * (uint32_t)(base + lane)), ring.peek(5u + 2u * (uint32_t)(base + lane)));
            Row<C> cvec[kWarpBatch];  // context_vec = w[i] - w[j] (proNet.cpp:1373)
            int64_t jid[kWarpBatch];
            T f[kWarpBatch];
#pragma unroll
            for (int r = 0; r < kWarpBatch; ++r) {
                jid[r] = __shfl_sync(kFull, cand, r);
                if (r < nb) {
                    cvec[r].load_ca(W + jid[r] * dim, lane, dim);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec[r].x[e] = A::sub(ri.x[e], cvec[r].x[e]);
                }
            }
            dots<C, kWarpBatch>(v, cvec, nb, f);
#pragma unroll
            for (int r = 0; r < kWarpBatch; ++r) {
                if (r < nb && !hit) {
                    ++scanned;
                    if (f[r] < (T)1) {
                        hit = true;
                        const int64_t j = jid[r];
                        T* pj = W + j * dim;
                        const T c = A::mul(alpha, (T)0.0025);
                        // Opt_BPRSGD recomputes f from the same (unchanged) rows, so f[r] is its value (:1380 -> :1059)
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f[r])), alpha);
                        if constexpr (kAtomicRows<C>) {
                            // fp32 tables: the three rows take their deltas with red.global.add; the negative's row is
                            // item - difference, no reload
#pragma unroll
                            for (int e = 0; e < C::EPL; ++e) {
                                const T verr = A::mul(gg, cvec[r].x[e]);
                                const T cerr = A::mul(gg, v.x[e]);
                                const T rje = A::sub(ri.x[e], cvec[r].x[e]);
                                cvec[r].x[e] = A::sub(A::mul(-c, rje), cerr);  // negative row: -c*w - cerr
                                ri.x[e] = A::msub(cerr, c, ri.x[e]);           // item row:     -c*w + cerr
                                v.x[e] = A::msub(verr, c, v.x[e]);             // user row:     -c*w + verr
                            }
                            row_red_add<C>(pi, ri, lane, dim);
                            row_red_add<C>(pj, cvec[r], lane, dim);
                            row_red_add<C>(pv, v, lane, dim);
                        } else if (j != v2 && j != v1 && v1 != v2) {
                            Row<C> rj;
                            rj.load_ca(pj, lane, dim);  // only the difference was kept; the row is an L2 hit now
#pragma unroll
                            for (int e = 0; e < C::EPL; ++e) {
                                const T verr = A::mul(gg, cvec[r].x[e]);
                                const T cerr = A::mul(gg, v.x[e]);
                                ri.x[e] = A::msub(ri.x[e], c, ri.x[e]);
                                rj.x[e] = A::msub(rj.x[e], c, rj.x[e]);
                                v.x[e] = A::msub(v.x[e], c, v.x[e]);
                                ri.x[e] = A::add(ri.x[e], cerr);
                                rj.x[e] = A::sub(rj.x[e], cerr);
                                v.x[e] = A::add(v.x[e], verr);
                            }
                            ri.store(pi, lane, dim);
                            rj.store(pj, lane, dim);
                            v.store(pv, lane, dim);
                        } else {
                            // coinciding rows: the reference's per-element order through memory (proNet.cpp:1382-1391)
                            for_owned<C>(lane, dim, [&](int e, int idx) {
                                const T verr = A::mul(gg, cvec[r].x[e]);
                                const T cerr = A::mul(gg, v.x[e]);
                                stv(pi + idx, A::msub(ldv(pi + idx), c, ldv(pi + idx)));
                                stv(pj + idx, A::msub(ldv(pj + idx), c, ldv(pj + idx)));
                                stv(pv + idx, A::msub(ldv(pv + idx), c, ldv(pv + idx)));
                                stv(pi + idx, A::add(ldv(pi + idx), cerr));
                                stv(pj + idx, A::sub(ldv(pj + idx), cerr));
                                stv(pv + idx, A::add(ldv(pv + idx), verr));
                            });
                        }
                    }
                }
            }
        }
        ring.advance(4u + 2u * (uint32_t)scanned);
        st.count++;
        st.pairs++;
        st.tries += (uint64_t)scanned;
        sched_tick(st, a.sched);
    }
    st.pos = ring.pos;
    if (lane == 0) a.state[w] = st;
}

// ---------------------------------------------------------------------------------------------------------------
// HOP-Rec: HBPR::Train (src/model/HBPR.cpp:93-126) + UpdateFBPRPair (src/proNet.cpp:1458-1515).
// Per sampled user: for hop w = 1..walk_steps: item at hop w, field-matched negative, 5 margin-gated BPR rounds
// (rounds > 0 draw UNIFORM negatives with field rejection), learning rate alpha/w, margin 1/w.
// ---------------------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(kBlockThreads) k_hoprec(TrainArgs<typename C::T> a) {
    using T = typename C::T;
    using A = Ar<T>;
    uint32_t* rings = reinterpret_cast<uint32_t*>(smem_raw);
    T* lut_s = reinterpret_cast<T*>(smem_raw + kWarpsPerBlock * 256 * sizeof(uint32_t));
    const T* lut = stage_lut<T>(a.lut, lut_s);
    int lane = threadIdx.x & 31;
    int wib = threadIdx.x >> 5;
    int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= a.n_warps) return;
    WarpState st = a.state[w];
    DrawRing ring;
    ring.init(rings + wib * 256, a.seed, a.stream_base + (uint64_t)w, st.pos, lane);
    const GraphDev& g = a.g;
    const int dim = a.dim;
    T* W = a.Wv;
    const uint32_t V32 = (uint32_t)g.V;
    for (uint64_t it = 0; it < a.jobs; ++it) {
        // user: SourceSample until field == 0 (HBPR.cpp:97-99)
        int64_t vid = reject_sample(ring, lane, 2u,
                                    [&](uint32_t w0, uint32_t w1) { return (int64_t)source_sample(g, w0, w1); },
                                    [&](int64_t c) { return __ldg(g.field + c) == 0; });
        int64_t cid = -1;
        bool dead = false;
        for (int hop = 1; hop <= a.steps && !dead; ++hop) {
            // item at this hop: 1 TargetSample for hop 1, two more for every further hop (HBPR.cpp:100,105-109)
            ring.ensure();
            int64_t c = cid;
            uint32_t used = 0;
            if (lane == 0) {
                int u;
                if (hop == 1) {
                    c = target_sample(g, vid, ring.peek(0), ring.peek(1), u);
                    used = 2;
                } else {
                    c = target_sample(g, c, ring.peek(0), ring.peek(1), u);
                    used = 2;
                    if (c >= 0) {
                        c = target_sample(g, c, ring.peek(2), ring.peek(3), u);
                        used = 4;
                    }
                }
            }
            cid = __shfl_sync(kFull, c, 0);
            ring.advance(__shfl_sync(kFull, used, 0));
            if (cid < 0) {  // sink: the reference would index row -1; abandon the sample
                dead = true;
                break;
            }
            const int cfield = __ldg(g.field + cid);
            // caller's negative: NegativeSample until same field as the item (HBPR.cpp:110-112)
            int64_t jid[5];
            jid[0] = reject_sample(ring, lane, 2u,
                                   [&](uint32_t w0, uint32_t w1) { return (int64_t)negative_sample(g, w0, w1); },
                                   [&](int64_t x) { return __ldg(g.field + x) == cfield; });
            // rounds 1..4: uniform vertex until same field (proNet.cpp:1477-1482), 1 word per attempt
#pragma unroll
            for (int r = 1; r < 5; ++r)
                jid[r] = reject_sample(ring, lane, 1u,
                                       [&](uint32_t w0, uint32_t) { return (int64_t)index_draw(w0, V32); },
                                       [&](int64_t x) { return __ldg(g.field + x) == cfield; });
            const T alpha = (T)(st.alpha / (double)hop);  // _alpha/w, margin/w (HBPR.cpp:113)
            const T margin = (T)(1.0 / (double)hop);
            const T cdec = A::mul(alpha, (T)0.0025);
            const T cv = A::mul(alpha, (T)0.025);
            T* pv = W + vid * dim;
            T* pi = W + cid * dim;
            bool dup = vid == cid;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                dup = dup || jid[r] == vid || jid[r] == cid;
#pragma unroll
                for (int s = 0; s < r; ++s) dup = dup || jid[r] == jid[s];
            }
            Row<C> verr;
            verr.zero();
            T up = 0;
            if constexpr (kAtomicRows<C>) {
                // fp32 tables: deltas with red.global.add, the item row's running value tracked in registers
                Row<C> v, ri, rj[5], di;
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int r = 0; r < 5; ++r) rj[r].load_ca(W + jid[r] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int r = 0; r < 5; ++r) pin(rj[r]);
                di.zero();
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[r].x[e]);
                    const T f = dot(v, cvec);
                    if (!(f > margin)) {
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
                        up += (T)1;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) {
                            verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                            const T cerr = A::mul(gg, v.x[e]);
                            const T d = A::msub(cerr, cdec, ri.x[e]);
                            di.x[e] = A::add(di.x[e], d);
                            ri.x[e] = A::add(ri.x[e], d);
                            rj[r].x[e] = A::sub(A::mul(-cdec, rj[r].x[e]), cerr);
                        }
                        row_red_add<C>(W + jid[r] * dim, rj[r], lane, dim);
                    }
                }
                if (up > (T)0) {
                    row_red_add<C>(pi, di, lane, dim);
                    const T inv = A::div((T)1, up);
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) v.x[e] = A::msub(A::mul(verr.x[e], inv), cv, v.x[e]);
                    row_red_add<C>(pv, v, lane, dim);
                }
            } else if (!dup) {
                Row<C> v, ri, rj[5];
                v.load_ca(pv, lane, dim);
                ri.load_ca(pi, lane, dim);
#pragma unroll
                for (int r = 0; r < 5; ++r) rj[r].load_ca(W + jid[r] * dim, lane, dim);
                pin(v);
                pin(ri);
#pragma unroll
                for (int r = 0; r < 5; ++r) pin(rj[r]);  // all seven gathers in flight before the first round
                bool ri_dirty = false;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    Row<C> cvec;
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) cvec.x[e] = A::sub(ri.x[e], rj[r].x[e]);
                    const T f = dot(v, cvec);
                    if (!(f > margin)) {
                        const T gg = A::mul(fast_sigmoid<T>(lut, A::sub((T)0, f)), alpha);
                        up += (T)1;
#pragma unroll
                        for (int e = 0; e < C::EPL; ++e) {
                            verr.x[e] = A::madd(verr.x[e], gg, cvec.x[e]);
                            const T cerr = A::mul(gg, v.x[e]);
                            ri.x[e] = A::msub(ri.x[e], cdec, ri.x[e]);
                            rj[r].x[e] = A::msub(rj[r].x[e], cdec, rj[r].x[e]);
                            ri.x[e] = A::add(ri.x[e], cerr);
                            rj[r].x[e] = A::sub(rj[r].x[e], cerr);
                        }
                        rj[r].store(W + jid[r] * dim, lane, dim);
                        ri_dirty = true;
                    }
                }
                if (ri_dirty) ri.store(pi, lane, dim);
                if (up > (T)0) {
#pragma unroll
                    for (int e = 0; e < C::EPL; ++e) {
                        v.x[e] = A::msub(v.x[e], cv, v.x[e]);
                        v.x[e] = A::add(v.x[e], A::div(verr.x[e], up));
                    }
                    v.store(pv, lane, dim);
                }
            } else {
                for (int r = 0; r < 5; ++r)
                    if (ordered_round<C>(pv, pi, W + jid[r] * dim, dim, lane, lut, alpha, true, margin, verr)) up += (T)1;
                if (up > (T)0) {
                    for_owned<C>(lane, dim, [&](int e, int idx) {
                        stv(pv + idx, A::msub(ldv(pv + idx), cv, ldv(pv + idx)));
                        stv(pv + idx, A::add(ldv(pv + idx), A::div(verr.x[e], up)));
                    });
                }
            }
            st.pairs += 5;
        }
        st.count++;
        sched_tick(st, a.sched);
    }
    st.pos = ring.pos;
    if (lane == 0) a.state[w] = st;
}

}  // namespace smore
